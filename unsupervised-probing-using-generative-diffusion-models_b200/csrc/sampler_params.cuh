// Kernel-side view of one sampler launch and the row <-> (window,row,sample,position) bookkeeping.
#pragma once
#include "upd_common.cuh"

struct UpdSamplerParams {
  const void* packed;      // packed weight blob (device)
  const float* y0_hat;     // [n_win*B, O, F] or nullptr (zeros)
  const float* gx;         // [n_win*B, O, F] (NsDiff only)
  const float* noise;      // nullptr (Philox) or [n_win, K/S, T, B*S, O, F]
  float* out;              // [n_win*B, K, O, F]
  long long n_rows;        // n_win*B*K*O denoiser rows
  int n_win, B, K, S, O, T;
  unsigned long long seed, window_base;
#ifdef UPD_TRACE
  long long* trace;        // debug builds only (scratch/): [16 warps][T][16] clock64 stamps of CTA 0, first tile
#endif
};

// A denoiser row is one (r0 = w*B + b, sample k, position o); rows are numbered in output order
// ((r0*K + k)*O + o) so a tile of consecutive rows writes one contiguous span of `out`.
struct UpdRowIndex { long long r0; int w, b, k, o; };

__device__ __forceinline__ UpdRowIndex upd_row_index(const UpdSamplerParams& p, long long row) {
  UpdRowIndex ix;
  long long rk = row / p.O;
  ix.o = (int)(row - rk * p.O);
  ix.r0 = rk / p.K;
  ix.k = (int)(rk - ix.r0 * p.K);
  ix.w = (int)(ix.r0 / p.B);
  ix.b = (int)(ix.r0 - (long long)ix.w * p.B);
  return ix;
}

// N(0,1) draw number `draw` (0 = y_T, i = reverse step t = T-i) of element f of a row.
// Validation mode reads the tensor exactly where the reference's sequential torch.randn_like
// calls would have put it: chunk c = k / S, tile row b*S + (k % S)  (NsDiff_model.py:227-236,
// SURVEY A.4).  Otherwise Philox keyed by global indices.
__device__ __forceinline__ float upd_draw(const UpdSamplerParams& p, const UpdRowIndex& ix, int f, int F, int draw) {
  if (p.noise != nullptr) {
    int c = ix.k / p.S, s = ix.k - c * p.S;
    long long C = p.K / p.S;
    long long idx = ((((long long)ix.w * C + c) * p.T + draw) * ((long long)p.B * p.S) + ((long long)ix.b * p.S + s));
    idx = (idx * p.O + ix.o) * F + f;
    return p.noise[idx];
  }
  return upd_pick4(upd_gauss4(p.seed, p.window_base + (unsigned long long)ix.w, (uint32_t)ix.b, (uint32_t)ix.k,
                              (uint32_t)(ix.o * F + f), (uint32_t)draw >> 2), draw);
}

// The same stream for a sampler that consumes the draws of an element in order: `cache` holds the current group of four
// and is refilled (one Philox call) when `draw` enters a new group.  Injected noise bypasses it.
__device__ __forceinline__ float upd_draw_cached(const UpdSamplerParams& p, const UpdRowIndex& ix, int f, int F, int draw,
                                                 float4& cache) {
  if (p.noise != nullptr) return upd_draw(p, ix, f, F, draw);
  if ((draw & 3) == 0)
    cache = upd_gauss4(p.seed, p.window_base + (unsigned long long)ix.w, (uint32_t)ix.b, (uint32_t)ix.k,
                       (uint32_t)(ix.o * F + f), (uint32_t)draw >> 2);
  return upd_pick4(cache, draw);
}

cudaError_t upd_launch_sampler_simt(const UpdSamplerParams& p, int kind, int F, int sms, cudaStream_t stream);
// two-tile tcgen05 sampler (sampler_tc.cu): F <= 4; cudaErrorInvalidValue = shape not built / shared memory exceeded
cudaError_t upd_launch_sampler_tc(const UpdSamplerParams& p, int kind, int F, int sms, cudaStream_t stream);
// warp-specialised tcgen05 sampler (sampler_ws.cu): F <= 2; cudaErrorInvalidValue otherwise
cudaError_t upd_launch_sampler_ws(const UpdSamplerParams& p, int kind, int F, int sms, cudaStream_t stream);
