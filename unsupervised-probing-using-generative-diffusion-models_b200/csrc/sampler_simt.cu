// fp32 FFMA implementation of the fused reverse-diffusion sampler (UPD_IMPL_SIMT).
//
// Bring-up and cross-check path for the tcgen05 kernel: same inputs, same outputs, every
// contraction in plain fp32 FFMA.  One warp owns ROWS=8 denoiser rows (a row = one
// (window,row,sample,position) with F features) for all T reverse steps; lane l computes hidden
// units l, l+32, l+64, l+96 of each of its rows, weights are read k-major from shared memory
// (conflict-free), activations are broadcast from a per-warp shared tile.
#include "upd_common.cuh"
#include "sampler_params.cuh"

namespace {

constexpr int ROWS = 8;          // rows per warp
constexpr int SIMT_WARPS = 12;   // warps per CTA

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int KIND, int F>
__global__ void __launch_bounds__(SIMT_WARPS * 32, 1)
sampler_simt_kernel(const UpdSamplerParams p) {
  constexpr int IN = (KIND == 1) ? 2 * F : 3 * F;
  constexpr bool NS = (KIND == 0);
  const UpdPackLayout L = upd_make_layout(KIND, F, p.T);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* w2t = reinterpret_cast<float*>(smem_raw);          // [128][128] k-major
  float* w3t = w2t + 128 * 128;
  float* w1t = w3t + 128 * 128;                             // [IN][128]
  float* b1 = w1t + IN * 128;
  float* b2 = b1 + 128;
  float* b3 = b2 + 128;
  float* w4 = b3 + 128;                                     // [F][128]
  float* wsg = w4 + F * 128;                                // [F][128]
  float* sched = wsg + F * 128;                             // [n_sched][T]
  float* hbuf = sched + L.n_sched * p.T;                    // [warps][ROWS][128]
  const unsigned char* blob = reinterpret_cast<const unsigned char*>(p.packed);
  auto gf = [&](uint32_t off) { return reinterpret_cast<const float*>(blob + off); };
  for (int i = threadIdx.x; i < 128 * 128; i += blockDim.x) { w2t[i] = gf(L.w2t)[i]; w3t[i] = gf(L.w3t)[i]; }
  for (int i = threadIdx.x; i < IN * 128; i += blockDim.x) w1t[i] = gf(L.w1t)[i];
  for (int i = threadIdx.x; i < 128; i += blockDim.x) { b1[i] = gf(L.b1)[i]; b2[i] = gf(L.b2)[i]; b3[i] = gf(L.b3)[i]; }
  // the three [TE,128] step-embedding tables stay in global memory (one coalesced 512-byte row per layer and step, an
  // L1/L2 hit): the kernel's shared memory does not grow with T, so every T <= UPD_MAX_T runs here
  for (int i = threadIdx.x; i < F * 128; i += blockDim.x) { w4[i] = gf(L.w4)[i]; wsg[i] = NS ? gf(L.ws)[i] : 0.f; }
  for (int i = threadIdx.x; i < L.n_sched * p.T; i += blockDim.x) sched[i] = gf(L.sched)[i];
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* hb = hbuf + warp * ROWS * 128;
  float b4v[F], bsv[F];
#pragma unroll
  for (int f = 0; f < F; ++f) { b4v[f] = gf(L.b4)[f]; bsv[f] = NS ? gf(L.bs)[f] : 0.f; }

  const long long n_groups = (p.n_rows + ROWS - 1) / ROWS;
  const long long warps_total = (long long)gridDim.x * SIMT_WARPS;
  for (long long g = (long long)blockIdx.x * SIMT_WARPS + warp; g < n_groups; g += warps_total) {
    // lane (r*F+f) owns element f of row r of this group
    const int er = lane / F, ef = lane % F;
    const bool owner = lane < ROWS * F;
    const long long row = g * ROWS + er;
    const bool live = owner && row < p.n_rows;
    UpdRowIndex ix = upd_row_index(p, live ? row : 0);
    float y0h = 0.f, gxv = 1.f, y = 0.f;
    if (live) {
      long long cidx = (ix.r0 * p.O + ix.o) * F + ef;
      y0h = p.y0_hat ? p.y0_hat[cidx] : 0.f;
      gxv = NS ? p.gx[cidx] : 1.f;
      float z = upd_draw(p, ix, ef, F, 0);
      y = NS ? sqrtf(gxv) * z + y0h : z + y0h;
    }
    for (int t = p.T - 1; t >= 0; --t) {
      // ---- layer-1 input tile -> shared ----
      __syncwarp();
      if (owner) {
        hb[er * 128 + ef] = y;
        hb[er * 128 + F + ef] = y0h;
        if (NS) hb[er * 128 + 2 * F + ef] = gxv;
      }
      __syncwarp();
      float acc[ROWS][4];
      // ---- lin1 ----
#pragma unroll
      for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int m = 0; m < 4; ++m) acc[r][m] = 0.f;
#pragma unroll
      for (int i = 0; i < IN; ++i) {
        float w[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) w[m] = w1t[i * 128 + lane + 32 * m];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          float x = hb[r * 128 + i];
#pragma unroll
          for (int m = 0; m < 4; ++m) acc[r][m] = fmaf(x, w[m], acc[r][m]);
        }
      }
      float inv[ROWS];
      const float* bias = b1;
#pragma unroll 1
      for (int layer = 0; layer < 3; ++layer) {
        const float* e = gf(layer == 0 ? L.e1 : (layer == 1 ? L.e2 : L.e3)) + t * 128;
        if (layer > 0) {
          const float* wt = (layer == 1) ? w2t : w3t;
          bias = (layer == 1) ? b2 : b3;
          __syncwarp();
#pragma unroll
          for (int r = 0; r < ROWS; ++r)
#pragma unroll
            for (int m = 0; m < 4; ++m) { hb[r * 128 + lane + 32 * m] = acc[r][m]; acc[r][m] = 0.f; }
          __syncwarp();
#pragma unroll 2
          for (int k = 0; k < 128; k += 4) {
            float w[4][4];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
              for (int m = 0; m < 4; ++m) w[kk][m] = wt[(k + kk) * 128 + lane + 32 * m];
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
              float4 x = *reinterpret_cast<const float4*>(hb + r * 128 + k);
#pragma unroll
              for (int m = 0; m < 4; ++m) {
                acc[r][m] = fmaf(x.x, w[0][m], acc[r][m]);
                acc[r][m] = fmaf(x.y, w[1][m], acc[r][m]);
                acc[r][m] = fmaf(x.z, w[2][m], acc[r][m]);
                acc[r][m] = fmaf(x.w, w[3][m], acc[r][m]);
              }
            }
          }
        }
        // bias, step embedding, softplus, (NsDiff) L2 normalise -- denoise.py:14-20,45-49
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          float ss = 0.f;
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            int j = lane + 32 * m;
            float h = upd_softplus(e[j] * (acc[r][m] + bias[j]));
            acc[r][m] = h;
            ss = fmaf(h, h, ss);
          }
          if (NS) {
            ss = warp_sum(ss);
            inv[r] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
            for (int m = 0; m < 4; ++m) acc[r][m] *= inv[r];
          }
        }
      }
      // ---- heads: eps = lin4(h); sigma = softplus(sigma_lin(softplus(h))) -- denoise.py:50 ----
      float my_eps = 0.f, my_sig = 0.f;
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
#pragma unroll
        for (int f = 0; f < F; ++f) {
          float pe = 0.f, ps = 0.f;
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            int j = lane + 32 * m;
            pe = fmaf(w4[f * 128 + j], acc[r][m], pe);
            if (NS) ps = fmaf(wsg[f * 128 + j], upd_softplus(acc[r][m]), ps);
          }
          pe = warp_sum(pe);
          if (NS) ps = warp_sum(ps);
          if (lane == r * F + f) { my_eps = pe + b4v[f]; my_sig = NS ? upd_softplus_accurate(ps + bsv[f]) : 0.f; }
        }
      }
      // ---- posterior update ----
      if (live) {
        const bool last = (t == 0);
        float z = last ? 0.f : upd_draw(p, ix, ef, F, p.T - t);
        if (NS) {
          UpdNsStep st = upd_ns_step(sched, p.T, t);
          y = upd_ns_update(st, y, y0h, gxv, my_eps, my_sig, z, last);
        } else {
          UpdTmStep st = upd_tm_step(sched, p.T, t);
          y = upd_tm_update(st, y, y0h, my_eps, z, last);
        }
      }
    }
    if (live) p.out[row * F + ef] = y;
  }
}

template <int KIND, int F>
cudaError_t launch(const UpdSamplerParams& p, int sms, cudaStream_t stream) {
  const UpdPackLayout L = upd_make_layout(KIND, F, p.T);
  constexpr int IN = (KIND == 1) ? 2 * F : 3 * F;
  size_t smem = sizeof(float) * (2 * 128 * 128 + IN * 128 + 3 * 128 + 2 * F * 128 +
                                 L.n_sched * p.T + SIMT_WARPS * ROWS * 128);
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  auto kern = sampler_simt_kernel<KIND, F>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  long long groups = (p.n_rows + ROWS - 1) / ROWS;
  long long ctas = (groups + SIMT_WARPS - 1) / SIMT_WARPS;
  int grid = (int)(ctas < sms ? ctas : sms);
  if (grid < 1) grid = 1;
  kern<<<grid, SIMT_WARPS * 32, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace

cudaError_t upd_launch_sampler_simt(const UpdSamplerParams& p, int kind, int F, int sms, cudaStream_t stream) {
#define UPD_CASE(KK, FF) if (kind == KK && F == FF) return launch<KK, FF>(p, sms, stream);
  UPD_CASE(0, 1) UPD_CASE(0, 2) UPD_CASE(0, 3) UPD_CASE(0, 4)
  UPD_CASE(1, 1) UPD_CASE(1, 2) UPD_CASE(1, 3) UPD_CASE(1, 4)
#undef UPD_CASE
  return cudaErrorInvalidValue;
}
