// Fused persistent reverse-diffusion sampler on tcgen05 tensor cores, two row tiles per SM (UPD_IMPL_TCGEN05_X2).
//
// The round-1 orchestration, kept as the tcgen05 path for F = 3, 4 (whose four-way head-sum exchange does not fit the
// warp-specialised kernel's shared memory, sampler_ws.cu) and as an independent cross-check of that kernel: one CTA per
// SM, 2 tiles x 8 warps.  A tile is 128 denoiser rows (row = TMEM lane); each TMEM lane quadrant is served by TWO warps
// that split the 128 hidden columns in halves; per-row reductions (sum of squares for F.normalize, head sums) cross the
// halves through shared memory on the barrier that precedes each MMA anyway.  Each tile ping-pongs between two private
// 128-column TMEM buffers, and the tiles take the MUFU-heavy phases in strict turns (named-barrier hand-off): one tile's
// softplus epilogue runs while the other's MMAs, posterior and operand build are in flight.  Left alone the tiles fall
// into lock-step (both in their softplus epilogue at once, then both waiting on interleaved MMAs with the MUFU pipe
// idle -- measured with clock64 stamps in round 1).  The half-0 warp of a row owns its state (y, y0_hat, gx), the
// posterior algebra and the layer-1 operand; the half-1 warp draws the Philox noise for it.
//
// Arithmetic, operand encodings and the weight image are those of sampler_ws.cu (shared code: sampler_epi.cuh,
// sampler_math.cuh): three chained GEMMs per step (A from TMEM, B from shared memory, fp16/tf32 hi-lo split, three
// passes), base-2 softplus epilogues in packed fp32x2 math re-encoding the activations in place, head sums inside the
// layer-3 epilogue.  What bounds it (profiles/r02_sampler_tc2_*): inside a turn only the active tile's two warps per SMSP
// issue, and they spend 1.5x as many cycles in fixed-latency dependency stalls as issuing: the MUFU pipe stays 74 % busy.
//
// Limits (cudaErrorInvalidValue -> UPD_ERR_UNSUPPORTED / the caller's fallback): F <= 4; T such that the weight image
// with its three [T,128] step-embedding tables fits 227 KB of shared memory (T <= ~40).
#include "sampler_epi.cuh"
#include "sampler_math.cuh"
#include "sampler_params.cuh"
#include "tc_helpers.cuh"
#include "upd_common.cuh"

namespace {
using namespace epi;

constexpr int PP_BAR0 = 13;           // named barriers 13/14: MUFU turn of tile 0 / tile 1

// pairs of a 16-column group that take the one-MUFU (polynomial) softplus: bit i = pair i (layers 1-2 / layer 3)
#ifndef UPD_PMASK12
#define UPD_PMASK12 0x00
#endif
#ifndef UPD_PMASK3
#define UPD_PMASK3 0x00
#endif

struct __align__(8) TcSync {
  unsigned long long wbar;
  unsigned long long mma_bar[2];
  uint32_t tmem_base;
  uint32_t pad;
};

template <int KIND, int F>
struct SamplerShape {
  static constexpr bool NS = (KIND == 0);
  // exchange area per tile (floats per row): 2 layers x 2 halves of sums of squares; layer-3 sum of squares; per
  // feature: eps sum, noise draw and (NsDiff) the four sigma-head sums handed from the half-1 warp to the row's owner
  static constexpr uint32_t XCH_TILE_FLOATS = (2 * 2 + 1 + 2 * F + (NS ? 4 * F : 0)) * 128;
  static constexpr uint32_t STEP_BYTES = NS ? sizeof(UpdNsStep) : sizeof(UpdTmStep);
};

__device__ __forceinline__ void mufu_turn_begin(int tile_id) { tc::named_bar_sync(PP_BAR0 + tile_id, 512); }
__device__ __forceinline__ void mufu_turn_end(int tile_id) { tc::named_bar_arrive(PP_BAR0 + (tile_id ^ 1), 512); }

// This warp's 64 columns of one hidden layer -> activations -> fp16 hi/lo A operand, in place (K-slice j of the next
// operand = hi words in columns [16j,16j+8), lo words in [16j+8,16j+16)); hands the MUFU turn over after the last group.
template <bool FIRST, bool GUARD, bool SUMSQ>
__device__ __forceinline__ float epilogue_half(uint32_t buf, const float* __restrict__ e, const float* __restrict__ b,
                                               float inv, int tile_id) {
  float2 ss2 = make_float2(0.f, 0.f);
  const float2 inv2 = sm::splat(inv);
  uint32_t r[16], rn[16], o[16];
  tc::tmem_ld16(buf, r);
  tc::wait_ld();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (q < 3) tc::tmem_ld16(buf + 16u * (q + 1), rn);
    epilogue_group<FIRST, GUARD, SUMSQ, UPD_PMASK12>(r, o, e + 16 * q, b + 16 * q, inv2, ss2);
    tc::tmem_st16(buf + 16u * q, o);
    if (q == 3) mufu_turn_end(tile_id);
    if (q < 3) {
      tc::wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = rn[i];
    }
  }
  return ss2.x + ss2.y;
}

// Layer-3 epilogue of this warp's 64 columns: activations feed the head sums, nothing is written back.
template <bool NS, int F, bool GUARD>
__device__ __forceinline__ void heads_half(uint32_t buf, const float* __restrict__ e, const float* __restrict__ b,
                                           const float* __restrict__ w4, const float* __restrict__ ws, float inv,
                                           HeadSums<NS, F>& H, int tile_id) {
  const float2 inv2 = sm::splat(inv);
  uint32_t r[16], rn[16];
  tc::tmem_ld16(buf, r);
  tc::wait_ld();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (q < 3) tc::tmem_ld16(buf + 16u * (q + 1), rn);
    heads_group<NS, F, GUARD, UPD_PMASK3>(r, e + 16 * q, b + 16 * q, w4 + 16 * q, ws + 16 * q, inv2, H);
    if (q < 3) {
      tc::wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = rn[i];
    }
  }
  mufu_turn_end(tile_id);
}

#ifdef UPD_TRACE
#define UPD_STAMP(k) do { if (tracing && lane == 0) p.trace[(warp * p.T + (p.T - 1 - t)) * 16 + (k)] = clock64(); } while (0)
#else
#define UPD_STAMP(k) do { } while (0)
#endif

template <int KIND, int F>
__global__ void __launch_bounds__(512, 1)
sampler_tc_kernel(const UpdSamplerParams p) {
  using Shape = SamplerShape<KIND, F>;
  constexpr bool NS = Shape::NS;
  constexpr int TILES = 2, THREADS = 512;
  constexpr int IN = NS ? 3 * F : 2 * F;
  constexpr int K1 = ((IN + 1 + 7) / 8) * 8;
  const UpdPackLayout L = upd_make_layout(KIND, F, p.T);
  extern __shared__ __align__(128) unsigned char smem[];
  auto sf = [&](uint32_t off) { return reinterpret_cast<float*>(smem + off); };
  const uint32_t steps_off = upd_align128(L.tc_image_bytes);
  const uint32_t xch_off = upd_align128(steps_off + Shape::STEP_BYTES * p.T);
  const uint32_t sync_off = upd_align128(xch_off + TILES * Shape::XCH_TILE_FLOATS * 4);
  TcSync* sync = reinterpret_cast<TcSync*>(smem + sync_off);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    tc::mbar_init(tc::smem_u32(&sync->wbar), 1);
    for (int i = 0; i < 2; ++i) tc::mbar_init(tc::smem_u32(&sync->mma_bar[i]), 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc<512>(tc::smem_u32(&sync->tmem_base));
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = sync->tmem_base;
  if (tid == 0) {
    const uint32_t bar = tc::smem_u32(&sync->wbar);
    tc::mbar_expect_tx(bar, L.tc_image_bytes);
    const unsigned char* src = reinterpret_cast<const unsigned char*>(p.packed);
    for (uint32_t off = 0; off < L.tc_image_bytes; off += 16384u) {
      uint32_t n = L.tc_image_bytes - off < 16384u ? L.tc_image_bytes - off : 16384u;
      tc::bulk_g2s(tc::smem_u32(smem + off), src + off, n, bar);
    }
  }
  tc::mbar_wait(tc::smem_u32(&sync->wbar), 0);
  // step-embedding tables to base 2 (e * log2e), per-step posterior scalars: once per CTA
  for (int i = tid; i < L.TE * 128; i += THREADS) {
    sf(L.e1)[i] *= LOG2E; sf(L.e2)[i] *= LOG2E; sf(L.e3)[i] *= LOG2E;
  }
  if (tid < p.T) {
    if (NS) reinterpret_cast<UpdNsStep*>(smem + steps_off)[tid] = upd_ns_step(sf(L.sched), p.T, tid);
    else reinterpret_cast<UpdTmStep*>(smem + steps_off)[tid] = upd_tm_step(sf(L.sched), p.T, tid);
  }
  __syncthreads();

  const int tile_id = warp >> 3, half = (warp >> 2) & 1, quad = warp & 3;
  const int trow = quad * 32 + lane;                        // row within the tile = TMEM lane
  const bool owner = (half == 0);
  const bool issuer = owner && quad == 0 && lane == 0;
  const uint32_t lane_sel = (uint32_t)(quad * 32) << 16;
  // TMEM: each tile owns a private ping-pong pair of 128-column buffers; cur = the one that holds the next A operand
  uint32_t cur = 0;
  auto a_col = [&]() -> uint32_t { return tmem_base + (uint32_t)tile_id * 256u + 128u * cur; };
  auto d_col = [&]() -> uint32_t { return tmem_base + (uint32_t)tile_id * 256u + 128u * (cur ^ 1u); };
  const uint32_t bar = tc::smem_u32(&sync->mma_bar[tile_id]);
  uint32_t phase = 0;
  // every thread: wait for the layer just issued; returns this warp's column half of its accumulator
  auto wait_layer = [&]() -> uint32_t {
    const uint32_t acc = d_col() + lane_sel + 64u * half;
    tc::mbar_wait(bar, phase);
    phase ^= 1u;
    cur ^= 1u;
    tc::fence_after_sync();
    return acc;
  };
  const uint32_t img = tc::smem_u32(smem);
  float* ssx = sf(xch_off) + tile_id * Shape::XCH_TILE_FLOATS;     // [2 layers][2 halves][128]
  float* headx = ssx + 2 * 2 * 128;                                 // [1 + 2F (+ 4F)][128]
  // named barriers: 1..2 = all 256 threads of a tile (precede every MMA issue); 4.. = the two warps that share
  // a TMEM lane quadrant (64 threads), for the half<->half exchanges that need no tile-wide rendezvous
  const int full_bar = 1 + tile_id, pair_bar = 4 + tile_id * 4 + quad;
  const float inv_ws2 = sf(L.scales)[0] * (NS ? 1.0f : LN2), inv_ws3 = sf(L.scales)[1] * (NS ? 1.0f : LN2);
  // NsDiff: |z'| of layers 2 and 3 is bounded by (|W_row| + |b|) |e| log2e because their input is L2-normalised; the
  // packer stores that bound (scales[2]); below 120 the ex2 overflow guard is compiled out of those epilogues.
  const bool guard23 = !NS || !(sf(L.scales)[2] > 0.f && sf(L.scales)[2] < 120.f);
  const float* e1 = sf(L.e1) + 64 * half;
  const float* e2 = sf(L.e2) + 64 * half;
  const float* e3 = sf(L.e3) + 64 * half;
  const float* b2 = sf(L.b2) + 64 * half;
  const float* b3 = sf(L.b3) + 64 * half;
  const float* w4 = sf(L.w4) + 64 * half;
  const float* wsg = sf(L.ws) + 64 * half;
  // ln2 * sum_j ws[f][j]: the constant term of the sigma-head polynomial
  float ws_sum[F];
#pragma unroll
  for (int f = 0; f < F; ++f) {
    float a = 0.f;
    if (NS) for (int j = 0; j < 128; ++j) a += sf(L.ws)[f * 128 + j];
    ws_sum[f] = sm::SPH_LN2 * a;
  }

  const long long n_tiles = (p.n_rows + 127) / 128;
#ifdef UPD_TRACE
  bool tracing = false;
#endif
  // Every tile slot of every CTA runs the same number of iterations (slots past the end compute on a clamped
  // row and store nothing): the MUFU hand-off is a strict alternation and must never wait for a partner that has left.
  const long long n_iters = (n_tiles + (long long)TILES * gridDim.x - 1) / ((long long)TILES * gridDim.x);
  if (tile_id == 1) tc::named_bar_arrive(PP_BAR0, 512);          // tile 0 takes the first turn
  for (long long it = 0; it < n_iters; ++it) {
    const long long tile = (it * gridDim.x + blockIdx.x) * TILES + tile_id;
    const long long row = tile * 128 + trow;
    const bool live = row < p.n_rows;
    UpdRowIndex ix = upd_row_index(p, live ? row : p.n_rows - 1);
    float y[F], y0h[F], gxv[F];
#pragma unroll
    for (int f = 0; f < F; ++f) { y[f] = 0.f; y0h[f] = 0.f; gxv[f] = 1.f; }
    if (owner) {
      const long long cidx = (ix.r0 * p.O + ix.o) * F;
#pragma unroll
      for (int f = 0; f < F; ++f) {
        y0h[f] = p.y0_hat ? p.y0_hat[cidx + f] : 0.f;
        gxv[f] = NS ? p.gx[cidx + f] : 1.f;
        float z = upd_draw(p, ix, f, F, 0);
        y[f] = NS ? sqrtf(gxv[f]) * z + y0h[f] : z + y0h[f];       // nsdiff_utils.py:274 / tmdm_diffusion_utils.py:110
      }
    }
#ifdef UPD_TRACE
    tracing = (p.trace != nullptr) && blockIdx.x == 0 && it == 1;
#endif
    for (int t = p.T - 1; t >= 0; --t) {
      UPD_STAMP(0);
      // ---------------- layer 1: A1 = [y | y0_hat | gx | 1 | 0] as tf32 hi/lo (owner warps) ----------------
      if (owner) {
        float in[K1];
#pragma unroll
        for (int i = 0; i < K1; ++i) in[i] = 0.f;
#pragma unroll
        for (int f = 0; f < F; ++f) {
          in[f] = y[f];
          in[F + f] = y0h[f];
          if (NS) in[2 * F + f] = gxv[f];
        }
        in[IN] = 1.0f;
        if (K1 == 8) {
          uint32_t a[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float hi = tc::to_tf32(in[i]);
            a[i] = __float_as_uint(hi);
            a[8 + i] = __float_as_uint(tc::to_tf32(in[i] - hi));
          }
          tc::tmem_st16(a_col() + lane_sel, a);
        } else {
          uint32_t a[32];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float v = in[i % K1];
            float hi = tc::to_tf32(v);
            a[i] = __float_as_uint(hi);
            a[16 + i] = __float_as_uint(tc::to_tf32(v - hi));
          }
          tc::tmem_st32(a_col() + lane_sel, a);
        }
        tc::wait_st();
      }
      tc::fence_before_sync();
      tc::named_bar_sync(full_bar, 256);
      UPD_STAMP(1);
      if (issuer) {
        tc::fence_after_sync();
        tc::issue_layer_tf32x3(d_col(), a_col(), K1, img + L.u1hi, img + L.u1lo, UMMA_LBO, UMMA_SBO);
        tc::mma_commit(bar);
      }
      uint32_t acc = wait_layer();
      UPD_STAMP(2);

      // ---------------- layer 1 epilogue -> A2 (in place); layer 2 ----------------
      mufu_turn_begin(tile_id);
      UPD_STAMP(13);
      float ss = epilogue_half<true, true, NS>(acc, e1 + t * 128, nullptr, 1.f, tile_id);
      if (NS) ssx[(0 * 2 + half) * 128 + trow] = ss;
      tc::wait_st();
      UPD_STAMP(3);
      tc::fence_before_sync();
      tc::named_bar_sync(full_bar, 256);
      UPD_STAMP(4);
      if (issuer) {
        tc::fence_after_sync();
        tc::issue_layer_f16x3_g16(d_col(), a_col(), img + L.u2hi, img + L.u2lo, UMMA_LBO, UMMA_SBO);
        tc::mma_commit(bar);
      }
      float inv = inv_ws2;
      if (NS) inv = inv_ws2 / fmaxf(sqrtf(ssx[0 * 128 + trow] + ssx[1 * 128 + trow]), 1e-12f);   // F.normalize, folded past the GEMM
      acc = wait_layer();
      UPD_STAMP(5);

      // ---------------- layer 2 epilogue -> A3 (in place); layer 3 ----------------
      mufu_turn_begin(tile_id);
      UPD_STAMP(14);
      if (guard23) ss = epilogue_half<false, true, NS>(acc, e2 + t * 128, b2, inv, tile_id);
      else ss = epilogue_half<false, false, NS>(acc, e2 + t * 128, b2, inv, tile_id);
      if (NS) ssx[(1 * 2 + half) * 128 + trow] = ss;
      tc::wait_st();
      UPD_STAMP(6);
      tc::fence_before_sync();
      tc::named_bar_sync(full_bar, 256);
      UPD_STAMP(7);
      if (issuer) {
        tc::fence_after_sync();
        tc::issue_layer_f16x3_g16(d_col(), a_col(), img + L.u3hi, img + L.u3lo, UMMA_LBO, UMMA_SBO);
        tc::mma_commit(bar);
      }
      inv = inv_ws3;
      if (NS) inv = inv_ws3 / fmaxf(sqrtf(ssx[2 * 128 + trow] + ssx[3 * 128 + trow]), 1e-12f);
      acc = wait_layer();
      UPD_STAMP(8);

      // ---------------- layer 3 epilogue + heads (denoise.py:50 / tmdm_model.py:63) ----------------
      // NsDiff heads read hn = h/||h|| (= L/||L||): eps = lin4(hn), sigma = softplus(sigma_lin(softplus(hn))).
      // ||L|| is only known once the whole row is done, so the layer-3 epilogue accumulates everything the heads need
      // as sums that are rescaled afterwards:
      //   lin4(hn)                 = inv3 * sum_j w4_j L_j
      //   sigma_lin(softplus(hn))  = sum_j ws_j (hn_j/2 + ln2 + u_j/8 + C2 u_j^2 + C3 u_j^3),   u_j = hn_j^2
      //                            = inv3/2 * PB + ln2 * sum_j ws_j + inv3^2/8 * M1 + C2 inv3^4 * M2 + C3 inv3^6 * M3
      // with PB = sum_j ws_j L_j and the weighted power sums M_k = sum_j ws_j L_j^(2k).
      HeadSums<NS, F> hs;
      hs.clear();
      mufu_turn_begin(tile_id);
      UPD_STAMP(15);
      if (guard23) heads_half<NS, F, true>(acc, e3 + t * 128, b3, w4, wsg, inv, hs, tile_id);
      else heads_half<NS, F, false>(acc, e3 + t * 128, b3, w4, wsg, inv, hs, tile_id);
      UPD_STAMP(9);
      const bool last = (t == 0);
      if (!owner) {
        // the half that owns no row state hands over its partial sums and draws this step's noise
        headx[0 * 128 + trow] = hs.ss.x + hs.ss.y;
#pragma unroll
        for (int f = 0; f < F; ++f) {
          headx[(1 + f) * 128 + trow] = hs.pe[f].x + hs.pe[f].y;
          headx[(1 + F + f) * 128 + trow] = last ? 0.f : upd_draw(p, ix, f, F, p.T - t);
          if (NS) {
            headx[(1 + 2 * F + f) * 128 + trow] = hs.pb[f].x + hs.pb[f].y;
            headx[(1 + 3 * F + f) * 128 + trow] = hs.m1[f].x + hs.m1[f].y;
            headx[(1 + 4 * F + f) * 128 + trow] = hs.m2[f].x + hs.m2[f].y;
            headx[(1 + 5 * F + f) * 128 + trow] = hs.m3[f].x + hs.m3[f].y;
          }
        }
        __threadfence_block();
        tc::named_bar_arrive(pair_bar, 64);
      } else {
        tc::named_bar_sync(pair_bar, 64);
        UPD_STAMP(10);
        // ---------------- posterior update (owner warps) ----------------
        if (NS) {
          const UpdNsStep st = reinterpret_cast<const UpdNsStep*>(smem + steps_off)[t];
          const float ss3 = (hs.ss.x + hs.ss.y) + headx[trow];
#pragma unroll
          for (int f = 0; f < F; ++f) {
            float eps, sig;
            ns_heads(ss3, (hs.pe[f].x + hs.pe[f].y) + headx[(1 + f) * 128 + trow],
                     (hs.pb[f].x + hs.pb[f].y) + headx[(1 + 2 * F + f) * 128 + trow],
                     (hs.m1[f].x + hs.m1[f].y) + headx[(1 + 3 * F + f) * 128 + trow],
                     (hs.m2[f].x + hs.m2[f].y) + headx[(1 + 4 * F + f) * 128 + trow],
                     (hs.m3[f].x + hs.m3[f].y) + headx[(1 + 5 * F + f) * 128 + trow], ws_sum[f], sf(L.b4)[f], sf(L.bs)[f],
                     eps, sig);
            y[f] = upd_ns_update(st, y[f], y0h[f], gxv[f], eps, sig, headx[(1 + F + f) * 128 + trow], last);
          }
        } else {
          const UpdTmStep st = reinterpret_cast<const UpdTmStep*>(smem + steps_off)[t];
#pragma unroll
          for (int f = 0; f < F; ++f) {
            float eps = ((hs.pe[f].x + hs.pe[f].y) + headx[(1 + f) * 128 + trow]) * LN2 + sf(L.b4)[f];
            y[f] = upd_tm_update(st, y[f], y0h[f], eps, headx[(1 + F + f) * 128 + trow], last);
          }
        }
      }
      UPD_STAMP(12);
    }
    if (owner && live) {
#pragma unroll
      for (int f = 0; f < F; ++f) p.out[row * F + f] = y[f];
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem_base);
}

template <int KIND, int F>
cudaError_t launch(const UpdSamplerParams& p, int sms, cudaStream_t stream) {
  using Shape = SamplerShape<KIND, F>;
  const UpdPackLayout L = upd_make_layout(KIND, F, p.T);
  size_t smem = upd_align128(upd_align128(upd_align128(L.tc_image_bytes) + Shape::STEP_BYTES * p.T) +
                             2 * Shape::XCH_TILE_FLOATS * 4) + sizeof(TcSync) + 128;
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  auto kern = sampler_tc_kernel<KIND, F>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  long long n_tiles = (p.n_rows + 127) / 128;
  long long ctas = (n_tiles + 1) / 2;
  int grid = (int)(ctas < sms ? ctas : sms);
  if (grid < 1) grid = 1;
  kern<<<grid, 512, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace

cudaError_t upd_launch_sampler_tc(const UpdSamplerParams& p, int kind, int F, int sms, cudaStream_t stream) {
#define UPD_CASE(KK, FF) if (kind == KK && F == FF) return launch<KK, FF>(p, sms, stream);
  UPD_CASE(0, 1) UPD_CASE(0, 2) UPD_CASE(0, 3) UPD_CASE(0, 4)
  UPD_CASE(1, 1) UPD_CASE(1, 2) UPD_CASE(1, 3) UPD_CASE(1, 4)
#undef UPD_CASE
  return cudaErrorInvalidValue;
}
