// Fused persistent reverse-diffusion sampler: three tiles per SM, eight warps per tile (UPD_IMPL_TCGEN05_X3W).
//
// sampler_tc.cu's tile (128 rows, two warps per TMEM lane quadrant splitting the 128 hidden columns, half<->half
// exchanges through shared memory) combined with sampler_tc3.cu's rotation of four 128-column TMEM buffers among three
// tiles (MMA number m of the CTA-wide sequence reads A from buffer (m+2) % 4 and accumulates into (m+1) % 4, the A buffer
// of MMA m-1, once that MMA has completed).  24 warps = 768 threads at 80 registers, no MUFU turn-taking: with three tiles
// free-running there are always softplus epilogues of other tiles to fill the MUFU pipe while one tile waits for its
// MMAs, and up to six warps per scheduler to hide the dependency chains (the X3 measurements: the more warps share the
// MUFU pipe, the better).  Arithmetic, operand encodings, weight image: identical to sampler_tc.cu, except that the NsDiff
// heads always take the two-pass form (pass 1 leaves L3 in TMEM, pass 2 evaluates the heads outside the MUFU-heavy part):
// at 80 registers the single-pass power sums cost more in spills than the second TMEM read (3.63 -> 3.74 G row-steps/s).
#include "sampler_params.cuh"
#include "tc_helpers.cuh"
#include "upd_common.cuh"

namespace {

constexpr int TC_THREADS = 768;
constexpr uint32_t UMMA_LBO = 2048;   // K-adjacent core matrices (layout in upd_common.cuh)
constexpr uint32_t UMMA_SBO = 128;    // N-adjacent core matrices
#ifndef UPD_HANDOFF_GROUP
#define UPD_HANDOFF_GROUP 4
#endif
constexpr int PP_BAR0 = 13;           // named barriers 13/14: MUFU turn of tile 0 / tile 1 (see mufu_turn_*)
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

struct __align__(8) TcSync {
  unsigned long long wbar;
  unsigned long long mma_bar[3];
  uint32_t tmem_base;
  uint32_t pad;
};

// lg2(1 + 2^z): softplus(z*ln2)/ln2.  Two MUFU ops; z is clamped where the caller cannot bound it.
__device__ __forceinline__ float lg2_1p_ex2(float z) {
  float u, l;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(u) : "f"(z));
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(1.0f + u));
  return l;
}

// The same function with ONE MUFU op: lg2(1 + 2^z) = max(z,0) + lg2(1 + u), u = 2^-|z| in (0,1], and lg2(1+u) as a
// degree-8 minimax polynomial on the FMA pipe (|error| 4e-8 exact, 1.8e-7 in fp32 Horner -- the size of lg2.approx's own
// error on these arguments).  Inside a MUFU turn the epilogue is MUFU-bound (32768 MUFU ops per tile-phase = 2048 clk
// at 16/clk/SM); evaluating UPD_POLY_LG2 of every 4 elements this way trades 1 MUFU op for 10 FMA-pipe instructions.
// MEASURED (B200, bench shape, parity green in every variant): 0 of 4: 3.72 G row-steps/s, 1 of 4: 3.73, 2 of 4: 3.48,
// 3 of 4: 3.29, 4 of 4: 3.14 -- inside a MUFU turn the issue slots are as full as the MUFU pipe, so the trade does not
// pay.  Kept (default off) as the record of that experiment.
#ifndef UPD_POLY_LG2
#define UPD_POLY_LG2 0
#endif
__device__ __forceinline__ float lg2_1p_ex2_poly(float z) {
  float u;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(u) : "f"(-fabsf(z)));
  float p = -9.0889083222e-03f;
  p = fmaf(p, u, 5.1134437323e-02f);
  p = fmaf(p, u, -1.3592693210e-01f);
  p = fmaf(p, u, 2.4041023850e-01f);
  p = fmaf(p, u, -3.4654855728e-01f);
  p = fmaf(p, u, 4.7846421599e-01f);
  p = fmaf(p, u, -7.2113221884e-01f);
  p = fmaf(p, u, 1.4426876307e+00f);
  p = fmaf(p, u, 4.2314418636e-08f);
  return fmaxf(z, 0.0f) + p;
}

// softplus(x) for x in [0,1] without MUFU: x/2 + P(x^2), P = degree-4 near-minimax fit of log(2 cosh(sqrt(u)/2))
// on u in [0,1] (|error| < 4e-9 before fp32 rounding, 1e-7 after).  The sigma head applies softplus to the
// L2-normalised, non-negative hidden vector, whose components always lie in [0,1]; taking those 128 of the 514
// softplus per row-step off the MUFU pipe removes a quarter of the kernel's transcendental work.
constexpr float SPU_C0 = 0.6931471824645996f, SPU_C1 = 0.12499982863664627f, SPU_C2 = -0.005206969100981951f,
                SPU_C3 = 0.0003433137317188084f, SPU_C4 = -2.16761418414535e-05f;

// MUFU hand-off between the two tiles of a CTA.  Left alone the tiles fall into lock-step (measured with clock64
// stamps: both in their softplus epilogue at once, each at half MUFU rate, then both waiting on interleaved MMAs
// with the MUFU pipe idle: 24.8k cycles per step).  A tile therefore takes the MUFU-heavy phases (the three
// softplus epilogues of a step) in turns: wait for the partner to finish its phase, run, hand over.  While one
// tile computes softplus the other has its MMAs, head FMAs and posterior algebra in flight.
__device__ __forceinline__ void mufu_turn_begin(int) {}      // three tiles free-running: no turn-taking in this kernel
__device__ __forceinline__ void mufu_turn_end(int) {}

// One 16-column group of an accumulator -> activations -> fp16 hi/lo A operand, in place.
// FIRST: layer 1 (bias rides in the GEMM).  CLAMP: guard ex2 overflow where inputs are unbounded.
template <bool FIRST, bool CLAMP>
__device__ __forceinline__ float epilogue_group(const uint32_t (&r)[16], uint32_t (&o)[16], const float* __restrict__ e,
                                                const float* __restrict__ b, float inv) {
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 16; j += 4) {
    float4 e4 = *reinterpret_cast<const float4*>(e + j);
    float4 b4 = FIRST ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(b + j);
    // (acc*inv + b) as one FMA: the pre-activation differs from the reference's two roundings by < 1 ulp, far
    // below the reordering of the 128-term sums it comes from
    float z0 = FIRST ? __uint_as_float(r[j]) * e4.x : fmaf(__uint_as_float(r[j]), inv, b4.x) * e4.x;
    float z1 = FIRST ? __uint_as_float(r[j + 1]) * e4.y : fmaf(__uint_as_float(r[j + 1]), inv, b4.y) * e4.y;
    float z2 = FIRST ? __uint_as_float(r[j + 2]) * e4.z : fmaf(__uint_as_float(r[j + 2]), inv, b4.z) * e4.z;
    float z3 = FIRST ? __uint_as_float(r[j + 3]) * e4.w : fmaf(__uint_as_float(r[j + 3]), inv, b4.w) * e4.w;
    if (CLAMP) { z0 = fminf(z0, 126.f); z1 = fminf(z1, 126.f); z2 = fminf(z2, 126.f); z3 = fminf(z3, 126.f); }
    // UPD_POLY_LG2 of these four go through the one-MUFU form (0: none, 1: h3, 2: h1 and h3, 4: all)
    float h0 = (UPD_POLY_LG2 >= 4) ? lg2_1p_ex2_poly(z0) : lg2_1p_ex2(z0);
    float h1 = (UPD_POLY_LG2 >= 2) ? lg2_1p_ex2_poly(z1) : lg2_1p_ex2(z1);
    float h2 = (UPD_POLY_LG2 >= 3) ? lg2_1p_ex2_poly(z2) : lg2_1p_ex2(z2);
    float h3 = (UPD_POLY_LG2 >= 1) ? lg2_1p_ex2_poly(z3) : lg2_1p_ex2(z3);
    ss = fmaf(h0, h0, ss); ss = fmaf(h1, h1, ss); ss = fmaf(h2, h2, ss); ss = fmaf(h3, h3, ss);
    tc::split_f16x2(h0, h1, o[j / 2], o[8 + j / 2]);
    tc::split_f16x2(h2, h3, o[j / 2 + 1], o[8 + j / 2 + 1]);
  }
  return ss;
}

// This warp's 64 columns of one hidden layer: 4 groups, TMEM loads software-pipelined one group ahead.
constexpr int HANDOFF_GROUP = UPD_HANDOFF_GROUP;   // the MUFU turn is handed over after this many of the 4 groups

// UPD_PRELOAD: the first 16-column group of the accumulator is fetched from TMEM BEFORE the tile asks for its MUFU turn
// (the MMAs are complete by then), so that the first MUFU instruction issues right at the hand-over instead of a TMEM
// round trip later.  MEASURED (B200, bench shape, parity green): 3.715 vs 3.721 G row-steps/s without it -- the hand-over
// gap is not the TMEM load.  Kept (default off) as the record of that experiment.
#ifndef UPD_PRELOAD
#define UPD_PRELOAD 0
#endif

template <bool FIRST, bool CLAMP>
__device__ __forceinline__ float epilogue_half(uint32_t buf, const float* __restrict__ e, const float* __restrict__ b,
                                               float inv, int tile_id) {
  float ss = 0.f;
  uint32_t r[16], rn[16], o[16];
  tc::tmem_ld16(buf, r);
  tc::wait_ld();
#if UPD_PRELOAD
  mufu_turn_begin(tile_id);
#endif
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (q < 3) tc::tmem_ld16(buf + 16u * (q + 1), rn);
    ss += epilogue_group<FIRST, CLAMP>(r, o, e + 16 * q, b + 16 * q, inv);
    tc::tmem_st16(buf + 16u * q, o);
    if (q == HANDOFF_GROUP - 1) mufu_turn_end(tile_id);
    if (q < 3) {
      tc::wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = rn[i];
    }
  }
  return ss;
}

#define UPD_STAMP(k) do { } while (0)

template <int KIND, int F>
__global__ void __launch_bounds__(TC_THREADS, 1)
sampler_tc3w_kernel(const UpdSamplerParams p) {
  constexpr bool NS = (KIND == 0);
  constexpr int IN = NS ? 3 * F : 2 * F;
  constexpr int K1 = ((IN + 1 + 7) / 8) * 8;
  const UpdPackLayout L = upd_make_layout(KIND, F, p.T);
  extern __shared__ __align__(128) unsigned char smem[];
  auto sf = [&](uint32_t off) { return reinterpret_cast<float*>(smem + off); };
  constexpr uint32_t STEP_BYTES = NS ? sizeof(UpdNsStep) : sizeof(UpdTmStep);
  const uint32_t steps_off = upd_align128(L.tc_image_bytes);
  const uint32_t xch_off = upd_align128(steps_off + STEP_BYTES * p.T);
  // exchange area per tile: ssx[2 layers][2 halves][128 rows]; headx[1 + 7F][128 rows] = layer-3 sum of squares,
  // F eps sums, F noise draws, 5F sigma-head sums handed from the half-1 warp to the row's owner
  constexpr uint32_t XCH_TILE_FLOATS = (KIND == 0) ? (3 * 2 + 3 * F) * 128 : (2 * 2 + 1 + 7 * F) * 128;
  const uint32_t sync_off = upd_align128(xch_off + 3 * XCH_TILE_FLOATS * 4);
  TcSync* sync = reinterpret_cast<TcSync*>(smem + sync_off);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    tc::mbar_init(tc::smem_u32(&sync->wbar), 1);
    tc::mbar_init(tc::smem_u32(&sync->mma_bar[0]), 1);
    tc::mbar_init(tc::smem_u32(&sync->mma_bar[1]), 1);
    tc::mbar_init(tc::smem_u32(&sync->mma_bar[2]), 1);
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
  for (int i = tid; i < L.TE * 128; i += TC_THREADS) {
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
  // rotating TMEM buffers: m = index of this tile's next MMA in the CTA-wide sequence (see the header)
  long long m = tile_id;
  auto bufcol = [&](long long idx) -> uint32_t { return tmem_base + 128u * (uint32_t)(idx & 3); };
  auto wait_prev_mma = [&]() {       // issuer only: the accumulator of MMA m is the A buffer of MMA m-1
    if (m > 0) {
      const long long pm = m - 1;
      tc::mbar_wait(tc::smem_u32(&sync->mma_bar[pm % 3]), (uint32_t)((pm / 3) & 1));
      tc::fence_after_sync();
    }
  };
  uint32_t acc = 0;                  // this warp's column half of the accumulator it reads next
  const uint32_t bar = tc::smem_u32(&sync->mma_bar[tile_id]);
  const uint32_t img = tc::smem_u32(smem);
  float* ssx = sf(xch_off) + tile_id * XCH_TILE_FLOATS;     // [2][2][128]
  float* headx = ssx + (NS ? 3 : 2) * 2 * 128;   // [1 + 7F][128] (single-pass heads) or [3F][128]
  // named barriers: 1,2 = all 256 threads of a tile (precede every MMA issue); 5..12 = the two warps that share
  // a TMEM lane quadrant (64 threads), for the half<->half exchanges that need no tile-wide rendezvous
  const int full_bar = 1 + tile_id, pair_bar = 4 + tile_id * 4 + quad;    // ids 1..3 and 4..15
  const float inv_ws2 = sf(L.scales)[0] * (NS ? 1.0f : LN2), inv_ws3 = sf(L.scales)[1] * (NS ? 1.0f : LN2);
  const float* e1 = sf(L.e1) + 64 * half;
  const float* e2 = sf(L.e2) + 64 * half;
  const float* e3 = sf(L.e3) + 64 * half;
  const float* b2 = sf(L.b2) + 64 * half;
  const float* b3 = sf(L.b3) + 64 * half;
  const float* w4 = sf(L.w4) + 64 * half;
  const float* wsg = sf(L.ws) + 64 * half;
  // c0 * sum_j ws[f][j]: the constant term of the sigma-head polynomial (see layer-3 epilogue)
  float ws_sum[F];
#pragma unroll
  for (int f = 0; f < F; ++f) {
    float a = 0.f;
    if (NS) for (int j = 0; j < 128; ++j) a += sf(L.ws)[f * 128 + j];
    ws_sum[f] = SPU_C0 * a;
  }

  const long long n_tiles = (p.n_rows + 127) / 128;
  // Every tile slot of every CTA runs the same number of iterations (slots past the end compute on a clamped
  // row and store nothing): the MUFU hand-off below is a strict alternation and must never wait for a
  // partner that has already left.
  const long long n_iters = (n_tiles + 3LL * gridDim.x - 1) / (3LL * gridDim.x);
  for (long long it = 0; it < n_iters; ++it) {
    const long long tile = (it * gridDim.x + blockIdx.x) * 3 + tile_id;
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
          tc::tmem_st16(bufcol(m + 2) + lane_sel, a);
        } else {
          uint32_t a[32];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float v = in[i % K1];
            float hi = tc::to_tf32(v);
            a[i] = __float_as_uint(hi);
            a[16 + i] = __float_as_uint(tc::to_tf32(v - hi));
          }
          tc::tmem_st32(bufcol(m + 2) + lane_sel, a);
        }
        tc::wait_st();
      }
      tc::fence_before_sync();
      tc::named_bar_sync(full_bar, 256);
      UPD_STAMP(1);
      if (issuer) {
        tc::fence_after_sync();
        wait_prev_mma();
        tc::issue_layer_tf32x3(bufcol(m + 1), bufcol(m + 2), K1, img + L.u1hi, img + L.u1lo, UMMA_LBO, UMMA_SBO);
        tc::mma_commit(bar);
      }
      tc::mbar_wait(bar, (uint32_t)((m / 3) & 1));
      acc = bufcol(m + 1) + lane_sel + 64u * half;
      m += 3;
      tc::fence_after_sync();
      UPD_STAMP(2);

      // ---------------- layer 1 epilogue -> A2 (in place, buf1); layer 2 ----------------
#if !UPD_PRELOAD
      mufu_turn_begin(tile_id);
#endif
      UPD_STAMP(13);
      float ss = epilogue_half<true, true>(acc, e1 + t * 128, nullptr, 1.f, tile_id);
      if (NS) ssx[(0 * 2 + half) * 128 + trow] = ss;
      tc::wait_st();
      UPD_STAMP(3);
      tc::fence_before_sync();
      tc::named_bar_sync(full_bar, 256);
      UPD_STAMP(4);
      if (issuer) {
        tc::fence_after_sync();
        wait_prev_mma();
        tc::issue_layer_f16x3_g16(bufcol(m + 1), bufcol(m + 2), img + L.u2hi, img + L.u2lo, UMMA_LBO, UMMA_SBO);
        tc::mma_commit(bar);
      }
      float inv = inv_ws2;
      if (NS) inv = inv_ws2 / fmaxf(sqrtf(ssx[0 * 128 + trow] + ssx[1 * 128 + trow]), 1e-12f);   // F.normalize, folded past the GEMM
      tc::mbar_wait(bar, (uint32_t)((m / 3) & 1));
      acc = bufcol(m + 1) + lane_sel + 64u * half;
      m += 3;
      tc::fence_after_sync();
      UPD_STAMP(5);

      // ---------------- layer 2 epilogue -> A3 (in place, buf0); layer 3 ----------------
#if !UPD_PRELOAD
      mufu_turn_begin(tile_id);
#endif
      UPD_STAMP(14);
      ss = epilogue_half<false, !NS>(acc, e2 + t * 128, b2, inv, tile_id);
      if (NS) ssx[(1 * 2 + half) * 128 + trow] = ss;
      tc::wait_st();
      UPD_STAMP(6);
      tc::fence_before_sync();
      tc::named_bar_sync(full_bar, 256);
      UPD_STAMP(7);
      if (issuer) {
        tc::fence_after_sync();
        wait_prev_mma();
        tc::issue_layer_f16x3_g16(bufcol(m + 1), bufcol(m + 2), img + L.u3hi, img + L.u3lo, UMMA_LBO, UMMA_SBO);
        tc::mma_commit(bar);
      }
      inv = inv_ws3;
      if (NS) inv = inv_ws3 / fmaxf(sqrtf(ssx[2 * 128 + trow] + ssx[3 * 128 + trow]), 1e-12f);
      tc::mbar_wait(bar, (uint32_t)((m / 3) & 1));
      acc = bufcol(m + 1) + lane_sel + 64u * half;
      m += 3;
      tc::fence_after_sync();
      UPD_STAMP(8);

      if constexpr (!NS) {
      // ---------------- layer 3 epilogue + heads (denoise.py:50 / tmdm_model.py:63) ----------------
      // NsDiff heads read hn = h/||h|| (= L/||L||): eps = lin4(hn), sigma = softplus(sigma_lin(softplus(hn))).
      // ||L|| is only known once the whole row is done, so instead of a second pass over the row the layer-3
      // epilogue accumulates everything the heads need as sums that are rescaled afterwards:
      //   lin4(hn)                 = inv3 * sum_j w4_j L_j
      //   sigma_lin(softplus(hn))  = sum_j ws_j (hn_j/2 + P(hn_j^2))          (hn_j in [0,1], P above)
      //                            = inv3/2 * sum_j ws_j L_j + c0 * sum_j ws_j + sum_k c_k inv3^(2k) * sum_j ws_j L_j^(2k)
      // i.e. the weighted power sums M_k = sum_j ws_j L_j^(2k), k = 1..4.  These FMAs ride in the MUFU-bound
      // phase, where issue slots are free; no TMEM round trip of the row, no second pass.
      float pe[F], pb[F], m1[F], m2[F], m3[F], m4[F];
#pragma unroll
      for (int f = 0; f < F; ++f) { pe[f] = 0.f; pb[f] = 0.f; m1[f] = 0.f; m2[f] = 0.f; m3[f] = 0.f; m4[f] = 0.f; }
      const float* e3t = e3 + t * 128;
      ss = 0.f;
#if !UPD_PRELOAD
      mufu_turn_begin(tile_id);
#endif
      UPD_STAMP(15);
      {
        uint32_t r[16], rn[16];
        tc::tmem_ld16(acc, r);
        tc::wait_ld();
#if UPD_PRELOAD
        mufu_turn_begin(tile_id);
#endif
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q < 3) tc::tmem_ld16(acc + 16u * (q + 1), rn);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int c = 16 * q + j;
            float h = lg2_1p_ex2(fminf(fmaf(__uint_as_float(r[j]), inv, b3[c]) * e3t[c], 126.f));
            if (NS) {
              float u = h * h;
              ss += u;
              float u2 = u * u, u3 = u2 * u, u4 = u2 * u2;
#pragma unroll
              for (int f = 0; f < F; ++f) {
                const float wv = wsg[f * 128 + c];
                pe[f] = fmaf(w4[f * 128 + c], h, pe[f]);
                pb[f] = fmaf(wv, h, pb[f]);
                m1[f] = fmaf(wv, u, m1[f]);
                m2[f] = fmaf(wv, u2, m2[f]);
                m3[f] = fmaf(wv, u3, m3[f]);
                m4[f] = fmaf(wv, u4, m4[f]);
              }
            } else {
#pragma unroll
              for (int f = 0; f < F; ++f) pe[f] = fmaf(w4[f * 128 + c], h, pe[f]);
            }
          }
          if (q < 3) {
            tc::wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = rn[i];
          }
        }
      }
      mufu_turn_end(tile_id);
      UPD_STAMP(9);
      const bool last = (t == 0);
      if (!owner) {
        // the half that owns no row state hands over its partial sums and draws this step's noise
        headx[0 * 128 + trow] = ss;
#pragma unroll
        for (int f = 0; f < F; ++f) {
          headx[(1 + f) * 128 + trow] = pe[f];
          headx[(1 + F + f) * 128 + trow] = last ? 0.f : upd_draw(p, ix, f, F, p.T - t);
          if (NS) {
            headx[(1 + 2 * F + f) * 128 + trow] = pb[f];
            headx[(1 + 3 * F + f) * 128 + trow] = m1[f];
            headx[(1 + 4 * F + f) * 128 + trow] = m2[f];
            headx[(1 + 5 * F + f) * 128 + trow] = m3[f];
            headx[(1 + 6 * F + f) * 128 + trow] = m4[f];
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
          const float inv3 = 1.0f / fmaxf(sqrtf(ss + headx[trow]), 1e-12f);
          const float i2 = inv3 * inv3, i4 = i2 * i2;
#pragma unroll
          for (int f = 0; f < F; ++f) {
            float eps = (pe[f] + headx[(1 + f) * 128 + trow]) * inv3 + sf(L.b4)[f];
            float lin = 0.5f * inv3 * (pb[f] + headx[(1 + 2 * F + f) * 128 + trow]);
            float poly = fmaf(i2, SPU_C1 * (m1[f] + headx[(1 + 3 * F + f) * 128 + trow]), ws_sum[f]);
            poly = fmaf(i4, SPU_C2 * (m2[f] + headx[(1 + 4 * F + f) * 128 + trow]), poly);
            poly = fmaf(i4 * i2, SPU_C3 * (m3[f] + headx[(1 + 5 * F + f) * 128 + trow]), poly);
            poly = fmaf(i4 * i4, SPU_C4 * (m4[f] + headx[(1 + 6 * F + f) * 128 + trow]), poly);
            float sig = upd_softplus_accurate(lin + poly + sf(L.bs)[f]);
            y[f] = upd_ns_update(st, y[f], y0h[f], gxv[f], eps, sig, headx[(1 + F + f) * 128 + trow], last);
          }
        } else {
          const UpdTmStep st = reinterpret_cast<const UpdTmStep*>(smem + steps_off)[t];
#pragma unroll
          for (int f = 0; f < F; ++f) {
            float eps = (pe[f] + headx[(1 + f) * 128 + trow]) * LN2 + sf(L.b4)[f];
            y[f] = upd_tm_update(st, y[f], y0h[f], eps, headx[(1 + F + f) * 128 + trow], last);
          }
        }
      }
      } else {
      // ---------------- layer 3 epilogue + heads, two passes (NsDiff, F > 1) ----------------
      // With several features the head sums cost 6F FMAs per element, which would make the MUFU turn issue-bound;
      // here pass 1 (in the MUFU turn) only produces L3 and its sum of squares, and pass 2 (outside the turn,
      // overlapping the other tile's softplus) evaluates the heads on hn = L3/||L3|| re-read from TMEM.
      float pe[F], ps[F];
#pragma unroll
      for (int f = 0; f < F; ++f) { pe[f] = 0.f; ps[f] = 0.f; }
      const float* e3t = e3 + t * 128;
      mufu_turn_begin(tile_id);
      UPD_STAMP(15);
      ss = 0.f;
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        uint32_t r[16];
        tc::tmem_ld16(acc + 16u * q, r);
        tc::wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float h = lg2_1p_ex2(fminf(fmaf(__uint_as_float(r[j]), inv, b3[16 * q + j]) * e3t[16 * q + j], 126.f));
          ss = fmaf(h, h, ss);
          r[j] = __float_as_uint(h);
        }
        tc::tmem_st16(acc + 16u * q, r);
      }
      ssx[(2 * 2 + half) * 128 + trow] = ss;
      mufu_turn_end(tile_id);
      tc::wait_st();
      UPD_STAMP(9);
      tc::named_bar_sync(pair_bar, 64);
      UPD_STAMP(10);
      const float inv3 = 1.0f / fmaxf(sqrtf(ssx[4 * 128 + trow] + ssx[5 * 128 + trow]), 1e-12f);
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        uint32_t r[16];
        tc::tmem_ld16(acc + 16u * q, r);
        tc::wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float hn = __uint_as_float(r[j]) * inv3, u = hn * hn;
          float sp = fmaf(SPU_C4, u, SPU_C3);
          sp = fmaf(sp, u, SPU_C2);
          sp = fmaf(sp, u, SPU_C1);
          sp = fmaf(sp, u, SPU_C0);
          sp = fmaf(0.5f, hn, sp);
#pragma unroll
          for (int f = 0; f < F; ++f) {
            pe[f] = fmaf(w4[f * 128 + 16 * q + j], hn, pe[f]);
            ps[f] = fmaf(wsg[f * 128 + 16 * q + j], sp, ps[f]);
          }
        }
      }
      const bool last = (t == 0);
      UPD_STAMP(11);
      if (!owner) {
#pragma unroll
        for (int f = 0; f < F; ++f) {
          headx[f * 128 + trow] = pe[f];
          headx[(F + f) * 128 + trow] = ps[f];
          headx[(2 * F + f) * 128 + trow] = last ? 0.f : upd_draw(p, ix, f, F, p.T - t);
        }
        __threadfence_block();
        tc::named_bar_arrive(pair_bar, 64);
      } else {
        tc::named_bar_sync(pair_bar, 64);
        const UpdNsStep st = reinterpret_cast<const UpdNsStep*>(smem + steps_off)[t];
#pragma unroll
        for (int f = 0; f < F; ++f) {
          float eps = (pe[f] + headx[f * 128 + trow]) + sf(L.b4)[f];
          float sig = upd_softplus_accurate((ps[f] + headx[(F + f) * 128 + trow]) + sf(L.bs)[f]);
          y[f] = upd_ns_update(st, y[f], y0h[f], gxv[f], eps, sig, headx[(2 * F + f) * 128 + trow], last);
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
  const UpdPackLayout L = upd_make_layout(KIND, F, p.T);
  constexpr uint32_t STEP_BYTES = (KIND == 0) ? sizeof(UpdNsStep) : sizeof(UpdTmStep);
  constexpr uint32_t XCH_TILE_FLOATS = (KIND == 0) ? (3 * 2 + 3 * F) * 128 : (2 * 2 + 1 + 7 * F) * 128;
  size_t smem = upd_align128(upd_align128(upd_align128(L.tc_image_bytes) + STEP_BYTES * p.T) + 3 * XCH_TILE_FLOATS * 4) +
                sizeof(TcSync) + 128;
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  auto kern = sampler_tc3w_kernel<KIND, F>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  long long n_tiles = (p.n_rows + 127) / 128;
  long long ctas = (n_tiles + 2) / 3;
  int grid = (int)(ctas < sms ? ctas : sms);
  if (grid < 1) grid = 1;
  kern<<<grid, TC_THREADS, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace

cudaError_t upd_launch_sampler_tc3w(const UpdSamplerParams& p, int kind, int F, int sms, cudaStream_t stream) {
#define UPD_CASE(KK, FF) if (kind == KK && F == FF) return launch<KK, FF>(p, sms, stream);
  UPD_CASE(0, 1) UPD_CASE(0, 2) UPD_CASE(0, 3) UPD_CASE(0, 4)
  UPD_CASE(1, 1) UPD_CASE(1, 2) UPD_CASE(1, 3) UPD_CASE(1, 4)
#undef UPD_CASE
  return cudaErrorInvalidValue;
}
