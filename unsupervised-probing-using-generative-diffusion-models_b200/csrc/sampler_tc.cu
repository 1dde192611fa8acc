// Fused persistent reverse-diffusion sampler on tcgen05 tensor cores (UPD_IMPL_TCGEN05 / UPD_IMPL_TCGEN05_X3W).
//
// One launch carries every (window,row,sample,position) of a sweep through all T reverse steps without leaving
// the SM: weights resident in shared memory (bulk-copied once per CTA), activations in TMEM, three chained GEMMs per
// step on tcgen05 (A from TMEM, B from shared memory, fp16/tf32 hi-lo split with three passes = fp32-grade accuracy),
// NsDiff / TMDM posterior algebra and Philox noise in registers.  It is organised around the pipe that actually
// bounds this MLP -- MUFU (386 softplus = 772 ex2/lg2 per denoiser row-step) -- not the tensor pipe:
//
//   * one CTA per SM, TILES x 8 warps.  A tile is 128 denoiser rows (row = TMEM lane).  Each TMEM lane quadrant is
//     served by TWO warps that split the 128 hidden columns in halves; per-row reductions (sum of squares for
//     F.normalize, head sums) cross the halves through shared memory on the barrier that precedes each MMA anyway.
//   * TILES = 2 (UPD_IMPL_TCGEN05): each tile ping-pongs between two private 128-column TMEM buffers and the tiles
//     take the MUFU-heavy phases in strict turns (named-barrier hand-off): one tile's softplus epilogue runs while the
//     other's MMAs, posterior and operand build are in flight.
//   * TILES = 3 (UPD_IMPL_TCGEN05_X3W): TMEM (512 columns) has no room for three ping-pong pairs, but a tile needs both
//     buffers only while its MMAs are in flight, so four 128-column buffers rotate among three tiles: MMA number m of
//     the CTA-wide sequence (tile m % 3) reads A from buffer (m+2) % 4 and accumulates into (m+1) % 4 -- the A buffer
//     of MMA m-1, free once that MMA has completed, which the issuer checks on the other tile's mbarrier.  24 warps,
//     free-running.
//   * softplus is evaluated in base 2: L = lg2(1 + 2^z'), z' = (acc*inv + b) * (e*log2e).  NsDiff L2-normalises every
//     hidden layer, so the factor ln2 between softplus and L cancels; for TMDM and the heads it is folded into the
//     scalar applied to the next accumulator.  The epilogue arithmetic is packed fp32x2 (sampler_math.cuh): 6-7 warp
//     instructions per hidden element, and a compile-time share of the elements (UPD_PMASK*) evaluates lg2(1+u) as a
//     polynomial on the FMA pipe so that the MUFU pipe and the issue port run out together.
//   * hidden activations are re-encoded IN PLACE over the accumulator columns they came from, 16 columns at a time:
//     K-slice j of the next A operand = fp16 hi in columns [16j,16j+8), lo in [16j+8,16j+16).  The normalisation of
//     layer l is a scalar applied to the accumulator of layer l+1 (W(h/|h|) = (Wh)/|h|).
//   * the sigma head's inner softplus acts on the L2-normalised hidden vector (components in [0,1]): a polynomial,
//     and because a polynomial of hn = L/||L|| is a sum of power sums of L, both heads are accumulated inside the
//     layer-3 epilogue and rescaled once ||L|| is known -- no second pass over the row.
//   * row state (y, y0_hat, gx), the posterior algebra and the A1 operand belong to the half-0 warp of each row; the
//     half-1 warp draws the Philox noise for it while it would otherwise idle.
//
// Limits (surfaced as UPD_ERR_UNSUPPORTED by the launcher, include/upd_b200.h): F <= 4; T such that the weight image
// with its three [T,128] step-embedding tables fits 227 KB of shared memory (T <= ~40 with two tiles).
#include "sampler_math.cuh"
#include "sampler_params.cuh"
#include "tc_helpers.cuh"
#include "upd_common.cuh"

namespace {

constexpr uint32_t UMMA_LBO = 2048;   // K-adjacent core matrices (layout in upd_common.cuh)
constexpr uint32_t UMMA_SBO = 128;    // N-adjacent core matrices
constexpr int PP_BAR0 = 13;           // named barriers 13/14: MUFU turn of tile 0 / tile 1 (two-tile kernel only)
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

// Which of the 8 column pairs of a 16-column group take the one-MUFU (polynomial) softplus: bit i = pair i.
// Layers 1 and 2 (plain epilogue) and layer 3 (epilogue + head sums) are balanced separately.
#ifndef UPD_PMASK12
#define UPD_PMASK12 0x00
#endif
#ifndef UPD_PMASK3
#define UPD_PMASK3 0x00
#endif
// MUFU turn-taking of the two-tile kernel (0 = free-running)
#ifndef UPD_TURNS
#define UPD_TURNS 1
#endif

struct __align__(8) TcSync {
  unsigned long long wbar;
  unsigned long long mma_bar[3];
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

// z' of a pair of columns.  FIRST: layer 1 (bias rides in the GEMM).
template <bool FIRST>
__device__ __forceinline__ float2 preact2(uint32_t a0, uint32_t a1, float2 inv2, float2 b, float2 e) {
  const float2 acc = make_float2(__uint_as_float(a0), __uint_as_float(a1));
  // (acc*inv + b) as one FMA: the pre-activation differs from the reference's two roundings by < 1 ulp, far below the
  // reordering of the 128-term sums it comes from
  return FIRST ? sm::fmul2(acc, e) : sm::fmul2(sm::ffma2(acc, inv2, b), e);
}

// The epilogues are software pipelines over UNITS of 8 accumulator columns (4 pairs) with three stages,
//   P: accumulator -> pre-activation z' (FMA pipe),   A: z' -> w (ex2),   B: w -> L (lg2 or its polynomial), consume,
// arranged as a ROLLED loop whose body holds B(i), A(i+1), P(i+2), B(i+1), A(i+2), P(i+3): every MUFU instruction of the
// body has its input ready when the body starts, so a warp's MUFU stream never drains while it waits for a TMEM load, a
// shared-memory operand or an ex2 result.  (The first packed-math version ran the stages of a 16-column group back to
// back, fully unrolled: ptxas sank the TMEM prefetch below the arithmetic, finished one group before touching the next,
// and the two warps of an SMSP fell into lock-step, both in their MUFU-free load/split sections at once -- clock64
// stamps showed 2950 cycles per phase for 2048 cycles of MUFU work.  ptxas does not move code across a loop edge.)
// PMASK: bit i = pair i of the unit takes the one-MUFU (polynomial) softplus.
struct UnitZ { float2 z[4]; };
struct UnitW {
  float2 w[4];   // MUFU form: 1 + 2^z';  polynomial form: 2^-|z'|
  float2 z[4];   // z' (dead, and dropped by the compiler, for unguarded MUFU pairs)
};

template <bool FIRST>
__device__ __forceinline__ void stage_p(const uint32_t* __restrict__ r, const float* __restrict__ e,
                                        const float* __restrict__ b, float2 inv2, UnitZ& Z) {
  const float4 e0 = *reinterpret_cast<const float4*>(e), e1 = *reinterpret_cast<const float4*>(e + 4);
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 b0 = FIRST ? z4 : *reinterpret_cast<const float4*>(b), b1 = FIRST ? z4 : *reinterpret_cast<const float4*>(b + 4);
  Z.z[0] = preact2<FIRST>(r[0], r[1], inv2, make_float2(b0.x, b0.y), make_float2(e0.x, e0.y));
  Z.z[1] = preact2<FIRST>(r[2], r[3], inv2, make_float2(b0.z, b0.w), make_float2(e0.z, e0.w));
  Z.z[2] = preact2<FIRST>(r[4], r[5], inv2, make_float2(b1.x, b1.y), make_float2(e1.x, e1.y));
  Z.z[3] = preact2<FIRST>(r[6], r[7], inv2, make_float2(b1.z, b1.w), make_float2(e1.z, e1.w));
}

template <bool GUARD, int PMASK>
__device__ __forceinline__ void stage_a(const UnitZ& Z, UnitW& W) {
  sm::static_for<4>([&](auto ii) {
    constexpr int i = decltype(ii)::value;
    W.z[i] = Z.z[i];
    if constexpr (((PMASK >> i) & 1) != 0) W.w[i] = sm::softplus2_poly_a(Z.z[i]);
    else W.w[i] = sm::softplus2_mufu_a<GUARD>(Z.z[i]);
  });
}

template <bool GUARD, int PMASK>
__device__ __forceinline__ void stage_b(const UnitW& W, float2 (&h)[4]) {
  sm::static_for<4>([&](auto ii) {
    constexpr int i = decltype(ii)::value;
    if constexpr (((PMASK >> i) & 1) != 0) h[i] = sm::softplus2_poly_b(W.w[i], W.z[i]);
    else h[i] = sm::softplus2_mufu_b<GUARD>(W.w[i], W.z[i]);
  });
}

// MUFU hand-off between the two tiles of a CTA (TILES == 2).  Left alone the tiles fall into lock-step (measured with
// clock64 stamps in round 1: both in their softplus epilogue at once, each at half MUFU rate, then both waiting on
// interleaved MMAs with the MUFU pipe idle).  A tile therefore takes the MUFU-heavy phases (the three softplus
// epilogues of a step) in turns: wait for the partner to finish its phase, run, hand over.
template <int TILES>
__device__ __forceinline__ void mufu_turn_begin(int tile_id) {
  if (TILES == 2 && UPD_TURNS) tc::named_bar_sync(PP_BAR0 + tile_id, 512);
}
template <int TILES>
__device__ __forceinline__ void mufu_turn_end(int tile_id) {
  if (TILES == 2 && UPD_TURNS) tc::named_bar_arrive(PP_BAR0 + (tile_id ^ 1), 512);
}

// Stage B's consumer for the hidden layers: sum of squares + fp16 hi/lo split of one unit, written back in place.
// K-slice j of the next A operand = hi words in columns [16j,16j+8), lo words in [16j+8,16j+16); unit u holds the
// elements 8u .. 8u+7 = words 4(u&1) .. 4(u&1)+3 of slice u/2.
template <bool SUMSQ>
__device__ __forceinline__ void store_unit(uint32_t slice_addr, int odd, const float2 (&h)[4], float2& ss2) {
  uint32_t hi[4], lo[4];
  sm::static_for<4>([&](auto ii) {
    constexpr int i = decltype(ii)::value;
    if (SUMSQ) ss2 = sm::ffma2(h[i], h[i], ss2);
    sm::split_f16x2(h[i].x, h[i].y, hi[i], lo[i]);
  });
  tc::tmem_st4(slice_addr + 4u * odd, hi[0], hi[1], hi[2], hi[3]);
  tc::tmem_st4(slice_addr + 8u + 4u * odd, lo[0], lo[1], lo[2], lo[3]);
}

// This warp's 64 columns of one hidden layer -> activations -> fp16 hi/lo A operand, in place.
template <int TILES, bool FIRST, bool GUARD, bool SUMSQ>
__device__ __forceinline__ float epilogue_half(uint32_t buf, const float* __restrict__ e, const float* __restrict__ b,
                                               float inv, int tile_id) {
  constexpr int PM0 = UPD_PMASK12 & 15, PM1 = (UPD_PMASK12 >> 4) & 15;     // even / odd units
  float2 ss2 = make_float2(0.f, 0.f);
  const float2 inv2 = sm::splat(inv);
  uint32_t r[16];
  UnitZ z0, z1, zn;
  UnitW w0, w1;
  float2 h[4];
  tc::tmem_ld16(buf, r);
  tc::wait_ld();
  stage_p<FIRST>(r, e, b, inv2, z0);
  stage_p<FIRST>(r + 8, e + 8, b + 8, inv2, z1);
  tc::tmem_ld16(buf + 16u, r);
  stage_a<GUARD, PM0>(z0, w0);
#pragma unroll 1
  for (int q = 0; q < 3; ++q) {
    // carried in: w0 = A(unit 2q), z1 = P(unit 2q+1), r <- group q+1 in flight
    tc::wait_ld();
    stage_p<FIRST>(r, e + 16 * (q + 1), b + 16 * (q + 1), inv2, z0);            // unit 2q+2
    stage_p<FIRST>(r + 8, e + 16 * (q + 1) + 8, b + 16 * (q + 1) + 8, inv2, zn); // unit 2q+3
    tc::tmem_ld16_if(q < 2, buf + 16u * (q + 2), r);
    stage_b<GUARD, PM0>(w0, h);
    store_unit<SUMSQ>(buf + 16u * q, 0, h, ss2);
    stage_a<GUARD, PM1>(z1, w1);
    stage_b<GUARD, PM1>(w1, h);
    store_unit<SUMSQ>(buf + 16u * q, 1, h, ss2);
    stage_a<GUARD, PM0>(z0, w0);
    z1 = zn;
  }
  // tail: units 6 and 7
  stage_b<GUARD, PM0>(w0, h);
  store_unit<SUMSQ>(buf + 48u, 0, h, ss2);
  stage_a<GUARD, PM1>(z1, w1);
  mufu_turn_end<TILES>(tile_id);
  stage_b<GUARD, PM1>(w1, h);
  store_unit<SUMSQ>(buf + 48u, 1, h, ss2);
  return ss2.x + ss2.y;
}

// Head sums of one warp's 64 columns (see the header and sampler_math.cuh): pe = sum w4 L, and for NsDiff
// pb = sum ws L, m_k = sum ws L^(2k), ss = sum L^2 -- all as packed pairs (even / odd columns), folded at the end.
template <bool NS, int F>
struct HeadSums {
  float2 ss, pe[F], pb[F], m1[F], m2[F], m3[F];
  __device__ __forceinline__ void clear() {
    ss = make_float2(0.f, 0.f);
#pragma unroll
    for (int f = 0; f < F; ++f) {
      pe[f] = make_float2(0.f, 0.f);
      if (NS) { pb[f] = make_float2(0.f, 0.f); m1[f] = pb[f]; m2[f] = pb[f]; m3[f] = pb[f]; }
    }
  }
  __device__ __forceinline__ void add(float2 h, const float* __restrict__ w4, const float* __restrict__ ws) {
    // w4 / ws point at this pair's two columns of feature 0; features are 128 floats apart
    if (NS) {
      const float2 u = sm::fmul2(h, h);
      ss = sm::fadd2(ss, u);
      const float2 u2 = sm::fmul2(u, u), u3 = sm::fmul2(u2, u);
#pragma unroll
      for (int f = 0; f < F; ++f) {
        const float2 a = *reinterpret_cast<const float2*>(w4 + f * 128);
        const float2 s = *reinterpret_cast<const float2*>(ws + f * 128);
        pe[f] = sm::ffma2(a, h, pe[f]);
        pb[f] = sm::ffma2(s, h, pb[f]);
        m1[f] = sm::ffma2(s, u, m1[f]);
        m2[f] = sm::ffma2(s, u2, m2[f]);
        m3[f] = sm::ffma2(s, u3, m3[f]);
      }
    } else {
#pragma unroll
      for (int f = 0; f < F; ++f) pe[f] = sm::ffma2(*reinterpret_cast<const float2*>(w4 + f * 128), h, pe[f]);
    }
  }
};

// Layer-3 epilogue of this warp's 64 columns: the same pipeline; stage B feeds the head sums, nothing is written back.
template <int TILES, bool NS, int F, bool GUARD>
__device__ __forceinline__ void heads_half(uint32_t buf, const float* __restrict__ e, const float* __restrict__ b,
                                           const float* __restrict__ w4, const float* __restrict__ ws, float inv,
                                           HeadSums<NS, F>& H, int tile_id) {
  constexpr int PM0 = UPD_PMASK3 & 15, PM1 = (UPD_PMASK3 >> 4) & 15;
  const float2 inv2 = sm::splat(inv);
  uint32_t r[16];
  UnitZ z0, z1, zn;
  UnitW w0, w1;
  float2 h[4];
  auto consume = [&](int col) {
    sm::static_for<4>([&](auto ii) {
      constexpr int i = decltype(ii)::value;
      H.add(h[i], w4 + col + 2 * i, ws + col + 2 * i);
    });
  };
  tc::tmem_ld16(buf, r);
  tc::wait_ld();
  stage_p<false>(r, e, b, inv2, z0);
  stage_p<false>(r + 8, e + 8, b + 8, inv2, z1);
  tc::tmem_ld16(buf + 16u, r);
  stage_a<GUARD, PM0>(z0, w0);
#pragma unroll 1
  for (int q = 0; q < 3; ++q) {
    tc::wait_ld();
    stage_p<false>(r, e + 16 * (q + 1), b + 16 * (q + 1), inv2, z0);
    stage_p<false>(r + 8, e + 16 * (q + 1) + 8, b + 16 * (q + 1) + 8, inv2, zn);
    tc::tmem_ld16_if(q < 2, buf + 16u * (q + 2), r);
    stage_b<GUARD, PM0>(w0, h);
    consume(16 * q);
    stage_a<GUARD, PM1>(z1, w1);
    stage_b<GUARD, PM1>(w1, h);
    consume(16 * q + 8);
    stage_a<GUARD, PM0>(z0, w0);
    z1 = zn;
  }
  stage_b<GUARD, PM0>(w0, h);
  consume(48);
  stage_a<GUARD, PM1>(z1, w1);
  mufu_turn_end<TILES>(tile_id);
  stage_b<GUARD, PM1>(w1, h);
  consume(56);
}

#ifdef UPD_TRACE
#define UPD_STAMP(k) do { if (tracing && lane == 0) p.trace[(warp * p.T + (p.T - 1 - t)) * 16 + (k)] = clock64(); } while (0)
#else
#define UPD_STAMP(k) do { } while (0)
#endif

template <int KIND, int F, int TILES>
__global__ void __launch_bounds__(TILES * 256, 1)
sampler_tc_kernel(const UpdSamplerParams p) {
  using Shape = SamplerShape<KIND, F>;
  constexpr bool NS = Shape::NS;
  constexpr bool ROT = (TILES == 3);
  constexpr int THREADS = TILES * 256;
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
    for (int i = 0; i < 3; ++i) tc::mbar_init(tc::smem_u32(&sync->mma_bar[i]), 1);
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
  // TMEM buffers.  Two tiles: private ping-pong pair (cur = the buffer that holds the next A operand).  Three tiles:
  // m = index of this tile's next MMA in the CTA-wide sequence, A in buffer (m+2)%4, accumulator in (m+1)%4.
  long long m = tile_id;
  uint32_t cur = 0;
  auto a_col = [&]() -> uint32_t { return ROT ? tmem_base + 128u * (uint32_t)((m + 2) & 3) : tmem_base + (uint32_t)tile_id * 256u + 128u * cur; };
  auto d_col = [&]() -> uint32_t { return ROT ? tmem_base + 128u * (uint32_t)((m + 1) & 3) : tmem_base + (uint32_t)tile_id * 256u + 128u * (cur ^ 1u); };
  const uint32_t bar = tc::smem_u32(&sync->mma_bar[tile_id]);
  uint32_t phase = 0;
  // issuer: launch one layer's MMAs (A operand complete in a_col) and commit to this tile's mbarrier
  auto wait_prev_mma = [&]() {       // three tiles: the accumulator of MMA m is the A buffer of MMA m-1
    if (ROT && m > 0) {
      const long long pm = m - 1;
      tc::mbar_wait(tc::smem_u32(&sync->mma_bar[pm % 3]), (uint32_t)((pm / 3) & 1));
      tc::fence_after_sync();
    }
  };
  // every thread: wait for the layer just issued; returns this warp's column half of its accumulator
  auto wait_layer = [&]() -> uint32_t {
    const uint32_t acc = d_col() + lane_sel + 64u * half;
    if (ROT) { tc::mbar_wait(bar, (uint32_t)((m / 3) & 1)); m += 3; }
    else { tc::mbar_wait(bar, phase); phase ^= 1u; cur ^= 1u; }
    tc::fence_after_sync();
    return acc;
  };
  const uint32_t img = tc::smem_u32(smem);
  float* ssx = sf(xch_off) + tile_id * Shape::XCH_TILE_FLOATS;     // [2 layers][2 halves][128]
  float* headx = ssx + 2 * 2 * 128;                                 // [1 + 2F (+ 4F)][128]
  // named barriers: 1..3 = all 256 threads of a tile (precede every MMA issue); 4.. = the two warps that share
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
  // row and store nothing): the MUFU hand-off and the buffer rotation are strict alternations and must never wait
  // for a partner that has already left.
  const long long n_iters = (n_tiles + (long long)TILES * gridDim.x - 1) / ((long long)TILES * gridDim.x);
  if (TILES == 2 && UPD_TURNS && tile_id == 1) tc::named_bar_arrive(PP_BAR0, 512);          // tile 0 takes the first turn
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
        wait_prev_mma();
        tc::issue_layer_tf32x3(d_col(), a_col(), K1, img + L.u1hi, img + L.u1lo, UMMA_LBO, UMMA_SBO);
        tc::mma_commit(bar);
      }
      uint32_t acc = wait_layer();
      UPD_STAMP(2);

      // ---------------- layer 1 epilogue -> A2 (in place); layer 2 ----------------
      mufu_turn_begin<TILES>(tile_id);
      UPD_STAMP(13);
      float ss = epilogue_half<TILES, true, true, NS>(acc, e1 + t * 128, nullptr, 1.f, tile_id);
      if (NS) ssx[(0 * 2 + half) * 128 + trow] = ss;
      tc::wait_st();
      UPD_STAMP(3);
      tc::fence_before_sync();
      tc::named_bar_sync(full_bar, 256);
      UPD_STAMP(4);
      if (issuer) {
        tc::fence_after_sync();
        wait_prev_mma();
        tc::issue_layer_f16x3_g16(d_col(), a_col(), img + L.u2hi, img + L.u2lo, UMMA_LBO, UMMA_SBO);
        tc::mma_commit(bar);
      }
      float inv = inv_ws2;
      if (NS) inv = inv_ws2 / fmaxf(sqrtf(ssx[0 * 128 + trow] + ssx[1 * 128 + trow]), 1e-12f);   // F.normalize, folded past the GEMM
      acc = wait_layer();
      UPD_STAMP(5);

      // ---------------- layer 2 epilogue -> A3 (in place); layer 3 ----------------
      mufu_turn_begin<TILES>(tile_id);
      UPD_STAMP(14);
      if (guard23) ss = epilogue_half<TILES, false, true, NS>(acc, e2 + t * 128, b2, inv, tile_id);
      else ss = epilogue_half<TILES, false, false, NS>(acc, e2 + t * 128, b2, inv, tile_id);
      if (NS) ssx[(1 * 2 + half) * 128 + trow] = ss;
      tc::wait_st();
      UPD_STAMP(6);
      tc::fence_before_sync();
      tc::named_bar_sync(full_bar, 256);
      UPD_STAMP(7);
      if (issuer) {
        tc::fence_after_sync();
        wait_prev_mma();
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
      mufu_turn_begin<TILES>(tile_id);
      UPD_STAMP(15);
      if (guard23) heads_half<TILES, NS, F, true>(acc, e3 + t * 128, b3, w4, wsg, inv, hs, tile_id);
      else heads_half<TILES, NS, F, false>(acc, e3 + t * 128, b3, w4, wsg, inv, hs, tile_id);
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
          const float inv3 = 1.0f / fmaxf(sqrtf((hs.ss.x + hs.ss.y) + headx[trow]), 1e-12f);
          const float i2 = inv3 * inv3, i4 = i2 * i2;
#pragma unroll
          for (int f = 0; f < F; ++f) {
            float eps = ((hs.pe[f].x + hs.pe[f].y) + headx[(1 + f) * 128 + trow]) * inv3 + sf(L.b4)[f];
            float lin = 0.5f * inv3 * ((hs.pb[f].x + hs.pb[f].y) + headx[(1 + 2 * F + f) * 128 + trow]);
            float poly = fmaf(i2, sm::SPH_C1 * ((hs.m1[f].x + hs.m1[f].y) + headx[(1 + 3 * F + f) * 128 + trow]), ws_sum[f]);
            poly = fmaf(i4, sm::SPH_C2 * ((hs.m2[f].x + hs.m2[f].y) + headx[(1 + 4 * F + f) * 128 + trow]), poly);
            poly = fmaf(i4 * i2, sm::SPH_C3 * ((hs.m3[f].x + hs.m3[f].y) + headx[(1 + 5 * F + f) * 128 + trow]), poly);
            float sig = upd_softplus_accurate(lin + poly + sf(L.bs)[f]);
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

template <int KIND, int F, int TILES>
cudaError_t launch(const UpdSamplerParams& p, int sms, cudaStream_t stream) {
  using Shape = SamplerShape<KIND, F>;
  const UpdPackLayout L = upd_make_layout(KIND, F, p.T);
  size_t smem = upd_align128(upd_align128(upd_align128(L.tc_image_bytes) + Shape::STEP_BYTES * p.T) +
                             TILES * Shape::XCH_TILE_FLOATS * 4) + sizeof(TcSync) + 128;
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  auto kern = sampler_tc_kernel<KIND, F, TILES>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  long long n_tiles = (p.n_rows + 127) / 128;
  long long ctas = (n_tiles + TILES - 1) / TILES;
  int grid = (int)(ctas < sms ? ctas : sms);
  if (grid < 1) grid = 1;
  kern<<<grid, TILES * 256, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace

cudaError_t upd_launch_sampler_tc(const UpdSamplerParams& p, int kind, int F, int tiles, int sms, cudaStream_t stream) {
#define UPD_CASE(KK, FF, TT) if (kind == KK && F == FF && tiles == TT) return launch<KK, FF, TT>(p, sms, stream);
  UPD_CASE(0, 1, 2) UPD_CASE(0, 2, 2) UPD_CASE(0, 3, 2) UPD_CASE(0, 4, 2)
  UPD_CASE(1, 1, 2) UPD_CASE(1, 2, 2) UPD_CASE(1, 3, 2) UPD_CASE(1, 4, 2)
  UPD_CASE(0, 1, 3) UPD_CASE(0, 2, 3) UPD_CASE(1, 1, 3) UPD_CASE(1, 2, 3)
#undef UPD_CASE
  return cudaErrorInvalidValue;
}
