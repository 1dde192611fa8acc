// Fused persistent reverse-diffusion sampler, warp-specialised (UPD_IMPL_TCGEN05_WS; the product path for F <= 2).
//
// Same arithmetic, operand encodings and weight image as sampler_tc.cu (three chained tcgen05 GEMMs per reverse step, A
// from TMEM, B from shared memory, fp16/tf32 hi-lo split; base-2 softplus epilogues in packed fp32x2 math; head sums
// inside the layer-3 epilogue), organised differently.  What the two-tile kernel's ncu capture showed after the
// epilogues went to packed math (profiles/r02_*): inside a MUFU turn only the active tile's two warps per SMSP issue, and
// they spend 1.5x as many cycles in fixed-latency dependency stalls as issuing -- two warps cannot feed the MUFU pipe
// (74 % busy) although neither the issue port (52 %) nor any pipe is full.  Here every epilogue warp always has work:
//
//   * 16 EPILOGUE warps = 4 TMEM lane quadrants x 4 column quarters.  A warp walks the phases (layer 1..3) x (tile A, B)
//     of a step in order, serving its 32 columns x 32 rows of whichever tile's accumulator is ready next; four warps per
//     SMSP interleave their dependency chains, and while tile A's next GEMM runs the same warps are busy with tile B.
//   * 4 ROW warps (one per lane quadrant) own the trajectory state of both tiles: they draw the Philox noise, build the
//     layer-1 operand, run the posterior algebra on the head sums, and one of their threads issues every tcgen05.mma.
//     None of this sits on an epilogue warp's instruction stream any more.
//   * hand-offs never stop an epilogue warp: epilogue warps -> issuer ("operand quarter written") and epilogue warps ->
//     row warps ("head sums stored") are named barriers on which the epilogue side only ARRIVES; tcgen05.commit ->
//     epilogue warps ("accumulator ready") is an mbarrier that is normally complete when a warp asks for it.  Per-row
//     partial sums (sum of squares for F.normalize, head sums) cross the four column quarters through shared memory and
//     are added in a fixed order (bit-reproducible).
//
// TMEM: each tile ping-pongs between two private 128-column buffers (A1 and A3 in buffer 0, A2 in buffer 1).
// Limits: F <= 2 (the four-way head-sum exchange of larger F does not fit next to the weight image in shared memory;
// the launcher returns cudaErrorInvalidValue and the caller falls back to sampler_tc.cu).
#include "sampler_epi.cuh"
#include "sampler_math.cuh"
#include "sampler_params.cuh"
#include "tc_helpers.cuh"
#include "upd_common.cuh"

namespace {
using namespace epi;

constexpr int EPI_WARPS = 16, ROW_WARPS = 4;
constexpr int WS_THREADS = (EPI_WARPS + ROW_WARPS) * 32;
// Named barriers.  The row warps WAIT on hardware barriers (bar.sync: a blocked warp issues nothing), never by polling
// an mbarrier: four polling warps have the highest warp ids of the CTA and take issue slots from the epilogue warps of
// their SMSP every time a try_wait returns (measured: the layer-3 phases, the most issue-hungry ones, ran 20 % slower).
constexpr int ROW_BAR = 1;            // the 128 row-warp threads (layer-1 operand complete)
constexpr int EPI_BAR0 = 2;           // + 2*tile + half: 16 epilogue warps arrive, the issuer warp syncs (the first / second 16 of
                                      //   every warp's 32 operand columns are written: even / odd K-slices of the next GEMM)
constexpr int HEADS_BAR0 = 6;         // + tile: 16 epilogue warps arrive, the 4 row warps sync (head sums stored)

// Pairs of a 16-column group that take the one-MUFU softplus (lg2(1 + u) as a packed-FMA polynomial, sampler_math.cuh):
// bit i = pair i; layers 1-2 / layer 3, per model kind.  With the two-pass contractions the MUFU pipe is what the
// epilogue warps wait for (ncu: XU 84 %, issue 48 %, FMA 20 %), so moving part of the lg2 work to the FMA pipe pays:
// measured on B200 (profiles/r02_pmask_sweep.txt) NsDiff F = 1 4.61 -> 4.82, F = 2 3.84 -> 3.94, TMDM 5.23 -> 5.73 G
// row-steps/s.  NsDiff's layer 3 keeps the MUFU form (its FMA pipe carries the head sums: any mask there loses);
// more than half of the pairs in layers 1-2 loses again (the FMA pipe becomes the longer one).
// -DUPD_WS_PMASK12=m / -DUPD_WS_PMASK3=m override both kinds (0 = the all-MUFU build).
#ifdef UPD_WS_PMASK12
#define UPD_WS_PMASK12_NS UPD_WS_PMASK12
#define UPD_WS_PMASK12_TM UPD_WS_PMASK12
#endif
#ifdef UPD_WS_PMASK3
#define UPD_WS_PMASK3_NS UPD_WS_PMASK3
#define UPD_WS_PMASK3_TM UPD_WS_PMASK3
#endif
#ifndef UPD_WS_PMASK12_NS
#define UPD_WS_PMASK12_NS 0x33
#endif
#ifndef UPD_WS_PMASK3_NS
#define UPD_WS_PMASK3_NS 0x00
#endif
#ifndef UPD_WS_PMASK12_TM
#define UPD_WS_PMASK12_TM 0x55
#endif
#ifndef UPD_WS_PMASK3_TM
#define UPD_WS_PMASK3_TM 0xff
#endif
#ifndef UPD_WS_PMASK1_NS               // layer 1 alone (its MUFU form always carries the overflow guard)
#define UPD_WS_PMASK1_NS UPD_WS_PMASK12_NS
#endif
#ifndef UPD_WS_PMASK1_TM
#ifdef UPD_WS_PMASK12
#define UPD_WS_PMASK1_TM UPD_WS_PMASK12
#else
#define UPD_WS_PMASK1_TM 0x77            // TMDM layer 1: 6 of 8 pairs (5.74 -> 5.86 G row-steps/s); NsDiff: no gain from any layer-1 mask
#endif
#endif
// Passes of the layer-2 / layer-3 contractions.  The activation operand is ONE fp16 word per element (round to nearest:
// an unbiased 2^-12 relative perturbation, independent per element and step), the weights stay hi + lo: hi*hi + hi*lo.
// Measured against the oracle on whole windows (tests/test_gpu_parity_full.py, the same bounds as before): NsDiff configs
// 1 and 2 max |d|/rms 3.0-5.5e-6, MPV 1e-7..4e-7 -- indistinguishable from the three-pass form (2.6-6.3e-6; both sit on
// the fp32 reordering floor) -- and TMDM 1.2e-5 / 4.5e-7 (three passes: 6e-6), 10x inside the per-value bound; the third
// pass cost 8 % (F = 1), 15 % (TMDM), 3.5 % (F = 2) of the kernel.  -DUPD_WS_A_LO=1 builds the three-pass form.
#ifndef UPD_WS_A_LO
#define UPD_WS_A_LO 0
#endif
#ifndef UPD_WS_B_LO
#define UPD_WS_B_LO 1
#endif
constexpr bool WS_A_LO = UPD_WS_A_LO != 0, WS_B_LO = UPD_WS_B_LO != 0;

struct __align__(8) WsSync {
  unsigned long long wbar;
  unsigned long long mma_bar[2];     // accumulator of the tile's current layer is complete (tcgen05.commit)
  uint32_t tmem_base;
  uint32_t pad;
};

#ifndef UPD_WS_NO_SETMAXNREG
template <int F> struct WS_SETMAXNREG { static constexpr bool value = (F == 1); };
#else
template <int F> struct WS_SETMAXNREG { static constexpr bool value = false; };
#endif

template <int KIND, int F>
struct WsShape {
  static constexpr bool NS = (KIND == 0);
  static constexpr int HEAD_ROWS = NS ? 1 + 5 * F : F;                       // ss, F x (pe, pb, m1, m2, m3) | F x pe
  static constexpr uint32_t SSX_FLOATS = 2 * 4 * 128;                        // [layer 1,2][quarter][row]
  static constexpr uint32_t XCH_TILE_FLOATS = SSX_FLOATS + 4 * HEAD_ROWS * 128;
  static constexpr uint32_t STEP_BYTES = NS ? sizeof(UpdNsStep) : sizeof(UpdTmStep);
};

// This warp's 32 columns of a hidden layer, in place: K-slice j of the next A operand = hi words in columns
// [16j,16j+8), lo words in [16j+8,16j+16).  Returns the partial sum of squares of the 32 activations.
// half_bar: named barrier on which the warp arrives once its first 16 columns (an even K-slice of the next GEMM) are in
// TMEM; the caller arrives on half_bar + 1 after the second 16.
template <bool FIRST, bool GUARD, bool SUMSQ>
__device__ __forceinline__ float epilogue_quarter(uint32_t acc, const float* __restrict__ e, const float* __restrict__ b,
                                                  float inv, int half_bar) {
  constexpr int PMASK = FIRST ? (SUMSQ ? UPD_WS_PMASK1_NS : UPD_WS_PMASK1_TM)     // SUMSQ <=> NsDiff (L2-normalised layers)
                              : (SUMSQ ? UPD_WS_PMASK12_NS : UPD_WS_PMASK12_TM);
  float2 ss2 = make_float2(0.f, 0.f);
  const float2 inv2 = sm::splat(inv);
  uint32_t r[32], o[16];
  tc::tmem_ld32(acc, r);
  tc::wait_ld();
  epilogue_group<FIRST, GUARD, SUMSQ, PMASK, WS_A_LO>(r, o, e, b, inv2, ss2);
  if (WS_A_LO) tc::tmem_st16(acc, o);
  else tc::tmem_st8(acc, *reinterpret_cast<uint32_t (*)[8]>(&o[0]));
  tc::wait_st();
  tc::fence_before_sync();
  tc::named_bar_arrive(half_bar, EPI_WARPS * 32 + 32);
  epilogue_group<FIRST, GUARD, SUMSQ, PMASK, WS_A_LO>(r + 16, o, e + 16, b + 16, inv2, ss2);
  if (WS_A_LO) tc::tmem_st16(acc + 16u, o);
  else tc::tmem_st8(acc + 16u, *reinterpret_cast<uint32_t (*)[8]>(&o[0]));
  return ss2.x + ss2.y;
}

// Layer 3: activations feed the head sums, nothing is written back.
template <bool NS, int F, bool GUARD>
__device__ __forceinline__ void heads_quarter(uint32_t acc, const float* __restrict__ e, const float* __restrict__ b,
                                              const float* __restrict__ w4, const float* __restrict__ ws, float inv,
                                              HeadSums<NS, F>& H) {
  const float2 inv2 = sm::splat(inv);
  uint32_t r[32];
  tc::tmem_ld32(acc, r);
  tc::wait_ld();
  heads_group<NS, F, GUARD, (NS ? UPD_WS_PMASK3_NS : UPD_WS_PMASK3_TM)>(r, e, b, w4, ws, inv2, H);
  heads_group<NS, F, GUARD, (NS ? UPD_WS_PMASK3_NS : UPD_WS_PMASK3_TM)>(r + 16, e + 16, b + 16, w4 + 16, ws + 16, inv2, H);
}

#ifdef UPD_TRACE
#define UPD_STAMP(k) do { if (tracing && lane == 0) p.trace[(warp * p.T + (p.T - 1 - t)) * 16 + (k)] = clock64(); } while (0)
#else
#define UPD_STAMP(k) do { } while (0)
#endif

template <int KIND, int F>
__global__ void __launch_bounds__(WS_THREADS, 1)
sampler_ws_kernel(const UpdSamplerParams p) {
  using Shape = WsShape<KIND, F>;
  constexpr bool NS = Shape::NS;
  constexpr int HR = Shape::HEAD_ROWS;
  constexpr int IN = NS ? 3 * F : 2 * F;
  constexpr int K1 = ((IN + 1 + 7) / 8) * 8;
  static_assert(K1 == 8, "the warp-specialised sampler is built for F <= 2 (one 8-wide tf32 K-slice in layer 1)");
  const UpdPackLayout L = upd_make_layout(KIND, F, p.T);
  extern __shared__ __align__(128) unsigned char smem[];
  auto sf = [&](uint32_t off) { return reinterpret_cast<float*>(smem + off); };
  const uint32_t steps_off = upd_align128(L.tc_image_bytes);
  const uint32_t xch_off = upd_align128(steps_off + Shape::STEP_BYTES * p.T);
  const uint32_t sync_off = upd_align128(xch_off + 2 * Shape::XCH_TILE_FLOATS * 4);
  WsSync* sync = reinterpret_cast<WsSync*>(smem + sync_off);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    tc::mbar_init(tc::smem_u32(&sync->wbar), 1);
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(tc::smem_u32(&sync->mma_bar[i]), 1);
    }
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
  for (int i = tid; i < L.TE * 128; i += WS_THREADS) {
    sf(L.e1)[i] *= LOG2E; sf(L.e2)[i] *= LOG2E; sf(L.e3)[i] *= LOG2E;
  }
  if (tid < p.T) {
    if (NS) reinterpret_cast<UpdNsStep*>(smem + steps_off)[tid] = upd_ns_step(sf(L.sched), p.T, tid);
    else reinterpret_cast<UpdTmStep*>(smem + steps_off)[tid] = upd_tm_step(sf(L.sched), p.T, tid);
  }
  __syncthreads();

  const uint32_t img = tc::smem_u32(smem);
  const long long n_tiles = (p.n_rows + 127) / 128;
  // every CTA runs the same number of tile pairs (slots past the end compute on a clamped row and store nothing)
  const long long n_iters = (n_tiles + 2LL * gridDim.x - 1) / (2LL * gridDim.x);
  const float inv_ws2 = sf(L.scales)[0] * (NS ? 1.0f : LN2), inv_ws3 = sf(L.scales)[1] * (NS ? 1.0f : LN2);
  // NsDiff: |z'| of layers 2 and 3 is bounded by (|W_row| + |b|) |e| log2e because their input is L2-normalised; the
  // packer stores that bound (scales[2]); below 120 the ex2 overflow guard is compiled out of those epilogues.
  const bool guard23 = !NS || !(sf(L.scales)[2] > 0.f && sf(L.scales)[2] < 120.f);
  auto xch = [&](int tile) { return sf(xch_off) + tile * Shape::XCH_TILE_FLOATS; };

  if (warp < EPI_WARPS) {
    // =========================================== epilogue warps ===========================================
    // Register re-allocation between the roles (F = 1): the kernel is compiled for 96 registers per thread (640 threads);
    // the four epilogue warpgroups take 104 and the row warpgroup gives back down to 64 (512 x 104 + 128 x 64 = 640 x 96).
    // Measured: 4.03 -> 4.21 G row-steps/s at F = 1 (the epilogue's spills go away, the posterior chain fits in 64);
    // at F = 2 the row warps' per-feature state spills at 64 registers and the kernel loses 8 % (2 % without their noise
    // cache), so it stays off there.
    if constexpr (WS_SETMAXNREG<F>::value) asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int quad = warp & 3, cq = warp >> 2;
    const int trow = quad * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(quad * 32) << 16;
    const int c0 = 32 * cq;                                          // first hidden column of this warp's quarter
    uint32_t mma_phase[2] = {0u, 0u};
    auto wait_acc = [&](int tile) {
      tc::mbar_wait(tc::smem_u32(&sync->mma_bar[tile]), mma_phase[tile]);
      mma_phase[tile] ^= 1u;
      tc::fence_after_sync();
    };

    for (long long it = 0; it < n_iters; ++it) {
#ifdef UPD_TRACE
      const bool tracing = (p.trace != nullptr) && blockIdx.x == 0 && it == 1;
#endif
      for (int t = p.T - 1; t >= 0; --t) {
        UPD_STAMP(12);
        // ---- layer 1 epilogue: accumulator in buffer 1 -> A2 in place ----
#pragma unroll
        for (int tile = 0; tile < 2; ++tile) {
          const uint32_t acc = tmem_base + lane_sel + (uint32_t)tile * 256u + 128u + (uint32_t)c0;
          wait_acc(tile);
          UPD_STAMP(0 + 2 * tile);
          const float ss = epilogue_quarter<true, true, NS>(acc, sf(L.e1) + t * 128 + c0, nullptr, 1.f, EPI_BAR0 + 2 * tile);
          if (NS) xch(tile)[(0 * 4 + cq) * 128 + trow] = ss;
          tc::wait_st();
          tc::fence_before_sync();
          tc::named_bar_arrive(EPI_BAR0 + 2 * tile + 1, EPI_WARPS * 32 + 32);
          UPD_STAMP(1 + 2 * tile);
        }
        // ---- layer 2 epilogue: accumulator in buffer 0 -> A3 in place ----
#pragma unroll
        for (int tile = 0; tile < 2; ++tile) {
          const uint32_t acc = tmem_base + lane_sel + (uint32_t)tile * 256u + (uint32_t)c0;
          wait_acc(tile);
          UPD_STAMP(4 + 2 * tile);
          float inv = inv_ws2;
          if (NS) {
            const float* s = xch(tile) + trow;
            inv = inv_ws2 / fmaxf(sqrtf(((s[0] + s[128]) + s[256]) + s[384]), 1e-12f);     // F.normalize, folded past the GEMM
          }
          float ss;
          if (guard23) ss = epilogue_quarter<false, true, NS>(acc, sf(L.e2) + t * 128 + c0, sf(L.b2) + c0, inv, EPI_BAR0 + 2 * tile);
          else ss = epilogue_quarter<false, false, NS>(acc, sf(L.e2) + t * 128 + c0, sf(L.b2) + c0, inv, EPI_BAR0 + 2 * tile);
          if (NS) xch(tile)[(1 * 4 + cq) * 128 + trow] = ss;
          tc::wait_st();
          tc::fence_before_sync();
          tc::named_bar_arrive(EPI_BAR0 + 2 * tile + 1, EPI_WARPS * 32 + 32);
          UPD_STAMP(5 + 2 * tile);
        }
        // ---- layer 3 epilogue + head sums: accumulator in buffer 1 ----
#pragma unroll
        for (int tile = 0; tile < 2; ++tile) {
          const uint32_t acc = tmem_base + lane_sel + (uint32_t)tile * 256u + 128u + (uint32_t)c0;
          wait_acc(tile);
          UPD_STAMP(8 + 2 * tile);
          float inv = inv_ws3;
          if (NS) {
            const float* s = xch(tile) + 4 * 128 + trow;
            inv = inv_ws3 / fmaxf(sqrtf(((s[0] + s[128]) + s[256]) + s[384]), 1e-12f);
          }
          HeadSums<NS, F> hs;
          hs.clear();
          if (guard23) heads_quarter<NS, F, true>(acc, sf(L.e3) + t * 128 + c0, sf(L.b3) + c0, sf(L.w4) + c0, sf(L.ws) + c0, inv, hs);
          else heads_quarter<NS, F, false>(acc, sf(L.e3) + t * 128 + c0, sf(L.b3) + c0, sf(L.w4) + c0, sf(L.ws) + c0, inv, hs);
          float* hx = xch(tile) + Shape::SSX_FLOATS + cq * HR * 128 + trow;
          if (NS) {
            hx[0] = hs.ss.x + hs.ss.y;
#pragma unroll
            for (int f = 0; f < F; ++f) {
              hx[(1 + 5 * f + 0) * 128] = hs.pe[f].x + hs.pe[f].y;
              hx[(1 + 5 * f + 1) * 128] = hs.pb[f].x + hs.pb[f].y;
              hx[(1 + 5 * f + 2) * 128] = hs.m1[f].x + hs.m1[f].y;
              hx[(1 + 5 * f + 3) * 128] = hs.m2[f].x + hs.m2[f].y;
              hx[(1 + 5 * f + 4) * 128] = hs.m3[f].x + hs.m3[f].y;
            }
          } else {
#pragma unroll
            for (int f = 0; f < F; ++f) hx[f * 128] = hs.pe[f].x + hs.pe[f].y;
          }
          tc::fence_before_sync();            // this warp's TMEM reads of the accumulator are complete (wait::ld above)
          __threadfence_block();
          tc::named_bar_arrive(HEADS_BAR0 + tile, EPI_WARPS * 32 + ROW_WARPS * 32);
          UPD_STAMP(9 + 2 * tile);
        }
      }
    }
  } else {
    // ============================================== row warps ==============================================
    if constexpr (WS_SETMAXNREG<F>::value) asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    const int quad = warp - EPI_WARPS;                               // = warp % 4: this warp's TMEM lane quadrant
    const int trow = quad * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(quad * 32) << 16;
    const bool issuer_warp = (quad == 0);
    const bool issuer = issuer_warp && lane == 0;
    // ln2 * sum_j ws[f][j]: the constant term of the sigma-head polynomial
    float ws_sum[F];
#pragma unroll
    for (int f = 0; f < F; ++f) {
      float a = 0.f;
      if (NS) for (int j = 0; j < 128; ++j) a += sf(L.ws)[f * 128 + j];
      ws_sum[f] = sm::SPH_LN2 * a;
    }
    float b4v[F], bsv[F];
#pragma unroll
    for (int f = 0; f < F; ++f) { b4v[f] = sf(L.b4)[f]; bsv[f] = NS ? sf(L.bs)[f] : 0.f; }
    // A1 = [y | y0_hat | gx | 1 | 0] as tf32 hi/lo into buffer 0 of the tile; then every row warp meets, and one thread
    // issues the layer-1 GEMM (accumulator in buffer 1)
    auto build_a1 = [&](int tile, const float (&y)[F], const float (&y0h)[F], const float (&gxv)[F]) {
      float in[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) in[i] = 0.f;
#pragma unroll
      for (int f = 0; f < F; ++f) {
        in[f] = y[f];
        in[F + f] = y0h[f];
        if (NS) in[2 * F + f] = gxv[f];
      }
      in[IN] = 1.0f;
      uint32_t a[16];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float hi = tc::to_tf32(in[i]);
        a[i] = __float_as_uint(hi);
        a[8 + i] = __float_as_uint(tc::to_tf32(in[i] - hi));
      }
      tc::tmem_st16(tmem_base + lane_sel + (uint32_t)tile * 256u, a);
      tc::wait_st();
      tc::fence_before_sync();
      tc::named_bar_sync(ROW_BAR, ROW_WARPS * 32);
      if (issuer) {
        tc::fence_after_sync();
        const uint32_t b0 = tmem_base + (uint32_t)tile * 256u;
        tc::issue_layer_tf32x3(b0 + 128u, b0, K1, img + L.u1hi, img + L.u1lo, UMMA_LBO, UMMA_SBO);
        tc::mma_commit(tc::smem_u32(&sync->mma_bar[tile]));
      }
    };
    for (long long it = 0; it < n_iters; ++it) {
      long long row[2];
      bool live[2];
      UpdRowIndex ix[2];
      float y[2][F], y0h[2][F], gxv[2][F], zn[2][F];
      float4 zc[2][F];                                   // the current group of four draws of every element
#pragma unroll
      for (int tile = 0; tile < 2; ++tile) {
        row[tile] = ((it * gridDim.x + blockIdx.x) * 2 + tile) * 128 + trow;
        live[tile] = row[tile] < p.n_rows;
        ix[tile] = upd_row_index(p, live[tile] ? row[tile] : p.n_rows - 1);
        const long long cidx = (ix[tile].r0 * p.O + ix[tile].o) * F;
#pragma unroll
        for (int f = 0; f < F; ++f) {
          y0h[tile][f] = p.y0_hat ? p.y0_hat[cidx + f] : 0.f;
          gxv[tile][f] = NS ? p.gx[cidx + f] : 1.f;
          const float z = upd_draw_cached(p, ix[tile], f, F, 0, zc[tile][f]);
          y[tile][f] = NS ? sqrtf(gxv[tile][f]) * z + y0h[tile][f] : z + y0h[tile][f];   // nsdiff_utils.py:274 / tmdm_diffusion_utils.py:110
        }
        build_a1(tile, y[tile], y0h[tile], gxv[tile]);
      }
#ifdef UPD_TRACE
      const bool tracing = (p.trace != nullptr) && blockIdx.x == 0 && it == 1;
#endif
      for (int t = p.T - 1; t >= 0; --t) {
        const bool last = (t == 0);
        UPD_STAMP(0);
        // layers 2 and 3: issue each tile's GEMM as soon as all 16 epilogue warps have written its operand
#pragma unroll
        for (int layer = 2; layer <= 3; ++layer) {
#pragma unroll
          for (int tile = 0; tile < 2; ++tile) {
            if (issuer_warp) {
#pragma unroll
             for (int half = 0; half < 2; ++half) {
              tc::named_bar_sync(EPI_BAR0 + 2 * tile + half, EPI_WARPS * 32 + 32);
              if (lane == 0) {
                tc::fence_after_sync();
                const uint32_t b0 = tmem_base + (uint32_t)tile * 256u;
                const uint32_t d = (layer == 2) ? b0 : b0 + 128u, a = (layer == 2) ? b0 + 128u : b0;
                const uint32_t whi = img + (layer == 2 ? L.u2hi : L.u3hi), wlo = img + (layer == 2 ? L.u2lo : L.u3lo);
                tc::issue_kparity_f16x3_g16<WS_A_LO, WS_B_LO>(d, a, whi, wlo, UMMA_LBO, UMMA_SBO, half);
                if (half == 1) tc::mma_commit(tc::smem_u32(&sync->mma_bar[tile]));
              }
              __syncwarp();
             }
            }
            UPD_STAMP(1 + 2 * (layer - 2) + tile);
          }
          if (layer == 2) {
            // this step's noise, drawn while the epilogue warps are busy
#pragma unroll
            for (int tile = 0; tile < 2; ++tile)
#pragma unroll
              for (int f = 0; f < F; ++f) zn[tile][f] = last ? 0.f : upd_draw_cached(p, ix[tile], f, F, p.T - t, zc[tile][f]);
            UPD_STAMP(5);
          }
        }
        // heads -> posterior update -> next step's layer-1 operand.  Everything of the NsDiff posterior that does not
        // depend on the heads is formed before the wait.
        UpdNsStep st;
        NsRowPre pre[2][F];
        float inv_two_lam0 = 0.f;
        if (NS) {
          st = reinterpret_cast<const UpdNsStep*>(smem + steps_off)[t];
          inv_two_lam0 = 1.0f / st.two_lam0;
#pragma unroll
          for (int tile = 0; tile < 2; ++tile)
#pragma unroll
            for (int f = 0; f < F; ++f) pre[tile][f] = ns_row_pre(st, y[tile][f], y0h[tile][f], gxv[tile][f]);
        }
#pragma unroll
        for (int tile = 0; tile < 2; ++tile) {
          tc::named_bar_sync(HEADS_BAR0 + tile, EPI_WARPS * 32 + ROW_WARPS * 32);
          UPD_STAMP(6 + 3 * tile);
          const float* hx = xch(tile) + Shape::SSX_FLOATS + trow;
          auto tot = [&](int r) {      // fixed-order sum over the four column quarters
            return ((hx[(0 * HR + r) * 128] + hx[(1 * HR + r) * 128]) + hx[(2 * HR + r) * 128]) + hx[(3 * HR + r) * 128];
          };
          if (NS) {
            const float ss3 = tot(0);
#pragma unroll
            for (int f = 0; f < F; ++f) {
              float eps, sig;
              ns_heads_fast(ss3, tot(1 + 5 * f), tot(2 + 5 * f), tot(3 + 5 * f), tot(4 + 5 * f), tot(5 + 5 * f), ws_sum[f],
                            b4v[f], bsv[f], eps, sig);
              y[tile][f] = ns_update_fast(st, pre[tile][f], inv_two_lam0, eps, sig, zn[tile][f], last);
            }
          } else {
            const UpdTmStep st = reinterpret_cast<const UpdTmStep*>(smem + steps_off)[t];
#pragma unroll
            for (int f = 0; f < F; ++f) {
              const float eps = tot(f) * LN2 + sf(L.b4)[f];
              y[tile][f] = upd_tm_update(st, y[tile][f], y0h[tile][f], eps, zn[tile][f], last);
            }
          }
          UPD_STAMP(7 + 3 * tile);
          if (!last) build_a1(tile, y[tile], y0h[tile], gxv[tile]);
          UPD_STAMP(8 + 3 * tile);
        }
      }
#pragma unroll
      for (int tile = 0; tile < 2; ++tile)
        if (live[tile]) {
#pragma unroll
          for (int f = 0; f < F; ++f) p.out[row[tile] * F + f] = y[tile][f];
        }
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem_base);
}

template <int KIND, int F>
cudaError_t launch(const UpdSamplerParams& p, int sms, cudaStream_t stream) {
  using Shape = WsShape<KIND, F>;
  const UpdPackLayout L = upd_make_layout(KIND, F, p.T);
  size_t smem = upd_align128(upd_align128(upd_align128(L.tc_image_bytes) + Shape::STEP_BYTES * p.T) +
                             2 * Shape::XCH_TILE_FLOATS * 4) + sizeof(WsSync) + 128;
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  auto kern = sampler_ws_kernel<KIND, F>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  long long n_tiles = (p.n_rows + 127) / 128;
  long long ctas = (n_tiles + 1) / 2;
  int grid = (int)(ctas < sms ? ctas : sms);
  if (grid < 1) grid = 1;
  kern<<<grid, WS_THREADS, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace

cudaError_t upd_launch_sampler_ws(const UpdSamplerParams& p, int kind, int F, int sms, cudaStream_t stream) {
#define UPD_CASE(KK, FF) if (kind == KK && F == FF) return launch<KK, FF>(p, sms, stream);
  UPD_CASE(0, 1) UPD_CASE(0, 2) UPD_CASE(1, 1) UPD_CASE(1, 2)
#undef UPD_CASE
  return cudaErrorInvalidValue;
}
