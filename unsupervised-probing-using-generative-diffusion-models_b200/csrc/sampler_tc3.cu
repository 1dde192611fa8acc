// Fused persistent reverse-diffusion sampler, three tiles per SM (UPD_IMPL_TCGEN05_X3).
//
// Same arithmetic, operand encodings and weight image as sampler_tc.cu; what changes is the orchestration.
// TMEM has no room for three ping-pong pairs (3 x 256 > 512 columns), but a tile needs both of its buffers only while
// its MMAs are in flight: during its epilogue and while it waits, one 128-column buffer is live (the accumulator,
// re-encoded in place as the next A operand).  So four 128-column buffers rotate among three tiles: MMA number m of the
// CTA-wide sequence (issued by tile m % 3) reads A from buffer (m+2) % 4 and accumulates into buffer (m+1) % 4, which
// is the A buffer of MMA m-1 -- free as soon as that MMA has completed, which the issuing thread checks on the other
// tile's mbarrier before it issues.  Four buffers = the whole TMEM.
//
// One warp per TMEM lane quadrant (3 tiles x 4 warps = 384 threads, up to 168 registers): a thread owns a full row
// -- all 128 hidden columns, the row state, the posterior, its Philox draws -- so the two-tile kernel's half<->half
// exchanges (sum of squares, head partial sums, noise) and their barriers disappear.
//
// MEASURED on B200 (G row-steps/s; two-tile kernel: 3.69 at the bench shape F=1, 3.21 at config 1's F=2):
//   strict MUFU turns (one tile in its softplus at a time)   2.42 / 2.82   one warp per scheduler cannot feed the MUFU pipe
//   hand-over after 6 / 4 / 2 of the 8 column groups         2.96 / 3.22 / 3.55   (F=1);  3.11 / 3.37 / 3.37 (F=2)
//   no turn-taking, three tiles free-running (the default)    3.58 / 3.38
// i.e. the more warps share the MUFU pipe the better; with 12 warps it matches the 16-warp kernel at F=1 and beats it
// by 5 % at F >= 2 (no second-pass exchange), which is where the host layer selects it.
#include "sampler_params.cuh"
#include "tc_helpers.cuh"
#include "upd_common.cuh"

namespace {

constexpr int X3_THREADS = 384;
constexpr uint32_t UMMA_LBO = 2048;
constexpr uint32_t UMMA_SBO = 128;
constexpr int TURN_BAR0 = 8;          // named barriers 8,9,10: MUFU turn of tile 0,1,2 (256 = 128 waiting + 128 arriving)
constexpr int TILE_BAR0 = 1;          // named barriers 1,2,3: the 128 threads of a tile
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
// Turn policy.  UPD_X3_TURNS = 0: no MUFU turn-taking at all (three tiles free-running).  Otherwise a tile hands the
// turn to the next one after UPD_X3_HANDOFF of its 8 column groups: 8 = strictly one tile in its softplus at a time,
// 4 = two tiles overlap by half an epilogue (two warps per scheduler on the MUFU pipe at any time).
#ifndef UPD_X3_TURNS
#define UPD_X3_TURNS 0
#endif
#ifndef UPD_X3_HANDOFF
#define UPD_X3_HANDOFF 8
#endif
__device__ __forceinline__ void turn_begin(int bar) { if (UPD_X3_TURNS) tc::named_bar_sync(bar, 256); }
__device__ __forceinline__ void turn_end(int bar) { if (UPD_X3_TURNS) tc::named_bar_arrive(bar, 256); }

struct __align__(8) X3Sync {
  unsigned long long wbar;
  unsigned long long mma_bar[3];
  uint32_t tmem_base;
  uint32_t pad;
};

__device__ __forceinline__ float lg2_1p_ex2(float z) {
  float u, l;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(u) : "f"(z));
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(1.0f + u));
  return l;
}

// softplus(x) for x in [0,1] on the FMA pipe (see sampler_tc.cu)
constexpr float SPU_C0 = 0.6931471824645996f, SPU_C1 = 0.12499982863664627f, SPU_C2 = -0.005206969100981951f,
                SPU_C3 = 0.0003433137317188084f, SPU_C4 = -2.16761418414535e-05f;

template <bool FIRST, bool CLAMP>
__device__ __forceinline__ float epilogue_group(const uint32_t (&r)[16], uint32_t (&o)[16], const float* __restrict__ e,
                                                const float* __restrict__ b, float inv) {
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 16; j += 4) {
    float4 e4 = *reinterpret_cast<const float4*>(e + j);
    float4 b4 = FIRST ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(b + j);
    float z0 = FIRST ? __uint_as_float(r[j]) * e4.x : fmaf(__uint_as_float(r[j]), inv, b4.x) * e4.x;
    float z1 = FIRST ? __uint_as_float(r[j + 1]) * e4.y : fmaf(__uint_as_float(r[j + 1]), inv, b4.y) * e4.y;
    float z2 = FIRST ? __uint_as_float(r[j + 2]) * e4.z : fmaf(__uint_as_float(r[j + 2]), inv, b4.z) * e4.z;
    float z3 = FIRST ? __uint_as_float(r[j + 3]) * e4.w : fmaf(__uint_as_float(r[j + 3]), inv, b4.w) * e4.w;
    if (CLAMP) { z0 = fminf(z0, 126.f); z1 = fminf(z1, 126.f); z2 = fminf(z2, 126.f); z3 = fminf(z3, 126.f); }
    float h0 = lg2_1p_ex2(z0), h1 = lg2_1p_ex2(z1), h2 = lg2_1p_ex2(z2), h3 = lg2_1p_ex2(z3);
    ss = fmaf(h0, h0, ss); ss = fmaf(h1, h1, ss); ss = fmaf(h2, h2, ss); ss = fmaf(h3, h3, ss);
    tc::split_f16x2(h0, h1, o[j / 2], o[8 + j / 2]);
    tc::split_f16x2(h2, h3, o[j / 2 + 1], o[8 + j / 2 + 1]);
  }
  return ss;
}

// The whole row of one hidden layer (8 groups of 16 columns), TMEM loads one group ahead; in place.
template <bool FIRST, bool CLAMP>
__device__ __forceinline__ float epilogue_row(uint32_t buf, const float* __restrict__ e, const float* __restrict__ b,
                                              float inv, int next_turn) {
  float ss = 0.f;
  uint32_t r[16], rn[16], o[16];
  tc::tmem_ld16(buf, r);
  tc::wait_ld();
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    if (q < 7) tc::tmem_ld16(buf + 16u * (q + 1), rn);
    ss += epilogue_group<FIRST, CLAMP>(r, o, e + 16 * q, b + 16 * q, inv);
    tc::tmem_st16(buf + 16u * q, o);
    if (q == UPD_X3_HANDOFF - 1) turn_end(next_turn);
    if (q < 7) {
      tc::wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = rn[i];
    }
  }
  return ss;
}

template <int KIND, int F>
__global__ void __launch_bounds__(X3_THREADS, 1)
sampler_tc3_kernel(const UpdSamplerParams p) {
  constexpr bool NS = (KIND == 0);
  constexpr int IN = NS ? 3 * F : 2 * F;
  constexpr int K1 = ((IN + 1 + 7) / 8) * 8;
  const UpdPackLayout L = upd_make_layout(KIND, F, p.T);
  extern __shared__ __align__(128) unsigned char smem[];
  auto sf = [&](uint32_t off) { return reinterpret_cast<float*>(smem + off); };
  constexpr uint32_t STEP_BYTES = NS ? sizeof(UpdNsStep) : sizeof(UpdTmStep);
  const uint32_t steps_off = upd_align128(L.tc_image_bytes);
  const uint32_t sync_off = upd_align128(steps_off + STEP_BYTES * p.T);
  X3Sync* sync = reinterpret_cast<X3Sync*>(smem + sync_off);

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
  for (int i = tid; i < L.TE * 128; i += X3_THREADS) {
    sf(L.e1)[i] *= LOG2E; sf(L.e2)[i] *= LOG2E; sf(L.e3)[i] *= LOG2E;
  }
  if (tid < p.T) {
    if (NS) reinterpret_cast<UpdNsStep*>(smem + steps_off)[tid] = upd_ns_step(sf(L.sched), p.T, tid);
    else reinterpret_cast<UpdTmStep*>(smem + steps_off)[tid] = upd_tm_step(sf(L.sched), p.T, tid);
  }
  __syncthreads();

  const int tile_id = warp >> 2, quad = warp & 3;
  const int trow = quad * 32 + lane;                        // row within the tile = TMEM lane
  const bool issuer = quad == 0 && lane == 0;
  const uint32_t lane_sel = (uint32_t)(quad * 32) << 16;
  const uint32_t own_bar = tc::smem_u32(&sync->mma_bar[tile_id]);
  const uint32_t img = tc::smem_u32(smem);
  const int tile_bar = TILE_BAR0 + tile_id;
  const int my_turn = TURN_BAR0 + tile_id, next_turn = TURN_BAR0 + (tile_id + 1) % 3;
  const float inv_ws2 = sf(L.scales)[0] * (NS ? 1.0f : LN2), inv_ws3 = sf(L.scales)[1] * (NS ? 1.0f : LN2);
  const float* e1 = sf(L.e1);
  const float* e2 = sf(L.e2);
  const float* e3 = sf(L.e3);
  const float* b2 = sf(L.b2);
  const float* b3 = sf(L.b3);
  const float* w4 = sf(L.w4);
  const float* wsg = sf(L.ws);
  float ws_sum[F];
#pragma unroll
  for (int f = 0; f < F; ++f) {
    float a = 0.f;
    if (NS) for (int j = 0; j < 128; ++j) a += wsg[f * 128 + j];
    ws_sum[f] = SPU_C0 * a;
  }

  // MMA bookkeeping.  m = index of this tile's next MMA in the CTA-wide sequence (tile m % 3 issues MMA m):
  //   A operand in buffer (m+2)&3, accumulator in buffer (m+1)&3; before issuing, MMA m-1 must have completed.
  long long m = tile_id;
  auto buf_of = [&](long long idx) { return tmem_base + 128u * (uint32_t)(idx & 3); };
  // Issue MMA m (called by the tile's issuer thread after the tile barrier): LAYER 1 = tf32 K1, 2/3 = fp16 K 128.
  auto issue = [&](int layer) {
    tc::fence_after_sync();
    if (m > 0) {
      const long long pm = m - 1;
      tc::mbar_wait(tc::smem_u32(&sync->mma_bar[pm % 3]), (uint32_t)((pm / 3) & 1));
      tc::fence_after_sync();
    }
    const uint32_t a = buf_of(m + 2), d = buf_of(m + 1);
    if (layer == 1) tc::issue_layer_tf32x3(d, a, K1, img + L.u1hi, img + L.u1lo, UMMA_LBO, UMMA_SBO);
    else if (layer == 2) tc::issue_layer_f16x3_g16(d, a, img + L.u2hi, img + L.u2lo, UMMA_LBO, UMMA_SBO);
    else tc::issue_layer_f16x3_g16(d, a, img + L.u3hi, img + L.u3lo, UMMA_LBO, UMMA_SBO);
    tc::mma_commit(own_bar);
  };
  // Everything a tile does between two of its MMAs: publish the operand, rendezvous, issue, wait for completion.
  // Returns the (lane-selected) address of the accumulator the tile will read next.
  auto run_mma = [&](int layer) -> uint32_t {
    tc::wait_st();
    tc::fence_before_sync();
    tc::named_bar_sync(tile_bar, 128);
    if (issuer) issue(layer);
    const uint32_t d = buf_of(m + 1) + lane_sel;
    tc::mbar_wait(own_bar, (uint32_t)((m / 3) & 1));
    tc::fence_after_sync();
    m += 3;
    return d;
  };

  const long long n_tiles = (p.n_rows + 127) / 128;
  // every tile slot of every CTA runs the same number of iterations: the turn rotation must never wait for a
  // partner that has already left (slots past the end compute on a clamped row and store nothing)
  const long long n_iters = (n_tiles + 3LL * gridDim.x - 1) / (3LL * gridDim.x);
  if (tile_id == 2) turn_end(TURN_BAR0);                      // tile 0 takes the first turn
  for (long long it = 0; it < n_iters; ++it) {
    const long long tile = (it * gridDim.x + blockIdx.x) * 3 + tile_id;
    const long long row = tile * 128 + trow;
    const bool live = row < p.n_rows;
    UpdRowIndex ix = upd_row_index(p, live ? row : p.n_rows - 1);
    float y[F], y0h[F], gxv[F];
    {
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
      // ---------------- layer 1: A1 = [y | y0_hat | gx | 1 | 0] as tf32 hi/lo ----------------
      {
        const uint32_t a1 = buf_of(m + 2) + lane_sel;
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
          tc::tmem_st16(a1, a);
        } else {
          uint32_t a[32];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float v = in[i % K1];
            float hi = tc::to_tf32(v);
            a[i] = __float_as_uint(hi);
            a[16 + i] = __float_as_uint(tc::to_tf32(v - hi));
          }
          tc::tmem_st32(a1, a);
        }
      }
      uint32_t acc = run_mma(1);
      // noise of this step, drawn while the other tiles hold the MUFU turn
      float zn[F];
#pragma unroll
      for (int f = 0; f < F; ++f) zn[f] = (t == 0) ? 0.f : upd_draw(p, ix, f, F, p.T - t);

      // ---------------- layer 1 epilogue -> A2 (in place); layer 2 ----------------
      turn_begin(my_turn);
      float ss = epilogue_row<true, true>(acc, e1 + t * 128, nullptr, 1.f, next_turn);
      acc = run_mma(2);
      float inv = inv_ws2;
      if (NS) inv = inv_ws2 / fmaxf(sqrtf(ss), 1e-12f);        // F.normalize, folded past the GEMM

      // ---------------- layer 2 epilogue -> A3 (in place); layer 3 ----------------
      turn_begin(my_turn);
      ss = epilogue_row<false, !NS>(acc, e2 + t * 128, b2, inv, next_turn);
      acc = run_mma(3);
      inv = inv_ws3;
      if (NS) inv = inv_ws3 / fmaxf(sqrtf(ss), 1e-12f);

      // ---------------- layer 3 epilogue + heads ----------------
      const bool last = (t == 0);
      const float* e3t = e3 + t * 128;
      if constexpr (!(NS && F > 1)) {
        // single pass: heads as power sums of L (see sampler_tc.cu for the algebra)
        float pe[F], pb[F], m1[F], m2[F], m3[F], m4[F];
#pragma unroll
        for (int f = 0; f < F; ++f) { pe[f] = 0.f; pb[f] = 0.f; m1[f] = 0.f; m2[f] = 0.f; m3[f] = 0.f; m4[f] = 0.f; }
        ss = 0.f;
        turn_begin(my_turn);
        {
          uint32_t r[16], rn[16];
          tc::tmem_ld16(acc, r);
          tc::wait_ld();
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (q < 7) tc::tmem_ld16(acc + 16u * (q + 1), rn);
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
            if (q == UPD_X3_HANDOFF - 1) turn_end(next_turn);
            if (q < 7) {
              tc::wait_ld();
#pragma unroll
              for (int i = 0; i < 16; ++i) r[i] = rn[i];
            }
          }
        }
        if (NS) {
          const UpdNsStep st = reinterpret_cast<const UpdNsStep*>(smem + steps_off)[t];
          const float inv3 = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
          const float i2 = inv3 * inv3, i4 = i2 * i2;
#pragma unroll
          for (int f = 0; f < F; ++f) {
            float eps = pe[f] * inv3 + sf(L.b4)[f];
            float lin = 0.5f * inv3 * pb[f];
            float poly = fmaf(i2, SPU_C1 * m1[f], ws_sum[f]);
            poly = fmaf(i4, SPU_C2 * m2[f], poly);
            poly = fmaf(i4 * i2, SPU_C3 * m3[f], poly);
            poly = fmaf(i4 * i4, SPU_C4 * m4[f], poly);
            float sig = upd_softplus_accurate(lin + poly + sf(L.bs)[f]);
            y[f] = upd_ns_update(st, y[f], y0h[f], gxv[f], eps, sig, zn[f], last);
          }
        } else {
          const UpdTmStep st = reinterpret_cast<const UpdTmStep*>(smem + steps_off)[t];
#pragma unroll
          for (int f = 0; f < F; ++f) {
            float eps = pe[f] * LN2 + sf(L.b4)[f];
            y[f] = upd_tm_update(st, y[f], y0h[f], eps, zn[f], last);
          }
        }
      } else {
        // NsDiff with several features: pass 1 (in the turn) leaves L3 in TMEM, pass 2 (outside) evaluates the heads
        float pe[F], ps[F];
#pragma unroll
        for (int f = 0; f < F; ++f) { pe[f] = 0.f; ps[f] = 0.f; }
        ss = 0.f;
        turn_begin(my_turn);
#pragma unroll 1
        for (int q = 0; q < 8; ++q) {
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
          if (q == UPD_X3_HANDOFF - 1) turn_end(next_turn);
        }
        tc::wait_st();
        const float inv3 = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll 1
        for (int q = 0; q < 8; ++q) {
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
        const UpdNsStep st = reinterpret_cast<const UpdNsStep*>(smem + steps_off)[t];
#pragma unroll
        for (int f = 0; f < F; ++f) {
          float eps = pe[f] + sf(L.b4)[f];
          float sig = upd_softplus_accurate(ps[f] + sf(L.bs)[f]);
          y[f] = upd_ns_update(st, y[f], y0h[f], gxv[f], eps, sig, zn[f], last);
        }
      }
    }
    if (live) {
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
  size_t smem = upd_align128(upd_align128(L.tc_image_bytes) + STEP_BYTES * p.T) + sizeof(X3Sync) + 128;
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  auto kern = sampler_tc3_kernel<KIND, F>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  long long n_tiles = (p.n_rows + 127) / 128;
  long long ctas = (n_tiles + 2) / 3;
  int grid = (int)(ctas < sms ? ctas : sms);
  if (grid < 1) grid = 1;
  kern<<<grid, X3_THREADS, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace

cudaError_t upd_launch_sampler_tc3(const UpdSamplerParams& p, int kind, int F, int sms, cudaStream_t stream) {
#define UPD_CASE(KK, FF) if (kind == KK && F == FF) return launch<KK, FF>(p, sms, stream);
  UPD_CASE(0, 1) UPD_CASE(0, 2) UPD_CASE(0, 3) UPD_CASE(0, 4)
  UPD_CASE(1, 1) UPD_CASE(1, 2) UPD_CASE(1, 3) UPD_CASE(1, 4)
#undef UPD_CASE
  return cudaErrorInvalidValue;
}
