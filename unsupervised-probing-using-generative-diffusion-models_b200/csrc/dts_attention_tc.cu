// DiffusionTS attention on tcgen05 tensor cores (FullAttention / CrossAttention,
// models/Diffusion_model/DiffusionTS/diffusionts_transformer.py:126-203): softmax(q k^T / sqrt(hs)) v with head size 16,
// forward and backward (the backward feeds the Langevin refinement gradient, DiffusionTS.py:384-399).
//
// Head size 16 is exactly one kind::f16 K-step, so every contraction of the attention and of its gradient is a short
// chain of tcgen05.mma with fp32 accumulation in TMEM on error-compensated fp16 hi/lo operands (hi*hi + lo*hi + hi*lo,
// ~22 mantissa bits per factor -- the same encoding as the fused sampler and fx_attention.cu).  One CTA per (row, head).
//
// FORWARD (256 threads = two warps per TMEM lane quadrant, two CTAs per SM; thread (quadrant, lane) = query row of a block of
// 128, the two warps of a quadrant take alternate groups of 16 keys):
//   S = Q K^T (3 MMAs, N = keys padded to 16) -> partial row max, exchanged through shared memory -> 2^(s - max) and
//   partial row sums; un-normalised P re-encoded IN PLACE as the A operand of O = P V -> O / sum and the base-2
//   log-sum-exp (saved for the backward) to global memory.
//
// BACKWARD (256 threads, two CTAs per SM).  Like the FFMA kernel it replaces (dts_attention.cu) it never reduces across
// rows: probabilities are recomputed once with queries as rows (dQ) and once with keys as rows (dK, dV), so every output
// row is private to one TMEM lane.  A pass over a block of 128 rows walks the columns in chunks of <= 96 (TMEM: 64
// accumulator columns + 2 x 96 = 256 per CTA, so that two heads are in flight per SM and one CTA's elementwise phase
// overlaps the other's MMAs):
//     S  = X  Yn^T     (X = scaled Q | K rows,  Yn = K | scaled Q of the chunk's columns)   3 MMAs, both operands in
//     dP = G  Ygn^T    (G = dO | V rows,        Ygn = V | dO)                               3 MMAs  shared memory
//     p  = 2^(S - lse_query),  dS = p (dP - D_query)      packed fp32x2 math; dS (and p) re-encoded in place as fp16
//                                                         hi/lo A operands
//     query rows:  dQ += dS K          key rows:  dV += p^T dO,  dK += dS^T Q               2 MMAs per 16 columns and output
//   (lse, D = dO.O per query: per-lane scalars when queries are rows, broadcast from shared memory when they are
//   columns.)  dO is tiny (1e-6 in the refinement loop, below fp16's normal range), and every gradient is linear in it:
//   the CTA scales its dO tile by a power of two so that max|dO| lands in [4, 8) and un-scales the three outputs.
//
// Small-N MMAs.  The contractions over keys / queries have N = 16 outputs; a tcgen05.mma costs ~50 clk however small N is
// (measured: 39 N = 16 MMAs per 2000 clk, with dependent and with independent accumulators alike), so the hi*hi and hi*lo
// passes share ONE N = 32 instruction on a combined [hi | lo] operand and only lo*hi needs its own: 2 instead of 3.
//
// Operand forms in shared memory (K-major, no swizzle, core matrix = 8 rows x 16 bytes):
//   N-form of X [n x 16]  (B operand with N = n, K = 16; also the A operand of a 128-row block):
//                          elem(s, c) at (c/8)*LBO + s*16 + (c%8)*2, LBO = NP*16 + 16 (the 16 spare bytes spread a
//                          quarter-warp's stores over the banks)
//   combined T-form of X^T [32 x n] (B operand with N = 32 | 16, K = n): rows 0..15 hi, rows 16..31 lo,
//                          elem(c', s) at (s/8)*512 + c'*16 + (s%8)*2
// Limits: keys S <= 224 (forward), S and Lq <= 224 (backward); beyond that the host entry points run the FFMA kernels of
// dts_attention.cu.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "sampler_math.cuh"
#include "tc_helpers.cuh"

namespace {

constexpr int HS = 16;
constexpr int MAX_NP = 224;
constexpr float LOG2E = 1.4426950408889634f;

struct __align__(8) TcSync {
  unsigned long long mma_bar;
  uint32_t tmem_base;
  uint32_t pad;
};

struct DtsTcParams {
  const float* q; long long q_stride;       // row (r*Lq + i) at q + row*q_stride, head h at + h*16
  const float* k; const float* v; long long kv_stride;   // row (r*S + j)
  int R, H, Lq, S;
  float scale;
  float* o;                                 // [R*Lq, H*16]
  float* lse;                               // [R*H, Lq] base-2 log-sum-exp of the scaled scores (may be null in the forward)
  __half* a3;                               // forward, optional: split operand [R*Lq, 3*H*16 + 8] of the out-projection
  const float* d_o;                         // [R*Lq, H*16]
  float* dq; long long dq_stride;
  float* dk; float* dv; long long dkv_stride;
};

__device__ __forceinline__ uint32_t idesc_f16(int n) {      // D fp32, A/B fp16 K-major, M = 128, N = n
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | (8u << 24);
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t nform_lbo(int NP) { return (uint32_t)NP * 16u + 16u; }
__device__ __forceinline__ uint32_t nform_bytes(int NP) { return 2u * nform_lbo(NP); }

// Staging is split into a LOAD half (all global loads of a phase are issued before any is consumed: a CTA pays one
// memory round trip per phase, not one per loop iteration -- the first version spent a third of the backward kernel in
// long-scoreboard stalls) and a STORE half (convert, write the operand form).  Both kernels run 256 threads and
// NP <= 224, so a thread has at most two tasks per matrix.
constexpr int STG_THREADS = 256, STG_TASKS = 2;

// N-form tasks: (row s, 8 consecutive channels c8) -> one 16-byte core-matrix row each for hi and lo.
struct NTasks { float4 a[STG_TASKS], b[STG_TASKS]; };
__device__ __forceinline__ void load_ntasks(const float* __restrict__ src, long long stride, int n, int NP, NTasks& t) {
#pragma unroll
  for (int u = 0; u < STG_TASKS; ++u) {
    const int i = threadIdx.x + u * STG_THREADS, s = i >> 1, c8 = i & 1;
    t.a[u] = t.b[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < NP * 2 && s < n) {
      t.a[u] = *reinterpret_cast<const float4*>(src + (long long)s * stride + c8 * 8);
      t.b[u] = *reinterpret_cast<const float4*>(src + (long long)s * stride + c8 * 8 + 4);
    }
  }
}
// N-form of mult * X[n x 16] (rows >= n are zero), hi and lo parts.
__device__ __forceinline__ void store_nform(const NTasks& t, int NP, float mult, unsigned char* hi, unsigned char* lo) {
  const uint32_t LBO = nform_lbo(NP);
#pragma unroll
  for (int u = 0; u < STG_TASKS; ++u) {
    const int i = threadIdx.x + u * STG_THREADS, s = i >> 1, c8 = i & 1;
    if (i < NP * 2) {
      const float4 a = t.a[u], b = t.b[u];
      uint4 h, l;
      tc::split_f16x2(a.x * mult, a.y * mult, h.x, l.x); tc::split_f16x2(a.z * mult, a.w * mult, h.y, l.y);
      tc::split_f16x2(b.x * mult, b.y * mult, h.z, l.z); tc::split_f16x2(b.z * mult, b.w * mult, h.w, l.w);
      const uint32_t off = (uint32_t)c8 * LBO + (uint32_t)s * 16u;
      *reinterpret_cast<uint4*>(hi + off) = h;
      *reinterpret_cast<uint4*>(lo + off) = l;
    }
  }
}

// T-form tasks: (channel c, 8 consecutive rows s8) -> one 16-byte row of X^T each for hi and lo.
struct TTasks { float x[STG_TASKS][8]; };
__device__ __forceinline__ void load_ttasks(const float* __restrict__ src, long long stride, int n, int NP, TTasks& t) {
#pragma unroll
  for (int u = 0; u < STG_TASKS; ++u) {
    const int i = threadIdx.x + u * STG_THREADS, c = i & 15, s8 = i >> 4;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int s = s8 * 8 + j;
      t.x[u][j] = (i < NP * 2 && s < n) ? src[(long long)s * stride + c] : 0.0f;
    }
  }
}
// Combined T-form of mult * X^T: rows 0..15 = hi parts, rows 16..31 = lo parts (a B operand with N = 32 that yields
// [A hi(X)^T | A lo(X)^T] in one MMA; its first 16 rows alone are the N = 16 operand hi(X)^T):
// elem(c', s) at (s/8)*512 + c'*16 + (s%8)*2.
__device__ __forceinline__ void store_tform32(const TTasks& t, int NP, float mult, unsigned char* dst) {
#pragma unroll
  for (int u = 0; u < STG_TASKS; ++u) {
    const int i = threadIdx.x + u * STG_THREADS, c = i & 15, s8 = i >> 4;
    if (i < NP * 2) {
      const float* x = t.x[u];
      uint4 h, l;
      tc::split_f16x2(x[0] * mult, x[1] * mult, h.x, l.x); tc::split_f16x2(x[2] * mult, x[3] * mult, h.y, l.y);
      tc::split_f16x2(x[4] * mult, x[5] * mult, h.z, l.z); tc::split_f16x2(x[6] * mult, x[7] * mult, h.w, l.w);
      const uint32_t off = (uint32_t)s8 * 512u + (uint32_t)c * 16u;
      *reinterpret_cast<uint4*>(dst + off) = h;
      *reinterpret_cast<uint4*>(dst + off + 256u) = l;
    }
  }
}

// D (+)= A B^T with A = one hi/lo K-slice in TMEM (hi at a, lo at a + 8) and B = hi/lo descriptors: small terms first.
__device__ __forceinline__ void mma3(uint32_t d, uint32_t a, uint64_t b_hi, uint64_t b_lo, uint32_t idesc, bool acc) {
  tc::mma_f16_ts(d, a + 8u, b_hi, idesc, acc);
  tc::mma_f16_ts(d, a, b_lo, idesc, true);
  tc::mma_f16_ts(d, a, b_hi, idesc, true);
}

// ================================================ forward ==============================================================
constexpr int FWD_THREADS = 256;                           // two warps per TMEM lane quadrant (each half of the key groups)

__global__ void __launch_bounds__(FWD_THREADS, 2) dts_attn_tc_fwd_kernel(const DtsTcParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ TcSync sync;
  __shared__ float pm[2][128], ps[2][128];                  // partial row max / row sum of the two column halves
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, part = warp >> 2, row = quad * 32 + lane;
  const int rh = blockIdx.x, r = rh / p.H, h = rh - r * p.H;
  const int S = p.S, SP = (S + 15) & ~15, NIT = SP >> 4;
  unsigned char* k_hi = smem;
  unsigned char* k_lo = k_hi + nform_bytes(SP);
  unsigned char* vt = k_lo + nform_bytes(SP);               // combined T-form, 64 * SP bytes
  if (tid == 0) { tc::mbar_init(tc::smem_u32(&sync.mma_bar), 1); tc::fence_mbar_init(); }
  if (warp == 0) tc::tmem_alloc<256>(tc::smem_u32(&sync.tmem_base));
  const float* kb = p.k + (long long)r * S * p.kv_stride + h * HS;
  const float* vb = p.v + (long long)r * S * p.kv_stride + h * HS;
  float4 qv0[4];                                            // the first query block's rows, requested with the K / V tiles
  {
    const bool v0 = part == 0 && row < p.Lq;
    const float* qr = p.q + ((long long)r * p.Lq + (v0 ? row : 0)) * p.q_stride + h * HS;
#pragma unroll
    for (int c = 0; c < 4; ++c) qv0[c] = v0 ? *reinterpret_cast<const float4*>(qr + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  {
    NTasks tk;
    TTasks tv;
    load_ntasks(kb, p.kv_stride, S, SP, tk);
    load_ttasks(vb, p.kv_stride, S, SP, tv);
    store_nform(tk, SP, 1.0f, k_hi, k_lo);
    store_tform32(tv, SP, 1.0f, vt);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor-core reads
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = sync.tmem_base;
  const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
  const uint32_t o_cols = lane_base;                        // [0,16): Q operand, then [0,32): O accumulator [hi*hi | small terms]
  const uint32_t s_cols = lane_base + 32u;                  // [32, 32+SP): scores, then the P operand
  const uint32_t bar = tc::smem_u32(&sync.mma_bar);
  const float qs = p.scale * LOG2E;
  const int d = p.H * HS;
  uint32_t parity = 0;

  for (int q0 = 0; q0 < p.Lq; q0 += 128) {
    const int l = q0 + row;
    const bool valid = l < p.Lq;
    if (part == 0) {
      const float* qr = p.q + ((long long)r * p.Lq + (valid ? l : 0)) * p.q_stride + h * HS;
      float4 qv[4];
      if (q0 == 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) qv[c] = qv0[c];
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) qv[c] = valid ? *reinterpret_cast<const float4*>(qr + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      uint32_t o[16];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tc::split_f16x2(qv[c].x * qs, qv[c].y * qs, o[2 * c], o[8 + 2 * c]);
        tc::split_f16x2(qv[c].z * qs, qv[c].w * qs, o[2 * c + 1], o[8 + 2 * c + 1]);
      }
      tc::tmem_st16(o_cols, o);
      tc::wait_st();
    }
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 0) {                                         // S = Q K^T
      tc::fence_after_sync();
      const uint32_t lbo = nform_lbo(SP);
      mma3(tmem_base + 32u, tmem_base, tc::smem_desc(tc::smem_u32(k_hi), lbo, 128u), tc::smem_desc(tc::smem_u32(k_lo), lbo, 128u),
           idesc_f16(SP), false);
      tc::mma_commit(bar);
    }
    tc::mbar_wait(bar, parity);
    parity ^= 1u;
    tc::fence_after_sync();

    // ---- row max over this warp's groups (part, part + 2, ..); padding keys exist only in the last group ----
    float m = -INFINITY;
    for (int g = part; g < NIT; g += 2) {
      uint32_t ra[16];
      tc::tmem_ld16(s_cols + 16u * g, ra);
      tc::wait_ld();
      if (16 * g + 16 > S) {
#pragma unroll
        for (int j = 0; j < 16; ++j) if (16 * g + j < S) m = fmaxf(m, __uint_as_float(ra[j]));
      } else {
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          m = fmaxf(fmaxf(m, fmaxf(__uint_as_float(ra[j]), __uint_as_float(ra[j + 1]))),
                    fmaxf(__uint_as_float(ra[j + 2]), __uint_as_float(ra[j + 3])));
      }
    }
    pm[part][row] = m;
    __syncthreads();
    m = fmaxf(pm[0][row], pm[1][row]);
    // ---- p = 2^(s - m) in place as the fp16 hi/lo operand of O = P V, partial row sums ----
    float2 sum2 = make_float2(0.f, 0.f);
    const float2 nm2 = sm::splat(-m);
    for (int g = part; g < NIT; g += 2) {
      uint32_t ra[16], o[16];
      tc::tmem_ld16(s_cols + 16u * g, ra);
      tc::wait_ld();
      const bool ragged = 16 * g + 16 > S;
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const float2 x = sm::fadd2(make_float2(__uint_as_float(ra[j]), __uint_as_float(ra[j + 1])), nm2);
        float2 e = make_float2(ex2f(x.x), ex2f(x.y));
        if (ragged) { e.x = 16 * g + j < S ? e.x : 0.0f; e.y = 16 * g + j + 1 < S ? e.y : 0.0f; }
        sum2 = sm::fadd2(sum2, e);
        tc::split_f16x2(e.x, e.y, o[j / 2], o[8 + j / 2]);
      }
      tc::tmem_st16(s_cols + 16u * g, o);
    }
    ps[part][row] = sum2.x + sum2.y;
    tc::wait_st();
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 0) {                                         // O = P V: hi * [hi | lo] (N = 32), lo * hi (N = 16) per 16 keys
      tc::fence_after_sync();
      const uint32_t id32 = idesc_f16(32), id16 = idesc_f16(16);
      for (int j = 0; j < NIT; ++j) {
        const uint64_t b = tc::smem_desc(tc::smem_u32(vt) + (uint32_t)j * 1024u, 512u, 128u);
        const uint32_t a = tmem_base + 32u + 16u * j;
        tc::mma_f16_ts(tmem_base, a, b, id32, j > 0);
        tc::mma_f16_ts(tmem_base + 16u, a + 8u, b, id16, true);
      }
      tc::mma_commit(bar);
    }
    tc::mbar_wait(bar, parity);
    parity ^= 1u;
    tc::fence_after_sync();
    if (part == 0) {
      uint32_t big[16], small[16];
      tc::tmem_ld16(o_cols, big);
      tc::tmem_ld16(o_cols + 16u, small);
      tc::wait_ld();
      if (valid) {
        const float sum = ps[0][row] + ps[1][row];
        const float inv = 1.0f / sum;
        float* op = p.o + ((long long)r * p.Lq + l) * d + h * HS;
#pragma unroll
        for (int c = 0; c < HS; c += 4)
          *reinterpret_cast<float4*>(op + c) =
              make_float4((__uint_as_float(small[c]) + __uint_as_float(big[c])) * inv, (__uint_as_float(small[c + 1]) + __uint_as_float(big[c + 1])) * inv,
                          (__uint_as_float(small[c + 2]) + __uint_as_float(big[c + 2])) * inv, (__uint_as_float(small[c + 3]) + __uint_as_float(big[c + 3])) * inv);
        if (p.lse) p.lse[(long long)rh * p.Lq + l] = m + log2f(sum);
        if (p.a3) {                                          // the out-projection's operand, heads merged
          __half* arow = p.a3 + ((long long)r * p.Lq + l) * (3 * d + 8) + h * HS;
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int c = 0; c < HS; c += 2)
            tc::split_f16x2((__uint_as_float(small[c]) + __uint_as_float(big[c])) * inv,
                            (__uint_as_float(small[c + 1]) + __uint_as_float(big[c + 1])) * inv, hi[c / 2], lo[c / 2]);
          const uint4 h0 = make_uint4(hi[0], hi[1], hi[2], hi[3]), h1 = make_uint4(hi[4], hi[5], hi[6], hi[7]);
          reinterpret_cast<uint4*>(arow)[0] = h0; reinterpret_cast<uint4*>(arow)[1] = h1;
          reinterpret_cast<uint4*>(arow + d)[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          reinterpret_cast<uint4*>(arow + d)[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
          reinterpret_cast<uint4*>(arow + 2 * d)[0] = h0; reinterpret_cast<uint4*>(arow + 2 * d)[1] = h1;
          if (h == 0) *reinterpret_cast<uint4*>(arow + 3 * d) = make_uint4(0x3C003C00u, 0u, 0u, 0u);
        }
      }
    }
    tc::fence_before_sync();                                 // this block's TMEM reads precede the next block's writes
  }
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<256>(tmem_base);
}

// ================================================ backward =============================================================
constexpr int BWD_THREADS = 256;                           // two warps per TMEM lane quadrant, two CTAs per SM
constexpr int CHUNK_GROUPS = 6;                            // columns are processed in chunks of <= 6 groups of 16
constexpr uint32_t COL_OUT1 = 0, COL_OUT2 = 32;            // second-stage accumulators, 32 columns each: [hi*hi | small terms]
constexpr uint32_t COL_S = 64, COL_P = 64 + 16 * CHUNK_GROUPS;   // S / dS chunk and dP / p chunk (256 columns in all)

__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)acc) : "memory");
}

struct BwdPass {
  uint32_t x_hi, x_lo, g_hi, g_lo, lbo_r;                  // row operands: N-forms in shared memory (A of the first stage)
  uint32_t yn_hi, yn_lo, ygn_hi, ygn_lo, lbo_c; int NPc;   // column operands (B of the first stage)
  uint32_t t1, t2;                                         // combined T-forms: out1 = A1 t1, out2 = dS t2 (keys as rows only)
  const float* colLD; int ncols;                           // keys as rows: (-lse, -D) per column (query); else: valid columns (keys)
};

// First stage of one chunk (columns c0 .. c0 + nc): S = X Yn^T, dP = G Ygn^T, both operands from shared memory.
__device__ __forceinline__ void bwd_first_stage(const BwdPass& a, int row0, int c0, int nc, uint32_t tmem_base) {
  const uint32_t id = idesc_f16(nc), ro = (uint32_t)row0 * 16u, co = (uint32_t)c0 * 16u;
  const uint64_t xh = tc::smem_desc(a.x_hi + ro, a.lbo_r, 128u), xl = tc::smem_desc(a.x_lo + ro, a.lbo_r, 128u);
  const uint64_t gh = tc::smem_desc(a.g_hi + ro, a.lbo_r, 128u), gl = tc::smem_desc(a.g_lo + ro, a.lbo_r, 128u);
  const uint64_t yh = tc::smem_desc(a.yn_hi + co, a.lbo_c, 128u), yl = tc::smem_desc(a.yn_lo + co, a.lbo_c, 128u);
  const uint64_t vh = tc::smem_desc(a.ygn_hi + co, a.lbo_c, 128u), vl = tc::smem_desc(a.ygn_lo + co, a.lbo_c, 128u);
  mma_f16_ss(tmem_base + COL_S, xl, yh, id, false);
  mma_f16_ss(tmem_base + COL_S, xh, yl, id, true);
  mma_f16_ss(tmem_base + COL_S, xh, yh, id, true);
  mma_f16_ss(tmem_base + COL_P, gl, vh, id, false);
  mma_f16_ss(tmem_base + COL_P, gh, vl, id, true);
  mma_f16_ss(tmem_base + COL_P, gh, vh, id, true);
}

// Second stage of one chunk: slices j0 .. j0 + ns of the contraction.  A small-N tcgen05.mma costs ~50 clk whatever N
// is (measured: 39 N = 16 MMAs per 2000 clk, independent accumulators or not), so the hi*hi and hi*lo passes share one
// N = 32 instruction on the combined T-form: 2 instead of 3 MMAs per slice and output.
template <bool KEYS>
__device__ __forceinline__ void bwd_second_stage(const BwdPass& a, int j0, int ns, uint32_t tmem_base) {
  const uint32_t id32 = idesc_f16(32), id16 = idesc_f16(16);
  for (int jj = 0; jj < ns; ++jj) {
    const int j = j0 + jj;
    const bool acc = j > 0;
    const uint32_t a_s = tmem_base + COL_S + 16u * jj, a_p = tmem_base + COL_P + 16u * jj;
    const uint64_t b1 = tc::smem_desc(a.t1 + (uint32_t)j * 1024u, 512u, 128u);
    if (!KEYS) {                                            // dQ = dS K
      tc::mma_f16_ts(tmem_base + COL_OUT1, a_s, b1, id32, acc);
      tc::mma_f16_ts(tmem_base + COL_OUT1 + 16u, a_s + 8u, b1, id16, true);
    } else {                                                // dV = p^T dO, dK = dS^T Q
      const uint64_t b2 = tc::smem_desc(a.t2 + (uint32_t)j * 1024u, 512u, 128u);
      tc::mma_f16_ts(tmem_base + COL_OUT1, a_p, b1, id32, acc);
      tc::mma_f16_ts(tmem_base + COL_OUT2, a_s, b2, id32, acc);
      tc::mma_f16_ts(tmem_base + COL_OUT1 + 16u, a_p + 8u, b1, id16, true);
      tc::mma_f16_ts(tmem_base + COL_OUT2 + 16u, a_s + 8u, b2, id16, true);
    }
  }
}

// One block of 128 rows against all columns, chunk by chunk.  KEYS = rows are keys (out1 = p^T dO needs p as an operand,
// out2 = dS^T Q); else rows are queries (out1 = dS K).  nrowL / nrowD: minus this lane's lse / D (queries as rows; -inf
// switches a padding row off), unused otherwise.  On return the accumulators are complete (the caller reads them).
template <bool KEYS>
__device__ __forceinline__ void bwd_block(const BwdPass& a, int row0, float nrowL, float nrowD, uint32_t tmem_base, uint32_t bar,
                                          uint32_t& parity) {
  const float2 nrowL2 = sm::splat(nrowL), nrowD2 = sm::splat(nrowD);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int quad = warp & 3, part = warp >> 2;
  const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
  const int NIT = a.NPc >> 4;
  const int NCH = (NIT + CHUNK_GROUPS - 1) / CHUNK_GROUPS, base = NIT / NCH, rem = NIT - base * NCH;
  if (tid == 0) {
    tc::fence_after_sync();
    bwd_first_stage(a, row0, 0, 16 * (base + (rem > 0 ? 1 : 0)), tmem_base);
    tc::mma_commit(bar);
  }
  int g0 = 0;
  for (int c = 0; c < NCH; ++c) {
    const int ng = base + (c < rem ? 1 : 0);
    tc::mbar_wait(bar, parity);
    parity ^= 1u;
    tc::fence_after_sync();
    for (int g = part; g < ng; g += 2) {
      uint32_t rs[16], rp[16], ods[16], op[16];
      tc::tmem_ld16(lane_base + COL_S + 16u * g, rs);
      tc::tmem_ld16(lane_base + COL_P + 16u * g, rp);
      tc::wait_ld();
      const int col0 = 16 * (g0 + g);
      if (KEYS) {                                           // columns are queries: (-lse, -D) per column from shared memory
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const float4 c = *reinterpret_cast<const float4*>(a.colLD + 2 * (col0 + j));    // (-L0, -D0, -L1, -D1)
          const float2 x = sm::fadd2(make_float2(__uint_as_float(rs[j]), __uint_as_float(rs[j + 1])), make_float2(c.x, c.z));
          const float2 pp = make_float2(ex2f(x.x), ex2f(x.y));
          const float2 t = sm::fadd2(make_float2(__uint_as_float(rp[j]), __uint_as_float(rp[j + 1])), make_float2(c.y, c.w));
          const float2 ds = sm::fmul2(pp, t);
          tc::split_f16x2(ds.x, ds.y, ods[j / 2], ods[8 + j / 2]);
          tc::split_f16x2(pp.x, pp.y, op[j / 2], op[8 + j / 2]);
        }
        tc::tmem_st16(lane_base + COL_S + 16u * g, ods);
        tc::tmem_st16(lane_base + COL_P + 16u * g, op);
      } else {                                              // columns are keys: per-lane (-lse, -D); padding keys in the last group
        const bool ragged = col0 + 16 > a.ncols;
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const float2 x = sm::fadd2(make_float2(__uint_as_float(rs[j]), __uint_as_float(rs[j + 1])), nrowL2);
          float2 pp = make_float2(ex2f(x.x), ex2f(x.y));
          if (ragged) { pp.x = col0 + j < a.ncols ? pp.x : 0.0f; pp.y = col0 + j + 1 < a.ncols ? pp.y : 0.0f; }
          const float2 t = sm::fadd2(make_float2(__uint_as_float(rp[j]), __uint_as_float(rp[j + 1])), nrowD2);
          const float2 ds = sm::fmul2(pp, t);
          tc::split_f16x2(ds.x, ds.y, ods[j / 2], ods[8 + j / 2]);
        }
        tc::tmem_st16(lane_base + COL_S + 16u * g, ods);
      }
    }
    tc::wait_st();
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      // the next chunk's first stage overwrites the operand columns this chunk's second stage reads: tcgen05.mma
      // instructions of one thread execute in issue order
      tc::fence_after_sync();
      bwd_second_stage<KEYS>(a, g0, ng, tmem_base);
      if (c + 1 < NCH) bwd_first_stage(a, row0, 16 * (g0 + ng), 16 * (base + (c + 1 < rem ? 1 : 0)), tmem_base);
      tc::mma_commit(bar);
    }
    g0 += ng;
  }
  tc::mbar_wait(bar, parity);
  parity ^= 1u;
  tc::fence_after_sync();
}

// dst[0..16) = (small + big) * mult: the two halves of a 32-column accumulator, small terms first
__device__ __forceinline__ void store_row16(float* dst, const uint32_t (&b)[16], const uint32_t (&s)[16], float mult) {
#pragma unroll
  for (int c = 0; c < HS; c += 4)
    *reinterpret_cast<float4*>(dst + c) =
        make_float4((__uint_as_float(s[c]) + __uint_as_float(b[c])) * mult, (__uint_as_float(s[c + 1]) + __uint_as_float(b[c + 1])) * mult,
                    (__uint_as_float(s[c + 2]) + __uint_as_float(b[c + 2])) * mult, (__uint_as_float(s[c + 3]) + __uint_as_float(b[c + 3])) * mult);
}

__global__ void __launch_bounds__(BWD_THREADS, 2) dts_attn_tc_bwd_kernel(const DtsTcParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ TcSync sync;
  __shared__ float red[BWD_THREADS / 32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rh = blockIdx.x, r = rh / p.H, h = rh - r * p.H;
  const int S = p.S, SP = (S + 15) & ~15, Lq = p.Lq, LP = (Lq + 15) & ~15;
  const int d = p.H * HS;
  unsigned char* ptr = smem;
  auto take = [&](uint32_t bytes) { unsigned char* q = ptr; ptr += bytes; return q; };
  unsigned char *kn_hi = take(nform_bytes(SP)), *kn_lo = take(nform_bytes(SP));
  unsigned char *vn_hi = take(nform_bytes(SP)), *vn_lo = take(nform_bytes(SP));
  unsigned char *qn_hi = take(nform_bytes(LP)), *qn_lo = take(nform_bytes(LP));
  unsigned char *gn_hi = take(nform_bytes(LP)), *gn_lo = take(nform_bytes(LP));
  unsigned char* kt = take(64u * SP);
  unsigned char* qt = take(64u * LP);
  unsigned char* gt = take(64u * LP);
  float* sLD = reinterpret_cast<float*>(take(8u * LP));     // per query: (-lse (-inf for padding), -D in scaled units), D = dO . O
  if (tid == 0) { tc::mbar_init(tc::smem_u32(&sync.mma_bar), 1); tc::fence_mbar_init(); }
  if (warp == 0) tc::tmem_alloc<256>(tc::smem_u32(&sync.tmem_base));

  const float* qb = p.q + (long long)r * Lq * p.q_stride + h * HS;
  const float* kb = p.k + (long long)r * S * p.kv_stride + h * HS;
  const float* vb = p.v + (long long)r * S * p.kv_stride + h * HS;
  const float* gb = p.d_o + (long long)r * Lq * d + h * HS;
  const float* ob = p.o + (long long)r * Lq * d + h * HS;
  float gs, gs_inv;
  const float qs = p.scale * LOG2E;
  float amax = 0.0f;
  {
    // phase 1: every element of K, V, Q, dO, O once (N-form tasks), all loads in flight together
    NTasks tk, tv, tq, tg, to;
    load_ntasks(kb, p.kv_stride, S, SP, tk);
    load_ntasks(vb, p.kv_stride, S, SP, tv);
    load_ntasks(qb, p.q_stride, Lq, LP, tq);
    load_ntasks(gb, d, Lq, LP, tg);
    load_ntasks(ob, d, Lq, LP, to);
    float lse_u[STG_TASKS];
#pragma unroll
    for (int u = 0; u < STG_TASKS; ++u) {
      const int i = (tid + u * STG_THREADS) >> 1;
      lse_u[u] = i < Lq ? p.lse[(long long)rh * Lq + i] : INFINITY;
    }
    store_nform(tk, SP, 1.0f, kn_hi, kn_lo);
    store_nform(tv, SP, 1.0f, vn_hi, vn_lo);
    store_nform(tq, LP, qs, qn_hi, qn_lo);
    // D = dO . O per query (a row = two tasks of neighbouring lanes) and max |dO| of the tile
#pragma unroll
    for (int u = 0; u < STG_TASKS; ++u) {
      const float4 ga = tg.a[u], gb4 = tg.b[u], oa = to.a[u], ob4 = to.b[u];
      float dsum = ga.x * oa.x;
      dsum = fmaf(ga.y, oa.y, dsum); dsum = fmaf(ga.z, oa.z, dsum); dsum = fmaf(ga.w, oa.w, dsum);
      dsum = fmaf(gb4.x, ob4.x, dsum); dsum = fmaf(gb4.y, ob4.y, dsum); dsum = fmaf(gb4.z, ob4.z, dsum); dsum = fmaf(gb4.w, ob4.w, dsum);
      dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
      amax = fmaxf(amax, fmaxf(fmaxf(fmaxf(fabsf(ga.x), fabsf(ga.y)), fmaxf(fabsf(ga.z), fabsf(ga.w))),
                               fmaxf(fmaxf(fabsf(gb4.x), fabsf(gb4.y)), fmaxf(fabsf(gb4.z), fabsf(gb4.w)))));
      const int t = tid + u * STG_THREADS, i = t >> 1;
      if (t < LP * 2 && (t & 1) == 0) { sLD[2 * i] = -lse_u[u]; sLD[2 * i + 1] = -dsum; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (lane == 0) red[warp] = amax;
    __syncthreads();
    amax = red[0];
#pragma unroll
    for (int w = 1; w < BWD_THREADS / 32; ++w) amax = fmaxf(amax, red[w]);
    // power of two that brings max|dO| into [4, 8); 1 for an all-zero (or non-finite) tile
    int e = 0;
    if (amax > 0.0f && amax < INFINITY) {
      e = 2 - ilogbf(amax);
      e = max(-120, min(120, e));
    }
    gs = ldexpf(1.0f, e);
    gs_inv = ldexpf(1.0f, -e);
    store_nform(tg, LP, gs, gn_hi, gn_lo);
#pragma unroll
    for (int u = 0; u < STG_TASKS; ++u) {
      const int t = tid + u * STG_THREADS;
      if (t < LP * 2 && (t & 1) == 0) sLD[2 * (t >> 1) + 1] *= gs;          // same thread wrote it
    }
  }
  {
    // phase 2: the transposed forms (the tiles were just read: L1 / L2 hits)
    TTasks xk, xq, xg;
    load_ttasks(kb, p.kv_stride, S, SP, xk);
    load_ttasks(qb, p.q_stride, Lq, LP, xq);
    load_ttasks(gb, d, Lq, LP, xg);
    store_tform32(xk, SP, 1.0f, kt);
    store_tform32(xq, LP, qs, qt);
    store_tform32(xg, LP, gs, gt);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = sync.tmem_base;
  const uint32_t bar = tc::smem_u32(&sync.mma_bar);
  const int quad = warp & 3, part = warp >> 2;
  const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
  uint32_t parity = 0;

  // ---- queries as rows: dQ ----
  {
    BwdPass a;
    a.x_hi = tc::smem_u32(qn_hi); a.x_lo = tc::smem_u32(qn_lo); a.g_hi = tc::smem_u32(gn_hi); a.g_lo = tc::smem_u32(gn_lo);
    a.lbo_r = nform_lbo(LP);
    a.yn_hi = tc::smem_u32(kn_hi); a.yn_lo = tc::smem_u32(kn_lo); a.ygn_hi = tc::smem_u32(vn_hi); a.ygn_lo = tc::smem_u32(vn_lo);
    a.lbo_c = nform_lbo(SP); a.NPc = SP;
    a.t1 = tc::smem_u32(kt); a.t2 = 0;
    a.colLD = nullptr; a.ncols = S;
    for (int row0 = 0; row0 < Lq; row0 += 128) {
      const int i = row0 + quad * 32 + lane;
      const bool valid = i < Lq;
      bwd_block<false>(a, row0, valid ? sLD[2 * i] : -INFINITY, valid ? sLD[2 * i + 1] : 0.0f, tmem_base, bar, parity);
      if (part == 0) {
        uint32_t big[16], small[16];
        tc::tmem_ld16(lane_base + COL_OUT1, big);
        tc::tmem_ld16(lane_base + COL_OUT1 + 16u, small);
        tc::wait_ld();
        if (valid) store_row16(p.dq + ((long long)r * Lq + i) * p.dq_stride + h * HS, big, small, p.scale * gs_inv);
      }
      tc::fence_before_sync();
      __syncthreads();                                      // accumulators read before the next block's first MMA may run
    }
  }
  // ---- keys as rows: dV, dK ----
  {
    BwdPass a;
    a.x_hi = tc::smem_u32(kn_hi); a.x_lo = tc::smem_u32(kn_lo); a.g_hi = tc::smem_u32(vn_hi); a.g_lo = tc::smem_u32(vn_lo);
    a.lbo_r = nform_lbo(SP);
    a.yn_hi = tc::smem_u32(qn_hi); a.yn_lo = tc::smem_u32(qn_lo); a.ygn_hi = tc::smem_u32(gn_hi); a.ygn_lo = tc::smem_u32(gn_lo);
    a.lbo_c = nform_lbo(LP); a.NPc = LP;
    a.t1 = tc::smem_u32(gt); a.t2 = tc::smem_u32(qt);
    a.colLD = sLD; a.ncols = Lq;
    for (int row0 = 0; row0 < S; row0 += 128) {
      const int j = row0 + quad * 32 + lane;
      bwd_block<true>(a, row0, 0.0f, 0.0f, tmem_base, bar, parity);
      if (part == 0) {
        uint32_t b1[16], s1[16], b2[16], s2[16];
        tc::tmem_ld16(lane_base + COL_OUT1, b1);
        tc::tmem_ld16(lane_base + COL_OUT1 + 16u, s1);
        tc::tmem_ld16(lane_base + COL_OUT2, b2);
        tc::tmem_ld16(lane_base + COL_OUT2 + 16u, s2);
        tc::wait_ld();
        if (j < S) {
          store_row16(p.dv + ((long long)r * S + j) * p.dkv_stride + h * HS, b1, s1, gs_inv);
          store_row16(p.dk + ((long long)r * S + j) * p.dkv_stride + h * HS, b2, s2, gs_inv * (1.0f / LOG2E));
        }
      }
      tc::fence_before_sync();
      __syncthreads();
    }
  }
  if (warp == 0) tc::tmem_dealloc<256>(tmem_base);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// the first-stage A operand of the last row block reads up to 128 rows past row0: keep that inside the allocation
size_t bwd_smem_bytes(int SP, int LP) {
  return (size_t)4 * (2 * ((size_t)SP * 16 + 16)) + (size_t)4 * (2 * ((size_t)LP * 16 + 16)) + (size_t)64 * SP +
         (size_t)2 * 64 * LP + (size_t)8 * LP + 4096;
}

}  // namespace

// cudaErrorInvalidValue = shape outside the tensor-core kernels' limits (the caller runs the FFMA kernels instead).
cudaError_t upd_launch_dts_attention_tc(const float* q, long long q_stride, const float* k, const float* v, long long kv_stride,
                                        int R, int H, int Lq, int S, float scale, float* o, float* lse, void* a3,
                                        cudaStream_t stream) {
  const int SP = (S + 15) & ~15;
  if (a3 && (((H * HS) & 7) || !aligned16(a3))) return cudaErrorInvalidValue;
  if (SP > MAX_NP || (q_stride & 3) || (kv_stride & 3) || !aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(o))
    return cudaErrorInvalidValue;
  DtsTcParams p = {};
  p.q = q; p.q_stride = q_stride; p.k = k; p.v = v; p.kv_stride = kv_stride; p.R = R; p.H = H; p.Lq = Lq; p.S = S;
  p.scale = scale; p.o = o; p.lse = lse; p.a3 = reinterpret_cast<__half*>(a3);
  const size_t smem = (size_t)2 * (2 * ((size_t)SP * 16 + 16)) + (size_t)64 * SP;
  cudaError_t e = cudaFuncSetAttribute(dts_attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dts_attn_tc_fwd_kernel<<<(unsigned)(R * H), FWD_THREADS, smem, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t upd_launch_dts_attention_tc_bwd(const float* q, long long q_stride, const float* k, const float* v,
                                            long long kv_stride, int R, int H, int Lq, int S, float scale, const float* o,
                                            const float* lse, const float* d_o, float* dq, long long dq_stride, float* dk,
                                            float* dv, long long dkv_stride, cudaStream_t stream) {
  const int SP = (S + 15) & ~15, LP = (Lq + 15) & ~15;
  if (SP > MAX_NP || LP > MAX_NP || (q_stride & 3) || (kv_stride & 3) || (dq_stride & 3) || (dkv_stride & 3) || !aligned16(q) ||
      !aligned16(k) || !aligned16(v) || !aligned16(o) || !aligned16(d_o) || !aligned16(dq) || !aligned16(dk) || !aligned16(dv))
    return cudaErrorInvalidValue;
  DtsTcParams p = {};
  p.q = q; p.q_stride = q_stride; p.k = k; p.v = v; p.kv_stride = kv_stride; p.R = R; p.H = H; p.Lq = Lq; p.S = S;
  p.scale = scale; p.o = const_cast<float*>(o); p.lse = const_cast<float*>(lse); p.d_o = d_o; p.dq = dq;
  p.dq_stride = dq_stride; p.dk = dk; p.dv = dv; p.dkv_stride = dkv_stride;
  const size_t smem = bwd_smem_bytes(SP, LP);
  cudaError_t e = cudaFuncSetAttribute(dts_attn_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dts_attn_tc_bwd_kernel<<<(unsigned)(R * H), BWD_THREADS, smem, stream>>>(p);
  return cudaGetLastError();
}
