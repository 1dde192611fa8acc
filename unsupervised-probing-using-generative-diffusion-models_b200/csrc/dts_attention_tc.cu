// DiffusionTS attention on tcgen05 tensor cores (FullAttention / CrossAttention,
// models/Diffusion_model/DiffusionTS/diffusionts_transformer.py:126-203): softmax(q k^T / sqrt(hs)) v with head size 16,
// forward and backward (the backward feeds the Langevin refinement gradient, DiffusionTS.py:384-399).
//
// Head size 16 is exactly one kind::f16 K-step, so every contraction of the attention and of its gradient is a short
// chain of tcgen05.mma with fp32 accumulation in TMEM on error-compensated fp16 hi/lo operands (hi*hi + lo*hi + hi*lo,
// ~22 mantissa bits per factor -- the same encoding as the fused sampler and fx_attention.cu).  One CTA per (row, head).
//
// FORWARD (128 threads, thread i = query i of a block of 128 = TMEM lane i, two CTAs per SM):
//   S = Q K^T (3 MMAs, N = keys padded to 16) -> row max / ex2 / row sum inside the thread's TMEM lane -> un-normalised
//   P re-encoded IN PLACE as the A operand of O = P V (3 MMAs per 16 keys, N = 16) -> O / sum and the base-2
//   log-sum-exp (saved for the backward) to global memory.
//
// BACKWARD (512 threads = 4 warps per TMEM lane quadrant, one CTA per SM).  Like the FFMA kernel it replaces
// (dts_attention.cu) it never reduces across rows: probabilities are recomputed once with queries as rows (dQ) and once
// with keys as rows (dK, dV), so every output row is private to one TMEM lane.  A pass over a block of 128 rows:
//     S  = X  Yn^T     (X = scaled Q | K rows,  Yn = K | scaled Q of all columns)      3 MMAs, N = columns
//     dP = G  Ygn^T    (G = dO | V rows,        Ygn = V | dO)                          3 MMAs
//     p  = ex2(S - lse_query),  dS = p (dP - D_query)          16 warps, each a quarter of the 16-column groups,
//                                                              dS (and p) re-encoded in place as fp16 hi/lo A operands
//     query rows:  dQ = dS K            key rows:  dV = p^T dO,  dK = dS^T Q           3 MMAs per 16 columns, N = 16
//   (lse, D = dO.O per query: per-lane scalars when queries are rows, broadcast from shared memory when they are
//   columns.)  dO is tiny (1e-6 in the refinement loop, below fp16's normal range), and every gradient is linear in it:
//   the CTA scales its dO tile by a power of two so that max|dO| lands in [4, 8) and un-scales the three outputs.
//
// Operand forms in shared memory (K-major, no swizzle, core matrix = 8 rows x 16 bytes):
//   N-form of X [n x 16]  (B operand with N = n, K = 16):  elem(s, c) at (c/8)*LBO + s*16 + (c%8)*2, LBO = NP*16 + 16
//                          (the 16 spare bytes spread a quarter-warp's stores over the banks); one row = the two 16-byte
//                          words groups a TMEM A operand row needs, so row operands are copied from it
//   T-form of X^T [16 x n] (B operand with N = 16, K = n):  elem(c, s) at (s/8)*256 + c*16 + (s%8)*2
// Limits: keys S <= 224 (forward), S and Lq <= 224 (backward: TMEM holds S and dP side by side); beyond that the host
// entry points run the FFMA kernels of dts_attention.cu.
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

// N-form of mult * X[n x 16] (rows >= n are zero), hi and lo parts.
__device__ __forceinline__ void stage_nform(const float* __restrict__ src, long long stride, int n, int NP, float mult,
                                            unsigned char* hi, unsigned char* lo) {
  const uint32_t LBO = nform_lbo(NP);
  for (int i = threadIdx.x; i < NP * 2; i += blockDim.x) {
    const int s = i >> 1, c8 = i & 1;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (s < n) {
      a = *reinterpret_cast<const float4*>(src + (long long)s * stride + c8 * 8);
      b = *reinterpret_cast<const float4*>(src + (long long)s * stride + c8 * 8 + 4);
    }
    uint4 h, l;
    tc::split_f16x2(a.x * mult, a.y * mult, h.x, l.x); tc::split_f16x2(a.z * mult, a.w * mult, h.y, l.y);
    tc::split_f16x2(b.x * mult, b.y * mult, h.z, l.z); tc::split_f16x2(b.z * mult, b.w * mult, h.w, l.w);
    const uint32_t off = (uint32_t)c8 * LBO + (uint32_t)s * 16u;
    *reinterpret_cast<uint4*>(hi + off) = h;
    *reinterpret_cast<uint4*>(lo + off) = l;
  }
}

// T-form of mult * X^T [16 x n] (columns >= n are zero).
__device__ __forceinline__ void stage_tform(const float* __restrict__ src, long long stride, int n, int NP, float mult,
                                            unsigned char* hi, unsigned char* lo) {
  for (int i = threadIdx.x; i < NP * 2; i += blockDim.x) {
    const int c = i & 15, s8 = i >> 4;
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int s = s8 * 8 + j;
      x[j] = s < n ? src[(long long)s * stride + c] * mult : 0.0f;
    }
    uint4 h, l;
    tc::split_f16x2(x[0], x[1], h.x, l.x); tc::split_f16x2(x[2], x[3], h.y, l.y);
    tc::split_f16x2(x[4], x[5], h.z, l.z); tc::split_f16x2(x[6], x[7], h.w, l.w);
    const uint32_t off = (uint32_t)s8 * 256u + (uint32_t)c * 16u;
    *reinterpret_cast<uint4*>(hi + off) = h;
    *reinterpret_cast<uint4*>(lo + off) = l;
  }
}

// Row r of an N-form -> the 16 TMEM words of an A operand K-slice (hi words 0..7, lo words 8..15); zeros past NP.
__device__ __forceinline__ void row_operand(const unsigned char* hi, const unsigned char* lo, int NP, int r, uint32_t (&o)[16]) {
  if (r < NP) {
    const uint32_t LBO = nform_lbo(NP);
    const uint4 h0 = *reinterpret_cast<const uint4*>(hi + (uint32_t)r * 16u);
    const uint4 h1 = *reinterpret_cast<const uint4*>(hi + LBO + (uint32_t)r * 16u);
    const uint4 l0 = *reinterpret_cast<const uint4*>(lo + (uint32_t)r * 16u);
    const uint4 l1 = *reinterpret_cast<const uint4*>(lo + LBO + (uint32_t)r * 16u);
    o[0] = h0.x; o[1] = h0.y; o[2] = h0.z; o[3] = h0.w; o[4] = h1.x; o[5] = h1.y; o[6] = h1.z; o[7] = h1.w;
    o[8] = l0.x; o[9] = l0.y; o[10] = l0.z; o[11] = l0.w; o[12] = l1.x; o[13] = l1.y; o[14] = l1.z; o[15] = l1.w;
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = 0u;
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

// Combined T-form of mult * X^T: rows 0..15 = hi parts, rows 16..31 = lo parts (a B operand with N = 32 that yields
// [A hi(X)^T | A lo(X)^T] in one MMA; its first 16 rows alone are the N = 16 operand hi(X)^T):
// elem(c', s) at (s/8)*512 + c'*16 + (s%8)*2.
__device__ __forceinline__ void stage_tform32(const float* __restrict__ src, long long stride, int n, int NP, float mult,
                                              unsigned char* dst) {
  for (int i = threadIdx.x; i < NP * 2; i += blockDim.x) {
    const int c = i & 15, s8 = i >> 4;
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int s = s8 * 8 + j;
      x[j] = s < n ? src[(long long)s * stride + c] * mult : 0.0f;
    }
    uint4 h, l;
    tc::split_f16x2(x[0], x[1], h.x, l.x); tc::split_f16x2(x[2], x[3], h.y, l.y);
    tc::split_f16x2(x[4], x[5], h.z, l.z); tc::split_f16x2(x[6], x[7], h.w, l.w);
    const uint32_t off = (uint32_t)s8 * 512u + (uint32_t)c * 16u;
    *reinterpret_cast<uint4*>(dst + off) = h;
    *reinterpret_cast<uint4*>(dst + off + 256u) = l;
  }
}

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
  stage_nform(kb, p.kv_stride, S, SP, 1.0f, k_hi, k_lo);
  stage_tform32(vb, p.kv_stride, S, SP, 1.0f, vt);
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
#pragma unroll
      for (int c = 0; c < 4; ++c) qv[c] = valid ? *reinterpret_cast<const float4*>(qr + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
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
  // ---- D = dO . O per query and max |dO| of the tile ----
  float amax = 0.0f;
  for (int i = tid; i < LP; i += BWD_THREADS) {
    float dsum = 0.0f;
    if (i < Lq) {
#pragma unroll
      for (int c = 0; c < HS; c += 4) {
        const float4 g4 = *reinterpret_cast<const float4*>(gb + (long long)i * d + c);
        const float4 o4 = *reinterpret_cast<const float4*>(ob + (long long)i * d + c);
        dsum = fmaf(g4.x, o4.x, dsum); dsum = fmaf(g4.y, o4.y, dsum); dsum = fmaf(g4.z, o4.z, dsum); dsum = fmaf(g4.w, o4.w, dsum);
        amax = fmaxf(amax, fmaxf(fmaxf(fabsf(g4.x), fabsf(g4.y)), fmaxf(fabsf(g4.z), fabsf(g4.w))));
      }
    }
    sLD[2 * i] = i < Lq ? -p.lse[(long long)rh * Lq + i] : -INFINITY;
    sLD[2 * i + 1] = -dsum;
  }
  // K, V and Q do not depend on the scale of dO: stage them while the reduction's loads are in flight
  const float qs = p.scale * LOG2E;
  stage_nform(kb, p.kv_stride, S, SP, 1.0f, kn_hi, kn_lo);
  stage_nform(vb, p.kv_stride, S, SP, 1.0f, vn_hi, vn_lo);
  stage_nform(qb, p.q_stride, Lq, LP, qs, qn_hi, qn_lo);
  stage_tform32(kb, p.kv_stride, S, SP, 1.0f, kt);
  stage_tform32(qb, p.q_stride, Lq, LP, qs, qt);
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
  const float gs = ldexpf(1.0f, e), gs_inv = ldexpf(1.0f, -e);
  for (int i = tid; i < LP; i += BWD_THREADS) sLD[2 * i + 1] *= gs;    // same thread wrote it
  stage_nform(gb, d, Lq, LP, gs, gn_hi, gn_lo);
  stage_tform32(gb, d, Lq, LP, gs, gt);
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
                                        int R, int H, int Lq, int S, float scale, float* o, float* lse, cudaStream_t stream) {
  const int SP = (S + 15) & ~15;
  if (SP > MAX_NP || (q_stride & 3) || (kv_stride & 3) || !aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(o))
    return cudaErrorInvalidValue;
  DtsTcParams p = {};
  p.q = q; p.q_stride = q_stride; p.k = k; p.v = v; p.kv_stride = kv_stride; p.R = R; p.H = H; p.Lq = Lq; p.S = S;
  p.scale = scale; p.o = o; p.lse = lse;
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
