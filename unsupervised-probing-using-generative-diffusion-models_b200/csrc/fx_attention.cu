// De-stationary attention of the f(x) condition encoder on tcgen05 tensor cores:
//     out = softmax(scale * (tau_b * Q K^T + delta_b)) V        (DSAttention of torch-timeseries 0.1.10 as called from
//     models/Diffusion_model/NsDiff/mu_backbone.py:70-104; head size 64, sequences of 100..150 positions)
// One CTA = one (batch row, head): K and V are staged once, then blocks of 128 queries run through it;
// 256 threads = two warps per TMEM lane quadrant: thread (quadrant, lane) = query row of the block = TMEM lane, the two
// warps of a quadrant take alternate 16-key groups of the score row (partial max / sum exchanged through shared memory)
// and halves of the output columns -- round 1 ran one warp per quadrant and was bound by that warp's serial stream.
//
//   Q row -> fp16 hi/lo A operand in TMEM (scaled by tau*scale*log2e, so scores come out as base-2 exponents)
//   K, V^T of this (b,h) -> shared memory as fp16 hi/lo B operands (K-major, no-swizzle core-matrix layout)
//   S = Q K^T      12 tcgen05.mma (4 K-slices x hi*hi + lo*hi + hi*lo), accumulator in TMEM, N = S padded to 16
//   softmax        each thread owns a full score row in its TMEM lane: max, ex2, sum without any shuffle; the
//                  un-normalised probabilities are re-encoded IN PLACE as the fp16 hi/lo A operand of the next GEMM
//   O = P V        3 * S/16 tcgen05.mma, N = 64, accumulator over the (dead) Q columns
//   epilogue       O / rowsum -> written as the split operand [hi | lo | hi | 1 1 0..] of the out-projection GEMM,
//                  heads merged (row (b,l), columns h*64..h*64+63): no fp32 attention output ever reaches HBM.
//
// TMEM: 64 (Q, later O) + S_pad <= 256 columns -> two CTAs per SM overlap each other's serial phases.
// Same operand encodings / descriptors as the fused sampler (tc_helpers.cuh, checked by upd_selftest_umma).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tc_helpers.cuh"

namespace {

constexpr int DK = 64;            // head size the kernel is built for
constexpr int MAX_SP = 192;       // keys, padded to a multiple of 16
constexpr float LOG2E = 1.4426950408889634f;

struct __align__(8) AttnSync {
  unsigned long long mma_bar;
  uint32_t tmem_base;
  uint32_t pad;
};

struct FxAttnParams {
  const float* q; long long q_stride;       // row (b*Lq + l) at q + row*q_stride, head h at + h*64
  const float* k; const float* v; long long kv_stride;   // row (b*S + s)
  const float* tau;                         // [B] or null
  const float* delta; int delta_pitch;      // [B, pitch] already multiplied by `scale`, or null
  int B, H, Lq, S, causal;
  float scale;
  __half* a3;                               // [B*Lq, 3*H*64 + 8]
};

__device__ __forceinline__ uint32_t idesc_f16(int n) {      // D fp32, A/B fp16 K-major, M = 128, N = n
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | (8u << 24);
}

constexpr int FX_THREADS = 256;

template <bool CAUSAL>
__global__ void __launch_bounds__(FX_THREADS, 2) fx_attention_kernel(const FxAttnParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ AttnSync sync;
  __shared__ float pm[2][128], ps[2][128];                  // partial row max / row sum of the two warps of a quadrant
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, part = warp >> 2, row = quad * 32 + lane;
  const int bh = blockIdx.x, b = bh / p.H, h = bh - b * p.H;
  const int S = p.S, SP = (S + 15) & ~15, NIT = SP >> 4;
  // K [SP x 64]: elem(s,c) at (c/8)*KLBO + s*16 + (c%8)*2 with KLBO = SP*16 + 16: the extra 16 bytes between
  // K-adjacent core-matrix columns spread the eight 16-byte rows a quarter-warp stores over all banks (8-way conflicts
  // otherwise, 53 M per launch in the first ncu capture); the MMA sees it only as the descriptor's LBO.
  const uint32_t KLBO = (uint32_t)SP * 16u + 16u;
  unsigned char* k_hi = smem;
  unsigned char* k_lo = k_hi + 8 * KLBO;
  unsigned char* v_hi = k_lo + 8 * KLBO;                    // V^T [64 x SP] elem(c,s) at (s/8)*64*16 + c*16 + (s%8)*2
  unsigned char* v_lo = v_hi + SP * DK * 2;
  float* sdelta = reinterpret_cast<float*>(v_lo + SP * DK * 2);   // [SP] delta*log2e, -inf for the padding keys
  if (tid == 0) { tc::mbar_init(tc::smem_u32(&sync.mma_bar), 1); tc::fence_mbar_init(); }
  if (warp == 0) tc::tmem_alloc<256>(tc::smem_u32(&sync.tmem_base));

  // ---- stage K and V^T of this (b,h) as fp16 hi/lo B operands.  Both take exactly SP/16 iterations of the 128
  // threads; the loads of four iterations (64 registers) are issued before any is consumed, so a CTA pays two or
  // three HBM round trips here instead of one per iteration (the projections were just written: 2.5 GB, not in L2).
  //   K: task (s, 8 consecutive channels) -> one 16-byte core-matrix row each for hi and lo
  //   V: task (channel c, 8 consecutive keys) -> one 16-byte row of V^T; global reads coalesced over c
  // the first query block's rows are requested before the K / V staging, so that a CTA pays ONE memory round trip
  // before its first MMA instead of two (a CTA lives ~5 k clk, a round trip is ~1.5 k)
  float4 qv0[8];
  {
    const bool v0 = row < p.Lq;
    const float* qr = p.q + ((long long)b * p.Lq + (v0 ? row : 0)) * p.q_stride + h * DK + 32 * part;
#pragma unroll
    for (int c = 0; c < 8; ++c) qv0[c] = v0 ? *reinterpret_cast<const float4*>(qr + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int s0 = tid; s0 < SP; s0 += FX_THREADS)
    sdelta[s0] = s0 < S ? (p.delta ? p.delta[(long long)b * p.delta_pitch + s0] * LOG2E : 0.0f) : -INFINITY;
  const float* kb = p.k + (long long)b * S * p.kv_stride + h * DK;
  const float* vb = p.v + (long long)b * S * p.kv_stride + h * DK;
  const int NST = (NIT + 1) >> 1;                             // staging iterations of 256 tasks
  for (int bt = 0; bt < NST; bt += 4) {
    float4 ka[4], kc[4];
    float vx[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int it = bt + u;
      const int i = tid + FX_THREADS * it;
      const int s = i >> 3, c8 = i & 7;
      ka[u] = kc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (it < NST && s < S) {
        ka[u] = *reinterpret_cast<const float4*>(kb + (long long)s * p.kv_stride + c8 * 8);
        kc[u] = *reinterpret_cast<const float4*>(kb + (long long)s * p.kv_stride + c8 * 8 + 4);
      }
      const int s8 = i >> 6, c = i & 63;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int sv = s8 * 8 + j;
        vx[u][j] = (it < NST && sv < S) ? vb[(long long)sv * p.kv_stride + c] : 0.0f;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int it = bt + u;
      if (it >= NST) break;
      const int i = tid + FX_THREADS * it;
      if (i < SP * 8) {
        const int s = i >> 3, c8 = i & 7;
        uint4 hi, lo;
        tc::split_f16x2(ka[u].x, ka[u].y, hi.x, lo.x); tc::split_f16x2(ka[u].z, ka[u].w, hi.y, lo.y);
        tc::split_f16x2(kc[u].x, kc[u].y, hi.z, lo.z); tc::split_f16x2(kc[u].z, kc[u].w, hi.w, lo.w);
        const uint32_t off = (uint32_t)c8 * KLBO + (uint32_t)s * 16u;
        *reinterpret_cast<uint4*>(k_hi + off) = hi;
        *reinterpret_cast<uint4*>(k_lo + off) = lo;
      }
      if (i < SP * 8) {
        const int s8 = i >> 6, c = i & 63;
        uint4 hi, lo;
        tc::split_f16x2(vx[u][0], vx[u][1], hi.x, lo.x); tc::split_f16x2(vx[u][2], vx[u][3], hi.y, lo.y);
        tc::split_f16x2(vx[u][4], vx[u][5], hi.z, lo.z); tc::split_f16x2(vx[u][6], vx[u][7], hi.w, lo.w);
        const uint32_t off = (uint32_t)s8 * (DK * 16u) + (uint32_t)c * 16u;
        *reinterpret_cast<uint4*>(v_hi + off) = hi;
        *reinterpret_cast<uint4*>(v_lo + off) = lo;
      }
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor-core reads
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = sync.tmem_base;
  const uint32_t lane_sel = (uint32_t)(quad * 32) << 16;
  const uint32_t q_cols = tmem_base + lane_sel;             // columns [0,64): Q operand, later the O accumulator
  const uint32_t s_cols = q_cols + 64u;                     // columns [64, 64+SP): scores, then P operand
  const uint32_t bar = tc::smem_u32(&sync.mma_bar);
  const float qs = (p.tau ? p.tau[b] : 1.0f) * p.scale * LOG2E;
  const int Kd = p.H * DK;
  uint32_t parity = 0;

  for (int q0 = 0; q0 < p.Lq; q0 += 128) {                  // query blocks share the staged K / V
    const int l = q0 + row;
    const bool valid = l < p.Lq;
    // ---- the query row -> TMEM (hi words [16j,16j+8), lo words [16j+8,16j+16) per K-slice j): each warp of the
    // quadrant writes two of the four K-slices ----
    {
      const float* qr = p.q + ((long long)b * p.Lq + (valid ? l : 0)) * p.q_stride + h * DK + 32 * part;
      float4 qv[8];
      if (q0 == 0) {
#pragma unroll
        for (int c = 0; c < 8; ++c) qv[c] = qv0[c];
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) qv[c] = valid ? *reinterpret_cast<const float4*>(qr + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        uint32_t o[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 a = qv[4 * j + c];
          tc::split_f16x2(a.x * qs, a.y * qs, o[2 * c], o[8 + 2 * c]);
          tc::split_f16x2(a.z * qs, a.w * qs, o[2 * c + 1], o[8 + 2 * c + 1]);
        }
        tc::tmem_st16(q_cols + 16u * (2 * part + j), o);
      }
    }
    tc::wait_st();
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 0) {                                         // S = Q K^T
      tc::fence_after_sync();
      const uint32_t lbo = KLBO, id = idesc_f16(SP);
#pragma unroll
      for (int j = 0; j < DK / 16; ++j) {
        const uint32_t a_hi = tmem_base + 16u * j, a_lo = a_hi + 8u;
        const uint64_t b_hi = tc::smem_desc(tc::smem_u32(k_hi) + (uint32_t)(2 * j) * lbo, lbo, 128u);
        const uint64_t b_lo = tc::smem_desc(tc::smem_u32(k_lo) + (uint32_t)(2 * j) * lbo, lbo, 128u);
        tc::mma_f16_ts(tmem_base + 64u, a_lo, b_hi, id, j > 0);
        tc::mma_f16_ts(tmem_base + 64u, a_hi, b_lo, id, true);
        tc::mma_f16_ts(tmem_base + 64u, a_hi, b_hi, id, true);
      }
      tc::mma_commit(bar);
    }
    tc::mbar_wait(bar, parity);
    parity ^= 1u;
    tc::fence_after_sync();

    // ---- softmax over the row, this warp's 16-key groups (part, part + 2, ..).  Scores are base-2 exponents (Q was
    // pre-scaled); delta (and the -inf of the padding keys) comes from shared memory as a broadcast. ----
    const int s_end = CAUSAL ? min(S, l + 1) : S;            // causal: keys >= s_end are masked
    float m = -INFINITY;
    for (int g = part; g < NIT; g += 2) {
      uint32_t ra[16];
      tc::tmem_ld16(s_cols + 16u * g, ra);
      tc::wait_ld();
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 d4 = *reinterpret_cast<const float4*>(sdelta + 16 * g + j);
        float x0 = __uint_as_float(ra[j]) + d4.x, x1 = __uint_as_float(ra[j + 1]) + d4.y;
        float x2 = __uint_as_float(ra[j + 2]) + d4.z, x3 = __uint_as_float(ra[j + 3]) + d4.w;
        if (CAUSAL) {
          const int s = 16 * g + j;
          x0 = s < s_end ? x0 : -INFINITY; x1 = s + 1 < s_end ? x1 : -INFINITY;
          x2 = s + 2 < s_end ? x2 : -INFINITY; x3 = s + 3 < s_end ? x3 : -INFINITY;
        }
        m = fmaxf(fmaxf(m, fmaxf(x0, x1)), fmaxf(x2, x3));
      }
    }
    pm[part][row] = m;
    __syncthreads();
    m = fmaxf(pm[0][row], pm[1][row]);
    float sum = 0.0f;
    for (int g = part; g < NIT; g += 2) {
      uint32_t cur[16], o[16];
      tc::tmem_ld16(s_cols + 16u * g, cur);
      tc::wait_ld();
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const float2 d2 = *reinterpret_cast<const float2*>(sdelta + 16 * g + j);
        float e0, e1;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"((__uint_as_float(cur[j]) + d2.x) - m));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"((__uint_as_float(cur[j + 1]) + d2.y) - m));
        if (CAUSAL) {
          const int s = 16 * g + j;
          e0 = s < s_end ? e0 : 0.0f; e1 = s + 1 < s_end ? e1 : 0.0f;
        }
        sum += e0 + e1;
        tc::split_f16x2(e0, e1, o[j / 2], o[8 + j / 2]);
      }
      tc::tmem_st16(s_cols + 16u * g, o);
    }
    ps[part][row] = sum;
    tc::wait_st();
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 0) {                                         // O = P V   (accumulator over the Q columns)
      tc::fence_after_sync();
      const uint32_t lbo = (uint32_t)DK * 16u, id = idesc_f16(DK);
      for (int j = 0; j < NIT; ++j) {
        const uint32_t a_hi = tmem_base + 64u + 16u * j, a_lo = a_hi + 8u;
        const uint64_t b_hi = tc::smem_desc(tc::smem_u32(v_hi) + (uint32_t)(2 * j) * lbo, lbo, 128u);
        const uint64_t b_lo = tc::smem_desc(tc::smem_u32(v_lo) + (uint32_t)(2 * j) * lbo, lbo, 128u);
        tc::mma_f16_ts(tmem_base, a_lo, b_hi, id, j > 0);
        tc::mma_f16_ts(tmem_base, a_hi, b_lo, id, true);
        tc::mma_f16_ts(tmem_base, a_hi, b_hi, id, true);
      }
      tc::mma_commit(bar);
    }
    tc::mbar_wait(bar, parity);
    parity ^= 1u;
    tc::fence_after_sync();

    // ---- O / sum -> split operand of the out-projection, heads merged; each warp of the quadrant writes 32 of the
    // head's 64 columns ----
    const float inv = 1.0f / (ps[0][row] + ps[1][row]);
    __half* arow = p.a3 + ((long long)b * p.Lq + (valid ? l : 0)) * (3 * Kd + 8) + h * DK + 32 * part;
    {
      uint32_t r[2][16];
#pragma unroll
      for (int g = 0; g < 2; ++g) tc::tmem_ld16(q_cols + 16u * (2 * part + g), r[g]);
      tc::wait_ld();
      if (valid) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2)
            tc::split_f16x2(__uint_as_float(r[g][j]) * inv, __uint_as_float(r[g][j + 1]) * inv, hi[j / 2], lo[j / 2]);
          uint4* d_hi = reinterpret_cast<uint4*>(arow + 16 * g);
          uint4* d_lo = reinterpret_cast<uint4*>(arow + Kd + 16 * g);
          uint4* d_h2 = reinterpret_cast<uint4*>(arow + 2 * Kd + 16 * g);
          const uint4 h0 = make_uint4(hi[0], hi[1], hi[2], hi[3]), h1 = make_uint4(hi[4], hi[5], hi[6], hi[7]);
          d_hi[0] = h0; d_hi[1] = h1;
          d_lo[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]); d_lo[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
          d_h2[0] = h0; d_h2[1] = h1;
        }
        if (h == 0 && part == 0)                             // bias columns of the operand: 1, 1, 0 x 6
          *reinterpret_cast<uint4*>(p.a3 + ((long long)b * p.Lq + l) * (3 * Kd + 8) + 3 * Kd) = make_uint4(0x3C003C00u, 0u, 0u, 0u);
      }
    }
    tc::fence_before_sync();                                 // this block's TMEM reads precede the next block's writes
    __syncthreads();                                         // (both warps of a quadrant: the next Q write covers other columns)
  }
  if (warp == 0) tc::tmem_dealloc<256>(tmem_base);
}

}  // namespace

cudaError_t upd_launch_fx_attention(const float* q, long long q_stride, const float* k, const float* v, long long kv_stride,
                                    const float* tau, const float* delta, int delta_pitch, int B, int H, int Lq, int S,
                                    int causal, float scale, void* a3, cudaStream_t stream) {
  const int SP = (S + 15) & ~15;
  if (S < 1 || SP > MAX_SP || Lq < 1 || (q_stride & 3) || (kv_stride & 3)) return cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(q) & 15) || (reinterpret_cast<uintptr_t>(k) & 15) || (reinterpret_cast<uintptr_t>(v) & 15) ||
      (reinterpret_cast<uintptr_t>(a3) & 15))
    return cudaErrorInvalidValue;
  FxAttnParams p;
  p.q = q; p.q_stride = q_stride; p.k = k; p.v = v; p.kv_stride = kv_stride; p.tau = tau; p.delta = delta;
  p.delta_pitch = delta_pitch; p.B = B; p.H = H; p.Lq = Lq; p.S = S; p.causal = causal; p.scale = scale;
  p.a3 = reinterpret_cast<__half*>(a3);
  const size_t smem = (size_t)2 * 8 * (SP * 16 + 16) + (size_t)2 * SP * DK * 2 + (size_t)SP * sizeof(float);
  cudaError_t e;
  if (causal) {
    e = cudaFuncSetAttribute(fx_attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    fx_attention_kernel<true><<<(unsigned)(B * H), FX_THREADS, smem, stream>>>(p);
  } else {
    e = cudaFuncSetAttribute(fx_attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    fx_attention_kernel<false><<<(unsigned)(B * H), FX_THREADS, smem, stream>>>(p);
  }
  return cudaGetLastError();
}
