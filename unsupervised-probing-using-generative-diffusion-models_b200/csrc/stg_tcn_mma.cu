// Front half of a DiffSTG / NsDiff_spatial ResidualBlock on the tensor cores (models/Diffusion_model/DiffSTG/ugnet.py:117-129):
//     h1 = causal_conv3(x) + b1[step]   (TcnBlock 1: the 1x1 shortcut is folded into tap 2 by the host, + t_conv(time emb))
//     h2 = causal_conv3(h1) + b2        (TcnBlock 2)
//     hn = LayerNorm_c(h2) * g + beta   (nn.LayerNorm([1, c]) over the channel axis, eps 1e-5)
//     sc = W_sc x                       (the block's own 1x1 shortcut, optional)
// The FFMA kernel of stg_steps.cu does this at ~20 TFLOP/s fp32 (FMA pipe 43 % busy, latency-bound).  Here each
// convolution is a GEMM  D[position, c_out] = sum_{tap, c_in} x[c_in][position + tap - 2] * w[c_out][c_in][tap]  issued as
// warp-level mma.sync.m16n8k16 (fp16 operands, fp32 accumulate): M = 16 positions, N = 8 output channels, one K-slice =
// 16 input channels of one tap.  N <= 16 and K = 3 c_in <= 96 are far below a tcgen05 tile (M = 128, operands through
// shared-memory descriptors / TMEM) and the pass is bounded by HBM once the contraction leaves the FMA pipe, so the
// register-operand warp MMA is the matching instrument.  fp32-grade accuracy comes from the error-compensated split the
// other kernels use: x = hi + lo (fp16 each; values below fp16's normal range keep an absolute error of 2^-25), three
// MMAs per tile (lo*hi + hi*lo + hi*hi); ~1e-6 of the output rms against the fp32 library convolution.
//
// The split is done ONCE per element, on the way into shared memory: a shared-memory word pair holds the packed fp16
// hi parts and the packed lo parts of two adjacent channels at one position -- exactly one A-fragment register each --
// so the inner loop is 8-byte shared loads and MMAs only (the first version split tf32 operands in the loop: 46 warp
// instructions per position, 11 % of them MMAs).
//
// One CTA (4 warps) walks rows n = blockIdx.x, + gridDim.x, ...:
//   load x[n], split -> shared [c_in / 2][TP] (4 zero halo columns in front: both convolutions are causal, zero-padded)
//   phase 1: every warp takes pairs of 16-position tiles; A fragments = four 8-byte shared loads per K-slice (row pitch
//            = 4 mod 16 pairs: conflict-free), B fragments = one 16-byte load per lane of the pre-split weights; the
//            accumulator fragment of a thread IS a channel pair at one position: h1 is split and stored the same way
//   phase 2: the same on h1; LayerNorm over the channels of a position = the thread's own columns + two shuffles over the
//            four lanes that share a row; result -> shared staging
//   copy-out: coalesced stores of the fp32 row, or of the fp16 split operand [hi | lo | hi | 1 1 0..] of the
//            down-sampling GEMM that follows (upd_gemm3).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "tc_helpers.cuh"
#include "upd_common.cuh"

namespace {

constexpr int TCN_WARPS = 4;
constexpr int TCN_THREADS = TCN_WARPS * 32;
constexpr int MT = 2;                        // position tiles a warp holds at once (B fragments are reused across them)

__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// {packed fp16 hi parts, packed fp16 lo parts} of two values adjacent in K (the even one in the low half)
__device__ __forceinline__ uint2 split_pair(float a, float b) {
  uint2 r;
  tc::split_f16x2(a, b, r.x, r.y);
  return r;
}

// acc[m] += A[m] * B over one K-slice of 16, three passes, small terms first.  w = {hi(b0), hi(b1), lo(b0), lo(b1)}
__device__ __forceinline__ void mma3(float (&acc)[MT][4], const uint32_t (&ahi)[MT][4], const uint32_t (&alo)[MT][4], uint4 w) {
#pragma unroll
  for (int m = 0; m < MT; ++m) mma_f16(acc[m], alo[m], w.x, w.y);
#pragma unroll
  for (int m = 0; m < MT; ++m) mma_f16(acc[m], ahi[m], w.z, w.w);
#pragma unroll
  for (int m = 0; m < MT; ++m) mma_f16(acc[m], ahi[m], w.x, w.y);
}

// A fragments (hi / lo) of MT position tiles for one K-slice: src [channel pairs][TP] of {hi, lo} words, position p at
// column p + 4; p0 points at (first pair of the slice + tig, column P0 + g + tap + 2).  HALF: only the first 8 channels of
// the slice exist (the other fragment registers are zero).
template <bool HALF>
__device__ __forceinline__ void load_a(const uint2* __restrict__ p0, int TP, uint32_t (&ahi)[MT][4], uint32_t (&alo)[MT][4]) {
#pragma unroll
  for (int m = 0; m < MT; ++m) {
    const uint2* p = p0 + 16 * m;
    const uint2 v0 = p[0], v1 = p[8];
    ahi[m][0] = v0.x; alo[m][0] = v0.y;
    ahi[m][1] = v1.x; alo[m][1] = v1.y;
    if (HALF) {
      ahi[m][2] = alo[m][2] = ahi[m][3] = alo[m][3] = 0u;
    } else {
      const uint2 v2 = p[4 * TP], v3 = p[4 * TP + 8];
      ahi[m][2] = v2.x; alo[m][2] = v2.y;
      ahi[m][3] = v3.x; alo[m][3] = v3.y;
    }
  }
}

// B fragments of this lane for a K-slice whose 16 inputs are channels cb .. cb+15 of w [C][CIN] (element stride `es`)
template <int C>
__device__ __forceinline__ uint4 pack_b(const float* __restrict__ w, int CIN, int es, int off, int nt, int cb, int lane) {
  const int co = nt * 8 + (lane >> 2), ci = cb + 2 * (lane & 3);
  auto at = [&](int c) { return (co < C && c < CIN) ? w[(co * CIN + c) * es + off] : 0.0f; };
  const uint2 b0 = split_pair(at(ci), at(ci + 1)), b1 = split_pair(at(ci + 8), at(ci + 9));
  return make_uint4(b0.x, b1.x, b0.y, b1.y);
}

// C: c_out in {4, 8, 16}; KS: 16-channel K-slices of the input (c_in <= 16 KS); HALF1: c_in <= 8
template <int C, int KS, bool HALF1>
__global__ void __launch_bounds__(TCN_THREADS) stg_tcn_mma_kernel(
    const float* __restrict__ x, const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
    const float* __restrict__ b2, const float* __restrict__ gamma, const float* __restrict__ beta, long long N, int CI, int T,
    float* __restrict__ hn, __half* __restrict__ a3, const float* __restrict__ wsc, float* __restrict__ sc_out,
    const float* __restrict__ x2, int CI1) {
  constexpr int NT = (C + 7) / 8;            // n-tiles of 8 output channels
  constexpr int CP = NT * 8;                 // c_out padded
  constexpr bool HALF2 = (C <= 8);           // phase 2 reads 8 channels of h1
  constexpr int XPAIRS = HALF1 ? 4 : 8 * KS; // channel pairs of x held in shared memory
  constexpr int HPAIRS = HALF2 ? 4 : 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int pairs = (T + 16 * MT - 1) / (16 * MT);              // pairs of 16-position tiles in a row
  const int TPAD = pairs * 16 * MT;
  const int TP = TPAD + 4;                   // pitch in {hi, lo} word pairs, = 4 mod 16: conflict-free 8-byte fragment loads
  const int TPO = TPAD + 4;                  // staging pitch in floats, = 4 mod 32: conflict-free fragment stores
  uint2* sx = reinterpret_cast<uint2*>(smem_raw);               // [XPAIRS][TP]
  uint2* sh = sx + XPAIRS * TP;              // [HPAIRS][TP]
  float* so = reinterpret_cast<float*>(sh + HPAIRS * TP);       // [CP][TPO]
  uint4* sw1 = reinterpret_cast<uint4*>(so + CP * TPO);         // [3][KS][NT][32]
  uint4* sw2 = sw1 + 3 * KS * NT * 32;       // [3][NT][32]
  uint4* swsc = sw2 + 3 * NT * 32;           // [KS][NT][32]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;

  // ---- once per CTA: zero the activation buffers (halo, tail and padding rows stay zero), split the weights ----
  for (int i = tid; i < (XPAIRS + HPAIRS) * TP; i += TCN_THREADS) sx[i] = make_uint2(0u, 0u);
  for (int i = tid; i < 3 * KS * NT * 32; i += TCN_THREADS) {
    const int l = i & 31, nt = (i >> 5) % NT, s = (i >> 5) / NT, ks = s % KS, tap = s / KS;
    sw1[i] = pack_b<C>(w1, CI, 3, tap, nt, ks * 16, l);
  }
  for (int i = tid; i < 3 * NT * 32; i += TCN_THREADS) {
    const int l = i & 31, nt = (i >> 5) % NT, tap = (i >> 5) / NT;
    sw2[i] = pack_b<C>(w2, C, 3, tap, nt, 0, l);
  }
  if (wsc)
    for (int i = tid; i < KS * NT * 32; i += TCN_THREADS) {
      const int l = i & 31, nt = (i >> 5) % NT, ks = (i >> 5) / NT;
      swsc[i] = pack_b<C>(wsc, CI, 1, 0, nt, ks * 16, l);
    }
  // this thread's accumulator columns: channels nt*8 + 2*tig + {0, 1}
  float b1r[NT][2], b2r[NT][2], gr[NT][2], ber[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c = nt * 8 + 2 * tig + j;
      const bool ok = c < C;
      b1r[nt][j] = ok ? b1[c] : 0.0f;
      b2r[nt][j] = ok ? b2[c] : 0.0f;
      gr[nt][j] = ok ? gamma[c] : 0.0f;
      ber[nt][j] = ok ? beta[c] : 0.0f;
    }
  const int Q = T >> 2;                      // float4 groups per channel row (T % 4 == 0)
  const int K = C * T;
  const int items = ((CI + 1) >> 1) * Q;     // (channel pair, group of 4 positions) units of the x load
  __syncthreads();                           // the zero fill is complete before the first row lands on top of it

  for (long long n = blockIdx.x; n < N; n += gridDim.x) {
    // ---- x[n] -> split -> shared: all global loads of a batch are issued before the first is consumed ----
    const float* xr = x + n * (long long)CI1 * T;
    const float* xr2 = x2 ? x2 + n * (long long)(CI - CI1) * T : nullptr;
    constexpr int U = 4;
    for (int i0 = tid; i0 < items; i0 += TCN_THREADS * U) {
      float4 va[U], vb[U];
      int dst[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * TCN_THREADS;
        dst[u] = -1;
        va[u] = vb[u] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (i < items) {
          const int cp = i / Q, q = i - cp * Q, c0 = 2 * cp, c1 = c0 + 1;
          dst[u] = cp * TP + 4 + 4 * q;
          va[u] = __ldg(reinterpret_cast<const float4*>((c0 < CI1 ? xr + (long long)c0 * T : xr2 + (long long)(c0 - CI1) * T) + 4 * q));
          if (c1 < CI)
            vb[u] = __ldg(reinterpret_cast<const float4*>((c1 < CI1 ? xr + (long long)c1 * T : xr2 + (long long)(c1 - CI1) * T) + 4 * q));
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (dst[u] >= 0) {
          const uint2 e0 = split_pair(va[u].x, vb[u].x), e1 = split_pair(va[u].y, vb[u].y);
          const uint2 e2 = split_pair(va[u].z, vb[u].z), e3 = split_pair(va[u].w, vb[u].w);
          uint4* d = reinterpret_cast<uint4*>(sx + dst[u]);
          d[0] = make_uint4(e0.x, e0.y, e1.x, e1.y);
          d[1] = make_uint4(e2.x, e2.y, e3.x, e3.y);
        }
    }
    __syncthreads();
    // ---- phase 1: h1 = conv3(x) + b1 (and the 1x1 shortcut from the tap-2 fragments) ----
    for (int pi = warp; pi < pairs; pi += TCN_WARPS) {
      const int P0 = pi * 16 * MT;
      float acc[NT][MT][4], scv[NT][MT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          acc[nt][m][0] = acc[nt][m][2] = b1r[nt][0];
          acc[nt][m][1] = acc[nt][m][3] = b1r[nt][1];
          scv[nt][m][0] = scv[nt][m][1] = scv[nt][m][2] = scv[nt][m][3] = 0.0f;
        }
      const uint2* pa = sx + tig * TP + P0 + g + 2;
#pragma unroll
      for (int tap = 0; tap < 3; ++tap)
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          uint32_t ahi[MT][4], alo[MT][4];
          load_a<HALF1>(pa + ks * 8 * TP + tap, TP, ahi, alo);
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) mma3(acc[nt], ahi, alo, sw1[((tap * KS + ks) * NT + nt) * 32 + lane]);
          if (tap == 2 && wsc) {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) mma3(scv[nt], ahi, alo, swsc[(ks * NT + nt) * 32 + lane]);
          }
        }
      // the accumulator fragment holds channel pair nt*4 + tig at positions g and g + 8: h1 in the operand layout
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          uint2* d = sh + (nt * 4 + tig) * TP + 4 + P0 + 16 * m + g;
          d[0] = split_pair(acc[nt][m][0], acc[nt][m][1]);
          d[8] = split_pair(acc[nt][m][2], acc[nt][m][3]);
        }
      if (wsc) {
        float* so_g = sc_out + n * (long long)K;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const int c = nt * 8 + 2 * tig + (r & 1), p = P0 + 16 * m + g + 8 * (r >> 1);
              if (c < C && p < T) so_g[c * T + p] = scv[nt][m][r];
            }
      }
    }
    __syncthreads();
    // ---- phase 2: h2 = conv3(h1) + b2, LayerNorm over the channels ----
    for (int pi = warp; pi < pairs; pi += TCN_WARPS) {
      const int P0 = pi * 16 * MT;
      float acc[NT][MT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          acc[nt][m][0] = acc[nt][m][2] = b2r[nt][0];
          acc[nt][m][1] = acc[nt][m][3] = b2r[nt][1];
        }
      const uint2* pa = sh + tig * TP + P0 + g + 2;
#pragma unroll
      for (int tap = 0; tap < 3; ++tap) {
        uint32_t ahi[MT][4], alo[MT][4];
        load_a<HALF2>(pa + tap, TP, ahi, alo);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma3(acc[nt], ahi, alo, sw2[(tap * NT + nt) * 32 + lane]);
      }
#pragma unroll
      for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int h = 0; h < 2; ++h) {          // rows g and g + 8 of the tile
          float s = 0.0f;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int j = 0; j < 2; ++j)
              if (nt * 8 + 2 * tig + j < C) s += acc[nt][m][2 * h + j];
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          const float mean = s * (1.0f / C);
          float q = 0.0f;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int j = 0; j < 2; ++j)
              if (nt * 8 + 2 * tig + j < C) { const float d = acc[nt][m][2 * h + j] - mean; q = fmaf(d, d, q); }
          q += __shfl_xor_sync(0xffffffffu, q, 1);
          q += __shfl_xor_sync(0xffffffffu, q, 2);
          const float r = rsqrtf(q * (1.0f / C) + 1e-5f);
          const int p = P0 + 16 * m + g + 8 * h;
          if (p < T) {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const int c = nt * 8 + 2 * tig + j;
                if (c < C) so[c * TPO + p] = fmaf((acc[nt][m][2 * h + j] - mean) * r, gr[nt][j], ber[nt][j]);
              }
          }
        }
    }
    __syncthreads();
    // ---- copy-out ----
    if (a3 == nullptr) {
      float* out = hn + n * (long long)K;
      for (int i = tid; i < C * Q; i += TCN_THREADS) {
        const int c = i / Q, q = i - c * Q;
        *reinterpret_cast<float4*>(out + c * T + 4 * q) = *reinterpret_cast<const float4*>(so + c * TPO + 4 * q);
      }
    } else {
      // the row as the split operand [hi | lo | hi | 1 1 0..] (K = C*T) of the fp16 tensor-core GEMM that follows
      __half* row = a3 + n * (long long)(3 * K + 8);
      for (int i = tid; i < C * Q; i += TCN_THREADS) {
        const int c = i / Q, q = i - c * Q;
        const float4 v = *reinterpret_cast<const float4*>(so + c * TPO + 4 * q);
        const uint2 e01 = split_pair(v.x, v.y), e23 = split_pair(v.z, v.w);
        const uint2 hi = make_uint2(e01.x, e23.x), lo = make_uint2(e01.y, e23.y);
        const int col = c * T + 4 * q;
        *reinterpret_cast<uint2*>(row + col) = hi;
        *reinterpret_cast<uint2*>(row + K + col) = lo;
        *reinterpret_cast<uint2*>(row + 2 * K + col) = hi;
      }
      if (tid == 0) *reinterpret_cast<uint4*>(row + 3 * K) = make_uint4(0x3C003C00u, 0u, 0u, 0u);
    }
    // the next row's x load touches sx only; phase 1 (writes sh) and phase 2 (writes so) start behind its barriers
  }
}

template <int C, int KS, bool HALF1>
size_t tcn_mma_smem(int T) {
  constexpr int NT = (C + 7) / 8, CP = NT * 8;
  constexpr int XPAIRS = HALF1 ? 4 : 8 * KS, HPAIRS = (C <= 8) ? 4 : 8;
  const int pairs = (T + 16 * MT - 1) / (16 * MT), TPAD = pairs * 16 * MT, TP = TPAD + 4, TPO = TPAD + 4;
  return sizeof(uint2) * (size_t)(XPAIRS + HPAIRS) * TP + sizeof(float) * (size_t)CP * TPO +
         sizeof(uint4) * 32 * ((size_t)3 * KS * NT + 3 * NT + KS * NT);
}

template <int C, int KS, bool HALF1>
cudaError_t launch_tcn_mma(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                           const float* gamma, const float* beta, long long N, int CI, int T, float* hn, void* a3,
                           const float* wsc, float* sc_out, const float* x2, int CI1, int sms, cudaStream_t stream) {
  const size_t smem = tcn_mma_smem<C, KS, HALF1>(T);
  if (smem > 110 * 1024) return cudaErrorNotSupported;      // below two CTAs per SM the FFMA kernel's segmented walk is the better fit
  auto kern = stg_tcn_mma_kernel<C, KS, HALF1>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 1;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TCN_THREADS, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  long long grid = (long long)sms * per_sm;
  if (grid > N) grid = N;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, TCN_THREADS, smem, stream>>>(x, w1, b1, w2, b2, gamma, beta, N, CI, T, hn, (__half*)a3, wsc, sc_out,
                                                       x2, CI1);
  return cudaGetLastError();
}

template <int C>
cudaError_t dispatch_ci(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                        const float* gamma, const float* beta, long long N, int CI, int T, float* hn, void* a3,
                        const float* wsc, float* sc_out, const float* x2, int CI1, int sms, cudaStream_t stream) {
  if (CI <= 8) return launch_tcn_mma<C, 1, true>(x, w1, b1, w2, b2, gamma, beta, N, CI, T, hn, a3, wsc, sc_out, x2, CI1, sms, stream);
  if (CI <= 16) return launch_tcn_mma<C, 1, false>(x, w1, b1, w2, b2, gamma, beta, N, CI, T, hn, a3, wsc, sc_out, x2, CI1, sms, stream);
  return launch_tcn_mma<C, 2, false>(x, w1, b1, w2, b2, gamma, beta, N, CI, T, hn, a3, wsc, sc_out, x2, CI1, sms, stream);
}

}  // namespace

// Returns cudaErrorNotSupported when the shape is outside this kernel (the caller falls back to the FFMA kernel).
cudaError_t upd_launch_stg_tcn_mma(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                                   const float* gamma, const float* beta, long long N, int CI, int C, int T, float* hn,
                                   void* a3, const float* wsc, float* sc_out, const float* x2, int CI1, int sms,
                                   cudaStream_t stream) {
  if ((T & 3) != 0 || T < 4 || T > 512 || CI < 1 || CI > 32) return cudaErrorNotSupported;
  switch (C) {
    case 4: return dispatch_ci<4>(x, w1, b1, w2, b2, gamma, beta, N, CI, T, hn, a3, wsc, sc_out, x2, CI1, sms, stream);
    case 8: return dispatch_ci<8>(x, w1, b1, w2, b2, gamma, beta, N, CI, T, hn, a3, wsc, sc_out, x2, CI1, sms, stream);
    case 16: return dispatch_ci<16>(x, w1, b1, w2, b2, gamma, beta, N, CI, T, hn, a3, wsc, sc_out, x2, CI1, sms, stream);
    default: return cudaErrorNotSupported;
  }
}
