// f(x) condition encoder (ns-Transformer, models/Diffusion_model/NsDiff/mu_backbone.py:53-183,
// TMDM/tmdm_ns_transformer.py:40-174): the memory-bound glue between its GEMMs, fused so that every activation is
// read once and leaves in the form its consumer needs.
//
// Every dense layer of the encoder runs as ONE fp16 tensor-core GEMM with fp32 accumulation on an error-compensated
// operand:  y = x W^T + b  ~=  [x_hi | x_lo | x_hi | 1 1 0..0] . [W_hi | W_hi | W_lo | b_hi b_lo 0..0]^T
// (hi = fp16(x), lo = fp16(x - hi): 22 mantissa bits, the x_lo*W_lo term ~2^-22 is dropped; measured 3e-6 of max|y|).
// The kernels here produce that "A3" operand [rows, 3K+8] fp16 directly from whatever precedes the GEMM:
//   fx_split_kernel ........ plain / activation (ReLU, exact GELU) / head-merge transpose of the attention output
//   fx_add_ln_split_kernel . residual add + LayerNorm (+ the stack's final LayerNorm) -> fp32 y and A3(y)
// One warp per row, float4 loads, 8-byte fp16 stores; HBM-bound: 4 B read + 6 B (+4 B) written per element.
#include <cuda_fp16.h>

#include "tc_helpers.cuh"
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int FX_MAX_K = 1024;            // widths with a LayerNorm or an embedding row held in registers
constexpr int FX_SPLIT_MAX_K = 1 << 20;   // the plain split streams its row

__device__ __forceinline__ float fx_act(float v, int act) {
  if (act == 1) return fmaxf(v, 0.0f);
  if (act == 2) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));      // F.gelu (exact, erf form)
  return v;
}

// 4 consecutive values -> hi/lo halves stored at columns c (hi), K + c (lo), 2K + c (hi) of one A3 row.
__device__ __forceinline__ void fx_store4(__half* row, int K, int c, float4 v) {
  uint2 hi, lo;
  tc::split_f16x2(v.x, v.y, hi.x, lo.x);
  tc::split_f16x2(v.z, v.w, hi.y, lo.y);
  *reinterpret_cast<uint2*>(row + c) = hi;
  *reinterpret_cast<uint2*>(row + K + c) = lo;
  *reinterpret_cast<uint2*>(row + 2 * K + c) = hi;
}

__device__ __forceinline__ void fx_store_tail(__half* row, int K) {     // the bias columns: 1, 1, 0 x 6
  uint4 t;
  t.x = 0x3C003C00u; t.y = 0u; t.z = 0u; t.w = 0u;                       // half(1.0) = 0x3C00
  *reinterpret_cast<uint4*>(row + 3 * K) = t;
}

// Source element (row r, column k): H == 1: x[r*K + k]; H > 1 (attention output [B,H,L,dk], K = H*dk):
// r = b*L + l, k = h*dk + j -> x[((b*H + h)*L + l)*dk + j]   (the head merge `out.transpose(1,2).reshape`).
__global__ void fx_split_kernel(const float* __restrict__ x, long long rows, int K, int H, int L, int act,
                                __half* __restrict__ a3) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int KP = 3 * K + 8;
  __half* out = a3 + r * KP;
  const int dk = K / H;
  const long long b = (H > 1) ? r / L : 0, l = (H > 1) ? r - b * L : 0;
  for (int c = lane * 4; c < K; c += 128) {
    const float* src = (H > 1) ? x + ((b * H + c / dk) * L + l) * dk + (c % dk) : x + r * K + c;
    float4 v = *reinterpret_cast<const float4*>(src);
    v.x = fx_act(v.x, act); v.y = fx_act(v.y, act); v.z = fx_act(v.z, act); v.w = fx_act(v.w, act);
    fx_store4(out, K, c, v);
  }
  if (lane == 0) fx_store_tail(out, K);
}

// y = LN2?(LN1(x + res)):  LayerNorm over K with eps 1e-5 (two-pass mean / biased variance like torch), affine g,b.
// Writes y (fp32, optional) and A3(y) (optional).
template <int NV>   // float4 chunks per lane: K = NV * 128
__global__ void fx_add_ln_split_kernel(const float* __restrict__ x, const float* __restrict__ res,
                                       const float* __restrict__ g1, const float* __restrict__ b1,
                                       const float* __restrict__ g2, const float* __restrict__ b2, long long rows,
                                       float* __restrict__ y, __half* __restrict__ a3) {
  constexpr int K = NV * 128;
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane * 4 + i * 128;
    v[i] = *reinterpret_cast<const float4*>(x + r * K + c);
    if (res) {
      float4 q = *reinterpret_cast<const float4*>(res + r * K + c);
      v[i].x += q.x; v[i].y += q.y; v[i].z += q.z; v[i].w += q.w;
    }
  }
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const float* g = pass == 0 ? g1 : g2;
    const float* bb = pass == 0 ? b1 : b2;
    if (!g) break;
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / K);
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
      q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.0f / K) + 1e-5f);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane * 4 + i * 128;
      const float4 gg = *reinterpret_cast<const float4*>(g + c), be = *reinterpret_cast<const float4*>(bb + c);
      v[i].x = (v[i].x - mean) * rstd * gg.x + be.x;
      v[i].y = (v[i].y - mean) * rstd * gg.y + be.y;
      v[i].z = (v[i].z - mean) * rstd * gg.z + be.z;
      v[i].w = (v[i].w - mean) * rstd * gg.w + be.w;
    }
  }
  if (y) {
#pragma unroll
    for (int i = 0; i < NV; ++i) *reinterpret_cast<float4*>(y + r * K + lane * 4 + i * 128) = v[i];
  }
  if (a3) {
    __half* out = a3 + r * (3 * K + 8);
#pragma unroll
    for (int i = 0; i < NV; ++i) fx_store4(out, K, lane * 4 + i * 128, v[i]);
    if (lane == 0) fx_store_tail(out, K);
  }
}

// Narrow rows (K = 32 * NS, e.g. TMDM's d_model = 64): the same residual + LayerNorm (+ final norm) -> fp32 and split
// operand, one warp per row with NS scalars per lane.
template <int NS>
__global__ void fx_add_ln_split_small_kernel(const float* __restrict__ x, const float* __restrict__ res,
                                             const float* __restrict__ g1, const float* __restrict__ b1,
                                             const float* __restrict__ g2, const float* __restrict__ b2, long long rows,
                                             float* __restrict__ y, __half* __restrict__ a3) {
  constexpr int K = NS * 32;
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  float v[NS];
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    v[i] = x[r * K + lane + 32 * i];
    if (res) v[i] += res[r * K + lane + 32 * i];
  }
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const float* g = pass == 0 ? g1 : g2;
    const float* bb = pass == 0 ? b1 : b2;
    if (!g) break;
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < NS; ++i) s += v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / K);
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < NS; ++i) { const float d = v[i] - mean; q += d * d; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.0f / K) + 1e-5f);
#pragma unroll
    for (int i = 0; i < NS; ++i) v[i] = (v[i] - mean) * rstd * g[lane + 32 * i] + bb[lane + 32 * i];
  }
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    const int c = lane + 32 * i;
    if (y) y[r * K + c] = v[i];
    if (a3) {
      __half* out = a3 + r * (3 * K + 8);
      const __half h = __float2half_rn(v[i]);
      const __half l = __float2half_rn(v[i] - __half2float(h));
      out[c] = h; out[K + c] = l; out[2 * K + c] = h;
    }
  }
  if (a3 && lane == 0) fx_store_tail(a3 + r * (3 * K + 8), K);
}

// DataEmbedding of the condition encoder: circular Conv1d(c_in -> d, k=3, no bias) over the sequence + sinusoidal
// positional table, written as fp32 AND as the split operand of the first projection GEMM in one pass
// (TokenEmbedding / PositionalEmbedding of torch-timeseries as used at mu_backbone.py:66-69; three einsum + roll + add
// launches over [B,L,d] before).  One warp per (b,l) row; x [B,L,NF], w [d,NF,3], pe [>=L, d].
__global__ void fx_embed_split_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ pe,
                                      long long rows, int L, int NF, int K, float* __restrict__ y, __half* __restrict__ a3) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const long long b = r / L;
  const int l = (int)(r - b * L);
  const float* xb = x + b * (long long)L * NF;
  const int lm = (l + L - 1) % L, lp = (l + 1) % L;
  __half* out = a3 + r * (3 * K + 8);
  for (int c = lane * 4; c < K; c += 128) {
    float4 v = *reinterpret_cast<const float4*>(pe + (long long)l * K + c);
    for (int f = 0; f < NF; ++f) {
      const float x0 = xb[lm * NF + f], x1 = xb[l * NF + f], x2 = xb[lp * NF + f];
      const float* wc = w + ((long long)c * NF + f) * 3;          // w[c][f][0..2], next channel NF*3 further
      const int st = NF * 3;
      v.x += wc[0] * x0 + wc[1] * x1 + wc[2] * x2;
      v.y += wc[st] * x0 + wc[st + 1] * x1 + wc[st + 2] * x2;
      v.z += wc[2 * st] * x0 + wc[2 * st + 1] * x1 + wc[2 * st + 2] * x2;
      v.w += wc[3 * st] * x0 + wc[3 * st + 1] * x1 + wc[3 * st + 2] * x2;
    }
    *reinterpret_cast<float4*>(y + r * K + c) = v;
    fx_store4(out, K, c, v);
  }
  if (lane == 0) fx_store_tail(out, K);
}

}  // namespace

cudaError_t upd_launch_fx_split(const float* x, long long rows, int K, int H, int L, int act, void* a3,
                                cudaStream_t stream) {
  if (K < 4 || K > FX_SPLIT_MAX_K || (K & 3) || H < 1 || (K % H) || ((K / H) & 3) || (H > 1 && (L < 1 || rows % L)))
    return cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(a3) & 15)) return cudaErrorInvalidValue;
  const int wpb = 8;
  fx_split_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, stream>>>(x, rows, K, H, L, act, (__half*)a3);
  return cudaGetLastError();
}

cudaError_t upd_launch_fx_add_ln_split(const float* x, const float* res, const float* g1, const float* b1,
                                       const float* g2, const float* b2, long long rows, int K, float* y, void* a3,
                                       cudaStream_t stream) {
  const int wpb = 8;
  const unsigned grid = (unsigned)((rows + wpb - 1) / wpb);
  if (K == 32 || K == 64 || K == 96) {
    if (K == 32) fx_add_ln_split_small_kernel<1><<<grid, wpb * 32, 0, stream>>>(x, res, g1, b1, g2, b2, rows, y, (__half*)a3);
    else if (K == 64) fx_add_ln_split_small_kernel<2><<<grid, wpb * 32, 0, stream>>>(x, res, g1, b1, g2, b2, rows, y, (__half*)a3);
    else fx_add_ln_split_small_kernel<3><<<grid, wpb * 32, 0, stream>>>(x, res, g1, b1, g2, b2, rows, y, (__half*)a3);
    return cudaGetLastError();
  }
  if (K < 128 || K > FX_MAX_K || (K % 128)) return cudaErrorInvalidValue;
#define UPD_LN_CASE(NV)                                                                                    \
  case NV:                                                                                                 \
    fx_add_ln_split_kernel<NV><<<grid, wpb * 32, 0, stream>>>(x, res, g1, b1, g2, b2, rows, y, (__half*)a3); \
    break;
  switch (K / 128) {
    UPD_LN_CASE(1) UPD_LN_CASE(2) UPD_LN_CASE(3) UPD_LN_CASE(4) UPD_LN_CASE(5) UPD_LN_CASE(6) UPD_LN_CASE(7) UPD_LN_CASE(8)
    default: return cudaErrorInvalidValue;
  }
#undef UPD_LN_CASE
  return cudaGetLastError();
}

cudaError_t upd_launch_fx_embed_split(const float* x, const float* w, const float* pe, long long rows, int L, int NF, int K,
                                      float* y, void* a3, cudaStream_t stream) {
  if (K < 4 || K > FX_MAX_K || (K & 3) || L < 1 || NF < 1 || rows % L) return cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(pe) & 15) || (reinterpret_cast<uintptr_t>(y) & 15) || (reinterpret_cast<uintptr_t>(a3) & 15))
    return cudaErrorInvalidValue;
  const int wpb = 8;
  fx_embed_split_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, stream>>>(x, w, pe, rows, L, NF, K, y, (__half*)a3);
  return cudaGetLastError();
}
