// DiffusionTS attention (FullAttention / CrossAttention, models/Diffusion_model/DiffusionTS/diffusionts_transformer.py:
// 126-203): softmax(q k^T / sqrt(hs)) v with head size 16 and sequences of ~200 positions, forward and backward
// (the backward feeds the Langevin refinement gradient, DiffusionTS.py:384-399).
//
// Head size 16 is far below a tensor-core tile's K and the whole K/V of one (row, head) is 25 KB, so this is an fp32 FFMA
// kernel: one CTA per (row, head), K and V in shared memory, one thread per query position holding its q row, the
// running (max, sum) and the output row in registers; every K/V row is a shared-memory broadcast.  Nothing of size
// [seq, seq] ever exists (the library path materialised scores, probabilities and their gradients: 5 tensors of
// R*heads*seq*seq*4 B per attention).  The forward keeps the per-row log-sum-exp; the backward recomputes the
// probabilities twice -- once per query (dQ) and once per key (dK, dV) -- instead of reducing across threads.
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int HS = 16;                      // head size the kernels are built for
constexpr float LOG2E = 1.4426950408889634f;

struct DtsAttnParams {
  const float* q; long long q_stride;       // row (r*Lq + i) at q + row*q_stride, head h at + h*16
  const float* k; const float* v; long long kv_stride;   // row (r*S + j)
  int R, H, Lq, S;
  float scale;
  float* o;                                 // [R*Lq, H*16]
  float* lse;                               // [R*H, Lq] base-2 log-sum-exp of the scaled scores
  // backward only
  const float* d_o;                         // [R*Lq, H*16]
  float* dq; long long dq_stride;           // same addressing as q
  float* dk; float* dv; long long dkv_stride;
  // forward only, de-stationary attention of the TMDM / NsDiff condition encoder at head size 16:
  const float* tau;                         // [R] or null: scores *= tau[r]
  const float* delta; int delta_pitch;      // [R, pitch] ALREADY multiplied by scale, or null: added to the scaled scores
  int causal;                               // keys j > i masked
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void load16(const float* p, float (&x)[HS]) {
#pragma unroll
  for (int c = 0; c < HS; c += 4) {
    const float4 t = *reinterpret_cast<const float4*>(p + c);
    x[c] = t.x; x[c + 1] = t.y; x[c + 2] = t.z; x[c + 3] = t.w;
  }
}

__device__ __forceinline__ float dot16(const float (&a)[HS], const float* __restrict__ b) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int c = 0; c < HS; c += 4) {
    const float4 t = *reinterpret_cast<const float4*>(b + c);
    s0 = fmaf(a[c], t.x, s0); s1 = fmaf(a[c + 1], t.y, s1); s2 = fmaf(a[c + 2], t.z, s2); s3 = fmaf(a[c + 3], t.w, s3);
  }
  return (s0 + s1) + (s2 + s3);
}

__device__ __forceinline__ void axpy16(float a, const float* __restrict__ x, float (&y)[HS]) {
#pragma unroll
  for (int c = 0; c < HS; c += 4) {
    const float4 t = *reinterpret_cast<const float4*>(x + c);
    y[c] = fmaf(a, t.x, y[c]); y[c + 1] = fmaf(a, t.y, y[c + 1]); y[c + 2] = fmaf(a, t.z, y[c + 2]); y[c + 3] = fmaf(a, t.w, y[c + 3]);
  }
}

// stage n rows of 16 floats (row j at src + j*stride) into smem [n][16]
__device__ __forceinline__ void stage_rows(const float* __restrict__ src, long long stride, int n, float* __restrict__ dst) {
  for (int i = threadIdx.x; i < n * (HS / 4); i += blockDim.x) {
    const int j = i >> 2, c = (i & 3) * 4;
    *reinterpret_cast<float4*>(dst + j * HS + c) = *reinterpret_cast<const float4*>(src + (long long)j * stride + c);
  }
}

__global__ void dts_attn_fwd_kernel(const DtsAttnParams p) {
  extern __shared__ __align__(16) float sm[];
  float* sk = sm;                            // [S][16]
  float* sv = sk + p.S * HS;                 // [S][16]
  const int rh = blockIdx.x, r = rh / p.H, h = rh - r * p.H;
  stage_rows(p.k + (long long)r * p.S * p.kv_stride + h * HS, p.kv_stride, p.S, sk);
  stage_rows(p.v + (long long)r * p.S * p.kv_stride + h * HS, p.kv_stride, p.S, sv);
  __syncthreads();
  const float qs = p.scale * LOG2E * (p.tau ? p.tau[r] : 1.0f);
  const float* dl = p.delta ? p.delta + (long long)r * p.delta_pitch : nullptr;
  for (int i = threadIdx.x; i < p.Lq; i += blockDim.x) {
    float q[HS], o[HS];
    load16(p.q + ((long long)r * p.Lq + i) * p.q_stride + h * HS, q);
#pragma unroll
    for (int c = 0; c < HS; ++c) { q[c] *= qs; o[c] = 0.f; }
    float m = -INFINITY, l = 0.f;
    const int s_end = p.causal ? min(p.S, i + 1) : p.S;
    for (int j = 0; j < s_end; ++j) {
      float s = dot16(q, sk + j * HS);
      if (dl) s = fmaf(dl[j], LOG2E, s);
      if (s > m) {                           // rare after the first few keys
        const float f = ex2f(m - s);
        l *= f;
#pragma unroll
        for (int c = 0; c < HS; ++c) o[c] *= f;
        m = s;
      }
      const float e = ex2f(s - m);
      l += e;
      axpy16(e, sv + j * HS, o);
    }
    const float inv = 1.0f / l;
    float* op = p.o + ((long long)r * p.Lq + i) * (p.H * HS) + h * HS;
#pragma unroll
    for (int c = 0; c < HS; c += 4)
      *reinterpret_cast<float4*>(op + c) = make_float4(o[c] * inv, o[c + 1] * inv, o[c + 2] * inv, o[c + 3] * inv);
    if (p.lse) p.lse[(long long)rh * p.Lq + i] = m + log2f(l);
  }
}

__global__ void dts_attn_bwd_kernel(const DtsAttnParams p) {
  extern __shared__ __align__(16) float sm[];
  float* sk = sm;                            // [S][16]
  float* sv = sk + p.S * HS;                 // [S][16]
  float* sq = sv + p.S * HS;                 // [Lq][16]  q * scale * log2e
  float* sg = sq + p.Lq * HS;                // [Lq][16]  dO
  float* sl = sg + p.Lq * HS;                // [Lq]      lse (base 2)
  float* sd = sl + p.Lq;                     // [Lq]      D_i = dO_i . O_i
  const int rh = blockIdx.x, r = rh / p.H, h = rh - r * p.H;
  const int d = p.H * HS;
  stage_rows(p.k + (long long)r * p.S * p.kv_stride + h * HS, p.kv_stride, p.S, sk);
  stage_rows(p.v + (long long)r * p.S * p.kv_stride + h * HS, p.kv_stride, p.S, sv);
  stage_rows(p.q + (long long)r * p.Lq * p.q_stride + h * HS, p.q_stride, p.Lq, sq);
  stage_rows(p.d_o + (long long)r * p.Lq * d + h * HS, d, p.Lq, sg);
  const float qs = p.scale * LOG2E;
  for (int i = threadIdx.x; i < p.Lq; i += blockDim.x) {
    float ov[HS], gv[HS];
    load16(p.o + ((long long)r * p.Lq + i) * d + h * HS, ov);
    load16(p.d_o + ((long long)r * p.Lq + i) * d + h * HS, gv);
    float dsum = 0.f;
#pragma unroll
    for (int c = 0; c < HS; ++c) dsum = fmaf(ov[c], gv[c], dsum);
    sd[i] = dsum;
    sl[i] = p.lse[(long long)rh * p.Lq + i];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < p.Lq * HS; i += blockDim.x) sq[i] *= qs;
  __syncthreads();
  // ---- per query: dQ_i = scale * sum_j p_ij (dO_i.V_j - D_i) K_j ----
  for (int i = threadIdx.x; i < p.Lq; i += blockDim.x) {
    float q[HS], g[HS], acc[HS];
#pragma unroll
    for (int c = 0; c < HS; ++c) { q[c] = sq[i * HS + c]; g[c] = sg[i * HS + c]; acc[c] = 0.f; }
    const float lse = sl[i], di = sd[i];
    for (int j = 0; j < p.S; ++j) {
      const float pij = ex2f(dot16(q, sk + j * HS) - lse);
      const float ds = pij * (dot16(g, sv + j * HS) - di);
      axpy16(ds, sk + j * HS, acc);
    }
    float* dst = p.dq + ((long long)r * p.Lq + i) * p.dq_stride + h * HS;
#pragma unroll
    for (int c = 0; c < HS; c += 4)
      *reinterpret_cast<float4*>(dst + c) = make_float4(acc[c] * p.scale, acc[c + 1] * p.scale, acc[c + 2] * p.scale, acc[c + 3] * p.scale);
  }
  // ---- per key: dV_j = sum_i p_ij dO_i ;  dK_j = scale * sum_i p_ij (dO_i.V_j - D_i) Q_i  (sq holds q*scale*log2e) ----
  const float unscale = 1.0f / LOG2E;
  for (int j = threadIdx.x; j < p.S; j += blockDim.x) {
    float kk[HS], vv[HS], dk[HS], dv[HS];
#pragma unroll
    for (int c = 0; c < HS; ++c) { kk[c] = sk[j * HS + c]; vv[c] = sv[j * HS + c]; dk[c] = 0.f; dv[c] = 0.f; }
    for (int i = 0; i < p.Lq; ++i) {
      const float pij = ex2f(dot16(kk, sq + i * HS) - sl[i]);
      const float ds = pij * (dot16(vv, sg + i * HS) - sd[i]);
      axpy16(pij, sg + i * HS, dv);
      axpy16(ds, sq + i * HS, dk);
    }
    float* dkp = p.dk + ((long long)r * p.S + j) * p.dkv_stride + h * HS;
    float* dvp = p.dv + ((long long)r * p.S + j) * p.dkv_stride + h * HS;
#pragma unroll
    for (int c = 0; c < HS; c += 4) {
      *reinterpret_cast<float4*>(dkp + c) = make_float4(dk[c] * unscale, dk[c + 1] * unscale, dk[c + 2] * unscale, dk[c + 3] * unscale);
      *reinterpret_cast<float4*>(dvp + c) = make_float4(dv[c], dv[c + 1], dv[c + 2], dv[c + 3]);
    }
  }
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

cudaError_t upd_launch_dts_attention(const float* q, long long q_stride, const float* k, const float* v, long long kv_stride,
                                     int R, int H, int Lq, int S, float scale, float* o, float* lse, const float* tau,
                                     const float* delta, int delta_pitch, int causal, cudaStream_t stream) {
  if ((q_stride & 3) || (kv_stride & 3) || !aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(o))
    return cudaErrorInvalidValue;
  const size_t smem = sizeof(float) * 2 * (size_t)S * HS;
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  DtsAttnParams p = {};
  p.q = q; p.q_stride = q_stride; p.k = k; p.v = v; p.kv_stride = kv_stride; p.R = R; p.H = H; p.Lq = Lq; p.S = S;
  p.scale = scale; p.o = o; p.lse = lse; p.tau = tau; p.delta = delta; p.delta_pitch = delta_pitch; p.causal = causal;
  cudaError_t e = cudaFuncSetAttribute(dts_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int threads = Lq >= 256 ? 256 : ((Lq + 31) / 32) * 32;
  dts_attn_fwd_kernel<<<(unsigned)(R * H), threads, smem, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t upd_launch_dts_attention_bwd(const float* q, long long q_stride, const float* k, const float* v,
                                         long long kv_stride, int R, int H, int Lq, int S, float scale, const float* o,
                                         const float* lse, const float* d_o, float* dq, long long dq_stride, float* dk,
                                         float* dv, long long dkv_stride, cudaStream_t stream) {
  if ((q_stride & 3) || (kv_stride & 3) || (dq_stride & 3) || (dkv_stride & 3) || !aligned16(q) || !aligned16(k) ||
      !aligned16(v) || !aligned16(o) || !aligned16(d_o) || !aligned16(dq) || !aligned16(dk) || !aligned16(dv))
    return cudaErrorInvalidValue;
  const size_t smem = sizeof(float) * (2 * (size_t)S * HS + 2 * (size_t)Lq * HS + 2 * (size_t)Lq);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  DtsAttnParams p = {};
  p.q = q; p.q_stride = q_stride; p.k = k; p.v = v; p.kv_stride = kv_stride; p.R = R; p.H = H; p.Lq = Lq; p.S = S;
  p.scale = scale; p.o = const_cast<float*>(o); p.lse = const_cast<float*>(lse); p.d_o = d_o; p.dq = dq;
  p.dq_stride = dq_stride; p.dk = dk; p.dv = dv; p.dkv_stride = dkv_stride;
  cudaError_t e = cudaFuncSetAttribute(dts_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int n = Lq > S ? Lq : S;
  const int threads = n >= 256 ? 256 : ((n + 31) / 32) * 32;
  dts_attn_bwd_kernel<<<(unsigned)(R * H), threads, smem, stream>>>(p);
  return cudaGetLastError();
}
