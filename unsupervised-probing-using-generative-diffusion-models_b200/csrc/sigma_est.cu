// g(x) = SigmaEstimation (g_backbone.py:49-72, sigma.py:34-71): the NsDiff conditional-variance
// estimate, which is also the whole "gx" uncertainty path.  Once per window-row, ~0.8 MFLOP per
// (row, feature): latency/L2-bound, not on the roofline-critical path, so a plain fp32 kernel:
// one CTA handles NB window-rows (all F features of each, because LayerNorm couples them),
// activations live in shared memory, each warp produces one hidden unit at a time with a coalesced
// read of that unit's weight row and a shuffle tree.
#include "upd_b200.h"
#include "upd_common.cuh"

namespace {

constexpr int NB = 4;          // window-rows per CTA (weight rows are reused NB*F times)
constexpr int THREADS = 256;

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = wsum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < THREADS / 32; ++i) t += red[i];
  return t;
}

// out[v][j] = act( sum_i W[j][i] * in[v][i] + b[j] ),  v < NV vectors, j < n_out
template <int NV, bool RELU>
__device__ void dense(const float* __restrict__ W, const float* __restrict__ b, const float* in, float* out,
                      int n_in, int n_out, int in_stride, int out_stride) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int j = warp; j < n_out; j += THREADS / 32) {
    float acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = 0.f;
    const float* wr = W + (long long)j * n_in;
    for (int i = lane; i < n_in; i += 32) {
      float w = __ldg(wr + i);
#pragma unroll
      for (int v = 0; v < NV; ++v) acc[v] = fmaf(w, in[v * in_stride + i], acc[v]);
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = wsum(acc[v]);
    if (lane == 0) {
      float bj = b[j];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        float r = acc[v] + bj;
        out[v * out_stride + j] = RELU ? fmaxf(r, 0.f) : r;
      }
    }
  }
}

// LayerNorm over the (F,H) block of each window-row (nn.LayerNorm([F,H]), eps 1e-5), in place.
template <int F>
__device__ void layer_norm_rows(float* h, const float* __restrict__ gamma, const float* __restrict__ beta, int H,
                                int nb, float* red) {
  for (int r = 0; r < nb; ++r) {
    float* hr = h + r * F * H;
    const int n = F * H;
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += THREADS) s += hr[i];
    float mu = block_sum(s, red) / (float)n;
    float q = 0.f;
    for (int i = threadIdx.x; i < n; i += THREADS) { float d = hr[i] - mu; q = fmaf(d, d, q); }
    float var = block_sum(q, red) / (float)n;
    float rstd = 1.0f / sqrtf(var + 1e-5f);
    for (int i = threadIdx.x; i < n; i += THREADS) hr[i] = (hr[i] - mu) * rstd * gamma[i] + beta[i];
    __syncthreads();
  }
}

template <int F>
__global__ void __launch_bounds__(THREADS)
sigma_estimation_kernel(UpdSigmaWeights w, const float* __restrict__ x, int rows, int Lw, int R, int H, int O,
                        float add_eps, float* __restrict__ gx) {
  extern __shared__ __align__(16) float sm[];
  const int n_in = Lw - R;
  float* vin = sm;                         // [NB*F][n_in]
  float* h1 = vin + NB * F * n_in;         // [NB*F][H]
  float* h2 = h1 + NB * F * H;             // [NB*F][H]
  __shared__ float red[THREADS / 32];
  const int row0 = blockIdx.x * NB;
  const int nb = min(NB, rows - row0);

  // trailing biased variance of x[i+1 .. i+R] for i in [0, n_in): the slice of wv_sigma_trailing
  // the MLP consumes (sigma.py:64-70, g_backbone.py:64); torch reduces float var in double.
  for (int idx = threadIdx.x; idx < NB * F * n_in; idx += THREADS) {
    int i = idx % n_in, vf = idx / n_in, r = vf / F, f = vf % F;
    float out = 0.f;
    if (r < nb) {
      const float* xs = x + ((long long)(row0 + r) * Lw + (i + 1)) * F + f;
      double s = 0.0;
      for (int k = 0; k < R; ++k) s += (double)xs[(long long)k * F];
      double mu = s / R, q = 0.0;
      for (int k = 0; k < R; ++k) { double d = (double)xs[(long long)k * F] - mu; q += d * d; }
      out = (float)(q / R) + 10e-8f;
    }
    vin[idx] = out;
  }
  __syncthreads();
  dense<NB * F, true>(w.w0, w.b0, vin, h1, n_in, H, n_in, H);
  __syncthreads();
  layer_norm_rows<F>(h1, w.ln1_w, w.ln1_b, H, nb, red);
  dense<NB * F, true>(w.w3, w.b3, h1, h2, H, H, H, H);
  __syncthreads();
  layer_norm_rows<F>(h2, w.ln2_w, w.ln2_b, H, nb, red);
  dense<NB * F, false>(w.w6, w.b6, h2, h1, H, O, H, H);   // h1 reused as [NB*F][H>=O] output
  __syncthreads();
  for (int idx = threadIdx.x; idx < nb * O * F; idx += THREADS) {
    int f = idx % F, o = (idx / F) % O, r = idx / (F * O);
    float v = h1[(r * F + f) * H + o];
    float sp = (v > 20.f) ? v : log1pf(expf(v));
    gx[((long long)(row0 + r) * O + o) * F + f] = sp + add_eps;
  }
}

}  // namespace

cudaError_t upd_launch_sigma(const UpdSigmaWeights& w, const float* x, int rows, int Lw, int R, int F, int H,
                             int O, float add_eps, float* gx, cudaStream_t stream) {
  if (O > H) return cudaErrorInvalidValue;
  size_t smem = sizeof(float) * ((size_t)NB * F * (Lw - R) + 2 * (size_t)NB * F * H);
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  int grid = (rows + NB - 1) / NB;
#define UPD_SIG(FF)                                                                                        \
  if (F == FF) {                                                                                           \
    cudaError_t e = cudaFuncSetAttribute(sigma_estimation_kernel<FF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return e;                                                                        \
    sigma_estimation_kernel<FF><<<grid, THREADS, smem, stream>>>(w, x, rows, Lw, R, H, O, add_eps, gx);   \
    return cudaGetLastError();                                                                             \
  }
  UPD_SIG(1) UPD_SIG(2) UPD_SIG(3) UPD_SIG(4)
#undef UPD_SIG
  return cudaErrorInvalidValue;
}
