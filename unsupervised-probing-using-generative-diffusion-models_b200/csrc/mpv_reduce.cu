// Dispersion reduction: K trajectories -> per-position predictive variance/mean -> per-window MPV.
//
// Stage 1 (HBM-bound, the kernel the roofline is quoted on): one thread per group of VEC
// consecutive (position,feature) elements of one row; it streams the K samples of that group
// (stride O*F floats) with 8 independent 16-byte loads in flight and folds them with Welford's
// update.  Algorithmic traffic = 4*K bytes read per element + 8 bytes written.
// Stage 2: one CTA per window sums var / mean over (B,O,F) in a fixed order (double accumulators,
// warp-shuffle tree) -> mpv, pred_mean, per-feature mpv.  No atomics: results are bit-reproducible.
#include "upd_common.cuh"

namespace {

struct Welford4 { float mean[4]; float m2[4]; };

template <int VEC>
__global__ void __launch_bounds__(256)
welford_over_samples(const float* __restrict__ traj, const float* __restrict__ scale,
                     long long n_slots, int n_q, int K, int E, int F,
                     float* __restrict__ var_out, float* __restrict__ mean_out) {
  long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n_slots) return;
  long long r0 = slot / n_q;
  int q = (int)(slot - r0 * n_q);
  int e0 = q * VEC;
  const float* base = traj + (r0 * K) * (long long)E + e0;
  float sc_m[VEC], sc_s[VEC];
#pragma unroll
  for (int c = 0; c < VEC; ++c) {
    int f = (e0 + c) % F;
    sc_m[c] = scale ? scale[f] : 0.f;
    sc_s[c] = scale ? scale[F + f] : 1.f;
  }
  float mean[VEC], m2[VEC];
#pragma unroll
  for (int c = 0; c < VEC; ++c) { mean[c] = 0.f; m2[c] = 0.f; }
  constexpr int U = 8;
  int k = 0;
  for (; k + U <= K; k += U) {
    float x[U][VEC];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float* p = base + (long long)(k + u) * E;
      if (VEC == 4) {
        float4 v = __ldcs(reinterpret_cast<const float4*>(p));
        x[u][0] = v.x; x[u][1 % VEC] = v.y; x[u][2 % VEC] = v.z; x[u][3 % VEC] = v.w;
      } else {
        x[u][0] = __ldcs(p);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float rn = 1.0f / (float)(k + u + 1);
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        float xv = scale ? x[u][c] * sc_s[c] + sc_m[c] : x[u][c];
        float d = xv - mean[c];
        mean[c] += d * rn;
        m2[c] += d * (xv - mean[c]);
      }
    }
  }
  for (; k < K; ++k) {
    const float* p = base + (long long)k * E;
    float rn = 1.0f / (float)(k + 1);
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
      float xv = __ldcs(p + c);
      if (scale) xv = xv * sc_s[c] + sc_m[c];
      float d = xv - mean[c];
      mean[c] += d * rn;
      m2[c] += d * (xv - mean[c]);
    }
  }
  float invk = 1.0f / (float)K;
  long long o = r0 * E + e0;
  if (VEC == 4) {
    *reinterpret_cast<float4*>(var_out + o) = make_float4(m2[0] * invk, m2[1 % VEC] * invk, m2[2 % VEC] * invk, m2[3 % VEC] * invk);
    *reinterpret_cast<float4*>(mean_out + o) = make_float4(mean[0], mean[1 % VEC], mean[2 % VEC], mean[3 % VEC]);
  } else {
    var_out[o] = m2[0] * invk;
    mean_out[o] = mean[0];
  }
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One CTA (128 threads) per window: sums over the window's B*E elements, per feature.
__global__ void __launch_bounds__(128)
window_means(const float* __restrict__ var_in, const float* __restrict__ mean_in, int B, int E, int F,
             float* __restrict__ mpv, float* __restrict__ pmean, float* __restrict__ mpv_f) {
  const int w = blockIdx.x;
  const long long n = (long long)B * E;
  const float* v = var_in + (long long)w * n;
  const float* m = mean_in + (long long)w * n;
  double sv[UPD_MAX_F] = {0, 0, 0, 0};
  double sm = 0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    int f = (int)((i % E) % F);
    float vv = v[i];
#pragma unroll
    for (int c = 0; c < UPD_MAX_F; ++c) sv[c] += (c == f) ? (double)vv : 0.0;
    sm += (double)m[i];
  }
  __shared__ double red[4][UPD_MAX_F + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < UPD_MAX_F; ++c) sv[c] = warp_sum_d(sv[c]);
  sm = warp_sum_d(sm);
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < UPD_MAX_F; ++c) red[warp][c] = sv[c];
    red[warp][UPD_MAX_F] = sm;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0;
    double per_f = (double)n / F;
    for (int c = 0; c < F; ++c) {
      double s = red[0][c] + red[1][c] + red[2][c] + red[3][c];
      tot += s;
      if (mpv_f) mpv_f[(long long)w * F + c] = (float)(s / per_f);
    }
    if (mpv) mpv[w] = (float)(tot / (double)n);
    if (pmean) pmean[w] = (float)((red[0][UPD_MAX_F] + red[1][UPD_MAX_F] + red[2][UPD_MAX_F] + red[3][UPD_MAX_F]) / (double)n);
  }
}


// ---- SLBP extras ------------------------------------------------------------------------------------
// Centred Gram matrix of a window's K trajectories, G = C C^T / (K - 1) with C = X - mean_K(X), X [K, D = O*F]: the K x K
// matrix with the same non-zero spectrum as the (O*F)^2 sample covariance `_slbp_intrinsic_dimension` diagonalises
// (diffusion_model_uncertainy.py:686-698).  One CTA per window; X is walked in chunks of DC columns staged (centred, in
// double) in shared memory; every thread owns a fixed set of (i <= j) pairs and accumulates them in double.  The per-chunk
// column means make the result independent of how D is cut.  Algorithmic traffic: 4 K D bytes read, 8 K^2 written.
constexpr int GRAM_DC = 32, GRAM_MAXK = 128, GRAM_THREADS = 256, GRAM_PAIRS_PER_THREAD = (GRAM_MAXK * (GRAM_MAXK + 1) / 2 + GRAM_THREADS - 1) / GRAM_THREADS;

__global__ void __launch_bounds__(GRAM_THREADS)
gram_centered_kernel(const float* __restrict__ traj, int K, int D, double* __restrict__ gram) {
  __shared__ double c[GRAM_MAXK][GRAM_DC + 1];
  const float* x = traj + (long long)blockIdx.x * K * D;
  double* g = gram + (long long)blockIdx.x * K * K;
  const int n_pairs = K * (K + 1) / 2;
  double acc[GRAM_PAIRS_PER_THREAD];
#pragma unroll
  for (int p = 0; p < GRAM_PAIRS_PER_THREAD; ++p) acc[p] = 0.0;
  for (int d0 = 0; d0 < D; d0 += GRAM_DC) {
    const int dc = min(GRAM_DC, D - d0);
    for (int e = threadIdx.x; e < K * GRAM_DC; e += GRAM_THREADS) {
      const int k = e / GRAM_DC, d = e - k * GRAM_DC;
      c[k][d] = d < dc ? (double)x[(long long)k * D + d0 + d] : 0.0;
    }
    __syncthreads();
    if (threadIdx.x < GRAM_DC) {                         // centre the chunk: mean over the K samples, per column
      double m = 0.0;
      for (int k = 0; k < K; ++k) m += c[k][threadIdx.x];
      m /= K;
      for (int k = 0; k < K; ++k) c[k][threadIdx.x] -= m;
    }
    __syncthreads();
    int p = 0;
    for (int q = threadIdx.x; q < n_pairs; q += GRAM_THREADS, ++p) {
      // q -> (i, j), i <= j, row-major over the upper triangle
      int i = (int)((2.0 * K + 1.0 - sqrt((2.0 * K + 1.0) * (2.0 * K + 1.0) - 8.0 * q)) * 0.5);
      while (i * (2 * K - i + 1) / 2 > q) --i;
      while ((i + 1) * (2 * K - i) / 2 <= q) ++i;
      const int j = i + (q - i * (2 * K - i + 1) / 2);
      double s = 0.0;
#pragma unroll 8
      for (int d = 0; d < GRAM_DC; ++d) s += c[i][d] * c[j][d];
      acc[p] += s;
    }
    __syncthreads();
  }
  const double inv = 1.0 / (double)max(K - 1, 1);
  int p = 0;
  for (int q = threadIdx.x; q < n_pairs; q += GRAM_THREADS, ++p) {
    int i = (int)((2.0 * K + 1.0 - sqrt((2.0 * K + 1.0) * (2.0 * K + 1.0) - 8.0 * q)) * 0.5);
    while (i * (2 * K - i + 1) / 2 > q) --i;
    while ((i + 1) * (2 * K - i) / 2 <= q) ++i;
    const int j = i + (q - i * (2 * K - i + 1) / 2);
    g[(long long)i * K + j] = acc[p] * inv;
    g[(long long)j * K + i] = acc[p] * inv;
  }
}

// Prediction error of a window (diffusion_model_uncertainy.py:542-549): | mean_K - target | averaged over the pred_len
// positions, per feature.  mean [W, O, F] are the Welford means of upd_mpv_reduce, target [W, O, F] the (scaled) futures.
__global__ void __launch_bounds__(128)
prediction_error_kernel(const float* __restrict__ mean, const float* __restrict__ target, int O, int F,
                        float* __restrict__ err) {
  __shared__ double part[4][UPD_MAX_F];
  const long long base = (long long)blockIdx.x * O * F;
  double a[UPD_MAX_F];
#pragma unroll
  for (int f = 0; f < UPD_MAX_F; ++f) a[f] = 0.0;
  for (int o = threadIdx.x; o < O; o += blockDim.x)
    for (int f = 0; f < F; ++f) a[f] += fabs((double)mean[base + (long long)o * F + f] - (double)target[base + (long long)o * F + f]);
#pragma unroll
  for (int f = 0; f < UPD_MAX_F; ++f)
    for (int off = 16; off; off >>= 1) a[f] += __shfl_xor_sync(0xffffffffu, a[f], off);
  if ((threadIdx.x & 31) == 0)
    for (int f = 0; f < UPD_MAX_F; ++f) part[threadIdx.x >> 5][f] = a[f];
  __syncthreads();
  if (threadIdx.x < F) err[(long long)blockIdx.x * F + threadIdx.x] =
      (float)((part[0][threadIdx.x] + part[1][threadIdx.x] + part[2][threadIdx.x] + part[3][threadIdx.x]) / O);
}

}  // namespace

cudaError_t upd_launch_mpv(const float* traj, const float* scale, int n_win, int B, int K, int O, int F,
                           float* var_out, float* mean_out, float* mpv, float* pmean, float* mpv_f,
                           cudaStream_t stream) {
  const int E = O * F;
  const long long R0 = (long long)n_win * B;
  const bool vec4 = (E % 4 == 0) && ((reinterpret_cast<uintptr_t>(traj) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(var_out) & 15) == 0) && ((reinterpret_cast<uintptr_t>(mean_out) & 15) == 0);
  if (vec4) {
    int n_q = E / 4;
    long long n_slots = R0 * n_q;
    unsigned grid = (unsigned)((n_slots + 255) / 256);
    welford_over_samples<4><<<grid, 256, 0, stream>>>(traj, scale, n_slots, n_q, K, E, F, var_out, mean_out);
  } else {
    long long n_slots = R0 * E;
    unsigned grid = (unsigned)((n_slots + 255) / 256);
    welford_over_samples<1><<<grid, 256, 0, stream>>>(traj, scale, n_slots, E, K, E, F, var_out, mean_out);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  window_means<<<n_win, 128, 0, stream>>>(var_out, mean_out, B, E, F, mpv, pmean, mpv_f);
  return cudaGetLastError();
}

cudaError_t upd_launch_gram_centered(const float* traj, int W, int K, int D, double* gram, cudaStream_t stream) {
  if (K > GRAM_MAXK) return cudaErrorInvalidValue;
  gram_centered_kernel<<<W, GRAM_THREADS, 0, stream>>>(traj, K, D, gram);
  return cudaGetLastError();
}

cudaError_t upd_launch_prediction_error(const float* mean, const float* target, int W, int O, int F, float* err,
                                        cudaStream_t stream) {
  prediction_error_kernel<<<W, 128, 0, stream>>>(mean, target, O, F, err);
  return cudaGetLastError();
}
