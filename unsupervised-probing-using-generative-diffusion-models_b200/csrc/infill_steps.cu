// DiffusionTS conditional sampler: the per-step algebra around the x0-predicting transformer, and the
// Fourier seasonal head (rfft bins -> top-k -> resynthesis) forward/backward as single kernels.
//   model_predictions / DDIM mean ....... models/Diffusion_model/DiffusionTS/DiffusionTS.py:152-160, 294-303
//   langevin_fn (fresh Adagrad step) ..... DiffusionTS.py:359-407
//   q_sample + infill overwrite .......... DiffusionTS.py:232-237, 305-306
//   FourierLayer ......................... models/Diffusion_model/DiffusionTS/diffusionts_transformer.py:52-103
// All HBM-bound streaming kernels: one pass over their operands, coalesced along the innermost axis.
#include <cuda_runtime.h>
#include <stdint.h>

#include "upd_common.cuh"

namespace {

constexpr int TOPK_MAX = 8;

// x_start = clamp(x0_raw, -1, 1); pred_noise = (sqrt_recip*img - x_start) / sqrt_recipm1;
// pred_mean = x_start*sqrt(alpha_next) + c*pred_noise; img_out = pred_mean + sigma*noise.
// last == 1 (time_next < 0): img_out = x_start.  Unfused fp32 rounding like the reference's tensor expressions.
__global__ void dts_ddim_step_kernel(const float* __restrict__ x0_raw, const float* __restrict__ img, long long n,
                                     float sqrt_recip, float sqrt_recipm1, float sqrt_an, float c, float sigma,
                                     const float* __restrict__ noise, int last, float* __restrict__ x_start,
                                     float* __restrict__ pred_mean, float* __restrict__ img_out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float xs = fminf(fmaxf(x0_raw[i], -1.0f), 1.0f);
    if (x_start) x_start[i] = xs;
    if (last) { img_out[i] = xs; continue; }
    float pn = __fdiv_rn(__fsub_rn(__fmul_rn(sqrt_recip, img[i]), xs), sqrt_recipm1);
    float pm = __fadd_rn(__fmul_rn(xs, sqrt_an), __fmul_rn(c, pn));
    if (pred_mean) pred_mean[i] = pm;
    float z = noise ? noise[i] : 0.0f;
    img_out[i] = __fadd_rn(pm, __fmul_rn(sigma, z));
  }
}

// torch.optim.Adagrad created anew for every iteration: state_sum = g*g, p -= lr * g / (sqrt(state_sum) + 1e-10).
__global__ void dts_adagrad_kernel(float* __restrict__ p, const float* __restrict__ g, long long n, float lr) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i];
    float std = __fadd_rn(__fsqrt_rn(__fmul_rn(gi, gi)), 1e-10f);
    p[i] = __fadd_rn(p[i], __fmul_rn(-lr, __fdiv_rn(gi, std)));
  }
}

// img[r, s, :] = s < L_obs ? sqrt_ac*target[r,s,:] + sqrt_1mac*noise[r,s,:] : refined[r,s,:]
// target is the observed window only, [rows, L_obs, F]; noise covers the whole [rows, seq, F] draw of q_sample.
__global__ void dts_infill_kernel(float* __restrict__ img, const float* __restrict__ refined,
                                  const float* __restrict__ target, const float* __restrict__ noise, long long rows,
                                  int seq, int L_obs, int F, float sqrt_ac, float sqrt_1mac) {
  const long long row_elems = (long long)seq * F, obs_elems = (long long)L_obs * F, n = rows * row_elems;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long r = i / row_elems, e = i - r * row_elems;
    float v;
    if (e < obs_elems) {
      float tg = target[r * obs_elems + e];
      v = noise ? __fadd_rn(__fmul_rn(sqrt_ac, tg), __fmul_rn(sqrt_1mac, noise[i])) : tg;
    } else {
      v = refined[i];
    }
    img[i] = v;
  }
}

// N(0,1) keyed by (seed, global row, element, draw): independent of how rows are batched or sharded.
__global__ void gauss_fill_kernel(float* __restrict__ out, long long rows, long long row_elems, uint64_t seed,
                                  uint64_t row_base, uint32_t draw) {
  const long long n = rows * row_elems;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long r = i / row_elems;
    uint32_t e = (uint32_t)(i - r * row_elems);
    out[i] = upd_gauss(seed, row_base + (uint64_t)r, draw >> 8, 0x5eedu, e, draw & 0xffu);
  }
}

// ------------------------------------------------------------------------------------------------------
// Fourier seasonal head.  spec[r, j, e] (j < NF) = Re X_{j+low}, spec[r, NF + j, e] = Im X_{j+low} of the rfft
// along the sequence axis (produced upstream by one GEMM with the DFT folded into the 1x1 projection).
// One thread per (row, channel e): scan the NF bins for the top-k amplitudes (descending, first index wins
// ties like a stable sort), then resynthesise all `seq` positions:
//     season[r, t, e] (+)= sum over [k bins, k conjugates] of |X| * cos(2*pi*f*t + arg X)
// with the argument formed in fp32 exactly as the reference forms it (2*pi*f rounded, times t, plus phase).
// ------------------------------------------------------------------------------------------------------
__global__ void dts_fourier_topk_fwd_kernel(const float* __restrict__ spec, long long spec_row_stride, long long rows,
                                            int NF, int low, int seq, int D, int top_k, int accumulate,
                                            float* __restrict__ season, int* __restrict__ idx_out) {
  const long long n = rows * D;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long r = i / D;
  const int e = (int)(i - r * D);
  const float* re_p = spec + r * spec_row_stride + e;
  const float* im_p = re_p + (long long)NF * D;
  float best[TOPK_MAX];
  int bidx[TOPK_MAX];
#pragma unroll
  for (int k = 0; k < TOPK_MAX; ++k) { best[k] = -1.0f; bidx[k] = 0; }
  for (int j = 0; j < NF; ++j) {
    float re = re_p[(long long)j * D], im = im_p[(long long)j * D];
    float a = hypotf(re, im);
    // insertion into the descending list (strict > keeps the earlier index on ties)
    float ca = a; int cj = j;
#pragma unroll
    for (int k = 0; k < TOPK_MAX; ++k) {
      if (k < top_k && ca > best[k]) {
        float ta = best[k]; int tj = bidx[k];
        best[k] = ca; bidx[k] = cj; ca = ta; cj = tj;
      }
    }
  }
  float amp[TOPK_MAX], ph[TOPK_MAX], w[TOPK_MAX];
#pragma unroll
  for (int k = 0; k < TOPK_MAX; ++k) {
    if (k < top_k) {
      int j = bidx[k];
      float re = re_p[(long long)j * D], im = im_p[(long long)j * D];
      amp[k] = hypotf(re, im);
      ph[k] = atan2f(im, re);
      float f = __fdiv_rn((float)(j + low), (float)seq);           // torch.fft.rfftfreq(t)[bin]
      w[k] = __fmul_rn(6.283185307179586f, f);                     // 2*pi*f in fp32
      if (idx_out) idx_out[(r * top_k + k) * D + e] = j;
    }
  }
  float* out = season + r * (long long)seq * D + e;
  for (int t = 0; t < seq; ++t) {
    float s = 0.0f, tf = (float)t;
#pragma unroll
    for (int k = 0; k < TOPK_MAX; ++k)
      if (k < top_k) s = __fadd_rn(s, __fmul_rn(amp[k], cosf(__fadd_rn(__fmul_rn(w[k], tf), ph[k]))));
#pragma unroll
    for (int k = 0; k < TOPK_MAX; ++k)   // the conjugate half: cos(-(w t + phase)) = the same value, added again in order
      if (k < top_k) s = __fadd_rn(s, __fmul_rn(amp[k], cosf(__fadd_rn(__fmul_rn(w[k], tf), ph[k]))));
    long long o = (long long)t * D;
    out[o] = accumulate ? __fadd_rn(out[o], s) : s;
  }
}

// Backward of the above for the selected bins (the selection itself has zero gradient):
//   season = sum_k 2*(Re_k cos(w_k t) - Im_k sin(w_k t))  =>  dRe_k = 2 sum_t g_t cos(w_k t), dIm_k = -2 sum_t g_t sin(w_k t)
// gspec must be zero-filled by the caller; only the selected bins are written.
__global__ void dts_fourier_topk_bwd_kernel(const float* __restrict__ gseason, const int* __restrict__ idx,
                                            long long gspec_row_stride, long long rows, int NF, int low, int seq, int D,
                                            int top_k, float* __restrict__ gspec) {
  const long long n = rows * D;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long r = i / D;
  const int e = (int)(i - r * D);
  float w[TOPK_MAX], gre[TOPK_MAX], gim[TOPK_MAX];
  int bj[TOPK_MAX];
#pragma unroll
  for (int k = 0; k < TOPK_MAX; ++k) {
    gre[k] = 0.f; gim[k] = 0.f; w[k] = 0.f; bj[k] = 0;
    if (k < top_k) {
      bj[k] = idx[(r * top_k + k) * D + e];
      w[k] = __fmul_rn(6.283185307179586f, __fdiv_rn((float)(bj[k] + low), (float)seq));
    }
  }
  const float* g = gseason + r * (long long)seq * D + e;
  for (int t = 0; t < seq; ++t) {
    float gt = g[(long long)t * D], tf = (float)t;
#pragma unroll
    for (int k = 0; k < TOPK_MAX; ++k)
      if (k < top_k) {
        float s, c;
        sincosf(__fmul_rn(w[k], tf), &s, &c);
        gre[k] = fmaf(gt, c, gre[k]);
        gim[k] = fmaf(gt, s, gim[k]);
      }
  }
  float* gr = gspec + r * gspec_row_stride + e;
  float* gi = gr + (long long)NF * D;
#pragma unroll
  for (int k = 0; k < TOPK_MAX; ++k)
    if (k < top_k) {
      gr[(long long)bj[k] * D] = 2.0f * gre[k];
      gi[(long long)bj[k] * D] = -2.0f * gim[k];
    }
}

inline unsigned stream_grid(long long n, int block, int sms) {
  long long g = (n + block - 1) / block;
  long long cap = (long long)sms * 16;        // grid-stride: a few waves of resident CTAs
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

cudaError_t upd_launch_dts_ddim(const float* x0_raw, const float* img, long long n, float sqrt_recip, float sqrt_recipm1,
                                float sqrt_an, float c, float sigma, const float* noise, int last, float* x_start,
                                float* pred_mean, float* img_out, int sms, cudaStream_t stream) {
  dts_ddim_step_kernel<<<stream_grid(n, 256, sms), 256, 0, stream>>>(x0_raw, img, n, sqrt_recip, sqrt_recipm1, sqrt_an, c,
                                                                      sigma, noise, last, x_start, pred_mean, img_out);
  return cudaGetLastError();
}

cudaError_t upd_launch_dts_adagrad(float* p, const float* g, long long n, float lr, int sms, cudaStream_t stream) {
  dts_adagrad_kernel<<<stream_grid(n, 256, sms), 256, 0, stream>>>(p, g, n, lr);
  return cudaGetLastError();
}

cudaError_t upd_launch_dts_infill(float* img, const float* refined, const float* target, const float* noise,
                                  long long rows, int seq, int L_obs, int F, float sqrt_ac, float sqrt_1mac, int sms,
                                  cudaStream_t stream) {
  dts_infill_kernel<<<stream_grid(rows * seq * F, 256, sms), 256, 0, stream>>>(img, refined, target, noise, rows, seq,
                                                                                L_obs, F, sqrt_ac, sqrt_1mac);
  return cudaGetLastError();
}

cudaError_t upd_launch_gauss_fill(float* out, long long rows, long long row_elems, uint64_t seed, uint64_t row_base,
                                  uint32_t draw, int sms, cudaStream_t stream) {
  gauss_fill_kernel<<<stream_grid(rows * row_elems, 256, sms), 256, 0, stream>>>(out, rows, row_elems, seed, row_base, draw);
  return cudaGetLastError();
}

cudaError_t upd_launch_dts_fourier_fwd(const float* spec, long long spec_row_stride, long long rows, int NF, int low,
                                       int seq, int D, int top_k, int accumulate, float* season, int* idx,
                                       cudaStream_t stream) {
  if (top_k < 1 || top_k > TOPK_MAX || top_k > NF) return cudaErrorInvalidValue;
  long long n = rows * D;
  dts_fourier_topk_fwd_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(spec, spec_row_stride, rows, NF, low, seq, D,
                                                                                top_k, accumulate, season, idx);
  return cudaGetLastError();
}

cudaError_t upd_launch_dts_fourier_bwd(const float* gseason, const int* idx, long long gspec_row_stride, long long rows,
                                       int NF, int low, int seq, int D, int top_k, float* gspec, cudaStream_t stream) {
  if (top_k < 1 || top_k > TOPK_MAX || top_k > NF) return cudaErrorInvalidValue;
  long long n = rows * D;
  dts_fourier_topk_bwd_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(gseason, idx, gspec_row_stride, rows, NF, low,
                                                                                seq, D, top_k, gspec);
  return cudaGetLastError();
}
