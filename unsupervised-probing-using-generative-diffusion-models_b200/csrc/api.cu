// extern "C" entry points (include/upd_b200.h): argument validation, host-side weight packing,
// kernel launches on the caller's stream.  No global state, no allocation.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "upd_b200.h"
#include "sampler_params.cuh"
#include "upd_common.cuh"

cudaError_t upd_launch_mpv(const float* traj, const float* scale, int n_win, int B, int K, int O, int F,
                           float* var_out, float* mean_out, float* mpv, float* pmean, float* mpv_f,
                           cudaStream_t stream);
cudaError_t upd_launch_sigma(const UpdSigmaWeights& w, const float* x, int rows, int Lw, int R, int F, int H,
                             int O, float add_eps, float* gx, cudaStream_t stream);

cudaError_t upd_launch_dts_ddim(const float* x0_raw, const float* img, long long n, float sqrt_recip, float sqrt_recipm1,
                                float sqrt_an, float c, float sigma, const float* noise, int last, float* x_start,
                                float* pred_mean, float* img_out, int sms, cudaStream_t stream);
cudaError_t upd_launch_dts_adagrad(float* p, const float* g, long long n, float lr, int sms, cudaStream_t stream);
cudaError_t upd_launch_dts_infill(float* img, const float* refined, const float* target, const float* noise,
                                  long long rows, int seq, int L_obs, int F, float sqrt_ac, float sqrt_1mac, int sms,
                                  cudaStream_t stream);
cudaError_t upd_launch_gauss_fill(float* out, long long rows, long long row_elems, uint64_t seed, uint64_t row_base,
                                  uint32_t draw, int sms, cudaStream_t stream);
cudaError_t upd_launch_dts_fourier_fwd(const float* spec, long long spec_row_stride, long long rows, int NF, int low,
                                       int seq, int D, int top_k, int accumulate, float* season, int* idx,
                                       cudaStream_t stream);
cudaError_t upd_launch_dts_fourier_bwd(const float* gseason, const int* idx, long long gspec_row_stride, long long rows,
                                       int NF, int low, int seq, int D, int top_k, float* gspec, cudaStream_t stream);
cudaError_t upd_launch_nsx_step(const float* e, const float* w4, const float* b4, const float* ws, const float* bs,
                                const float* y, const float* yT, const float* gx, const float* z, const float* sched,
                                int n_steps, int t, long long N, int DH, int T, int F, float* out, float* eps_out,
                                float* sig_out, int sms, cudaStream_t stream);
cudaError_t upd_launch_stg_posterior(const float* xt, const float* pred, const float* z, long long n, float a, float b,
                                     float c, float* out, int sms, cudaStream_t stream);
cudaError_t upd_launch_stg_gated_aggregate(const float* kqvs, const int* rowptr, const int* col, const float* bias,
                                           long long N, int V, int C, int relu, float* out, int sms, cudaStream_t stream);
cudaError_t upd_launch_stg_tcn_ln(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                                  const float* gamma, const float* beta, long long N, int CI, int C, int T, float* hn,
                                  void* a3, const float* wsc, float* sc_out, const float* x2, int CI2, int sms,
                                  cudaStream_t stream);
cudaError_t upd_launch_stg_conv1d(const float* x, const float* w, const float* b, long long N, int CI, int CO, int Tin, int Tout,
                                   int K, int stride, int pad, int transposed, float* y, int sms, cudaStream_t stream);
cudaError_t upd_launch_fx_split(const float* x, long long rows, int K, int H, int L, int act, void* a3, cudaStream_t stream);
cudaError_t upd_launch_fx_add_ln_split(const float* x, const float* res, const float* g1, const float* b1,
                                       const float* g2, const float* b2, long long rows, int K, float* y, void* a3,
                                       cudaStream_t stream);
cudaError_t upd_launch_fx_attention(const float* q, long long q_stride, const float* k, const float* v, long long kv_stride,
                                    const float* tau, const float* delta, int delta_pitch, int B, int H, int Lq, int S,
                                    int causal, float scale, void* a3, cudaStream_t stream);
cudaError_t upd_launch_dts_attention(const float* q, long long q_stride, const float* k, const float* v, long long kv_stride,
                                     int R, int H, int Lq, int S, float scale, float* o, float* lse, const float* tau,
                                     const float* delta, int delta_pitch, int causal, cudaStream_t stream);
cudaError_t upd_launch_dts_attention_bwd(const float* q, long long q_stride, const float* k, const float* v,
                                         long long kv_stride, int R, int H, int Lq, int S, float scale, const float* o,
                                         const float* lse, const float* d_o, float* dq, long long dq_stride, float* dk,
                                         float* dv, long long dkv_stride, cudaStream_t stream);
cudaError_t upd_launch_dts_attention_tc(const float* q, long long q_stride, const float* k, const float* v, long long kv_stride,
                                        int R, int H, int Lq, int S, float scale, float* o, float* lse, void* a3,
                                        cudaStream_t stream);
cudaError_t upd_launch_dts_attention_tc_bwd(const float* q, long long q_stride, const float* k, const float* v,
                                            long long kv_stride, int R, int H, int Lq, int S, float scale, const float* o,
                                            const float* lse, const float* d_o, float* dq, long long dq_stride, float* dk,
                                            float* dv, long long dkv_stride, cudaStream_t stream);
cudaError_t upd_launch_fx_embed_split(const float* x, const float* w, const float* pe, long long rows, int L, int NF, int K,
                                      float* y, void* a3, cudaStream_t stream);
cudaError_t upd_launch_dts_layernorm(const float* x, const float* gamma, const float* beta, long long rows, int D, float* y,
                                     float* stats, void* a3, cudaStream_t stream);
cudaError_t upd_launch_dts_layernorm_bwd(const float* x, const float* dy, const float* gamma, const float* stats,
                                         long long rows, int D, float* dx, cudaStream_t stream);

cudaError_t upd_launch_gram_centered(const float* traj, int W, int K, int D, double* gram, cudaStream_t stream);
cudaError_t upd_launch_prediction_error(const float* mean, const float* target, int W, int O, int F, float* err,
                                        cudaStream_t stream);
cudaError_t upd_launch_gemm3(const void* a3, const void* w3, long long M, int Nw, int n_out, int Kp, float* out,
                             const float* addend, int sms, cudaStream_t stream);

namespace {

thread_local int g_last_cuda_error = 0;

int cuda_fail(cudaError_t e) {
  g_last_cuda_error = (int)e;
  return UPD_ERR_CUDA;
}

// Device checks shared by every launch: sm_100 family, SM count for persistent grids.  The two attributes are
// immutable per device, so they are looked up once per (thread, device): cudaDeviceGetAttribute costs ~15 us,
// more than the launch itself for the small per-step kernels.
int device_info(int* sms) {
  thread_local int cached_dev = -1, cached_major = 0, cached_sms = 0;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e);
  if (dev != cached_dev) {
    int major = 0, n = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) return cuda_fail(e);
    e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return cuda_fail(e);
    cached_dev = dev; cached_major = major; cached_sms = n;
  }
  if (cached_major != 10) return UPD_ERR_NO_DEVICE;
  *sms = cached_sms;
  return UPD_OK;
}

bool dims_ok(int kind, int F, int T) {
  return (kind == UPD_KIND_NSDIFF || kind == UPD_KIND_TMDM) && F >= 1 && F <= UPD_MAX_F && T >= 2 && T <= UPD_MAX_T;
}

float tf32_round(float x) {  // round-to-nearest (ties away) to 10 explicit mantissa bits, like cvt.rna.tf32
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return x;
  u = (u + 0x1000u) & 0xffffe000u;
  memcpy(&x, &u, 4);
  return x;
}

// Power-of-two scale that puts max|W| near 2^10: keeps the fp16 lo parts out of the subnormal
// range without risking overflow.  Exactly undone in the epilogue.
float pick_wscale(const float* w, int n) {
  float m = 0.f;
  for (int i = 0; i < n; ++i) m = fmaxf(m, fabsf(w[i]));
  if (!(m > 0.f) || !isfinite(m)) return 1.f;
  int e = (int)floorf(log2f(1024.f / m));
  if (e > 24) e = 24;
  if (e < -24) e = -24;
  return ldexpf(1.f, e);
}

void pack_umma_f16(const float* w /*[128][128] row-major n,k*/, float wscale, unsigned char* hi, unsigned char* lo) {
  for (int n = 0; n < 128; ++n)
    for (int k = 0; k < 128; ++k) {
      float v = w[n * 128 + k] * wscale;
      __half h = __float2half_rn(v);
      __half l = __float2half_rn(v - __half2float(h));
      size_t off = (size_t)(k / 8) * 2048 + (size_t)n * 16 + (size_t)(k % 8) * 2;
      memcpy(hi + off, &h, 2);
      memcpy(lo + off, &l, 2);
    }
}

}  // namespace

extern "C" {

const char* upd_error_string(int code) {
  switch (code) {
    case UPD_OK: return "ok";
    case UPD_ERR_BAD_ARG: return "bad argument";
    case UPD_ERR_UNSUPPORTED: return "unsupported shape";
    case UPD_ERR_CUDA: return "CUDA runtime error";
    case UPD_ERR_NO_DEVICE: return "no sm_100 device";
    default: return "unknown error";
  }
}

int upd_last_cuda_error(void) { return g_last_cuda_error; }
int upd_abi_version(void) { return UPD_ABI_VERSION; }

size_t upd_denoiser_pack_bytes(int kind, int F, int T) {
  if (!dims_ok(kind, F, T)) return 0;
  return upd_make_layout(kind, F, T).total_bytes;
}

int upd_denoiser_pack(const UpdDenoiserWeights* w, void* out_host, size_t capacity) {
  if (!w || !out_host) return UPD_ERR_BAD_ARG;
  if (!dims_ok(w->kind, w->F, w->T)) return UPD_ERR_UNSUPPORTED;
  const bool ns = (w->kind == UPD_KIND_NSDIFF);
  if (!w->lin1_w || !w->lin1_b || !w->embed1 || !w->lin2_w || !w->lin2_b || !w->embed2 || !w->lin3_w ||
      !w->lin3_b || !w->embed3 || !w->lin4_w || !w->lin4_b || !w->sched || (ns && (!w->sigma_w || !w->sigma_b)))
    return UPD_ERR_BAD_ARG;
  const UpdPackLayout L = upd_make_layout(w->kind, w->F, w->T);
  if (capacity < L.total_bytes) return UPD_ERR_BAD_ARG;
  unsigned char* o = (unsigned char*)out_host;
  memset(o, 0, L.total_bytes);
  auto fp = [&](uint32_t off) { return (float*)(o + off); };
  const int F = w->F, T = w->T, IN = L.in_dim;

  const float s2 = pick_wscale(w->lin2_w, 128 * 128), s3 = pick_wscale(w->lin3_w, 128 * 128);
  pack_umma_f16(w->lin2_w, s2, o + L.u2hi, o + L.u2lo);
  pack_umma_f16(w->lin3_w, s3, o + L.u3hi, o + L.u3lo);
  // lin1 as tf32 hi/lo with the bias as column IN (the A tile carries a constant 1 there)
  for (int n = 0; n < 128; ++n)
    for (int k = 0; k < L.K1; ++k) {
      float v = (k < IN) ? w->lin1_w[n * IN + k] : (k == IN ? w->lin1_b[n] : 0.f);
      float h = tf32_round(v), l = tf32_round(v - h);
      size_t off = (size_t)(k / 4) * 2048 + (size_t)n * 16 + (size_t)(k % 4) * 4;
      memcpy(o + L.u1hi + off, &h, 4);
      memcpy(o + L.u1lo + off, &l, 4);
    }
  memcpy(fp(L.b2), w->lin2_b, 128 * 4);
  memcpy(fp(L.b3), w->lin3_b, 128 * 4);
  memcpy(fp(L.e1), w->embed1, (size_t)L.TE * 128 * 4);
  memcpy(fp(L.e2), w->embed2, (size_t)L.TE * 128 * 4);
  memcpy(fp(L.e3), w->embed3, (size_t)L.TE * 128 * 4);
  memcpy(fp(L.w4), w->lin4_w, (size_t)F * 128 * 4);
  memcpy(fp(L.b4), w->lin4_b, (size_t)F * 4);
  if (ns) {
    memcpy(fp(L.ws), w->sigma_w, (size_t)F * 128 * 4);
    memcpy(fp(L.bs), w->sigma_b, (size_t)F * 4);
  }
  fp(L.scales)[0] = 1.f / s2;
  fp(L.scales)[1] = 1.f / s3;
  // NsDiff: layers 2 and 3 read an L2-normalised vector, so their base-2 pre-activation is bounded by
  // (|W_row|_2 + |b|) |e| log2e (+1 % for the rounding of the split operand).  The tcgen05 kernel drops its ex2
  // overflow guard when this bound is below 120; 0 = no bound (TMDM: no normalisation).
  if (ns) {
    double zmax = 0.0;
    const float* ws[2] = {w->lin2_w, w->lin3_w};
    const float* bs[2] = {w->lin2_b, w->lin3_b};
    const float* es[2] = {w->embed2, w->embed3};
    for (int l = 0; l < 2; ++l)
      for (int n = 0; n < 128; ++n) {
        double nrm = 0.0, emax = 0.0;
        for (int k = 0; k < 128; ++k) nrm += (double)ws[l][n * 128 + k] * ws[l][n * 128 + k];
        for (int t = 0; t < L.TE; ++t) emax = fmax(emax, fabs((double)es[l][t * 128 + n]));
        zmax = fmax(zmax, 1.01 * (sqrt(nrm) + fabs((double)bs[l][n])) * emax * 1.4426950408889634);
      }
    fp(L.scales)[2] = isfinite(zmax) ? (float)zmax : 0.f;
  }
  memcpy(fp(L.sched), w->sched, (size_t)L.n_sched * T * 4);
  // fp32 k-major transposes for the FFMA kernel
  for (int k = 0; k < IN; ++k)
    for (int n = 0; n < 128; ++n) fp(L.w1t)[k * 128 + n] = w->lin1_w[n * IN + k];
  memcpy(fp(L.b1), w->lin1_b, 128 * 4);
  for (int k = 0; k < 128; ++k)
    for (int n = 0; n < 128; ++n) {
      fp(L.w2t)[k * 128 + n] = w->lin2_w[n * 128 + k];
      fp(L.w3t)[k * 128 + n] = w->lin3_w[n * 128 + k];
    }
  return UPD_OK;
}

static int sample_common(int kind, const void* packed, const float* y0_hat, const float* gx, int n_win, int B,
                         int K, int S, int O, int F, int T, uint64_t seed, uint64_t window_base,
                         const float* noise, float* out, int impl, void* stream) {
  if (!packed || !out || n_win <= 0 || B <= 0 || K <= 0 || S <= 0 || O <= 0) return UPD_ERR_BAD_ARG;
  if (kind == UPD_KIND_NSDIFF && !gx) return UPD_ERR_BAD_ARG;
  if (kind == UPD_KIND_TMDM && !y0_hat) return UPD_ERR_BAD_ARG;
  if (!dims_ok(kind, F, T)) return UPD_ERR_UNSUPPORTED;
  if (noise && (K % S) != 0) return UPD_ERR_BAD_ARG;
  if (impl != UPD_IMPL_TCGEN05 && impl != UPD_IMPL_SIMT && impl != UPD_IMPL_TCGEN05_X2 && impl != UPD_IMPL_TCGEN05_WS)
    return UPD_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(packed) & 127) != 0) return UPD_ERR_BAD_ARG;
  int sms = 0;
  int rc = device_info(&sms);
  if (rc != UPD_OK) return rc;
  UpdSamplerParams p;
  p.packed = packed; p.y0_hat = y0_hat; p.gx = gx; p.noise = noise; p.out = out;
  p.n_win = n_win; p.B = B; p.K = K; p.S = S; p.O = O; p.T = T;
  p.n_rows = (long long)n_win * B * K * O;
  p.seed = seed; p.window_base = window_base;
#ifdef UPD_TRACE
  { const char* e = getenv("UPD_TRACE_PTR"); p.trace = e ? (long long*)strtoull(e, nullptr, 16) : nullptr; }
#endif
  cudaError_t e;
  if (impl == UPD_IMPL_SIMT) {
    e = upd_launch_sampler_simt(p, kind, F, sms, (cudaStream_t)stream);
  } else if (impl == UPD_IMPL_TCGEN05) {
    // the library's choice: the warp-specialised kernel where it is built (F <= 2), else the two-tile kernel; a step
    // count whose embedding tables do not fit next to the weight image in shared memory (T > ~40) runs on the FFMA kernel
    e = upd_launch_sampler_ws(p, kind, F, sms, (cudaStream_t)stream);
    if (e == cudaErrorInvalidValue) e = upd_launch_sampler_tc(p, kind, F, sms, (cudaStream_t)stream);
    if (e == cudaErrorInvalidValue) e = upd_launch_sampler_simt(p, kind, F, sms, (cudaStream_t)stream);
  } else if (impl == UPD_IMPL_TCGEN05_WS) {
    e = upd_launch_sampler_ws(p, kind, F, sms, (cudaStream_t)stream);
  } else {
    e = upd_launch_sampler_tc(p, kind, F, sms, (cudaStream_t)stream);
  }
  if (e == cudaErrorInvalidValue) return UPD_ERR_UNSUPPORTED;
  return e == cudaSuccess ? UPD_OK : cuda_fail(e);
}

int upd_nsdiff_sample(const void* packed_dev, const float* y0_hat_dev, const float* gx_dev, int n_win, int B, int K,
                      int S, int O, int F, int T, uint64_t seed, uint64_t window_base, const float* noise_dev,
                      float* out_dev, int impl, void* stream) {
  return sample_common(UPD_KIND_NSDIFF, packed_dev, y0_hat_dev, gx_dev, n_win, B, K, S, O, F, T, seed, window_base,
                       noise_dev, out_dev, impl, stream);
}

int upd_tmdm_sample(const void* packed_dev, const float* y0_hat_dev, int n_win, int B, int K, int S, int Lr, int F,
                    int T, uint64_t seed, uint64_t window_base, const float* noise_dev, float* out_dev, int impl,
                    void* stream) {
  return sample_common(UPD_KIND_TMDM, packed_dev, y0_hat_dev, nullptr, n_win, B, K, S, Lr, F, T, seed, window_base,
                       noise_dev, out_dev, impl, stream);
}

size_t upd_mpv_scratch_bytes(int n_win, int B, int O, int F) {
  if (n_win <= 0 || B <= 0 || O <= 0 || F <= 0) return 0;
  return 2 * sizeof(float) * (size_t)n_win * B * O * F + 256;
}

int upd_mpv_reduce(const float* traj_dev, const float* scale_dev, int n_win, int B, int K, int O, int F,
                   float* var_dev, float* mean_dev, float* mpv_dev, float* pmean_dev, float* mpv_f_dev,
                   void* scratch_dev, void* stream) {
  if (!traj_dev || n_win <= 0 || B <= 0 || K <= 0 || O <= 0) return UPD_ERR_BAD_ARG;
  if (F < 1 || F > UPD_MAX_F) return UPD_ERR_UNSUPPORTED;
  if ((!var_dev || !mean_dev) && !scratch_dev) return UPD_ERR_BAD_ARG;
  int sms = 0;
  int rc = device_info(&sms);
  if (rc != UPD_OK) return rc;
  size_t n = (size_t)n_win * B * O * F;
  float* sc = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(scratch_dev) + 127) & ~(uintptr_t)127);
  float* var = var_dev ? var_dev : sc;
  float* mean = mean_dev ? mean_dev : sc + ((n + 31) & ~(size_t)31);
  cudaError_t e = upd_launch_mpv(traj_dev, scale_dev, n_win, B, K, O, F, var, mean, mpv_dev, pmean_dev, mpv_f_dev,
                                 (cudaStream_t)stream);
  return e == cudaSuccess ? UPD_OK : cuda_fail(e);
}

int upd_sigma_estimation(const UpdSigmaWeights* w, const float* x_dev, int rows, int L, int R, int F, int H, int O,
                         float add_eps, float* gx_dev, void* stream) {
  if (!w || !x_dev || !gx_dev || rows <= 0 || L <= 0 || R < 1 || R >= L || H <= 0 || O <= 0) return UPD_ERR_BAD_ARG;
  if (!w->w0 || !w->b0 || !w->ln1_w || !w->ln1_b || !w->w3 || !w->b3 || !w->ln2_w || !w->ln2_b || !w->w6 || !w->b6)
    return UPD_ERR_BAD_ARG;
  if (F < 1 || F > UPD_MAX_F || O > H) return UPD_ERR_UNSUPPORTED;
  int sms = 0;
  int rc = device_info(&sms);
  if (rc != UPD_OK) return rc;
  cudaError_t e = upd_launch_sigma(*w, x_dev, rows, L, R, F, H, O, add_eps, gx_dev, (cudaStream_t)stream);
  if (e == cudaErrorInvalidValue) return UPD_ERR_UNSUPPORTED;
  return e == cudaSuccess ? UPD_OK : cuda_fail(e);
}

#define UPD_DEVICE_OR_RETURN() int sms = 0; { int rc_ = device_info(&sms); if (rc_ != UPD_OK) return rc_; }
#define UPD_FINISH(e) do { cudaError_t e_ = (e); if (e_ == cudaErrorInvalidValue) return UPD_ERR_UNSUPPORTED; \
                           return e_ == cudaSuccess ? UPD_OK : cuda_fail(e_); } while (0)

int upd_dts_ddim_step(const float* x0_raw_dev, const float* img_dev, long long n, float sqrt_recip_ac,
                      float sqrt_recipm1_ac, float sqrt_alpha_next, float c, float sigma, const float* noise_dev, int last,
                      float* x_start_dev, float* pred_mean_dev, float* img_out_dev, void* stream) {
  if (!x0_raw_dev || !img_out_dev || n <= 0) return UPD_ERR_BAD_ARG;
  if (!last && (!img_dev || (sigma != 0.0f && !noise_dev))) return UPD_ERR_BAD_ARG;
  UPD_DEVICE_OR_RETURN();
  UPD_FINISH(upd_launch_dts_ddim(x0_raw_dev, img_dev, n, sqrt_recip_ac, sqrt_recipm1_ac, sqrt_alpha_next, c, sigma,
                                 noise_dev, last, x_start_dev, pred_mean_dev, img_out_dev, sms, (cudaStream_t)stream));
}

int upd_dts_adagrad_step(float* p_dev, const float* grad_dev, long long n, float lr, void* stream) {
  if (!p_dev || !grad_dev || n <= 0) return UPD_ERR_BAD_ARG;
  UPD_DEVICE_OR_RETURN();
  UPD_FINISH(upd_launch_dts_adagrad(p_dev, grad_dev, n, lr, sms, (cudaStream_t)stream));
}

int upd_dts_infill(float* img_dev, const float* refined_dev, const float* target_dev, const float* noise_dev,
                   long long rows, int seq, int L_obs, int F, float sqrt_ac, float sqrt_one_minus_ac, void* stream) {
  if (!img_dev || !refined_dev || !target_dev || rows <= 0 || seq <= 0 || F <= 0 || L_obs < 0 || L_obs > seq)
    return UPD_ERR_BAD_ARG;
  UPD_DEVICE_OR_RETURN();
  UPD_FINISH(upd_launch_dts_infill(img_dev, refined_dev, target_dev, noise_dev, rows, seq, L_obs, F, sqrt_ac,
                                   sqrt_one_minus_ac, sms, (cudaStream_t)stream));
}

int upd_gauss_fill(float* out_dev, long long rows, long long row_elems, uint64_t seed, uint64_t row_base, uint32_t draw,
                   void* stream) {
  if (!out_dev || rows <= 0 || row_elems <= 0 || row_elems > 0xffffffffLL) return UPD_ERR_BAD_ARG;
  UPD_DEVICE_OR_RETURN();
  UPD_FINISH(upd_launch_gauss_fill(out_dev, rows, row_elems, seed, row_base, draw, sms, (cudaStream_t)stream));
}

int upd_dts_fourier_topk(const float* spec_dev, long long spec_row_stride, long long rows, int NF, int low, int seq, int D,
                         int top_k, int accumulate, float* season_dev, int* idx_dev, void* stream) {
  if (!spec_dev || !season_dev || rows <= 0 || NF <= 0 || low < 0 || seq <= 0 || D <= 0 ||
      spec_row_stride < 2LL * NF * D) return UPD_ERR_BAD_ARG;
  UPD_DEVICE_OR_RETURN();
  UPD_FINISH(upd_launch_dts_fourier_fwd(spec_dev, spec_row_stride, rows, NF, low, seq, D, top_k, accumulate, season_dev,
                                        idx_dev, (cudaStream_t)stream));
}

int upd_dts_fourier_topk_bwd(const float* gseason_dev, const int* idx_dev, long long gspec_row_stride, long long rows,
                             int NF, int low, int seq, int D, int top_k, float* gspec_dev, void* stream) {
  if (!gseason_dev || !idx_dev || !gspec_dev || rows <= 0 || NF <= 0 || low < 0 || seq <= 0 || D <= 0 ||
      gspec_row_stride < 2LL * NF * D) return UPD_ERR_BAD_ARG;
  UPD_DEVICE_OR_RETURN();
  UPD_FINISH(upd_launch_dts_fourier_bwd(gseason_dev, idx_dev, gspec_row_stride, rows, NF, low, seq, D, top_k, gspec_dev,
                                        (cudaStream_t)stream));
}

int upd_stg_posterior(const float* xt_dev, const float* pred_dev, const float* z_dev, long long n, float a, float b,
                      float c, float* out_dev, void* stream) {
  if (!xt_dev || !pred_dev || !out_dev || n <= 0) return UPD_ERR_BAD_ARG;
  UPD_DEVICE_OR_RETURN();
  UPD_FINISH(upd_launch_stg_posterior(xt_dev, pred_dev, z_dev, n, a, b, c, out_dev, sms, (cudaStream_t)stream));
}

int upd_nsx_step(const float* e_dev, const float* w4_dev, const float* b4_dev, const float* ws_dev, const float* bs_dev,
                 const float* y_dev, const float* yT_dev, const float* gx_dev, const float* z_dev, const float* sched_dev,
                 int n_steps, int t, long long N, int DH, int T, int F, float* out_dev, float* eps_out_dev,
                 float* sig_out_dev, void* stream) {
  if (!e_dev || !w4_dev || !b4_dev || !ws_dev || !bs_dev || !sched_dev || N <= 0 || DH <= 0 || T <= 0 || F <= 0 ||
      n_steps <= 0 || t < 0 || t >= n_steps) return UPD_ERR_BAD_ARG;
  if (y_dev) {
    if (!yT_dev || !gx_dev || !out_dev || (t > 0) != (z_dev != nullptr)) return UPD_ERR_BAD_ARG;
  } else if (!eps_out_dev && !sig_out_dev) {
    return UPD_ERR_BAD_ARG;
  }
  UPD_DEVICE_OR_RETURN();
  UPD_FINISH(upd_launch_nsx_step(e_dev, w4_dev, b4_dev, ws_dev, bs_dev, y_dev, yT_dev, gx_dev, z_dev, sched_dev, n_steps, t,
                                 N, DH, T, F, out_dev, eps_out_dev, sig_out_dev, sms, (cudaStream_t)stream));
}

int upd_stg_gated_aggregate(const float* kqvs_dev, const int* rowptr_dev, const int* col_dev, const float* bias_dev,
                            long long N, int V, int C, int relu, float* out_dev, void* stream) {
  if (!kqvs_dev || !rowptr_dev || !col_dev || !out_dev || N <= 0 || V <= 0 || C <= 0 || (N % V) != 0)
    return UPD_ERR_BAD_ARG;
  UPD_DEVICE_OR_RETURN();
  UPD_FINISH(upd_launch_stg_gated_aggregate(kqvs_dev, rowptr_dev, col_dev, bias_dev, N, V, C, relu, out_dev, sms,
                                            (cudaStream_t)stream));
}

int upd_stg_tcn_ln(const float* x_dev, const float* w1_dev, const float* b1_dev, const float* w2_dev, const float* b2_dev,
                   const float* gamma_dev, const float* beta_dev, long long N, int CI, int C, int T, float* hn_dev,
                   void* a3_dev, const float* wsc_dev, float* sc_dev, void* stream) {
  if (!x_dev || !w1_dev || !b1_dev || !w2_dev || !b2_dev || !gamma_dev || !beta_dev || (!hn_dev && !a3_dev) || N <= 0 ||
      T <= 0)
    return UPD_ERR_BAD_ARG;
  if ((C != 4 && C != 8 && C != 16) || CI < 1 || (T & 1) != 0) return UPD_ERR_UNSUPPORTED;
  UPD_DEVICE_OR_RETURN();
  UPD_FINISH(upd_launch_stg_tcn_ln(x_dev, w1_dev, b1_dev, w2_dev, b2_dev, gamma_dev, beta_dev, N, CI, C, T, hn_dev, a3_dev,
                                   wsc_dev, sc_dev, nullptr, 0, sms, (cudaStream_t)stream));
}

int upd_stg_tcn_ln_cat(const float* x_dev, int CI1, const float* x2_dev, int CI2, const float* w1_dev, const float* b1_dev,
                       const float* w2_dev, const float* b2_dev, const float* gamma_dev, const float* beta_dev, long long N,
                       int C, int T, float* hn_dev, void* a3_dev, const float* wsc_dev, float* sc_dev, void* stream) {
  if (!x_dev || !x2_dev || !w1_dev || !b1_dev || !w2_dev || !b2_dev || !gamma_dev || !beta_dev || (!hn_dev && !a3_dev) ||
      N <= 0 || T <= 0 || CI1 < 1 || CI2 < 1)
    return UPD_ERR_BAD_ARG;
  if ((C != 4 && C != 8 && C != 16) || (T & 1) != 0) return UPD_ERR_UNSUPPORTED;
  UPD_DEVICE_OR_RETURN();
  UPD_FINISH(upd_launch_stg_tcn_ln(x_dev, w1_dev, b1_dev, w2_dev, b2_dev, gamma_dev, beta_dev, N, CI1 + CI2, C, T, hn_dev,
                                   a3_dev, wsc_dev, sc_dev, x2_dev, CI2, sms, (cudaStream_t)stream));
}

int upd_stg_conv1d(const float* x_dev, const float* w_dev, const float* b_dev, long long N, int CI, int CO, int Tin, int K,
                   int stride, int pad, int transposed, float* y_dev, void* stream) {
  if (!x_dev || !w_dev || !y_dev || N <= 0 || CI < 1 || CO < 1 || Tin < 1 || K < 1 || stride < 1 || pad < 0)
    return UPD_ERR_BAD_ARG;
  const int Tout = transposed ? (Tin - 1) * stride - 2 * pad + K : (Tin + 2 * pad - K) / stride + 1;
  if (Tout < 1) return UPD_ERR_BAD_ARG;
  if ((long long)CI * K * CO > 12288) return UPD_ERR_UNSUPPORTED;
  UPD_DEVICE_OR_RETURN();
  UPD_FINISH(upd_launch_stg_conv1d(x_dev, w_dev, b_dev, N, CI, CO, Tin, Tout, K, stride, pad, transposed, y_dev, sms,
                                   (cudaStream_t)stream));
}

int upd_gram_centered(const float* traj_dev, int n_win, int K, int D, double* gram_dev, void* stream) {
  if (!traj_dev || !gram_dev || n_win <= 0 || K <= 1 || D <= 0) return UPD_ERR_BAD_ARG;
  int sms = 0;
  int rc = device_info(&sms);
  if (rc != UPD_OK) return rc;
  cudaError_t e = upd_launch_gram_centered(traj_dev, n_win, K, D, gram_dev, (cudaStream_t)stream);
  if (e == cudaErrorInvalidValue) return UPD_ERR_UNSUPPORTED;
  return e == cudaSuccess ? UPD_OK : cuda_fail(e);
}

int upd_prediction_error(const float* mean_dev, const float* target_dev, int n_win, int O, int F, float* err_dev,
                         void* stream) {
  if (!mean_dev || !target_dev || !err_dev || n_win <= 0 || O <= 0) return UPD_ERR_BAD_ARG;
  if (F < 1 || F > UPD_MAX_F) return UPD_ERR_UNSUPPORTED;
  int sms = 0;
  int rc = device_info(&sms);
  if (rc != UPD_OK) return rc;
  cudaError_t e = upd_launch_prediction_error(mean_dev, target_dev, n_win, O, F, err_dev, (cudaStream_t)stream);
  return e == cudaSuccess ? UPD_OK : cuda_fail(e);
}

int upd_gemm3(const void* a3_dev, const void* w3_dev, long long M, int Nw, int n_out, int Kp, float* out_dev,
              const float* addend_dev, void* stream) {
  if (!a3_dev || !w3_dev || !out_dev || M <= 0 || Nw <= 0 || n_out <= 0 || n_out > Nw || Kp <= 0) return UPD_ERR_BAD_ARG;
  int sms = 0;
  int rc = device_info(&sms);
  if (rc != UPD_OK) return rc;
  cudaError_t e = upd_launch_gemm3(a3_dev, w3_dev, M, Nw, n_out, Kp, out_dev, addend_dev, sms, (cudaStream_t)stream);
  if (e == cudaErrorInvalidValue) return UPD_ERR_UNSUPPORTED;
  return e == cudaSuccess ? UPD_OK : cuda_fail(e);
}

int upd_fx_split(const float* x_dev, long long rows, int K, int H, int L, int act, void* a3_dev, void* stream) {
  if (!x_dev || !a3_dev || rows <= 0 || act < 0 || act > 2) return UPD_ERR_BAD_ARG;
  UPD_DEVICE_OR_RETURN();
  UPD_FINISH(upd_launch_fx_split(x_dev, rows, K, H, L, act, a3_dev, (cudaStream_t)stream));
}

int upd_fx_add_ln_split(const float* x_dev, const float* res_dev, const float* g1_dev, const float* b1_dev,
                        const float* g2_dev, const float* b2_dev, long long rows, int K, float* y_dev, void* a3_dev,
                        void* stream) {
  if (!x_dev || !g1_dev || !b1_dev || rows <= 0 || (!y_dev && !a3_dev) || ((g2_dev == nullptr) != (b2_dev == nullptr)))
    return UPD_ERR_BAD_ARG;
  UPD_DEVICE_OR_RETURN();
  UPD_FINISH(upd_launch_fx_add_ln_split(x_dev, res_dev, g1_dev, b1_dev, g2_dev, b2_dev, rows, K, y_dev, a3_dev,
                                        (cudaStream_t)stream));
}

int upd_fx_attention(const float* q_dev, long long q_row_stride, const float* k_dev, const float* v_dev,
                     long long kv_row_stride, const float* tau_dev, const float* delta_dev, int delta_pitch, int B, int H,
                     int Lq, int S, int head_dim, int causal, float scale, void* a3_dev, void* stream) {
  if (!q_dev || !k_dev || !v_dev || !a3_dev || B <= 0 || H <= 0 || Lq <= 0 || S <= 0) return UPD_ERR_BAD_ARG;
  if (delta_dev && delta_pitch < S) return UPD_ERR_BAD_ARG;
  if (head_dim != 64 || S > 192 || (long long)B * H > 0x7fffffffLL) return UPD_ERR_UNSUPPORTED;
  UPD_DEVICE_OR_RETURN();
  UPD_FINISH(upd_launch_fx_attention(q_dev, q_row_stride, k_dev, v_dev, kv_row_stride, tau_dev, delta_dev, delta_pitch, B,
                                     H, Lq, S, causal, scale, a3_dev, (cudaStream_t)stream));
}

// UPD_DTS_ATTN_FFMA=1: run the fp32 FFMA attention kernels even where the tcgen05 ones apply (comparison runs, tests).
static bool dts_attention_ffma_forced() {
  const char* e = getenv("UPD_DTS_ATTN_FFMA");
  return e && e[0] == '1';
}

int upd_dts_attention(const float* q_dev, long long q_row_stride, const float* k_dev, const float* v_dev,
                      long long kv_row_stride, int R, int H, int Lq, int S, int head_dim, float scale, float* o_dev,
                      float* lse_dev, void* a3_dev, void* stream) {
  if (!q_dev || !k_dev || !v_dev || !o_dev || R <= 0 || H <= 0 || Lq <= 0 || S <= 0) return UPD_ERR_BAD_ARG;
  if (head_dim != 16 || (long long)R * H > 0x7fffffffLL) return UPD_ERR_UNSUPPORTED;
  UPD_DEVICE_OR_RETURN();
  if (!dts_attention_ffma_forced()) {       // tcgen05 kernel; sequences beyond its limits fall through to the FFMA kernel
    cudaError_t e = upd_launch_dts_attention_tc(q_dev, q_row_stride, k_dev, v_dev, kv_row_stride, R, H, Lq, S, scale, o_dev,
                                                lse_dev, a3_dev, (cudaStream_t)stream);
    if (e != cudaErrorInvalidValue) UPD_FINISH(e);
  }
  cudaError_t e = upd_launch_dts_attention(q_dev, q_row_stride, k_dev, v_dev, kv_row_stride, R, H, Lq, S, scale, o_dev, lse_dev,
                                           nullptr, nullptr, 0, 0, (cudaStream_t)stream);
  if (e == cudaSuccess && a3_dev)           // FFMA path: the out-projection's operand in a second pass
    e = upd_launch_fx_split(o_dev, (long long)R * Lq, H * 16, 1, 1, 0, a3_dev, (cudaStream_t)stream);
  UPD_FINISH(e);
}

int upd_fx_attention_hs16(const float* q_dev, long long q_row_stride, const float* k_dev, const float* v_dev,
                          long long kv_row_stride, const float* tau_dev, const float* delta_dev, int delta_pitch, int B, int H,
                          int Lq, int S, int causal, float scale, float* o_dev, void* stream) {
  if (!q_dev || !k_dev || !v_dev || !o_dev || B <= 0 || H <= 0 || Lq <= 0 || S <= 0) return UPD_ERR_BAD_ARG;
  if (delta_dev && delta_pitch < S) return UPD_ERR_BAD_ARG;
  UPD_DEVICE_OR_RETURN();
  UPD_FINISH(upd_launch_dts_attention(q_dev, q_row_stride, k_dev, v_dev, kv_row_stride, B, H, Lq, S, scale, o_dev, nullptr,
                                      tau_dev, delta_dev, delta_pitch, causal, (cudaStream_t)stream));
}

int upd_dts_attention_bwd(const float* q_dev, long long q_row_stride, const float* k_dev, const float* v_dev,
                          long long kv_row_stride, int R, int H, int Lq, int S, int head_dim, float scale,
                          const float* o_dev, const float* lse_dev, const float* do_dev, float* dq_dev,
                          long long dq_row_stride, float* dk_dev, float* dv_dev, long long dkv_row_stride, void* stream) {
  if (!q_dev || !k_dev || !v_dev || !o_dev || !lse_dev || !do_dev || !dq_dev || !dk_dev || !dv_dev || R <= 0 || H <= 0 ||
      Lq <= 0 || S <= 0)
    return UPD_ERR_BAD_ARG;
  if (head_dim != 16 || (long long)R * H > 0x7fffffffLL) return UPD_ERR_UNSUPPORTED;
  UPD_DEVICE_OR_RETURN();
  if (!dts_attention_ffma_forced()) {
    cudaError_t e = upd_launch_dts_attention_tc_bwd(q_dev, q_row_stride, k_dev, v_dev, kv_row_stride, R, H, Lq, S, scale, o_dev,
                                                    lse_dev, do_dev, dq_dev, dq_row_stride, dk_dev, dv_dev, dkv_row_stride,
                                                    (cudaStream_t)stream);
    if (e != cudaErrorInvalidValue) UPD_FINISH(e);
  }
  UPD_FINISH(upd_launch_dts_attention_bwd(q_dev, q_row_stride, k_dev, v_dev, kv_row_stride, R, H, Lq, S, scale, o_dev,
                                          lse_dev, do_dev, dq_dev, dq_row_stride, dk_dev, dv_dev, dkv_row_stride,
                                          (cudaStream_t)stream));
}

int upd_fx_embed_split(const float* x_dev, const float* w_dev, const float* pe_dev, long long rows, int L, int NF, int K,
                       float* y_dev, void* a3_dev, void* stream) {
  if (!x_dev || !w_dev || !pe_dev || !y_dev || !a3_dev || rows <= 0) return UPD_ERR_BAD_ARG;
  UPD_DEVICE_OR_RETURN();
  UPD_FINISH(upd_launch_fx_embed_split(x_dev, w_dev, pe_dev, rows, L, NF, K, y_dev, a3_dev, (cudaStream_t)stream));
}

int upd_dts_layernorm(const float* x_dev, const float* gamma_dev, const float* beta_dev, long long rows, int D,
                      float* y_dev, float* stats_dev, void* a3_dev, void* stream) {
  if (!x_dev || !gamma_dev || !beta_dev || (!y_dev && !a3_dev) || rows <= 0) return UPD_ERR_BAD_ARG;
  UPD_DEVICE_OR_RETURN();
  UPD_FINISH(upd_launch_dts_layernorm(x_dev, gamma_dev, beta_dev, rows, D, y_dev, stats_dev, a3_dev, (cudaStream_t)stream));
}

int upd_dts_layernorm_bwd(const float* x_dev, const float* dy_dev, const float* gamma_dev, const float* stats_dev,
                          long long rows, int D, float* dx_dev, void* stream) {
  if (!x_dev || !dy_dev || !gamma_dev || !stats_dev || !dx_dev || rows <= 0) return UPD_ERR_BAD_ARG;
  UPD_DEVICE_OR_RETURN();
  UPD_FINISH(upd_launch_dts_layernorm_bwd(x_dev, dy_dev, gamma_dev, stats_dev, rows, D, dx_dev, (cudaStream_t)stream));
}

}  // extern "C"
