// Shared host/device definitions for the fused reverse-diffusion samplers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define UPD_HID 128           // hidden width of the conditional MLP denoiser (denoise.py:28-30)
#define UPD_MAX_F 4
#define UPD_MAX_T 64
#define UPD_ABI_VERSION 8

#ifdef __CUDACC__
#define UPD_HD __host__ __device__ __forceinline__
#else
#define UPD_HD inline
#endif

// ---------------------------------------------------------------------------------------------
// Packed weight blob.  One layout function shared by the host packer and every kernel, so no
// header has to be read back from the device.  All offsets are bytes from the blob start and
// multiples of 128.
//
//   [tc image]  contiguous, bulk-copied verbatim into shared memory by the tcgen05 kernel:
//     u2hi,u2lo,u3hi,u3lo : lin2/lin3 weights * wscale, fp16 hi/lo split, UMMA K-major no-swizzle
//                           core-matrix layout: elem(n,k) at (k/8)*2048 + n*16 + (k%8)*2
//     u1hi,u1lo           : [lin1 | bias column | 0-pad] as tf32 hi/lo, K1 = roundup(in+1, 8):
//                           elem(n,k) at (k/4)*2048 + n*16 + (k%4)*4
//     fp32 tail           : b2,b3 [128]; e1,e2,e3 [TE,128]; w4 [F,128]; ws [F,128]; b4,bs [4];
//                           scales [4] = (1/wscale2, 1/wscale3, zmax23, 0) -- zmax23: bound of the base-2
//                           pre-activations of layers 2/3 (NsDiff; 0 = unbounded); sched [n_sched, T]
//   [simt extra] fp32, k-major transposes for the FFMA kernel: w1t [in,128], b1 [128],
//                           w2t [128,128], w3t [128,128]
// ---------------------------------------------------------------------------------------------
struct UpdPackLayout {
  int kind, F, T, TE, in_dim, K1, n_sched;
  uint32_t u2hi, u2lo, u3hi, u3lo, u1hi, u1lo;
  uint32_t b2, b3, e1, e2, e3, w4, ws, b4, bs, scales, sched;
  uint32_t tc_image_bytes;
  uint32_t w1t, b1, w2t, w3t;
  uint32_t total_bytes;
};

UPD_HD uint32_t upd_align128(uint32_t x) { return (x + 127u) & ~127u; }

UPD_HD UpdPackLayout upd_make_layout(int kind, int F, int T) {
  UpdPackLayout L;
  L.kind = kind; L.F = F; L.T = T;
  L.TE = (kind == 1) ? T + 1 : T;
  L.in_dim = (kind == 1) ? 2 * F : 3 * F;
  L.K1 = ((L.in_dim + 1 + 7) / 8) * 8;
  L.n_sched = (kind == 1) ? 2 : 10;
  uint32_t o = 0;
  L.u2hi = o; o += 128 * 128 * 2;
  L.u2lo = o; o += 128 * 128 * 2;
  L.u3hi = o; o += 128 * 128 * 2;
  L.u3lo = o; o += 128 * 128 * 2;
  L.u1hi = o; o += 128 * L.K1 * 4;
  L.u1lo = o; o += 128 * L.K1 * 4;
  L.b2 = o; o += 128 * 4;
  L.b3 = o; o += 128 * 4;
  L.e1 = o; o += upd_align128(L.TE * 128 * 4);
  L.e2 = o; o += upd_align128(L.TE * 128 * 4);
  L.e3 = o; o += upd_align128(L.TE * 128 * 4);
  L.w4 = o; o += UPD_MAX_F * 128 * 4;
  L.ws = o; o += UPD_MAX_F * 128 * 4;
  L.b4 = o; o += 128;
  L.bs = o; o += 128;
  L.scales = o; o += 128;
  L.sched = o; o += upd_align128(L.n_sched * T * 4);
  L.tc_image_bytes = o;
  L.w1t = o; o += upd_align128(L.in_dim * 128 * 4);
  L.b1 = o; o += 128 * 4;
  L.w2t = o; o += 128 * 128 * 4;
  L.w3t = o; o += 128 * 128 * 4;
  L.total_bytes = o;
  return L;
}

// Row order of the NsDiff schedule table inside the blob (nsdiff_utils.py:271 argument order).
enum {
  SCH_ALPHAS = 0, SCH_OM_ABAR_SQRT = 1, SCH_ACP = 2, SCH_ACP_SUM = 3, SCH_ACP_PREV = 4,
  SCH_ACP_SUM_PREV = 5, SCH_BT = 6, SCH_BB = 7, SCH_BT_M1 = 8, SCH_BB_M1 = 9
};

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011) keyed per trajectory element; one call = 4 uniforms.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void upd_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                  uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Two N(0,1) draws from two 32-bit words (Box-Muller, fp32).
__device__ __forceinline__ void upd_box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
  float u1 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f);   // (0,1)
  float u2 = ((float)(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
  float r = sqrtf(-2.0f * __logf(u1));
  float s, c;
  __sincosf(6.283185307179586f * u2, &s, &c);
  z0 = r * c; z1 = r * s;
}

// N(0,1) for (global window, row-in-window, sample, position, feature, draw).  Counter layout:
//   c0 = position*F + feature, c1 = sample, c2 = row-in-window, c3 = draw | (window_lo << 8);
//   key = seed ^ (window_hi...).  One Philox call per scalar keeps the stream independent of how
//   the sweep is tiled; its cost is noise next to the 386 softplus of the same row-step.
__device__ __forceinline__ float upd_gauss(uint64_t seed, uint64_t window, uint32_t row, uint32_t sample,
                                           uint32_t elem, uint32_t draw) {
  uint32_t r[4];
  upd_philox4x32_10(elem, sample, row, (draw & 0xFFu) | ((uint32_t)window << 8),
                    (uint32_t)seed ^ (uint32_t)(window >> 24), (uint32_t)(seed >> 32), r);
  float z0, z1;
  upd_box_muller(r[0], r[1], z0, z1);
  return z0;
}

// Four N(0,1) draws from ONE Philox call, for the fused samplers: the draws number 4g .. 4g+3 of a trajectory element
// (draw 0 = y_T, draw i = reverse step t = T - i) share the call keyed by the draw group g.  The stream stays a pure
// function of (seed, global window, row, sample, element, draw), so it is still independent of how a sweep is tiled,
// batched or sharded; a sampler that walks the steps in order pays a quarter of the Philox + Box-Muller work per draw.
__device__ __forceinline__ float4 upd_gauss4(uint64_t seed, uint64_t window, uint32_t row, uint32_t sample, uint32_t elem,
                                             uint32_t draw_group) {
  uint32_t r[4];
  upd_philox4x32_10(elem, sample, row, (draw_group & 0xFFu) | ((uint32_t)window << 8),
                    (uint32_t)seed ^ (uint32_t)(window >> 24) ^ 0x9E3779B9u, (uint32_t)(seed >> 32), r);
  float4 z;
  upd_box_muller(r[0], r[1], z.x, z.y);
  upd_box_muller(r[2], r[3], z.z, z.w);
  return z;
}
__device__ __forceinline__ float upd_pick4(const float4& z, int i) {
  return (i & 2) ? ((i & 1) ? z.w : z.z) : ((i & 1) ? z.y : z.x);
}

// softplus(z) = log1p(exp(z)), beta=1, threshold=20 (F.softplus defaults).
// max(z,0) + ln2*lg2(1 + ex2(-|z|*log2e)): two MUFU ops; for z > 20 the correction is below
// half an ulp of z, so the threshold branch of the reference is reproduced without a select.
__device__ __forceinline__ float upd_softplus(float z) {
  float u;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(u) : "f"(-fabsf(z) * 1.4426950408889634f));
  float l;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(1.0f + u));
  return fmaf(l, 0.6931471805599453f, fmaxf(z, 0.0f));
}

// Same function with full relative accuracy for small results.  The sigma head ends in a softplus whose
// value is ~1e-4 (log1p of a tiny number): there the relative error of lg2(1+u) would reach 1e-3 and
// feed straight into sqrt(sigma)*z and the Sigma_Y0 quadratic, so that one call per (row, feature, step)
// uses the libm-grade path.
__device__ __forceinline__ float upd_softplus_accurate(float z) {
  return z > 20.0f ? z : log1pf(expf(z));
}

// Per-step scalars of the NsDiff posterior, hoisted out of the element loop.  Every quantity is
// formed with the reference's own fp32 operation order (nsdiff_utils.py:132-147, :40-56, :80-92).
struct UpdNsStep {
  float a, one_m_a, bt, bb, btm, bbm;
  float lam0, c1a, c1b, c2a, c2b, two_lam0, bb_m_bt, bbm_m_btm;
  float sqrt_abar, inv_sqrt_abar, one_m_sqrt_abar;
  float one_m_a_sq, a_one_m_a, sqrt_a, sqrt_abar_prev, sqrt_a_am1, one_m_sqrt_abar_prev;
};

__device__ __forceinline__ UpdNsStep upd_ns_step(const float* __restrict__ sched, int T, int t) {
  UpdNsStep s;
  s.a = sched[SCH_ALPHAS * T + t];
  s.bt = sched[SCH_BT * T + t];
  s.bb = sched[SCH_BB * T + t];
  s.btm = sched[SCH_BT_M1 * T + t];
  s.bbm = sched[SCH_BB_M1 * T + t];
  float om = sched[SCH_OM_ABAR_SQRT * T + t];
  float acp_prev = sched[SCH_ACP_PREV * T + t];
  s.one_m_a = 1.0f - s.a;
  s.one_m_a_sq = s.one_m_a * s.one_m_a;
  s.a_one_m_a = s.a * s.one_m_a;
  s.bbm_m_btm = s.bbm - s.btm;
  s.bb_m_bt = s.bb - s.bt;
  s.lam0 = s.a_one_m_a * s.btm;                                  // alpha*(1-alpha)*bt_m1
  s.c1a = s.one_m_a_sq * s.btm + s.a_one_m_a * s.bbm_m_btm;      // coefficient of gx in lambda_1
  s.c1b = s.a * s.btm + s.a_one_m_a;                             // coefficient of sigma_theta in lambda_1
  s.c2a = s.bbm_m_btm;                                           // used as gx^2*(1-a)^2*c2a
  s.c2b = __fadd_rn(__fsub_rn(__fmul_rn(s.a, s.bbm), __fmul_rn(s.a, s.btm)), s.one_m_a_sq);
  s.two_lam0 = 2.0f * s.lam0;
  s.sqrt_abar = sqrtf(1.0f - om * om);
  s.inv_sqrt_abar = 1.0f / s.sqrt_abar;
  s.one_m_sqrt_abar = 1.0f - s.sqrt_abar;
  s.sqrt_a = sqrtf(s.a);
  s.sqrt_abar_prev = sqrtf(acp_prev);
  s.sqrt_a_am1 = s.sqrt_a * (s.a - 1.0f);
  s.one_m_sqrt_abar_prev = 1.0f - s.sqrt_abar_prev;
  return s;
}

// One element of p_sample (nsdiff_utils.py:139-157) / p_sample_t_1to0 (:225-238).
// last == true returns y_0 reparam without noise.  __f*_rn keep the reference's unfused rounding.
__device__ __forceinline__ float upd_ns_update(const UpdNsStep& s, float y, float yT, float gx, float eps,
                                               float sig, float z, bool last) {
  float lam1 = __fsub_rn(__fmul_rn(s.c1a, gx), __fmul_rn(sig, s.c1b));
  float lam2 = __fsub_rn(__fmul_rn(__fmul_rn(__fmul_rn(gx, gx), s.one_m_a_sq), s.c2a),
                         __fmul_rn(__fmul_rn(sig, gx), s.c2b));
  float disc = __fsub_rn(__fmul_rn(lam1, lam1), __fmul_rn(__fmul_rn(4.0f, s.lam0), lam2));
  float sy0 = __fdiv_rn(__fadd_rn(-lam1, __fsqrt_rn(disc)), s.two_lam0);
  float noise = __fadd_rn(__fmul_rn(s.bb_m_bt, gx), __fmul_rn(s.bt, sy0));
  float y0 = __fmul_rn(s.inv_sqrt_abar,
                       __fsub_rn(__fsub_rn(y, __fmul_rn(s.one_m_sqrt_abar, yT)), __fmul_rn(eps, __fsqrt_rn(noise))));
  if (last) return y0;
  float S1 = __fadd_rn(__fmul_rn(s.one_m_a_sq, gx), __fmul_rn(s.a_one_m_a, sy0));
  float S2 = __fadd_rn(__fmul_rn(s.bbm_m_btm, gx), __fmul_rn(s.btm, sy0));
  float den = __fadd_rn(__fmul_rn(s.a, S2), S1);
  float g0 = __fdiv_rn(__fmul_rn(s.sqrt_abar_prev, S1), den);
  float g1 = __fdiv_rn(__fmul_rn(s.sqrt_a, S2), den);
  float g2 = __fdiv_rn(__fadd_rn(__fmul_rn(s.sqrt_a_am1, S2), __fmul_rn(s.one_m_sqrt_abar_prev, S1)), den);
  float m = __fadd_rn(__fadd_rn(__fmul_rn(g0, y0), __fmul_rn(g1, y)), __fmul_rn(g2, yT));
  return __fadd_rn(m, __fmul_rn(__fsqrt_rn(sig), z));
}

// TMDM per-step scalars (tmdm_diffusion_utils.py:70-88).
struct UpdTmStep { float g0, g1, g2, inv_sab, one_m_sab, s1m, sqrt_bhat; };

__device__ __forceinline__ UpdTmStep upd_tm_step(const float* __restrict__ sched, int T, int t) {
  UpdTmStep s;
  float a = sched[0 * T + t];
  float s1m = sched[1 * T + t];
  float s1m_prev = sched[1 * T + (t > 0 ? t - 1 : 0)];
  float sab = sqrtf(1.0f - s1m * s1m);
  float sab_prev = sqrtf(1.0f - s1m_prev * s1m_prev);
  float s1m2 = s1m * s1m;
  s.g0 = (1.0f - a) * sab_prev / s1m2;
  s.g1 = (s1m_prev * s1m_prev) * sqrtf(a) / s1m2;
  s.g2 = 1.0f + (sab - 1.0f) * (sqrtf(a) + sab_prev) / s1m2;
  s.inv_sab = 1.0f / sab;
  s.one_m_sab = 1.0f - sab;
  s.s1m = s1m;
  s.sqrt_bhat = sqrtf((s1m_prev * s1m_prev) / s1m2 * (1.0f - a));
  return s;
}

__device__ __forceinline__ float upd_tm_update(const UpdTmStep& s, float y, float yT, float eps, float z, bool last) {
  float y0 = __fmul_rn(s.inv_sab, __fsub_rn(__fsub_rn(y, __fmul_rn(s.one_m_sab, yT)), __fmul_rn(eps, s.s1m)));
  if (last) return y0;
  float m = __fadd_rn(__fadd_rn(__fmul_rn(s.g0, y0), __fmul_rn(s.g1, y)), __fmul_rn(s.g2, yT));
  return __fadd_rn(m, __fmul_rn(s.sqrt_bhat, z));
}
#endif  // __CUDACC__
