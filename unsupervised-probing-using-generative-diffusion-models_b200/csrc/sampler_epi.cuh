// Pieces of the layer epilogues shared by the tcgen05 samplers (sampler_tc.cu: two / three tiles with MUFU turns or TMEM
// rotation; sampler_ws.cu: warp-specialised): the pre-activation of a column pair, the head sums of layer 3 and the
// algebra that turns them into (eps, sigma) once the row norm is known.
#pragma once
#include "sampler_math.cuh"
#include "upd_common.cuh"

namespace epi {

constexpr uint32_t UMMA_LBO = 2048;   // K-adjacent core matrices (layout in upd_common.cuh)
constexpr uint32_t UMMA_SBO = 128;    // N-adjacent core matrices
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

// z' of a pair of columns.  FIRST: layer 1 (bias rides in the GEMM).
template <bool FIRST>
__device__ __forceinline__ float2 preact2(uint32_t a0, uint32_t a1, float2 inv2, float2 b, float2 e) {
  const float2 acc = make_float2(__uint_as_float(a0), __uint_as_float(a1));
  // (acc*inv + b) as one FMA: the pre-activation differs from the reference's two roundings by < 1 ulp, far below the
  // reordering of the 128-term sums it comes from
  return FIRST ? sm::fmul2(acc, e) : sm::fmul2(sm::ffma2(acc, inv2, b), e);
}

// Head sums of one warp's share of the 128 hidden columns (sampler_math.cuh): pe = sum w4 L, and for NsDiff
// pb = sum ws L, m_k = sum ws L^(2k), ss = sum L^2 -- all as packed pairs (even / odd columns), folded at the end.
template <bool NS, int F>
struct HeadSums {
  float2 ss, pe[F], pb[F], m1[F], m2[F], m3[F];
  __device__ __forceinline__ void clear() {
    ss = make_float2(0.f, 0.f);
#pragma unroll
    for (int f = 0; f < F; ++f) {
      pe[f] = make_float2(0.f, 0.f);
      if (NS) { pb[f] = make_float2(0.f, 0.f); m1[f] = pb[f]; m2[f] = pb[f]; m3[f] = pb[f]; }
    }
  }
  __device__ __forceinline__ void add(float2 h, const float* __restrict__ w4, const float* __restrict__ ws) {
    // w4 / ws point at this pair's two columns of feature 0; features are 128 floats apart
    if (NS) {
      const float2 u = sm::fmul2(h, h);
      ss = sm::fadd2(ss, u);
      const float2 u2 = sm::fmul2(u, u), u3 = sm::fmul2(u2, u);
#pragma unroll
      for (int f = 0; f < F; ++f) {
        const float2 a = *reinterpret_cast<const float2*>(w4 + f * 128);
        const float2 s = *reinterpret_cast<const float2*>(ws + f * 128);
        pe[f] = sm::ffma2(a, h, pe[f]);
        pb[f] = sm::ffma2(s, h, pb[f]);
        m1[f] = sm::ffma2(s, u, m1[f]);
        m2[f] = sm::ffma2(s, u2, m2[f]);
        m3[f] = sm::ffma2(s, u3, m3[f]);
      }
    } else {
#pragma unroll
      for (int f = 0; f < F; ++f) pe[f] = sm::ffma2(*reinterpret_cast<const float2*>(w4 + f * 128), h, pe[f]);
    }
  }
};

// One 16-column group of an accumulator -> activations -> fp16 hi/lo A-operand words of one K-slice of the next GEMM
// (hi words in o[0..7], lo words in o[8..15]).  FIRST: layer 1 (bias rides in the GEMM).  PMASK: bit i = column pair i
// takes the one-MUFU (polynomial) softplus.  LO = false: only the hi words are formed (o[8..15] untouched) -- the
// warp-specialised sampler's two-pass contraction (activation as one fp16 word, weights hi + lo).
template <bool FIRST, bool GUARD, bool SUMSQ, int PMASK, bool LO = true>
__device__ __forceinline__ void epilogue_group(const uint32_t* __restrict__ r, uint32_t (&o)[16], const float* __restrict__ e,
                                               const float* __restrict__ b, float2 inv2, float2& ss2) {
  sm::static_for<4>([&](auto jj) {
    constexpr int j = 4 * decltype(jj)::value;
    const float4 e4 = *reinterpret_cast<const float4*>(e + j);
    const float4 b4 = FIRST ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(b + j);
    const float2 z0 = preact2<FIRST>(r[j], r[j + 1], inv2, make_float2(b4.x, b4.y), make_float2(e4.x, e4.y));
    const float2 z1 = preact2<FIRST>(r[j + 2], r[j + 3], inv2, make_float2(b4.z, b4.w), make_float2(e4.z, e4.w));
    const float2 h0 = sm::softplus2<((PMASK >> (j / 2)) & 1) != 0, GUARD>(z0);
    const float2 h1 = sm::softplus2<((PMASK >> (j / 2 + 1)) & 1) != 0, GUARD>(z1);
    if (SUMSQ) { ss2 = sm::ffma2(h0, h0, ss2); ss2 = sm::ffma2(h1, h1, ss2); }
    if (LO) {
      sm::split_f16x2(h0.x, h0.y, o[j / 2], o[8 + j / 2]);
      sm::split_f16x2(h1.x, h1.y, o[j / 2 + 1], o[8 + j / 2 + 1]);
    } else {
      asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(o[j / 2]) : "f"(h0.y), "f"(h0.x));
      asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(o[j / 2 + 1]) : "f"(h1.y), "f"(h1.x));
    }
  });
}

// One 16-column group of the layer-3 accumulator -> head sums (nothing is written back).  e, b, w4, ws point at the
// group's first column.
template <bool NS, int F, bool GUARD, int PMASK>
__device__ __forceinline__ void heads_group(const uint32_t* __restrict__ r, const float* __restrict__ e,
                                            const float* __restrict__ b, const float* __restrict__ w4,
                                            const float* __restrict__ ws, float2 inv2, HeadSums<NS, F>& H) {
  sm::static_for<4>([&](auto jj) {
    constexpr int j = 4 * decltype(jj)::value;
    const float4 e4 = *reinterpret_cast<const float4*>(e + j);
    const float4 b4 = *reinterpret_cast<const float4*>(b + j);
    const float2 z0 = preact2<false>(r[j], r[j + 1], inv2, make_float2(b4.x, b4.y), make_float2(e4.x, e4.y));
    const float2 z1 = preact2<false>(r[j + 2], r[j + 3], inv2, make_float2(b4.z, b4.w), make_float2(e4.z, e4.w));
    const float2 h0 = sm::softplus2<((PMASK >> (j / 2)) & 1) != 0, GUARD>(z0);
    const float2 h1 = sm::softplus2<((PMASK >> (j / 2 + 1)) & 1) != 0, GUARD>(z1);
    H.add(h0, w4 + j, ws + j);
    H.add(h1, w4 + j + 2, ws + j + 2);
  });
}

// NsDiff heads from the row totals of the sums above (denoise.py:50): hn = L/||L||,
//   eps   = lin4(hn)                              = inv3 * PE + b4
//   sigma = softplus(sigma_lin(softplus(hn)))     , sigma_lin(softplus(hn)) =
//           inv3/2 * PB + ln2 * sum_j ws_j + inv3^2/8 * M1 + C2 inv3^4 * M2 + C3 inv3^6 * M3 + bs
// ws_sum = ln2 * sum_j ws_j.  The outer softplus needs relative accuracy (its value is ~1e-4): upd_softplus_accurate.
__device__ __forceinline__ void ns_heads(float ss, float pe, float pb, float m1, float m2, float m3, float ws_sum, float b4,
                                         float bs, float& eps, float& sig) {
  const float inv3 = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
  const float i2 = inv3 * inv3, i4 = i2 * i2;
  eps = pe * inv3 + b4;
  const float lin = 0.5f * inv3 * pb;
  float poly = fmaf(i2, sm::SPH_C1 * m1, ws_sum);
  poly = fmaf(i4, sm::SPH_C2 * m2, poly);
  poly = fmaf(i4 * i2, sm::SPH_C3 * m3, poly);
  sig = upd_softplus_accurate(lin + poly + bs);
}


// ---- short-chain forms for the warp-specialised sampler's row warps ----------------------------------------------------
// The posterior update of a tile sits between its head sums and its next layer-1 GEMM: a serial chain of ~200 dependent
// instructions in upd_ns_update + ns_heads (IEEE divisions, log1pf(expf())), ~1500 clk during which the tile's epilogue
// warps have at most one phase of the other tile to work on.  The forms below compute the same quantities with everything
// that does not depend on the heads hoisted out (NsRowPre, filled while the row warp waits), one reciprocal instead of
// three divisions, a multiply by the per-step 1/(2 lambda_0), and the series of log1p for the sigma head's tiny argument.
// They differ from the reference's operation order by a few ulp -- two orders of magnitude below what the split-operand
// GEMMs and the polynomial head already contribute (1e-6), and the parity tests hold them to the same 2e-5 x rms.
struct NsRowPre { float c1a_gx, l2a, gx_c2b, n0, s1a, s2a, yy, sa_y, yT_c; };

__device__ __forceinline__ NsRowPre ns_row_pre(const UpdNsStep& s, float y, float yT, float gx) {
  NsRowPre q;
  q.c1a_gx = s.c1a * gx;
  q.l2a = ((gx * gx) * s.one_m_a_sq) * s.c2a;
  q.gx_c2b = gx * s.c2b;
  q.n0 = s.bb_m_bt * gx;
  q.s1a = s.one_m_a_sq * gx;
  q.s2a = s.bbm_m_btm * gx;
  q.yy = y - s.one_m_sqrt_abar * yT;
  q.sa_y = s.sqrt_a * y;
  q.yT_c = yT;
  return q;
}

// Square roots, the reciprocal and the exponential of the row warps' chain WITHOUT the MUFU pipe: the epilogue warps keep
// the SMSP's MUFU queue full, so every MUFU op in the posterior chain waits behind ~a dozen queued ex2/lg2 (clock64 stamps:
// 2000-2600 clk for a chain of ~130 instructions).  Bit-trick seeds + Newton steps on the FMA pipe cost more instructions
// but no queueing; all results are good to ~1e-7 relative.
__device__ __forceinline__ float rsqrt_fma(float x) {            // x > 0, normal
  float y = __uint_as_float(0x5f3759dfu - (__float_as_uint(x) >> 1));
  const float hx = 0.5f * x;
  y = y * fmaf(-hx * y, y, 1.5f);
  y = y * fmaf(-hx * y, y, 1.5f);
  y = y * fmaf(-hx * y, y, 1.5f);
  return y * fmaf(-hx * y, y, 1.5f);
}
__device__ __forceinline__ float sqrt_fma(float x) {             // x >= 0
  const float xs = fmaxf(x, 1e-30f);
  return xs * rsqrt_fma(xs);
}
__device__ __forceinline__ float rcp_fma(float d) {              // d > 0, normal
  float y = __uint_as_float(0x7EF311C7u - __float_as_uint(d));
  y = y * fmaf(-d, y, 2.0f);
  y = y * fmaf(-d, y, 2.0f);
  y = y * fmaf(-d, y, 2.0f);
  return y * fmaf(-d, y, 2.0f);
}
__device__ __forceinline__ float exp_fma(float x) {              // -80 < x < 80
  const float t = fmaf(x, 1.4426950408889634f, 12582912.0f);      // round(x log2e) in the low mantissa bits
  const float n = t - 12582912.0f;
  float f = fmaf(n, -0.693145751953125f, x);                      // Cody-Waite: ln2 = hi + lo
  f = fmaf(n, -1.428606765330187e-06f, f);
  float p = fmaf(f, 1.9841270e-04f, 1.3888889e-03f);              // Taylor degree 7 on |f| <= 0.347: error 5e-9
  p = fmaf(p, f, 8.3333333e-03f);
  p = fmaf(p, f, 4.1666668e-02f);
  p = fmaf(p, f, 1.6666667e-01f);
  p = fmaf(p, f, 0.5f);
  p = fmaf(p, f, 1.0f);
  p = fmaf(p, f, 1.0f);
  return __uint_as_float(__float_as_uint(p) + ((__float_as_uint(t) - 0x4B400000u) << 23));
}

// softplus with full relative accuracy for the sigma head: log1p(e^x) by its alternating series when e^x < 0.0184
// (x < -4: relative truncation error u^4/5 < 3e-8; trained sigma heads sit near x = -9), libm otherwise.
__device__ __forceinline__ float softplus_rel(float x) {
  if (x < -4.0f) {
    const float u = exp_fma(fmaxf(x, -80.0f));
    return u * fmaf(u, fmaf(u, fmaf(u, -0.25f, 0.33333334f), -0.5f), 1.0f);
  }
  return x > 20.0f ? x : log1pf(expf(x));
}

__device__ __forceinline__ void ns_heads_fast(float ss, float pe, float pb, float m1, float m2, float m3, float ws_sum,
                                              float b4, float bs, float& eps, float& sig) {
  const float inv3 = rsqrt_fma(fmaxf(ss, 1e-24f));
  const float i2 = inv3 * inv3, i4 = i2 * i2;
  eps = fmaf(pe, inv3, b4);
  float poly = fmaf(i2, sm::SPH_C1 * m1, ws_sum + bs);
  poly = fmaf(i4, sm::SPH_C2 * m2, poly);
  poly = fmaf(i4 * i2, sm::SPH_C3 * m3, poly);
  sig = softplus_rel(fmaf(0.5f * inv3, pb, poly));
}

// p_sample / p_sample_t_1to0 (nsdiff_utils.py:139-157, :225-238) from the hoisted terms.
__device__ __forceinline__ float ns_update_fast(const UpdNsStep& s, const NsRowPre& q, float inv_two_lam0, float eps,
                                                float sig, float z, bool last) {
  const float lam1 = q.c1a_gx - sig * s.c1b;
  const float lam2 = q.l2a - sig * q.gx_c2b;
  const float disc = lam1 * lam1 - (4.0f * s.lam0) * lam2;
  const float sy0 = (sqrt_fma(disc) - lam1) * inv_two_lam0;
  const float noise = q.n0 + s.bt * sy0;
  const float y0 = s.inv_sqrt_abar * (q.yy - eps * sqrt_fma(noise));
  if (last) return y0;
  const float S1 = q.s1a + s.a_one_m_a * sy0;
  const float S2 = q.s2a + s.btm * sy0;
  const float r = rcp_fma(s.a * S2 + S1);
  const float num = (s.sqrt_abar_prev * S1) * y0 + S2 * q.sa_y + (s.sqrt_a_am1 * S2 + s.one_m_sqrt_abar_prev * S1) * q.yT_c;
  return num * r + sqrt_fma(sig) * z;
}

}  // namespace epi
