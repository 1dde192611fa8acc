// Arithmetic shared by the fused tcgen05 samplers' epilogues: packed fp32x2 math (sm_100 FFMA2 / FMUL2 / FADD2: two
// fp32 lanes per issue slot), base-2 softplus with two or with one MUFU op, the fp16 hi/lo operand split, and the
// polynomial softplus of the sigma head.  Everything here runs once per hidden element per reverse step -- 384 times
// per denoiser row-step -- so it is written against the issue-slot budget: the sampler is bounded by the MUFU pipe
// (4 lanes/clk/SMSP: a warp-wide ex2 or lg2 occupies it for 8 clk) and by the one instruction per clock an SMSP can
// issue.  Round 1 spent 16.2 warp instructions per hidden element; the forms below need 6-7.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

#include <utility>

#include "tc_helpers.cuh"

namespace sm {

// compile-time loop: f(std::integral_constant<int, 0>) ... f(std::integral_constant<int, N-1>)
template <class Fn, int... Is>
__device__ __forceinline__ void static_for_impl(Fn&& f, std::integer_sequence<int, Is...>) {
  (f(std::integral_constant<int, Is>{}), ...);
}
template <int N, class Fn>
__device__ __forceinline__ void static_for(Fn&& f) { static_for_impl(f, std::make_integer_sequence<int, N>{}); }

// ---- packed fp32 pairs (PTX ISA 8.6, sm_100+) ---------------------------------------------------
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{ .reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2,%3};\n\tmov.b64 rb, {%4,%5};\n\tmov.b64 rc, {%6,%7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0,%1}, rd; }"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("{ .reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2,%3};\n\tmov.b64 rb, {%4,%5};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0,%1}, rd; }"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("{ .reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2,%3};\n\tmov.b64 rb, {%4,%5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0,%1}, rd; }"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 splat(float v) { return make_float2(v, v); }

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- base-2 softplus of a pair: L = lg2(1 + 2^z) = softplus(z ln2)/ln2 ---------------------------
// MUFU form: ex2, +1, lg2 (two MUFU ops per element).  GUARD: z is not bounded by construction (layer 1 reads the
// trajectory state; TMDM has no normalisation): ex2 would overflow beyond 2^127, so the argument is clamped and the
// result is max(z, .) -- exact in the clamped range too, because lg2(1 + 2^z) = z to fp32 precision for z > 25.
// Both forms come in two stages so that the epilogues can software-pipeline them (stage A of the next columns is
// issued with stage B of the current ones):  A: z' -> w (ex2 done),  B: (w, z') -> L.
template <bool GUARD>
__device__ __forceinline__ float2 softplus2_mufu_a(float2 z) {
  float2 u;
  u.x = ex2(GUARD ? fminf(z.x, 126.f) : z.x);
  u.y = ex2(GUARD ? fminf(z.y, 126.f) : z.y);
  return fadd2(u, splat(1.0f));
}
template <bool GUARD>
__device__ __forceinline__ float2 softplus2_mufu_b(float2 w, float2 z) {
  float2 l = make_float2(lg2(w.x), lg2(w.y));
  if (GUARD) { l.x = fmaxf(l.x, z.x); l.y = fmaxf(l.y, z.y); }
  return l;
}

// One-MUFU form: L = max(z,0) + lg2(1 + u), u = 2^-|z| in (0,1] (never overflows, no guard needed), and lg2(1+u) as a
// minimax polynomial on the FMA pipe in packed Horner form.  It trades one MUFU op (8 clk of the SMSP's MUFU pipe)
// for ~9 packed FMA-pipe instructions per pair (measured 2.1-2.4 clk each, profiles/r02_pipe_rates.txt); the epilogues
// mix the two forms pair by pair (PMASK) so that the MUFU pipe and the FMA pipe run out together.  |error| of the
// polynomial with these fp32 coefficients: degree 7: 3.0e-7 in exact arithmetic, 3.9e-7 as the fp32 FMA chain below;
// degree 8: 0.9e-7 / 1.9e-7 (tests/test_host_cpu.py evaluates both from this source; lg2.approx itself is good to ~1e-7
// on these arguments, and 1 + u rounds u to 6e-8 before it).  Degree 7 is
// the default: full-size oracle parity is the same to three digits with either (tests/test_gpu_parity_full.py: NsDiff
// 3.0-5.5e-6 of rms, TMDM 1.24e-5, MPV 1e-7..5e-7) and it is 1 % (NsDiff) / 2.6 % (TMDM) faster.
#ifndef UPD_LG2_DEG
#define UPD_LG2_DEG 7
#endif
__device__ __forceinline__ float2 softplus2_poly_a(float2 z) {
  return make_float2(ex2(-fabsf(z.x)), ex2(-fabsf(z.y)));
}
__device__ __forceinline__ float2 softplus2_poly_b(float2 u, float2 z) {
#if UPD_LG2_DEG == 8
  float2 p = ffma2(splat(-9.0886848e-03f), u, splat(5.1133554e-02f));
  p = ffma2(p, u, splat(-1.3592552e-01f));
  p = ffma2(p, u, splat(2.4040906e-01f));
  p = ffma2(p, u, splat(-3.4654802e-01f));
  p = ffma2(p, u, splat(4.7846410e-01f));
  p = ffma2(p, u, splat(-7.2113216e-01f));
  p = ffma2(p, u, splat(1.4426876e+00f));
  p = ffma2(p, u, splat(4.2320107e-08f));
#else
  float2 p = ffma2(splat(1.5124998e-02f), u, splat(-7.806088e-02f));
  p = ffma2(p, u, splat(1.9208784e-01f));
  p = ffma2(p, u, splat(-3.2425278e-01f));
  p = ffma2(p, u, splat(4.7289753e-01f));
  p = ffma2(p, u, splat(-7.2045296e-01f));
  p = ffma2(p, u, splat(1.4426563e+00f));
  p = ffma2(p, u, splat(2.7732642e-07f));
#endif
  return fadd2(make_float2(fmaxf(z.x, 0.f), fmaxf(z.y, 0.f)), p);
}

// unstaged forms (microbenchmarks, scratch/)
template <bool POLY, bool GUARD>
__device__ __forceinline__ float2 softplus2(float2 z) {
  if constexpr (POLY) return softplus2_poly_b(softplus2_poly_a(z), z);
  else return softplus2_mufu_b<GUARD>(softplus2_mufu_a<GUARD>(z), z);
}

// ---- fp16 hi/lo split of two consecutive K elements ---------------------------------------------
// hi = fp16(a), lo = fp16(a - hi): 22 mantissa bits between them.  Four instructions per pair: F2FP (pack hi),
// two FHADD (sm_100 mixed-precision add: fp32 + (-fp16) in one instruction, no unpack), F2FP (pack lo).
// Element 2c goes to bits [0,16), element 2c+1 to bits [16,32), as the A operand wants them in a TMEM column.
__device__ __forceinline__ void split_f16x2(float a, float b, uint32_t& hi, uint32_t& lo) { tc::split_f16x2(a, b, hi, lo); }

// ---- sigma head: softplus on [0,1] without MUFU ---------------------------------------------------
// The sigma head applies softplus to the L2-normalised, non-negative hidden vector hn (components in [0,1]):
//   softplus(x) = x/2 + ln2 + u/8 + C2 u^2 + C3 u^3 + r(u),  u = x^2,  |r(u)| <= 2.9e-6 u^2
// (the first three terms are the exact series; C2, C3 are a minimax fit of the remainder relative to u^2, so the error
// of a row is bounded by 2.9e-6 * sum_j |ws_j| u_j^2 with sum_j u_j = 1: ~1e-8 for a spread-out hidden vector, 2.9e-6
// |ws_j| in the one-hot worst case).  Because a polynomial of hn = L/||L|| is a sum of power sums of L, the layer-3
// epilogue accumulates M_k = sum_j ws_j L_j^(2k), k = 1..3, and rescales once ||L|| is known: no second pass.
constexpr float SPH_LN2 = 0.6931471805599453f, SPH_C1 = 0.125f, SPH_C2 = -0.005205402f, SPH_C3 = 0.00032283954f;

}  // namespace sm
