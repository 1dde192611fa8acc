// Pipe-rate microbenchmark for the sampler epilogue's instruction kinds (B200): cycles per warp instruction per SMSP.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../unsupervised-probing-using-generative-diffusion-models_b200/csrc/sampler_math.cuh"
#define N_ITER 2000
template <int KIND>
__global__ void __launch_bounds__(512, 1) bench(float* out, long long* cycles, float seed) {
  float2 a[8];
  for (int i = 0; i < 8; ++i) a[i] = make_float2(seed + i * 0.01f + threadIdx.x * 1e-4f, seed * 0.5f + i * 0.02f);
  const float2 k1 = make_float2(0.999f, 1.001f), k2 = make_float2(1e-3f, -1e-3f);
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < N_ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (KIND == 0) { a[i].x = sm::ex2(a[i].x); a[i].y = sm::ex2(a[i].y); }                       // 2 MUFU
      if (KIND == 1) { a[i].x = fmaf(a[i].x, k1.x, k2.x); a[i].y = fmaf(a[i].y, k1.y, k2.y); }     // 2 FFMA
      if (KIND == 2) { a[i] = sm::ffma2(a[i], k1, k2); a[i] = sm::ffma2(a[i], k1, k2); }           // 2 FFMA2
      if (KIND == 3) { a[i] = sm::fmul2(a[i], k1); a[i] = sm::fmul2(a[i], k1); }                   // 2 FMUL2
      if (KIND == 4) { a[i] = sm::fadd2(a[i], k2); a[i] = sm::fadd2(a[i], k2); }                   // 2 FADD2
      if (KIND == 5) { uint32_t h, l; sm::split_f16x2(a[i].x, a[i].y, h, l); a[i].x = __uint_as_float(h & 0x3fffffffu); a[i].y = __uint_as_float(l & 0x3fffffffu); }  // 2 F2FP + 2 FHADD + 2 LOP
      if (KIND == 6) { a[i].x = fminf(a[i].x, k1.x); a[i].y = fmaxf(a[i].y, k2.y); }               // 2 FMNMX
      if (KIND == 7) { a[i].x = sm::ex2(a[i].x); a[i].y = sm::ex2(a[i].y); a[i] = sm::ffma2(a[i], k1, k2); a[i] = sm::ffma2(a[i], k1, k2); a[i] = sm::ffma2(a[i], k1, k2); a[i] = sm::ffma2(a[i], k1, k2);}  // 2 MUFU + 4 FFMA2
      if (KIND == 8) { a[i].x = sm::ex2(a[i].x); a[i].y = sm::ex2(a[i].y); 
                       for (int r = 0; r < 8; ++r) a[i] = sm::ffma2(a[i], k1, k2); }  // 2 MUFU + 8 FFMA2
      if (KIND == 9) { a[i].x = sm::ex2(a[i].x); a[i].y = sm::ex2(a[i].y);
                       for (int r = 0; r < 6; ++r) { a[i].x = fmaf(a[i].x, k1.x, k2.x); a[i].y = fmaf(a[i].y, k1.y, k2.y);} }  // 2 MUFU + 12 FFMA
      if (KIND == 10) { float2 z = a[i]; float2 h = sm::softplus2<false, false>(z); a[i] = sm::ffma2(h, k1, k2); }  // softplus mufu
      if (KIND == 11) { float2 z = a[i]; float2 h = sm::softplus2<true, false>(z); a[i] = sm::ffma2(h, k1, k2); }   // softplus poly
      if (KIND == 12) { a[i].x = sm::ex2(a[i].x); a[i].y = sm::ex2(a[i].y);
                       for (int r = 0; r < 12; ++r) a[i] = sm::ffma2(a[i], k1, k2); }  // 2 MUFU + 12 FFMA2
    }
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template <int KIND> void run(const char* name, int instr_per_body, int threads) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  bench<KIND><<<148, threads>>>(out, cyc, 0.3f);
  cudaDeviceSynchronize();
  bench<KIND><<<148, threads>>>(out, cyc, 0.3f);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  double warps_per_smsp = threads / 32 / 4.0;
  printf("%-28s threads %4d: %.0f cycles; %.3f clk per warp-instr per SMSP (%d instr/body)\n", name, threads, c,
         c / (double(N_ITER) * 8 * instr_per_body * warps_per_smsp), instr_per_body);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int threads : {256, 512}) {
    run<0>("MUFU.EX2 x2", 2, threads);
    run<1>("FFMA x2", 2, threads);
    run<2>("FFMA2 x2", 2, threads);
    run<3>("FMUL2 x2", 2, threads);
    run<4>("FADD2 x2", 2, threads);
    run<5>("split (2 F2FP+2 FHADD+2 LOP)", 6, threads);
    run<6>("FMNMX x2", 2, threads);
    run<7>("2 MUFU + 4 FFMA2", 6, threads);
    run<8>("2 MUFU + 8 FFMA2", 10, threads);
    run<12>("2 MUFU + 12 FFMA2", 14, threads);
    run<9>("2 MUFU + 12 FFMA", 14, threads);
    run<10>("softplus2 mufu (+1 FFMA2)", 6, threads);
    run<11>("softplus2 poly (+1 FFMA2)", 14, threads);
  }
  return 0;
}
