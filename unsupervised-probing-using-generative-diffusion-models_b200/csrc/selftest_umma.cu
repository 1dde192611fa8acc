// Known-answer test of the tcgen05 plumbing the fused sampler relies on (upd_selftest_umma): shared-memory
// matrix descriptors (K-major, SWIZZLE_NONE), instruction descriptors, A operand staged in TMEM with the
// sampler's fp16 / tf32 hi-lo encodings, three-pass accumulation, TMEM read-back.  D = A * B^T, one CTA.
#include "tc_helpers.cuh"
#include "upd_common.cuh"

namespace {

constexpr uint32_t UMMA_LBO = 2048;
constexpr uint32_t UMMA_SBO = 128;

struct __align__(8) TcSync {
  unsigned long long wbar;
  unsigned long long mma_bar[2];
  uint32_t tmem_base;
  uint32_t pad;
};

// ------------------------------------------------------------------------------------------------
// Known-answer kernel for the descriptor / operand encodings above: D = A * B^T, one CTA.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
selftest_umma_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ D, int K, int mode,
                     int flags) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ TcSync sync;
  unsigned char* bhi = smem;
  unsigned char* blo = smem + 32768;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { tc::mbar_init(tc::smem_u32(&sync.mma_bar[0]), 1); tc::fence_mbar_init(); }
  if (warp == 0) tc::tmem_alloc<256>(tc::smem_u32(&sync.tmem_base));
  // B -> shared, UMMA K-major no-swizzle core-matrix layout (same element map as the host packer)
  for (int idx = tid; idx < 128 * K; idx += 128) {
    int n = idx / K, k = idx % K;
    float v = Bm[n * K + k];
    if (mode == 0) {
      __half h = __float2half_rn(v);
      __half l = __float2half_rn(v - __half2float(h));
      size_t off = (size_t)(k / 8) * 2048 + (size_t)n * 16 + (size_t)(k % 8) * 2;
      *reinterpret_cast<__half*>(bhi + off) = h;
      *reinterpret_cast<__half*>(blo + off) = l;
    } else {
      float h = tc::to_tf32(v), l = tc::to_tf32(v - h);
      size_t off = (size_t)(k / 4) * 2048 + (size_t)n * 16 + (size_t)(k % 4) * 4;
      *reinterpret_cast<float*>(bhi + off) = h;
      *reinterpret_cast<float*>(blo + off) = l;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor-core reads
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = sync.tmem_base;
  const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
  const uint32_t abuf = tmem_base + lane_sel, dbuf = abuf + 128u;
  // A row of this thread -> TMEM with the sampler's operand encodings
  if (mode == 0) {
    for (int q = 0; q < 8; ++q) {          // 16-column groups: hi words [16q,16q+8), lo words [16q+8,16q+16)
      uint32_t o[16];
      for (int j = 0; j < 16; j += 2)
        tc::split_f16x2(A[tid * K + 16 * q + j], A[tid * K + 16 * q + j + 1], o[j / 2], o[8 + j / 2]);
      tc::tmem_st16(abuf + 16u * q, o);
    }
  } else {
    uint32_t a[32];
    for (int i = 0; i < 32; ++i) a[i] = 0u;
    // hi in columns [0,K), lo in [K,2K); K <= 16 fits one x32 store, K = 24/32 needs two
    for (int half = 0; half < (K > 16 ? 2 : 1); ++half) {
      for (int i = 0; i < 32; ++i) {
        int col = 32 * half + i;
        float v = 0.f;
        bool is_lo = col >= K;
        int k = is_lo ? col - K : col;
        if (k < K) {
          float x = A[tid * K + k];
          float hi = tc::to_tf32(x);
          v = is_lo ? tc::to_tf32(x - hi) : hi;
        }
        a[i] = __float_as_uint(v);
      }
      tc::tmem_st32(abuf + 32u * half, a);
    }
  }
  tc::wait_st();
  tc::fence_before_sync();
  __syncthreads();
  const uint32_t lbo = (flags & 1) ? UMMA_SBO : UMMA_LBO, sbo = (flags & 1) ? UMMA_LBO : UMMA_SBO;
  if (tid == 0) {
    tc::fence_after_sync();
    if (mode == 0) tc::issue_layer_f16x3_g16(tmem_base + 128u, tmem_base, tc::smem_u32(bhi), tc::smem_u32(blo), lbo, sbo);
    else tc::issue_layer_tf32x3(tmem_base + 128u, tmem_base, K, tc::smem_u32(bhi), tc::smem_u32(blo), lbo, sbo);
    tc::mma_commit(tc::smem_u32(&sync.mma_bar[0]));
  }
  tc::mbar_wait(tc::smem_u32(&sync.mma_bar[0]), 0);
  tc::fence_after_sync();
  for (int c = 0; c < 4; ++c) {
    uint32_t r[32];
    tc::tmem_ld32(dbuf + 32u * c, r);
    tc::wait_ld();
    for (int j = 0; j < 32; ++j) D[tid * 128 + 32 * c + j] = __uint_as_float(r[j]);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<256>(tmem_base);
}

}  // namespace

// Test-only entry point (libupd_selftest.so, built by _build.build_selftest(); NOT part of the product ABI):
// D[128,128] = A[128,K] * B[128,K]^T with A staged in TMEM and B in shared memory (mode 0: fp16 hi/lo 3-pass, K = 128;
// mode 1: tf32 hi/lo 3-pass, K = 8, 16, 24, 32); flags bit 0 swaps the descriptor's LBO/SBO (negative control).
// Returns 0, 2 (unsupported K) or 3 (CUDA error).
extern "C" int upd_selftest_umma(const float* a_dev, const float* b_dev, float* d_dev, int K, int mode, int flags,
                                 void* stream) {
  if (!a_dev || !b_dev || !d_dev) return 1;
  if (mode == 0 ? (K != 128) : (K < 8 || K > 32 || K % 8)) return 2;
  cudaError_t e = cudaFuncSetAttribute(selftest_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  if (e != cudaSuccess) return 3;
  selftest_umma_kernel<<<1, 128, 65536, (cudaStream_t)stream>>>(a_dev, b_dev, d_dev, K, mode, flags);
  return cudaGetLastError() == cudaSuccess ? 0 : 3;
}
