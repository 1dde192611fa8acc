// Thin inline-PTX layer over tcgen05 / TMEM / mbarrier / bulk copy for sm_100a.
// Everything the fused sampler and the descriptor self-test issue goes through these wrappers,
// so the self-test (upd_selftest_umma) validates exactly the encodings the sampler relies on.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  // Bounded spin: a descriptor or protocol bug must surface as a trapped launch (cudaErrorLaunchFailure
  // -> UPD_ERR_CUDA), never as a hung GPU.  try_wait itself sleeps in hardware for a bounded time.
  uint32_t done, spins = 0;
  do {
    if (++spins > (1u << 26)) __trap();
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x2000;\n\t"   // suspend-time hint: sleep in hardware, do not spin on the issue port
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}

// Plain arrive (release at CTA scope): one count towards the current phase.
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// ---- 1-D bulk copy global -> shared (TMA engine, SASS UBLKCP), completes on an mbarrier --------
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src_gmem, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_smem), "l"(src_gmem), "r"(bytes), "r"(bar) : "memory");
}

// ---- TMEM allocation (one warp, all 32 lanes) ---------------------------------------------------
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// ---- tcgen05 fences / waits ---------------------------------------------------------------------
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Named barrier over `n` threads (ids 1.. ; id 0 is __syncthreads).
__device__ __forceinline__ void named_bar_sync(int id, int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}

// Producer side of a named barrier: signal without waiting (consumers use named_bar_sync with the same count).
__device__ __forceinline__ void named_bar_arrive(int id, int n) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
}

// ---- descriptors --------------------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major, SWIZZLE_NONE ("interleave"): core matrix = 8 rows x 16 B
// stored contiguously (128 B); LBO = byte distance between core matrices adjacent in K,
// SBO = byte distance between core matrices adjacent in M/N (cute/arch/mma_sm100_desc.hpp
// SmemDescriptor; canonical layout ((8,n),2):((1,SBO),LBO) in 16-byte units).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version for sm_100
  return d;         // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}

// Instruction descriptor (UMMA::InstrDescriptor): D fp32, A/B K-major, M=128, N=128.
constexpr uint32_t IDESC_F16_M128_N128 = (1u << 4) | (0u << 7) | (0u << 10) | (16u << 17) | (8u << 24);
constexpr uint32_t IDESC_TF32_M128_N128 = (1u << 4) | (2u << 7) | (2u << 10) | (16u << 17) | (8u << 24);

// ---- MMA issue (one thread), A from TMEM, B from shared memory ----------------------------------
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// Arrive on an mbarrier when all previously issued MMAs of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- TMEM <-> registers, shape 32x32b: lane i of the warp <-> TMEM lane (32*(warp%4) + i) -------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// width-generic forms (N = 8, 16, 32 columns of this thread's TMEM lane)
template <int N> __device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&r)[N]) {
  static_assert(N == 8 || N == 16 || N == 32, "TMEM 32x32b shapes used here: x8, x16, x32");
  if constexpr (N == 8) tmem_ld8(taddr, r);
  else if constexpr (N == 16) tmem_ld16(taddr, r);
  else tmem_ld32(taddr, r);
}
template <int N> __device__ __forceinline__ void tmem_st(uint32_t taddr, const uint32_t (&r)[N]) {
  static_assert(N == 8 || N == 16 || N == 32, "TMEM 32x32b shapes used here: x8, x16, x32");
  if constexpr (N == 8) tmem_st8(taddr, r);
  else if constexpr (N == 16) tmem_st16(taddr, r);
  else tmem_st32(taddr, r);
}

// ---- operand encodings --------------------------------------------------------------------------
// fp16 hi/lo split of two consecutive K elements, packed as the A operand wants them in a TMEM
// column: element 2c in bits [0,16), element 2c+1 in bits [16,32).
__device__ __forceinline__ void split_f16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  // four instructions per pair: F2FP (pack hi), two FHADD (sm_100 mixed-precision add: fp32 + (-fp16), no unpack),
  // F2FP (pack lo) -- the unpack-and-subtract form (HADD2.F32 + FADD per element) needs six
  uint32_t h;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(b), "f"(a));
  float la, lb;
  asm("{ .reg .b16 h0, h1, n0, n1;\n\t"
      "mov.b32 {h0,h1}, %2;\n\tneg.f16 n0, h0;\n\tneg.f16 n1, h1;\n\t"
      "add.rn.f32.f16 %0, n0, %3;\n\tadd.rn.f32.f16 %1, n1, %4; }"
      : "=f"(la), "=f"(lb) : "r"(h), "f"(a), "f"(b));
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(lb), "f"(la));
  hi = h;
}
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// Issue one 128x128x128 contraction as three passes (hi*hi + lo*hi + hi*lo) of eight K-slices.
// fp16: slice = 16 elements = 8 TMEM columns of A and two 2048-byte-apart core-matrix columns of B.
// A layout (in place over the previous accumulator, 16 columns at a time): K-slice j = fp16 hi words in
// columns [16j,16j+8), lo words in [16j+8,16j+16).
__device__ __forceinline__ void issue_layer_f16x3_g16(uint32_t d_tmem, uint32_t a_tmem, uint32_t bhi_smem,
                                                      uint32_t blo_smem, uint32_t lbo, uint32_t sbo) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint32_t a_hi = a_tmem + 16u * j;
    uint32_t a_lo = a_hi + 8u;
    uint64_t b_hi = smem_desc(bhi_smem + (uint32_t)(2 * j) * 2048u, lbo, sbo);
    uint64_t b_lo = smem_desc(blo_smem + (uint32_t)(2 * j) * 2048u, lbo, sbo);
    mma_f16_ts(d_tmem, a_lo, b_hi, IDESC_F16_M128_N128, j > 0);
    mma_f16_ts(d_tmem, a_hi, b_lo, IDESC_F16_M128_N128, true);
    mma_f16_ts(d_tmem, a_hi, b_hi, IDESC_F16_M128_N128, true);
  }
}
// The same contraction restricted to the even (parity 0) or odd (parity 1) K-slices: the warp-specialised sampler issues
// the even slices of a layer as soon as every epilogue warp has written the first 16 of its 32 columns, so that half of
// the layer's tensor time is spent while the epilogue of the other 16 columns is still running.
// A_LO / B_LO: which correction passes are issued besides hi*hi (A_LO: lo(A)*hi(B), B_LO: hi(A)*lo(B)).
template <bool A_LO = true, bool B_LO = true>
__device__ __forceinline__ void issue_kparity_f16x3_g16(uint32_t d_tmem, uint32_t a_tmem, uint32_t bhi_smem,
                                                        uint32_t blo_smem, uint32_t lbo, uint32_t sbo, int parity) {
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    const int j = 2 * jj + parity;
    uint32_t a_hi = a_tmem + 16u * j;
    uint32_t a_lo = a_hi + 8u;
    uint64_t b_hi = smem_desc(bhi_smem + (uint32_t)(2 * j) * 2048u, lbo, sbo);
    uint64_t b_lo = smem_desc(blo_smem + (uint32_t)(2 * j) * 2048u, lbo, sbo);
    const bool first = (parity == 0 && jj == 0);
    if (A_LO) mma_f16_ts(d_tmem, a_lo, b_hi, IDESC_F16_M128_N128, !first);
    if (B_LO) mma_f16_ts(d_tmem, a_hi, b_lo, IDESC_F16_M128_N128, !first || A_LO);
    mma_f16_ts(d_tmem, a_hi, b_hi, IDESC_F16_M128_N128, !first || A_LO || B_LO);
  }
}
//   tf32: slice = 8 elements = 8 TMEM columns; A holds hi in columns [0,K1), lo in [K1,2*K1).
__device__ __forceinline__ void issue_layer_tf32x3(uint32_t d_tmem, uint32_t a_tmem, int K1, uint32_t bhi_smem,
                                                   uint32_t blo_smem, uint32_t lbo, uint32_t sbo) {
  for (int j = 0; j < K1 / 8; ++j) {
    uint32_t a_hi = a_tmem + 8u * j;
    uint32_t a_lo = a_hi + (uint32_t)K1;
    uint64_t b_hi = smem_desc(bhi_smem + (uint32_t)(2 * j) * 2048u, lbo, sbo);
    uint64_t b_lo = smem_desc(blo_smem + (uint32_t)(2 * j) * 2048u, lbo, sbo);
    mma_tf32_ts(d_tmem, a_lo, b_hi, IDESC_TF32_M128_N128, j > 0);
    mma_tf32_ts(d_tmem, a_hi, b_lo, IDESC_TF32_M128_N128, true);
    mma_tf32_ts(d_tmem, a_hi, b_hi, IDESC_TF32_M128_N128, true);
  }
}

}  // namespace tc
