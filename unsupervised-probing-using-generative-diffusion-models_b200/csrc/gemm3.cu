// upd_gemm3 -- the dense layer of the condition encoders and of the transformer / graph denoisers on tcgen05.
//
//   C[M, n_out] (fp32)  (+)=  A3[M, Kp] (fp16, row-major)  x  W3[Nw, Kp]^T (fp16, row-major)
//
// A3 / W3 are the error-compensated split operands every dense layer of this package is phrased in (fx_encoder.py):
// [x_hi | x_lo | x_hi | 1 1 0..] against [W_hi | W_hi | W_lo | b_hi b_lo 0..], Kp = 3K + 8, so ONE fp16 tensor-core GEMM
// with fp32 accumulation yields x W^T + b to ~3e-6 (22 mantissa bits per factor, bias inside the contraction).  Round 1
// ran it as a cuBLAS call; this is the same contraction as a warp-specialised persistent sm_100a kernel:
//
//   * tiles of 128 rows x BN columns (BN = 64 / 128 / 256 by n_out), K in blocks of 64 halves (one 128-byte swizzle span);
//   * warp 0: TMA producer -- 2-D tensor maps (SWIZZLE_128B) for A3 and W3, a ring of STAGES shared-memory slots guarded
//     by full/empty mbarriers; rows and K columns past the end are zero-filled by the TMA unit, so M, n_out and Kp need
//     no padding (Kp = 3K+8 leaves an 8-wide tail block);
//   * warp 1: one thread issues tcgen05.mma (kind::f16, M = 128, N = BN, K = 16, A and B from shared memory through
//     SWIZZLE_128B K-major descriptors, 4 per K block); tcgen05.commit frees the slot / publishes the accumulator;
//   * warps 2-5: epilogue -- the accumulator lives in TMEM (two buffers of BN columns: the epilogue of tile i overlaps the
//     main loop of tile i+1), read with tcgen05.ld 32x32b (one thread = one output row), optionally added to an fp32
//     addend (the graph blocks' shortcut), stored as float4.
//   * persistent grid (one CTA per SM), tiles handed out round-robin with the N index fastest so that the CTAs working at
//     the same time share their A3 rows in L2;
//   * CTAs run as CLUSTERS OF TWO on neighbouring row blocks of the same column block: each loads its own A3 rows and
//     HALF of the W3 tile, multicast into both CTAs' shared memory (cp.async.bulk.tensor ... .multicast::cluster), and a
//     slot is released by both CTAs' MMA threads (tcgen05.commit ... .multicast::cluster on the peer's empty barrier too).
//     The first version, without it, re-read the whole W3 tile from L2 in every CTA for every K block -- 48 KB per 512
//     tensor clocks per SM -- and ran at 0.5-0.67x of the library GEMM, L2-bandwidth-bound; sharing W3 cuts that to 32 KB.
//
// Roofline: tensor pipe -- 2 M n Kp FLOP issued (3x the algorithmic 2 M n K: the price of fp32-grade accuracy on fp16
// tensor cores); HBM traffic M Kp 2 + M n 4 bytes is ~10x below the time the MMAs need at d_model = 512.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdlib.h>

#include "tc_helpers.cuh"

#ifndef UPD_GEMM3_DEFAULT_CL
#define UPD_GEMM3_DEFAULT_CL 2
#endif

namespace {

constexpr int BM = 128, BK = 64;                       // BK halves = 128 bytes = one swizzle span
constexpr int GEMM_THREADS = 192;                      // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr uint32_t A_STAGE_BYTES = BM * BK * 2;

template <int BN> struct Cfg {
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr uint32_t B_STAGE_BYTES = BN * BK * 2;
  static constexpr uint32_t STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr uint32_t TMEM_COLS = 2 * BN;        // two accumulator buffers (power of two >= 32)
  static constexpr size_t SMEM = 1024 + (size_t)STAGES * STAGE_BYTES + 1024 + 4 * 8192;
};

// Shared-memory matrix descriptor, K-major, SWIZZLE_128B: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO);
// the K step inside the swizzle span is taken on the start-address field (+32 bytes per K = 16 halves).
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)1 << 16;                              // LBO: unused for swizzled K-major layouts
  d |= (uint64_t)(1024u >> 4) << 32;                   // SBO
  d |= 1ull << 46;                                     // descriptor version (sm_100)
  d |= 2ull << 61;                                     // layout type SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// the same box delivered to the same shared-memory offset (and mbarrier) of every CTA of the cluster named in `mask`
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
               "[%0], [%1, {%3, %4}], [%2], %5;"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
// tcgen05.commit arriving on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void mma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)acc) : "memory");
}

struct Ring {
  int stage = 0;
  uint32_t phase = 0;
  template <int STAGES> __device__ __forceinline__ void advance() {
    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
  }
};

// Output through shared memory + TMA store.  One thread owns one accumulator row, so direct stores put the 32 lanes of a
// warp on 32 different 128-byte lines: 32 L1 wavefronts per STG.128, 8192 per 128 x 256 tile -- two thirds of the
// L1/shared-memory cycles the MMAs and the TMA refills of the same SM need (measured: every cluster size and the CTA-pair
// kernel plateaued at 0.77x of the library GEMM).  Instead each epilogue warp transposes 32 rows x 32 columns through an
// 4 KB staging tile in the 128-byte swizzle (conflict-free st.shared.v4) and one lane hands it to the TMA unit, which
// writes whole lines and clips rows / columns past the end of C.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
// The same with the TMA unit ADDING the tile to what C already holds (fp32 reduction in L2): out = addend + A3 W3^T with
// the addend resident in C -- no thread ever loads it.
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr uint32_t EPI_STAGE_BYTES = 2 * 32 * 128;      // per epilogue warp: two 32-row x 128-byte staging tiles

// One epilogue warp's 32 rows of a BN-wide accumulator -> global memory (+ addend), 32 columns at a time.
// tma_c != nullptr: staged TMA stores (stage = this warp's staging area); else direct stores (addend, n_out % 4 != 0).
template <int BN>
__device__ __forceinline__ void drain_accumulator(uint32_t src, float* __restrict__ C, const float* __restrict__ addend,
                                                  long long row, long long M, int n0, int n_out, const CUtensorMap* tma_c,
                                                  bool reduce_add, unsigned char* stage, int lane) {
  int buf = 0;
#pragma unroll 1
  for (int c = 0; c < BN; c += 32) {
    if (n0 + c >= n_out) break;                                         // warp-uniform
    uint32_t v[32];
    tc::tmem_ld32(src + c, v);
    tc::wait_ld();
    if (tma_c != nullptr) {
      if (lane == 0) tma_store_wait_read<1>();                          // the store that last read this staging tile is done
      __syncwarp();
      unsigned char* tile = stage + buf * (32 * 128) + lane * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(tile + ((j ^ (lane & 7)) << 4)) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (reduce_add) tma_reduce_add_2d(tma_c, tc::smem_u32(stage + buf * (32 * 128)), n0 + c, (int)(row - lane));
        else tma_store_2d(tma_c, tc::smem_u32(stage + buf * (32 * 128)), n0 + c, (int)(row - lane));
        tma_store_commit();
      }
      buf ^= 1;
    } else if (row < M) {
      float* out = C + row * (long long)n_out + n0 + c;
      const float* add = addend ? addend + row * (long long)n_out + n0 + c : nullptr;
      if (n0 + c + 32 <= n_out && (n_out & 3) == 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                 __uint_as_float(v[j + 3]));
          if (add) {
            const float4 a = *reinterpret_cast<const float4*>(add + j);
            o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
          }
          *reinterpret_cast<float4*>(out + j) = o;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n0 + c + j < n_out) out[j] = __uint_as_float(v[j]) + (add ? add[j] : 0.f);
      }
    }
  }
}

// CL = CTAs per cluster (1 or 2).  With CL = 2 a cluster works on the row blocks (2j, 2j+1) of one column block.
template <int BN, int CL>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm3_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
             const __grid_constant__ CUtensorMap tma_c, int use_tma_store, float* __restrict__ C,
             const float* __restrict__ addend, long long M, int n_out, int num_k_blocks, int tail_ksteps, long long num_m_blocks,
             int num_n_blocks) {
  using K = Cfg<BN>;
  constexpr int STAGES = K::STAGES;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + (size_t)STAGES * K::STAGE_BYTES);
  unsigned long long* full = bars;                       // [STAGES]  TMA -> MMA
  unsigned long long* empty = bars + STAGES;             // [STAGES]  MMA -> TMA
  unsigned long long* acc_full = bars + 2 * STAGES;      // [2]       MMA -> epilogue
  unsigned long long* acc_empty = bars + 2 * STAGES + 2; // [2]       epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  unsigned char* epi_stage = smem + (size_t)STAGES * K::STAGE_BYTES + 1024;      // 4 warps x EPI_STAGE_BYTES, 1024-aligned (swizzle atoms)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { tc::mbar_init(tc::smem_u32(&full[i]), 1); tc::mbar_init(tc::smem_u32(&empty[i]), CL); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(tc::smem_u32(&acc_full[i]), 1); tc::mbar_init(tc::smem_u32(&acc_empty[i]), 4); }
    tc::fence_mbar_init();
  }
  if (warp == 1) tc::tmem_alloc<K::TMEM_COLS>(tc::smem_u32(tmem_slot));
  tc::fence_before_sync();
  __syncthreads();
  if (CL > 1) cluster_sync_all();                        // the peer's barriers exist before anything is multicast to them
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // work units: (group of CL consecutive row blocks) x (column block); cluster c takes units c, c + #clusters, ...
  const uint32_t crank = (CL > 1) ? cluster_ctarank() : 0u;
  const long long num_groups = (num_m_blocks + CL - 1) / CL;
  const long long num_tiles = num_groups * num_n_blocks;
  const long long first = blockIdx.x / CL, stride = gridDim.x / CL;
  constexpr uint16_t MASK = (uint16_t)((1u << CL) - 1u);

  if (warp == 0) {
    // ------------------------------------------------ TMA producer ------------------------------------------------
    if (lane == 0) {
      Ring r;
      for (long long t = first; t < num_tiles; t += stride) {
        const int m0 = (int)((t / num_n_blocks) * CL + crank) * BM, n0 = (int)(t % num_n_blocks) * BN;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          tc::mbar_wait(tc::smem_u32(&empty[r.stage]), r.phase ^ 1u);       // freed by the MMA threads of all CL CTAs
          const uint32_t bar = tc::smem_u32(&full[r.stage]);
          const uint32_t sa = tc::smem_u32(smem + (size_t)r.stage * K::STAGE_BYTES);
          tc::mbar_expect_tx(bar, K::STAGE_BYTES);
          tma_load_2d(sa, &tma_a, bar, kb * BK, m0);
          if (CL == 1) {
            tma_load_2d(sa + A_STAGE_BYTES, &tma_b, bar, kb * BK, n0);
          } else {                                                        // my 1/CL of the W3 tile, to every CTA of the cluster
            const uint32_t part = (uint32_t)(BN / CL) * crank;
            tma_load_2d_mc(sa + A_STAGE_BYTES + part * (BK * 2), &tma_b, bar, kb * BK, n0 + (int)part, MASK);
          }
          r.advance<STAGES>();
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------- MMA issuer -------------------------------------------------
    if (lane == 0) {
      constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);   // f16 x f16 -> f32, K-major
      Ring r;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (long long t = first; t < num_tiles; t += stride) {
        tc::mbar_wait(tc::smem_u32(&acc_empty[acc]), acc_phase ^ 1u);      // the epilogue has drained this buffer
        tc::fence_after_sync();
        const uint32_t d = tmem_base + (uint32_t)acc * BN;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          tc::mbar_wait(tc::smem_u32(&full[r.stage]), r.phase);
          tc::fence_after_sync();
          const uint32_t sa = tc::smem_u32(smem + (size_t)r.stage * K::STAGE_BYTES);
          const uint64_t da = smem_desc_sw128(sa), db = smem_desc_sw128(sa + A_STAGE_BYTES);
#pragma unroll
          const int nk = kb + 1 < num_k_blocks ? BK / 16 : tail_ksteps;     // Kp = 3K + 8: the last block holds 8 columns
          for (int k = 0; k < nk; ++k) mma_f16_ss(d, da + 2u * k, db + 2u * k, IDESC, (kb | k) != 0);
          if (CL == 1) tc::mma_commit(tc::smem_u32(&empty[r.stage]));       // slot free once these MMAs have read it
          else mma_commit_mc(tc::smem_u32(&empty[r.stage]), MASK);           // ... in this CTA AND for the peer's producer
          r.advance<STAGES>();
        }
        tc::mma_commit(tc::smem_u32(&acc_full[acc]));                       // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // -------------------------------------------------- epilogue --------------------------------------------------
    const int quad = warp & 3;                                             // TMEM lane quadrant this warp may read
    int acc = 0;
    uint32_t acc_phase = 0;
    for (long long t = first; t < num_tiles; t += stride) {
      const long long row = ((t / num_n_blocks) * CL + crank) * BM + quad * 32 + lane;
      const int n0 = (int)(t % num_n_blocks) * BN;
      tc::mbar_wait(tc::smem_u32(&acc_full[acc]), acc_phase);
      tc::fence_after_sync();
      const uint32_t src = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc * BN;
      drain_accumulator<BN>(src, C, addend, row, M, n0, n_out, use_tma_store ? &tma_c : nullptr, use_tma_store == 2,
                            epi_stage + (size_t)quad * EPI_STAGE_BYTES, lane);
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tc::smem_u32(&acc_empty[acc]));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (lane == 0) tma_store_wait_read<0>();
  }
  tc::fence_before_sync();
  __syncthreads();
  if (CL > 1) cluster_sync_all();                        // no CTA leaves while its peer may still signal its barriers
  if (warp == 1) tc::tmem_dealloc<K::TMEM_COLS>(tmem_base);
}


// ---- CTA-pair form (tcgen05 cta_group::2): 256 x 256 tiles on two SMs ------------------------------------------------
// With both operands in shared memory a 128 x 256 x 16 MMA reads 12 KB per 128 tensor clocks (96 B/clk) while TMA refills
// the ring at the same rate: together well above the 128 B/clk an SM's shared memory delivers, which is what held the
// single-CTA kernel at ~0.75x of the library GEMM whatever the cluster size.  As a CTA pair each SM keeps only HALF of
// the W3 tile (the tensor cores fetch the other half from the peer), so the operand traffic per SM drops to 64 + 62 B/clk
// and the ring is 6 deep instead of 4.  Rank 0 of the pair issues every MMA; both CTAs load their own A3 rows and W3
// half, both drain their own 128 accumulator rows.
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void mma_f16_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)acc) : "memory");
}
__device__ __forceinline__ void mma_commit_pair(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t slot_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

template <int BN> struct PairCfg {
  static constexpr int STAGES = 6;
  static constexpr uint32_t B_STAGE_BYTES = (BN / 2) * BK * 2;      // this CTA's half of the W3 tile
  static constexpr uint32_t STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr uint32_t TMEM_COLS = 2 * BN;
  static constexpr size_t SMEM = 1024 + (size_t)STAGES * STAGE_BYTES + 1024 + 4 * 8192;
};

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm3_pair_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                  const __grid_constant__ CUtensorMap tma_c, int use_tma_store, float* __restrict__ C,
                  const float* __restrict__ addend, long long M, int n_out, int num_k_blocks, int tail_ksteps, long long num_m_blocks,
                  int num_n_blocks) {
  using K = PairCfg<BN>;
  constexpr int STAGES = K::STAGES;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + (size_t)STAGES * K::STAGE_BYTES);
  unsigned long long* full = bars;                       // [STAGES]  rank 0's: both CTAs' TMA -> the MMA thread
  unsigned long long* empty = bars + STAGES;             // [STAGES]  per CTA: MMA (rank 0, multicast) -> this CTA's producer
  unsigned long long* acc_full = bars + 2 * STAGES;      // [2]       per CTA: MMA (multicast) -> this CTA's epilogue
  unsigned long long* acc_empty = bars + 2 * STAGES + 2; // [2]       rank 0's: both CTAs' epilogue warps -> the MMA thread
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  unsigned char* epi_stage = smem + (size_t)STAGES * K::STAGE_BYTES + 1024;      // 4 warps x EPI_STAGE_BYTES, 1024-aligned (swizzle atoms)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { tc::mbar_init(tc::smem_u32(&full[i]), 1); tc::mbar_init(tc::smem_u32(&empty[i]), 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(tc::smem_u32(&acc_full[i]), 1); tc::mbar_init(tc::smem_u32(&acc_empty[i]), 8); }
    tc::fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_pair<K::TMEM_COLS>(tc::smem_u32(tmem_slot));
  tc::fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const long long num_groups = (num_m_blocks + 1) / 2;
  const long long num_tiles = num_groups * num_n_blocks;
  const long long first = blockIdx.x / 2, stride = gridDim.x / 2;

  if (warp == 0) {
    if (lane == 0) {
      Ring r;
      for (long long t = first; t < num_tiles; t += stride) {
        const int m0 = (int)((t / num_n_blocks) * 2 + crank) * BM, n0 = (int)(t % num_n_blocks) * BN + (int)crank * (BN / 2);
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          tc::mbar_wait(tc::smem_u32(&empty[r.stage]), r.phase ^ 1u);
          const uint32_t bar0 = mapa_rank(tc::smem_u32(&full[r.stage]), 0);       // the pair's barrier lives in rank 0
          const uint32_t sa = tc::smem_u32(smem + (size_t)r.stage * K::STAGE_BYTES);
          if (crank == 0) tc::mbar_expect_tx(tc::smem_u32(&full[r.stage]), 2 * K::STAGE_BYTES);
          tma_load_2d_pair(sa, &tma_a, bar0, kb * BK, m0);
          tma_load_2d_pair(sa + A_STAGE_BYTES, &tma_b, bar0, kb * BK, n0);
          r.advance<STAGES>();
        }
      }
    }
  } else if (warp == 1) {
    if (crank == 0 && lane == 0) {
      constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);   // M = 256 over the pair
      Ring r;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (long long t = first; t < num_tiles; t += stride) {
        tc::mbar_wait(tc::smem_u32(&acc_empty[acc]), acc_phase ^ 1u);      // both CTAs' epilogues have drained this buffer
        tc::fence_after_sync();
        const uint32_t d = tmem_base + (uint32_t)acc * BN;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          tc::mbar_wait(tc::smem_u32(&full[r.stage]), r.phase);
          tc::fence_after_sync();
          const uint32_t sa = tc::smem_u32(smem + (size_t)r.stage * K::STAGE_BYTES);
          const uint64_t da = smem_desc_sw128(sa), db = smem_desc_sw128(sa + A_STAGE_BYTES);
#pragma unroll
          const int nk = kb + 1 < num_k_blocks ? BK / 16 : tail_ksteps;     // Kp = 3K + 8: the last block holds 8 columns
          for (int k = 0; k < nk; ++k) mma_f16_ss_pair(d, da + 2u * k, db + 2u * k, IDESC, (kb | k) != 0);
          mma_commit_pair(tc::smem_u32(&empty[r.stage]), 3);                // frees the slot in both CTAs
          r.advance<STAGES>();
        }
        mma_commit_pair(tc::smem_u32(&acc_full[acc]), 3);                   // accumulator complete, in both CTAs
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    const int quad = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (long long t = first; t < num_tiles; t += stride) {
      const long long row = ((t / num_n_blocks) * 2 + crank) * BM + quad * 32 + lane;
      const int n0 = (int)(t % num_n_blocks) * BN;
      tc::mbar_wait(tc::smem_u32(&acc_full[acc]), acc_phase);
      tc::fence_after_sync();
      const uint32_t src = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc * BN;
      drain_accumulator<BN>(src, C, addend, row, M, n0, n_out, use_tma_store ? &tma_c : nullptr, use_tma_store == 2,
                            epi_stage + (size_t)quad * EPI_STAGE_BYTES, lane);
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_rank(tc::smem_u32(&acc_empty[acc]), 0));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (lane == 0) tma_store_wait_read<0>();
  }
  tc::fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair<K::TMEM_COLS>(tmem_base);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda at link time)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// [rows, cols] fp16 row-major, box = 64 columns x box_rows rows, 128-byte swizzle, zero fill out of bounds
bool make_map(CUtensorMap* m, const void* base, long long rows, int cols, int box_rows) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// C [M, n_out] fp32 row-major, box = 32 columns x 32 rows, 128-byte swizzle (the epilogue's staging tiles)
bool make_map_c(CUtensorMap* m, float* base, long long rows, int cols) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, int CL>
cudaError_t launch_cl(const CUtensorMap& ma, const CUtensorMap& mb, long long M, int n_out, int Kp, float* out,
                      const float* addend, int sms, cudaStream_t stream) {
  auto kern = gemm3_kernel<BN, CL>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<BN>::SMEM);
  if (e != cudaSuccess) return e;
  const long long mb_ = (M + BM - 1) / BM;
  const int nb = (n_out + BN - 1) / BN;
  const long long units = ((mb_ + CL - 1) / CL) * nb;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg<BN>::SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // clusters that can be co-resident (a cluster lives inside one GPC): the persistent grid is exactly that many
  static thread_local int max_clusters_cache[3][5] = {};
  int& maxc = max_clusters_cache[BN == 64 ? 0 : (BN == 128 ? 1 : 2)][CL];
  if (maxc == 0) {
    cfg.gridDim = dim3((unsigned)((sms / CL) * CL));
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) n = sms / CL;
    maxc = n;
  }
  long long clusters = maxc;
  if (units < clusters) clusters = units;
  cfg.gridDim = dim3((unsigned)(clusters * CL));
  CUtensorMap mc = ma;                                              // placeholder when the direct-store path is taken
  // 1 = staged TMA store; 2 = staged TMA reduce-add onto the addend, which is first copied into C unless it already is C
  int use_tma = ((n_out & 3) == 0 && make_map_c(&mc, out, M, n_out)) ? (addend ? 2 : 1) : 0;
  if (use_tma == 2 && addend != out) {
    cudaError_t ce = cudaMemcpyAsync(out, addend, (size_t)M * n_out * sizeof(float), cudaMemcpyDeviceToDevice, stream);
    if (ce != cudaSuccess) return ce;
  }
  return cudaLaunchKernelEx(&cfg, kern, ma, mb, mc, use_tma, out, use_tma ? nullptr : addend, M, n_out, (Kp + BK - 1) / BK,
                            ((Kp - ((Kp + BK - 1) / BK - 1) * BK) + 15) / 16, mb_, nb);
}

template <int BN>
cudaError_t launch_pair(const CUtensorMap& ma, const CUtensorMap& mb, long long M, int n_out, int Kp, float* out,
                        const float* addend, int sms, cudaStream_t stream) {
  auto kern = gemm3_pair_kernel<BN>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PairCfg<BN>::SMEM);
  if (e != cudaSuccess) return e;
  const long long mb_ = (M + BM - 1) / BM;
  const int nb = (n_out + BN - 1) / BN;
  const long long units = ((mb_ + 1) / 2) * nb;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = PairCfg<BN>::SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static thread_local int maxc = 0;
  if (maxc == 0) {
    cfg.gridDim = dim3((unsigned)((sms / 2) * 2));
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) n = sms / 2;
    maxc = n;
  }
  long long clusters = maxc;
  if (units < clusters) clusters = units;
  cfg.gridDim = dim3((unsigned)(clusters * 2));
  CUtensorMap mc = ma;
  // 1 = staged TMA store; 2 = staged TMA reduce-add onto the addend, which is first copied into C unless it already is C
  int use_tma = ((n_out & 3) == 0 && make_map_c(&mc, out, M, n_out)) ? (addend ? 2 : 1) : 0;
  if (use_tma == 2 && addend != out) {
    cudaError_t ce = cudaMemcpyAsync(out, addend, (size_t)M * n_out * sizeof(float), cudaMemcpyDeviceToDevice, stream);
    if (ce != cudaSuccess) return ce;
  }
  return cudaLaunchKernelEx(&cfg, kern, ma, mb, mc, use_tma, out, use_tma ? nullptr : addend, M, n_out, (Kp + BK - 1) / BK,
                            ((Kp - ((Kp + BK - 1) / BK - 1) * BK) + 15) / 16, mb_, nb);
}

template <int BN>
cudaError_t launch(const void* a3, const void* w3, long long M, int Nw, int n_out, int Kp, float* out, const float* addend,
                   int sms, cudaStream_t stream) {
  static int want = -1;                                            // UPD_GEMM3_CL = 1 / 2 / 4 overrides (experiments)
  if (want < 0) { const char* e = getenv("UPD_GEMM3_CL"); want = e ? atoi(e) : 0; }
  const long long mblocks = (M + BM - 1) / BM;
  if (BN == 256 && mblocks >= 2 && want != 1 && want != 2 && want != 4) {       // the product path of the wide layers
    CUtensorMap ma, mb;
    if (!make_map(&ma, a3, M, Kp, BM) || !make_map(&mb, w3, Nw, Kp, BN / 2)) return cudaErrorInvalidValue;
    return launch_pair<BN>(ma, mb, M, n_out, Kp, out, addend, sms, stream);
  }
  int cl = want > 0 ? want : UPD_GEMM3_DEFAULT_CL;
  while (cl > 1 && mblocks < cl) cl >>= 1;                         // fewer row blocks than CTAs in a cluster
  CUtensorMap ma, mb;
  if (!make_map(&ma, a3, M, Kp, BM) || !make_map(&mb, w3, Nw, Kp, BN / cl)) return cudaErrorInvalidValue;
  if (cl == 4) return launch_cl<BN, 4>(ma, mb, M, n_out, Kp, out, addend, sms, stream);
  if (cl == 2) return launch_cl<BN, 2>(ma, mb, M, n_out, Kp, out, addend, sms, stream);
  return launch_cl<BN, 1>(ma, mb, M, n_out, Kp, out, addend, sms, stream);
}

}  // namespace

// M x Kp fp16 times (Nw x Kp fp16)^T -> M x n_out fp32 (n_out <= Nw); cudaErrorInvalidValue = shape the TMA path cannot
// take (Kp not a multiple of 8 halves, operands not 16-byte aligned)
cudaError_t upd_launch_gemm3(const void* a3, const void* w3, long long M, int Nw, int n_out, int Kp, float* out,
                             const float* addend, int sms, cudaStream_t stream) {
  if ((Kp & 7) != 0 || (reinterpret_cast<uintptr_t>(a3) & 15) != 0 || (reinterpret_cast<uintptr_t>(w3) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(out) & 15) != 0 || (addend && (reinterpret_cast<uintptr_t>(addend) & 15) != 0))
    return cudaErrorInvalidValue;
  int bn = n_out <= 64 ? 64 : (n_out <= 128 ? 128 : 256);
  if (bn == 256) {                     // a ragged last column block: 128-wide blocks when they cut the padding (> 25 %)
    const int p256 = (n_out + 255) / 256 * 256, p128 = (n_out + 127) / 128 * 128;
    if (p128 < p256 && 4 * (p256 - n_out) > n_out) bn = 128;
  }
  if (const char* e = getenv("UPD_GEMM3_BN")) {                     // experiments: force the column block
    const int f = atoi(e);
    if (f == 64 || f == 128 || f == 256) bn = f;
  }
  if (bn == 64) return launch<64>(a3, w3, M, Nw, n_out, Kp, out, addend, sms, stream);
  if (bn == 128) return launch<128>(a3, w3, M, Nw, n_out, Kp, out, addend, sms, stream);
  return launch<256>(a3, w3, M, Nw, n_out, Kp, out, addend, sms, stream);
}
