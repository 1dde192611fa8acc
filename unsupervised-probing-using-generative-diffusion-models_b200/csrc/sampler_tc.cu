// Fused persistent reverse-diffusion sampler on tcgen05 tensor cores (UPD_IMPL_TCGEN05).
//
// One CTA per SM, 256 threads = two independent warpgroups.  A warpgroup owns a tile of 128
// denoiser rows (row = one (window,row,sample,position), thread = row = TMEM lane) and carries it
// through all T reverse steps without leaving the SM:
//
//   weights  : lin2/lin3 as fp16 hi/lo (x wscale) and lin1|bias as tf32 hi/lo, in UMMA K-major
//              core-matrix layout, plus the fp32 tables, bulk-copied (TMA 1-D) into shared memory
//              once per CTA and resident for the whole launch.
//   per step : A1 = [y | y0_hat | gx | 1] (tf32 hi/lo) -> TMEM;  D1 = A1 * W1'      (3 tf32 MMAs / K-slice)
//              epilogue: e[t] * D -> softplus -> sum of squares -> fp16 hi/lo, written IN PLACE
//                        over the accumulator columns it was read from (A operand of the next layer)
//              D2 = A2 * W2', D3 = A3 * W3'   (A from TMEM, B from smem, 24 fp16 MMAs each = hi*hi+lo*hi+hi*lo)
//              the L2 normalisation of layer l is applied as a scalar on the accumulator of layer l+1
//              heads (eps, sigma) as fp32 FMAs on the thread's own row; NsDiff/TMDM posterior algebra;
//              Philox (or injected) Gaussian noise.
//   TMEM     : 512 columns = 2 tiles x 2 ping-pong buffers x 128 columns.
// While one warpgroup waits for its MMAs the other one runs its softplus epilogue, so the MUFU pipe
// (the real bound of this MLP: 514 softplus per row-step) and the tensor pipe overlap.
#include "sampler_params.cuh"
#include "tc_helpers.cuh"
#include "upd_common.cuh"

namespace {

constexpr int TC_THREADS = 256;
constexpr uint32_t UMMA_LBO = 2048;   // K-adjacent core matrices (see upd_common.cuh layout)
constexpr uint32_t UMMA_SBO = 128;    // N-adjacent core matrices

struct __align__(8) TcSync {
  unsigned long long wbar;
  unsigned long long mma_bar[2];
  uint32_t tmem_base;
  uint32_t pad;
};

// e[t] * acc (+bias, *inv) -> softplus -> ss; result re-encoded in place as the next A operand.
template <bool FIRST>
__device__ __forceinline__ float epilogue_to_a(uint32_t buf, const float* __restrict__ e, const float* __restrict__ b,
                                               float inv) {
  float ss = 0.f;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    uint32_t r[32], o[32];
    tc::tmem_ld32(buf + 32u * c, r);
    tc::wait_ld();
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float4 e4 = *reinterpret_cast<const float4*>(e + 32 * c + j);
      float4 b4 = FIRST ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(b + 32 * c + j);
      float a0 = __uint_as_float(r[j]), a1 = __uint_as_float(r[j + 1]), a2 = __uint_as_float(r[j + 2]),
            a3 = __uint_as_float(r[j + 3]);
      float z0 = FIRST ? a0 * e4.x : (a0 * inv + b4.x) * e4.x;
      float z1 = FIRST ? a1 * e4.y : (a1 * inv + b4.y) * e4.y;
      float z2 = FIRST ? a2 * e4.z : (a2 * inv + b4.z) * e4.z;
      float z3 = FIRST ? a3 * e4.w : (a3 * inv + b4.w) * e4.w;
      float h0 = upd_softplus(z0), h1 = upd_softplus(z1), h2 = upd_softplus(z2), h3 = upd_softplus(z3);
      ss = fmaf(h0, h0, ss); ss = fmaf(h1, h1, ss); ss = fmaf(h2, h2, ss); ss = fmaf(h3, h3, ss);
      tc::split_f16x2(h0, h1, o[j / 2], o[16 + j / 2]);
      tc::split_f16x2(h2, h3, o[j / 2 + 1], o[16 + j / 2 + 1]);
    }
    tc::tmem_st32(buf + 32u * c, o);
  }
  return ss;
}

template <int KIND, int F>
__global__ void __launch_bounds__(TC_THREADS, 1)
sampler_tc_kernel(const UpdSamplerParams p) {
  constexpr bool NS = (KIND == 0);
  constexpr int IN = NS ? 3 * F : 2 * F;
  constexpr int K1 = ((IN + 1 + 7) / 8) * 8;
  const UpdPackLayout L = upd_make_layout(KIND, F, p.T);
  extern __shared__ __align__(128) unsigned char smem[];
  auto sf = [&](uint32_t off) { return reinterpret_cast<const float*>(smem + off); };
  const uint32_t steps_off = upd_align128(L.tc_image_bytes);
  constexpr uint32_t STEP_BYTES = NS ? sizeof(UpdNsStep) : sizeof(UpdTmStep);
  const uint32_t sync_off = upd_align128(steps_off + STEP_BYTES * p.T);
  TcSync* sync = reinterpret_cast<TcSync*>(smem + sync_off);

  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    tc::mbar_init(tc::smem_u32(&sync->wbar), 1);
    tc::mbar_init(tc::smem_u32(&sync->mma_bar[0]), 1);
    tc::mbar_init(tc::smem_u32(&sync->mma_bar[1]), 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc<512>(tc::smem_u32(&sync->tmem_base));
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = sync->tmem_base;
  if (tid == 0) {
    const uint32_t bar = tc::smem_u32(&sync->wbar);
    tc::mbar_expect_tx(bar, L.tc_image_bytes);
    const unsigned char* src = reinterpret_cast<const unsigned char*>(p.packed);
    for (uint32_t off = 0; off < L.tc_image_bytes; off += 16384u) {
      uint32_t n = L.tc_image_bytes - off < 16384u ? L.tc_image_bytes - off : 16384u;
      tc::bulk_g2s(tc::smem_u32(smem + off), src + off, n, bar);
    }
  }
  tc::mbar_wait(tc::smem_u32(&sync->wbar), 0);
  // per-step posterior scalars, computed once per CTA
  if (tid < p.T) {
    if (NS) reinterpret_cast<UpdNsStep*>(smem + steps_off)[tid] = upd_ns_step(sf(L.sched), p.T, tid);
    else reinterpret_cast<UpdTmStep*>(smem + steps_off)[tid] = upd_tm_step(sf(L.sched), p.T, tid);
  }
  __syncthreads();

  const int wg = tid >> 7, wtid = tid & 127, quad = (tid >> 5) & 3;
  const uint32_t col0 = (uint32_t)wg * 256u;
  const uint32_t lane_sel = (uint32_t)(quad * 32) << 16;
  const uint32_t buf0 = tmem_base + lane_sel + col0, buf1 = buf0 + 128u;          // this warp's lanes
  const uint32_t mma0 = tmem_base + col0, mma1 = mma0 + 128u;                      // lane 0 (MMA operands)
  const uint32_t bar = tc::smem_u32(&sync->mma_bar[wg]);
  const uint32_t img = tc::smem_u32(smem);
  const float inv_ws2 = sf(L.scales)[0], inv_ws3 = sf(L.scales)[1];
  uint32_t phase = 0;

  const long long n_tiles = (p.n_rows + 127) / 128;
  for (long long tile = (long long)blockIdx.x * 2 + wg; tile < n_tiles; tile += (long long)gridDim.x * 2) {
    const long long row = tile * 128 + wtid;
    const bool live = row < p.n_rows;
    const UpdRowIndex ix = upd_row_index(p, live ? row : p.n_rows - 1);
    float y[F], y0h[F], gxv[F];
    {
      const long long cidx = (ix.r0 * p.O + ix.o) * F;
#pragma unroll
      for (int f = 0; f < F; ++f) {
        y0h[f] = p.y0_hat ? p.y0_hat[cidx + f] : 0.f;
        gxv[f] = NS ? p.gx[cidx + f] : 1.f;
        float z = upd_draw(p, ix, f, F, 0);
        y[f] = NS ? sqrtf(gxv[f]) * z + y0h[f] : z + y0h[f];     // nsdiff_utils.py:274 / tmdm_diffusion_utils.py:110
      }
    }
    for (int t = p.T - 1; t >= 0; --t) {
      // ---------------- layer 1: A1 = [y | y0_hat | gx | 1 | 0] as tf32 hi/lo ----------------
      {
        float in[K1];
#pragma unroll
        for (int i = 0; i < K1; ++i) in[i] = 0.f;
#pragma unroll
        for (int f = 0; f < F; ++f) {
          in[f] = y[f];
          in[F + f] = y0h[f];
          if (NS) in[2 * F + f] = gxv[f];
        }
        in[IN] = 1.0f;
        if (K1 == 8) {
          uint32_t a[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float hi = tc::to_tf32(in[i]);
            a[i] = __float_as_uint(hi);
            a[8 + i] = __float_as_uint(tc::to_tf32(in[i] - hi));
          }
          tc::tmem_st16(buf0, a);
        } else {
          uint32_t a[32];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float v = in[i % K1];
            float hi = tc::to_tf32(v);
            a[i] = __float_as_uint(hi);
            a[16 + i] = __float_as_uint(tc::to_tf32(v - hi));
          }
          tc::tmem_st32(buf0, a);
        }
      }
      tc::wait_st();
      tc::fence_before_sync();
      tc::named_bar_sync(1 + wg, 128);
      if (wtid == 0) {
        tc::fence_after_sync();
        tc::issue_layer_tf32x3(mma1, mma0, K1, img + L.u1hi, img + L.u1lo, UMMA_LBO, UMMA_SBO);
        tc::mma_commit(bar);
      }
      tc::mbar_wait(bar, phase); phase ^= 1u;
      tc::fence_after_sync();

      // ---------------- layer 1 epilogue -> A2 (in place, buf1); layer 2 ----------------
      float ss = epilogue_to_a<true>(buf1, sf(L.e1) + t * 128, nullptr, 1.f);
      tc::wait_st();
      tc::fence_before_sync();
      tc::named_bar_sync(1 + wg, 128);
      if (wtid == 0) {
        tc::fence_after_sync();
        tc::issue_layer_f16x3(mma0, mma1, img + L.u2hi, img + L.u2lo, UMMA_LBO, UMMA_SBO);
        tc::mma_commit(bar);
      }
      float inv = NS ? inv_ws2 / fmaxf(sqrtf(ss), 1e-12f) : inv_ws2;      // F.normalize folded past the GEMM
      tc::mbar_wait(bar, phase); phase ^= 1u;
      tc::fence_after_sync();

      // ---------------- layer 2 epilogue -> A3 (in place, buf0); layer 3 ----------------
      ss = epilogue_to_a<false>(buf0, sf(L.e2) + t * 128, sf(L.b2), inv);
      tc::wait_st();
      tc::fence_before_sync();
      tc::named_bar_sync(1 + wg, 128);
      if (wtid == 0) {
        tc::fence_after_sync();
        tc::issue_layer_f16x3(mma1, mma0, img + L.u3hi, img + L.u3lo, UMMA_LBO, UMMA_SBO);
        tc::mma_commit(bar);
      }
      inv = NS ? inv_ws3 / fmaxf(sqrtf(ss), 1e-12f) : inv_ws3;
      tc::mbar_wait(bar, phase); phase ^= 1u;
      tc::fence_after_sync();

      // ---------------- layer 3 epilogue + heads (denoise.py:50 / tmdm_model.py:63) ----------------
      float eps[F], sig[F];
#pragma unroll
      for (int f = 0; f < F; ++f) { eps[f] = 0.f; sig[f] = 0.f; }
      const float* e3 = sf(L.e3) + t * 128;
      const float* b3 = sf(L.b3);
      const float* w4 = sf(L.w4);
      const float* wsg = sf(L.ws);
      if (NS) {
        // pass 1: h3 = softplus(.) kept in TMEM as fp32 (in place), sum of squares
        ss = 0.f;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tc::tmem_ld32(buf1 + 32u * c, r);
          tc::wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float h = upd_softplus((__uint_as_float(r[j]) * inv + b3[32 * c + j]) * e3[32 * c + j]);
            ss = fmaf(h, h, ss);
            r[j] = __float_as_uint(h);
          }
          tc::tmem_st32(buf1 + 32u * c, r);
        }
        tc::wait_st();
        const float inv3 = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
        // pass 2: normalised h -> eps head, softplus(h) -> sigma head
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tc::tmem_ld32(buf1 + 32u * c, r);
          tc::wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float hn = __uint_as_float(r[j]) * inv3;
            float sp = upd_softplus(hn);
#pragma unroll
            for (int f = 0; f < F; ++f) {
              eps[f] = fmaf(w4[f * 128 + 32 * c + j], hn, eps[f]);
              sig[f] = fmaf(wsg[f * 128 + 32 * c + j], sp, sig[f]);
            }
          }
        }
#pragma unroll
        for (int f = 0; f < F; ++f) {
          eps[f] += sf(L.b4)[f];
          sig[f] = upd_softplus_accurate(sig[f] + sf(L.bs)[f]);
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tc::tmem_ld32(buf1 + 32u * c, r);
          tc::wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float h = upd_softplus((__uint_as_float(r[j]) * inv + b3[32 * c + j]) * e3[32 * c + j]);
#pragma unroll
            for (int f = 0; f < F; ++f) eps[f] = fmaf(w4[f * 128 + 32 * c + j], h, eps[f]);
          }
        }
#pragma unroll
        for (int f = 0; f < F; ++f) eps[f] += sf(L.b4)[f];
      }

      // ---------------- posterior update ----------------
      const bool last = (t == 0);
      if (NS) {
        const UpdNsStep st = reinterpret_cast<const UpdNsStep*>(smem + steps_off)[t];
#pragma unroll
        for (int f = 0; f < F; ++f) {
          float z = last ? 0.f : upd_draw(p, ix, f, F, p.T - t);
          y[f] = upd_ns_update(st, y[f], y0h[f], gxv[f], eps[f], sig[f], z, last);
        }
      } else {
        const UpdTmStep st = reinterpret_cast<const UpdTmStep*>(smem + steps_off)[t];
#pragma unroll
        for (int f = 0; f < F; ++f) {
          float z = last ? 0.f : upd_draw(p, ix, f, F, p.T - t);
          y[f] = upd_tm_update(st, y[f], y0h[f], eps[f], z, last);
        }
      }
    }
    if (live) {
#pragma unroll
      for (int f = 0; f < F; ++f) p.out[row * F + f] = y[f];
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem_base);
}

template <int KIND, int F>
cudaError_t launch(const UpdSamplerParams& p, int sms, cudaStream_t stream) {
  const UpdPackLayout L = upd_make_layout(KIND, F, p.T);
  constexpr uint32_t STEP_BYTES = (KIND == 0) ? sizeof(UpdNsStep) : sizeof(UpdTmStep);
  size_t smem = upd_align128(upd_align128(L.tc_image_bytes) + STEP_BYTES * p.T) + sizeof(TcSync) + 128;
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  auto kern = sampler_tc_kernel<KIND, F>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  long long n_tiles = (p.n_rows + 127) / 128;
  long long ctas = (n_tiles + 1) / 2;
  int grid = (int)(ctas < sms ? ctas : sms);
  if (grid < 1) grid = 1;
  kern<<<grid, TC_THREADS, smem, stream>>>(p);
  return cudaGetLastError();
}

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
  if (mode == 0 && (flags & 4)) {
    for (int q = 0; q < 8; ++q) {          // 16-column groups: hi words [16q,16q+8), lo words [16q+8,16q+16)
      uint32_t o[16];
      for (int j = 0; j < 16; j += 2)
        tc::split_f16x2(A[tid * K + 16 * q + j], A[tid * K + 16 * q + j + 1], o[j / 2], o[8 + j / 2]);
      tc::tmem_st16(abuf + 16u * q, o);
    }
  } else if (mode == 0) {
    for (int c = 0; c < 4; ++c) {
      uint32_t o[32];
      for (int j = 0; j < 32; j += 2) {
        float a0 = A[tid * K + 32 * c + j], a1 = A[tid * K + 32 * c + j + 1];
        if (flags & 2) { float tmp = a0; a0 = a1; a1 = tmp; }
        tc::split_f16x2(a0, a1, o[j / 2], o[16 + j / 2]);
      }
      tc::tmem_st32(abuf + 32u * c, o);
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
    if (mode == 0 && (flags & 4)) tc::issue_layer_f16x3_g16(tmem_base + 128u, tmem_base, tc::smem_u32(bhi), tc::smem_u32(blo), lbo, sbo);
    else if (mode == 0) tc::issue_layer_f16x3(tmem_base + 128u, tmem_base, tc::smem_u32(bhi), tc::smem_u32(blo), lbo, sbo);
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

cudaError_t upd_launch_sampler_tc(const UpdSamplerParams& p, int kind, int F, int sms, cudaStream_t stream) {
#define UPD_CASE(KK, FF) if (kind == KK && F == FF) return launch<KK, FF>(p, sms, stream);
  UPD_CASE(0, 1) UPD_CASE(0, 2) UPD_CASE(0, 3) UPD_CASE(0, 4)
  UPD_CASE(1, 1) UPD_CASE(1, 2) UPD_CASE(1, 3) UPD_CASE(1, 4)
#undef UPD_CASE
  return cudaErrorInvalidValue;
}

cudaError_t upd_launch_selftest_umma(const float* a, const float* b, float* d, int K, int mode, int flags,
                                     cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(selftest_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  if (e != cudaSuccess) return e;
  selftest_umma_kernel<<<1, 128, 65536, stream>>>(a, b, d, K, mode, flags);
  return cudaGetLastError();
}
