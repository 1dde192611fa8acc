// DiffSTG sampler: the Gaussian posterior step and the gated graph aggregation of the spatial block.
//   gaussian_posterior (DDIM / DDPM branch) ... models/Diffusion_model/DiffSTG/graph_diffusion_model.py:46-73
//   SpatialBlock = relu(ResGatedGraphConv) ..... models/Diffusion_model/DiffSTG/ugnet.py:36-45, models/layer/gnn_conv.py:18-19
//   duplicate_edge_index ....................... graph_diffusion_model.py:77-84 (replicas share one CSR here)
// HBM/L2-bound gather kernels; no tensor-core work.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "upd_common.cuh"

namespace {

// out = a*(xt - b*pred) + c*w,  w = z (DDPM branch, t <= 1) or pred (DDIM).  a, b, c are the reference's
// Python floats (float64 schedule -> .item()), rounded to fp32 where they meet the fp32 tensors.
__global__ void stg_posterior_kernel(const float* __restrict__ xt, const float* __restrict__ pred,
                                     const float* __restrict__ z, long long n, float a, float b, float c,
                                     float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float p = pred[i];
    float m = __fmul_rn(a, __fsub_rn(xt[i], __fmul_rn(b, p)));
    float w = z ? z[i] : p;
    out[i] = __fadd_rn(m, __fmul_rn(c, w));
  }
}

// out[n, ch] = act( sum_{j in N_in(v)} sigmoid(k[n,ch] + q[rep*V + j, ch]) * v[rep*V + j, ch] + skip[n, ch] + bias[ch] )
// with n = rep*V + v.  kqvs is the output of ONE fused projection [N, 4C] = (key | query | value | skip).
// Neighbours are visited in edge order (the order a sequential scatter-add accumulates them).
__global__ void stg_gated_aggregate_kernel(const float* __restrict__ kqvs, const int* __restrict__ rowptr,
                                           const int* __restrict__ col, const float* __restrict__ bias, long long N, int V,
                                           int C, int relu, float* __restrict__ out) {
  const long long total = N * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / C;
    const int ch = (int)(i - n * C);
    const int v = (int)(n % V);
    const long long base = n - v;                        // first node of this replica
    const float ki = kqvs[n * 4 * C + ch];
    float acc = 0.0f;
    const int e0 = rowptr[v], e1 = rowptr[v + 1];
    for (int e = e0; e < e1; ++e) {
      const float* src = kqvs + (base + col[e]) * 4 * C;
      float g = __fadd_rn(ki, src[C + ch]);
      float s = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-g)));
      acc = __fadd_rn(acc, __fmul_rn(s, src[2 * C + ch]));
    }
    acc = __fadd_rn(acc, kqvs[n * 4 * C + 3 * C + ch]);
    if (bias) acc = __fadd_rn(acc, bias[ch]);
    out[i] = relu ? fmaxf(acc, 0.0f) : acc;
  }
}


// The same aggregation with a replica's query / value rows staged in shared memory: every node of a replica reads the
// q and v rows of ~2E/V neighbours, so the gather version above moves 2E/V times the tensor through L2 (measured: 11 % of
// a DiffSTG step, L2-bound).  One CTA per (replica, slab of CS channels): q and v of all V nodes of the slab -> shared
// (coalesced), then a warp per node, a lane per channel; arithmetic and edge order as above (bit-identical results).
template <int CS>
__global__ void __launch_bounds__(256) stg_gated_aggregate_smem_kernel(const float* __restrict__ kqvs, const int* __restrict__ rowptr,
                                                                       const int* __restrict__ col, const float* __restrict__ bias,
                                                                       int V, int C, int relu, float* __restrict__ out) {
  extern __shared__ float sqv[];                     // q [V][CS] | v [V][CS]
  float* sq = sqv;
  float* sv = sqv + V * CS;
  const long long base = (long long)blockIdx.x * V;  // first node of this replica
  const int c0 = blockIdx.y * CS;
  const int lane = threadIdx.x % CS, grp = threadIdx.x / CS, ngrp = blockDim.x / CS;
  const int ch = c0 + lane;
  const bool ok = ch < C;
  for (int v = grp; v < V; v += ngrp) {
    const float* src = kqvs + (base + v) * 4 * C;
    sq[v * CS + lane] = ok ? src[C + ch] : 0.0f;
    sv[v * CS + lane] = ok ? src[2 * C + ch] : 0.0f;
  }
  __syncthreads();
  if (!ok) return;
  for (int v = grp; v < V; v += ngrp) {
    const float* self = kqvs + (base + v) * 4 * C;
    const float ki = self[ch];
    float acc = 0.0f;
    const int e0 = rowptr[v], e1 = rowptr[v + 1];
    for (int e = e0; e < e1; ++e) {
      const int j = __ldg(col + e);
      float g = __fadd_rn(ki, sq[j * CS + lane]);
      float s = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-g)));
      acc = __fadd_rn(acc, __fmul_rn(s, sv[j * CS + lane]));
    }
    acc = __fadd_rn(acc, self[3 * C + ch]);
    if (bias) acc = __fadd_rn(acc, bias[ch]);
    out[(base + v) * C + ch] = relu ? fmaxf(acc, 0.0f) : acc;
  }
}

// ------------------------------------------------------------------------------------------------------
// Front half of a ResidualBlock (ugnet.py:117-127) in one pass over the activations:
//     h1 = causal_conv3(x) + b1[step]     (TcnBlock 1: conv + 1x1 shortcut folded into tap 2, + t_conv(time emb))
//     h2 = causal_conv3(h1) + b2          (TcnBlock 2, identity shortcut folded)
//     hn = LayerNorm_c(h2) * g + beta     (nn.LayerNorm([1, c]) over the channel axis, eps 1e-5)
// x [N, CI, T] -> hn [N, C, T].  One CTA per row (or 8 rows); the row is walked in segments of TT = 4*blockDim positions:
// x and h1 of a segment live in shared memory behind 4 halo columns -- zero at the start of a row (both convolutions
// are causal and zero-padded), else the last 4 columns of the previous segment.  Each thread owns 4 consecutive
// positions x all C channels in registers, so the LayerNorm needs no cross-thread reduction and the store is one
// float4 per channel (scalar loads / stores when T is not a multiple of 4: the global rows are then only 8-byte aligned).
// fp32 FFMA: C <= 16 channels is far below a tensor-core tile, and the pass is 3 B/FLOP away from HBM-bound.
// ------------------------------------------------------------------------------------------------------
template <int C, bool XG>   // XG: x is read through L1 straight from global memory instead of being staged (T % 4 == 0 only)
__global__ void __launch_bounds__(128) stg_tcn_ln_kernel(const float* __restrict__ x, const float* __restrict__ w1,
                                                         const float* __restrict__ b1, const float* __restrict__ w2,
                                                         const float* __restrict__ b2, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, long long N, int CI, int T,
                                                         int rows_per_cta, float* __restrict__ hn,
                                                         __half* __restrict__ a3, const float* __restrict__ wsc,
                                                         float* __restrict__ sc_out, const float* __restrict__ x2, int CI1) {
  // x2 != nullptr: the input is the channel concatenation of x [N, CI1, T] and x2 [N, CI - CI1, T] (U-Net skip connection,
  // ugnet.py:288-289) read from the two tensors in place
  extern __shared__ __align__(16) float smem[];
  const int TT = (T <= 4 * (int)blockDim.x) ? ((T + 3) & ~3) : 4 * (int)blockDim.x;   // positions per segment
  const int TP = TT + 4;                             // row pitch: data starts at column 4 (16-byte aligned),
                                                     // columns 0..3 are the halo of the causal convolutions
  float* sx = smem;                                  // [CI][TP]   (absent when XG)
  float* sh = XG ? smem : sx + CI * TP;              // [C][TP]
  float* sw1 = sh + C * TP;                          // [CI][3][C]  (tap-major inside a channel, channels contiguous)
  float* sw2 = sw1 + CI * 3 * C;                     // [C][3][C]
  float* swsc = sw2 + C * 3 * C;                     // [CI][C]   1x1 block shortcut (optional)
  // weights once per CTA (they were a third of a row's load traffic when staged per row)
  for (int i = threadIdx.x; i < CI * 3 * C; i += blockDim.x) {       // w1 [C][CI][3] -> [CI][3][C]
    int ci = i / (3 * C), r = i - ci * 3 * C, k = r / C, co = r - k * C;
    sw1[i] = w1[(co * CI + ci) * 3 + k];
  }
  for (int i = threadIdx.x; i < C * 3 * C; i += blockDim.x) {
    int ci = i / (3 * C), r = i - ci * 3 * C, k = r / C, co = r - k * C;
    sw2[i] = w2[(co * C + ci) * 3 + k];
  }
  if (wsc)
    for (int i = threadIdx.x; i < CI * C; i += blockDim.x) swsc[i] = wsc[(i % C) * CI + i / C];       // [C][CI] -> [CI][C]
  const int t0 = threadIdx.x * 4;                    // first of this thread's 4 positions inside a segment
  const bool vec = (T & 3) == 0;
  const int Q = TT >> 2;
  for (int rr = 0; rr < rows_per_cta; ++rr) {
    const long long n = (long long)blockIdx.x * rows_per_cta + rr;
    if (n >= N) break;
    const float* xr = x + n * (long long)CI1 * T;
    const float* xr2 = x2 ? x2 + n * (long long)(CI - CI1) * T : nullptr;
   for (int tb = 0; tb < T; tb += TT) {
    const bool active = tb + t0 < T;
    __syncthreads();                                 // previous segment's / row's readers of sx / sh are done
    for (int i = threadIdx.x; i < ((XG ? 0 : CI) + C) * 4; i += blockDim.x) {   // halos of sx and sh (contiguous rows)
      float* r = smem + (i >> 2) * TP + (i & 3);
      *r = tb == 0 ? 0.0f : r[TT];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (XG ? 0 : CI * Q); i += blockDim.x) {
      const int ci = i / Q, q = i - ci * Q, t = tb + 4 * q;
      float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      const float* src = (ci < CI1 ? xr + (long long)ci * T : xr2 + (long long)(ci - CI1) * T) + t;
      if (vec) {
        if (t < T) v = *reinterpret_cast<const float4*>(src);
      } else {
        if (t < T) v.x = src[0];
        if (t + 1 < T) v.y = src[1];
        if (t + 2 < T) v.z = src[2];
        if (t + 3 < T) v.w = src[3];
      }
      *reinterpret_cast<float4*>(sx + ci * TP + 4 + 4 * q) = v;
    }
    __syncthreads();
    float acc[4][C];
    if (active) {
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int c = 0; c < C; ++c) acc[p][c] = b1[c];
      for (int ci = 0; ci < CI; ++ci) {
        float xv[6];                                                                    // x[t0-2 .. t0+3]
        if (XG) {
          const float* xc = (ci < CI1 ? xr + (long long)ci * T : xr2 + (long long)(ci - CI1) * T) + tb + t0;
          const float4 b = __ldg(reinterpret_cast<const float4*>(xc));
          float4 a = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
          if (tb + t0 >= 4) a = __ldg(reinterpret_cast<const float4*>(xc - 4));        // the neighbour's quad: an L1 hit
          xv[0] = a.z; xv[1] = a.w; xv[2] = b.x; xv[3] = b.y; xv[4] = b.z; xv[5] = b.w;
        } else {
          const float4 a = *reinterpret_cast<const float4*>(sx + ci * TP + t0);        // columns t0..t0+3 = x[t0-4..t0-1]
          const float4 b = *reinterpret_cast<const float4*>(sx + ci * TP + t0 + 4);
          xv[0] = a.z; xv[1] = a.w; xv[2] = b.x; xv[3] = b.y; xv[4] = b.z; xv[5] = b.w;
        }
        const float* w = sw1 + ci * 3 * C;
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
          for (int c = 0; c < C; c += 4) {
            const float4 wv = *reinterpret_cast<const float4*>(w + k * C + c);
#pragma unroll
            for (int p = 0; p < 4; ++p) {
              acc[p][c] = fmaf(wv.x, xv[p + k], acc[p][c]);
              acc[p][c + 1] = fmaf(wv.y, xv[p + k], acc[p][c + 1]);
              acc[p][c + 2] = fmaf(wv.z, xv[p + k], acc[p][c + 2]);
              acc[p][c + 3] = fmaf(wv.w, xv[p + k], acc[p][c + 3]);
            }
          }
      }
#pragma unroll
      for (int c = 0; c < C; ++c)
        *reinterpret_cast<float4*>(sh + c * TP + 4 + t0) = make_float4(acc[0][c], acc[1][c], acc[2][c], acc[3][c]);
    }
    __syncthreads();
    if (!active) continue;
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int c = 0; c < C; ++c) acc[p][c] = b2[c];
    for (int ci = 0; ci < C; ++ci) {
      float hv[6];
      {
        const float4 a = *reinterpret_cast<const float4*>(sh + ci * TP + t0);
        const float4 b = *reinterpret_cast<const float4*>(sh + ci * TP + t0 + 4);
        hv[0] = a.z; hv[1] = a.w; hv[2] = b.x; hv[3] = b.y; hv[4] = b.z; hv[5] = b.w;
      }
      const float* w = sw2 + ci * 3 * C;
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int c = 0; c < C; c += 4) {
          const float4 wv = *reinterpret_cast<const float4*>(w + k * C + c);
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            acc[p][c] = fmaf(wv.x, hv[p + k], acc[p][c]);
            acc[p][c + 1] = fmaf(wv.y, hv[p + k], acc[p][c + 1]);
            acc[p][c + 2] = fmaf(wv.z, hv[p + k], acc[p][c + 2]);
            acc[p][c + 3] = fmaf(wv.w, hv[p + k], acc[p][c + 3]);
          }
        }
    }
    const int tg = tb + t0;                          // global position of this thread's quad
    float* out = hn + n * (long long)C * T;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float m = 0.0f;
#pragma unroll
      for (int c = 0; c < C; ++c) m += acc[p][c];
      m *= (1.0f / C);
      float v = 0.0f;
#pragma unroll
      for (int c = 0; c < C; ++c) { float d = acc[p][c] - m; v = fmaf(d, d, v); }
      float r = rsqrtf(v * (1.0f / C) + 1e-5f);
#pragma unroll
      for (int c = 0; c < C; ++c) acc[p][c] = fmaf((acc[p][c] - m) * r, gamma[c], beta[c]);
    }
    if (a3 == nullptr) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        if (vec) {
          *reinterpret_cast<float4*>(out + c * T + tg) = make_float4(acc[0][c], acc[1][c], acc[2][c], acc[3][c]);
        } else {
#pragma unroll
          for (int p = 0; p < 4; ++p)
            if (tg + p < T) out[c * T + tg + p] = acc[p][c];
        }
      }
    } else {
      // the row as the split operand [hi | lo | hi | 1 1 0..] (K = C*T) of the fp16 tensor-core GEMM that follows
      const int K = C * T;
      __half* row = a3 + n * (long long)(3 * K + 8);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        __half2 h01 = __floats2half2_rn(acc[0][c], acc[1][c]), h23 = __floats2half2_rn(acc[2][c], acc[3][c]);
        float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        __half2 l01 = __floats2half2_rn(acc[0][c] - f01.x, acc[1][c] - f01.y);
        __half2 l23 = __floats2half2_rn(acc[2][c] - f23.x, acc[3][c] - f23.y);
        uint2 hi, lo;
        hi.x = *reinterpret_cast<uint32_t*>(&h01); hi.y = *reinterpret_cast<uint32_t*>(&h23);
        lo.x = *reinterpret_cast<uint32_t*>(&l01); lo.y = *reinterpret_cast<uint32_t*>(&l23);
        const int col = c * T + tg;
        if (vec) {
          *reinterpret_cast<uint2*>(row + col) = hi;
          *reinterpret_cast<uint2*>(row + K + col) = lo;
          *reinterpret_cast<uint2*>(row + 2 * K + col) = hi;
        } else {
          const __half hv4[4] = {__low2half(h01), __high2half(h01), __low2half(h23), __high2half(h23)};
          const __half lv4[4] = {__low2half(l01), __high2half(l01), __low2half(l23), __high2half(l23)};
#pragma unroll
          for (int p = 0; p < 4; ++p)
            if (tg + p < T) {
              row[col + p] = hv4[p];
              row[K + col + p] = lv4[p];
              row[2 * K + col + p] = hv4[p];
            }
        }
      }
      if (threadIdx.x == 0 && tb == 0) *reinterpret_cast<uint4*>(row + 3 * K) = make_uint4(0x3C003C00u, 0u, 0u, 0u);
    }
    if (wsc) {
      // the block's 1x1 shortcut W_sc x (ugnet.py:129, c_in != c_out) while x is still in shared memory: it becomes the
      // accumulator input of the up-sampling GEMM instead of a batched [c_out x c_in] matmul over the whole tensor
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int c = 0; c < C; ++c) acc[p][c] = 0.0f;
      for (int ci = 0; ci < CI; ++ci) {
        const float4 xb = XG ? __ldg(reinterpret_cast<const float4*>((ci < CI1 ? xr + (long long)ci * T : xr2 + (long long)(ci - CI1) * T) + tb + t0))
                             : *reinterpret_cast<const float4*>(sx + ci * TP + t0 + 4);
        const float* w = swsc + ci * C;
#pragma unroll
        for (int c = 0; c < C; c += 4) {
          const float4 wv = *reinterpret_cast<const float4*>(w + c);
          acc[0][c] = fmaf(wv.x, xb.x, acc[0][c]); acc[1][c] = fmaf(wv.x, xb.y, acc[1][c]);
          acc[2][c] = fmaf(wv.x, xb.z, acc[2][c]); acc[3][c] = fmaf(wv.x, xb.w, acc[3][c]);
          acc[0][c + 1] = fmaf(wv.y, xb.x, acc[0][c + 1]); acc[1][c + 1] = fmaf(wv.y, xb.y, acc[1][c + 1]);
          acc[2][c + 1] = fmaf(wv.y, xb.z, acc[2][c + 1]); acc[3][c + 1] = fmaf(wv.y, xb.w, acc[3][c + 1]);
          acc[0][c + 2] = fmaf(wv.z, xb.x, acc[0][c + 2]); acc[1][c + 2] = fmaf(wv.z, xb.y, acc[1][c + 2]);
          acc[2][c + 2] = fmaf(wv.z, xb.z, acc[2][c + 2]); acc[3][c + 2] = fmaf(wv.z, xb.w, acc[3][c + 2]);
          acc[0][c + 3] = fmaf(wv.w, xb.x, acc[0][c + 3]); acc[1][c + 3] = fmaf(wv.w, xb.y, acc[1][c + 3]);
          acc[2][c + 3] = fmaf(wv.w, xb.z, acc[2][c + 3]); acc[3][c + 3] = fmaf(wv.w, xb.w, acc[3][c + 3]);
        }
      }
      float* so = sc_out + n * (long long)C * T;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        if (vec) {
          *reinterpret_cast<float4*>(so + c * T + tg) = make_float4(acc[0][c], acc[1][c], acc[2][c], acc[3][c]);
        } else {
#pragma unroll
          for (int p = 0; p < 4; ++p)
            if (tg + p < T) so[c * T + tg + p] = acc[p][c];
        }
      }
    }
   }
  }
}

// NsDiff_spatial: the two heads of the graph denoiser and one NsDiff posterior step, fused.
//   e [N, DH, T] = UGnet's out block, channel-major -> eps = lin4(e^T), sigma = softplus(sigma_lin(softplus(e^T)))
//   (models/Diffusion_model/NsDiff/ugnet.py:290-292), then p_sample / p_sample_t_1to0
//   (models/Diffusion_model/NsDiff/nsdiff_utils.py:111-158, 209-239) on [N, T, F].  y == nullptr: heads only.
__global__ void nsx_step_kernel(const float* __restrict__ e, const float* __restrict__ w4, const float* __restrict__ b4,
                                const float* __restrict__ ws, const float* __restrict__ bs, const float* __restrict__ y,
                                const float* __restrict__ yT, const float* __restrict__ gx, const float* __restrict__ z,
                                const float* __restrict__ sched, int n_steps, int t, long long N, int DH, int T, int F,
                                float* __restrict__ out, float* __restrict__ eps_out, float* __restrict__ sig_out) {
  const UpdNsStep st = upd_ns_step(sched, n_steps, t);
  const long long total = N * T * F;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i % F);
    const long long np = i / F;
    const int p = (int)(np % T);
    const long long n = np / T;
    const float* er = e + n * DH * T + p;
    float ae = 0.0f, as = 0.0f;
    for (int c = 0; c < DH; ++c) {
      const float v = er[(long long)c * T];
      ae = fmaf(v, w4[f * DH + c], ae);
      as = fmaf(upd_softplus_accurate(v), ws[f * DH + c], as);
    }
    const float eps = __fadd_rn(ae, b4[f]);
    const float sig = upd_softplus_accurate(__fadd_rn(as, bs[f]));
    if (eps_out) eps_out[i] = eps;
    if (sig_out) sig_out[i] = sig;
    if (y) out[i] = upd_ns_update(st, y[i], yT[i], gx[i], eps, sig, z ? z[i] : 0.0f, t == 0);
  }
}


// Narrow convolutions of the U-Net along the time axis (ugnet.py:152 DownSample Conv2d (1,3)/(1,2)/(0,1), :171 UpSample
// ConvTranspose2d (1,4)/(1,2)/(0,1), :245-246 the 1x1 x_proj / out.0): x [N, CI, Tin] -> y [N, CO, Tout], a handful of
// channels, so a thread owns one output position and up to 8 output channels; weights broadcast from shared memory.
//   transposed == 0:  y[co][t] = b[co] + sum_ci sum_k w[co][ci][k] * x[ci][t*stride + k - pad]
//   transposed == 1:  y[co][t] = b[co] + sum_ci sum_k w[ci][co][k] * x[ci][(t + pad - k) / stride]   (where divisible)
// HBM / L2-bound (a few FMAs per byte).  These were cuDNN implicit-GEMM / dgrad / magma launches (10 % of a DiffSTG step).
template <int CB>   // output channels a thread carries at once (1, 4 or 8)
__global__ void __launch_bounds__(128) stg_conv1d_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ b, long long N, int CI, int CO, int Tin,
                                                         int Tout, int K, int stride, int pad, int transposed,
                                                         int rows_per_cta, float* __restrict__ y) {
  extern __shared__ float sw[];                       // [CI][K][CO]: channels-out contiguous
  for (int i = threadIdx.x; i < CI * K * CO; i += blockDim.x) {
    const int co = i % CO, r = i / CO, k = r % K, ci = r / K;
    sw[i] = transposed ? w[(ci * CO + co) * K + k] : w[(co * CI + ci) * K + k];
  }
  __syncthreads();
  const long long n0 = (long long)blockIdx.x * rows_per_cta;
  const long long n1 = n0 + rows_per_cta < N ? n0 + rows_per_cta : N;
  // a CTA walks its rows as one flat index space of output positions
  const int total = (int)(n1 - n0) * Tout;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int r = idx / Tout, t = idx - r * Tout;
    const long long n = n0 + r;
    const float* xr = x + n * (long long)CI * Tin;
    float* yr = y + n * (long long)CO * Tout;
    // input positions of the K taps (-1: outside / not on the stride grid)
    for (int c0 = 0; c0 < CO; c0 += CB) {
      float acc[CB];
#pragma unroll
      for (int j = 0; j < CB; ++j) acc[j] = (b && c0 + j < CO) ? b[c0 + j] : 0.0f;
      for (int k = 0; k < K; ++k) {
        int src;
        if (!transposed) {
          src = t * stride + k - pad;
        } else {
          const int u = t + pad - k;
          src = (u >= 0 && u % stride == 0) ? u / stride : -1;
        }
        if (src < 0 || src >= Tin) continue;
        for (int ci = 0; ci < CI; ++ci) {
          const float xv = __ldg(xr + ci * Tin + src);
          const float* wk = sw + (ci * K + k) * CO + c0;
#pragma unroll
          for (int j = 0; j < CB; ++j)
            if (c0 + j < CO) acc[j] = fmaf(wk[j], xv, acc[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < CB; ++j)
        if (c0 + j < CO) yr[(c0 + j) * Tout + t] = acc[j];
    }
  }
}

inline unsigned stream_grid(long long n, int block, int sms) {
  long long g = (n + block - 1) / block;
  long long cap = (long long)sms * 16;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

cudaError_t upd_launch_stg_posterior(const float* xt, const float* pred, const float* z, long long n, float a, float b,
                                     float c, float* out, int sms, cudaStream_t stream) {
  stg_posterior_kernel<<<stream_grid(n, 256, sms), 256, 0, stream>>>(xt, pred, z, n, a, b, c, out);
  return cudaGetLastError();
}

cudaError_t upd_launch_nsx_step(const float* e, const float* w4, const float* b4, const float* ws, const float* bs,
                                const float* y, const float* yT, const float* gx, const float* z, const float* sched,
                                int n_steps, int t, long long N, int DH, int T, int F, float* out, float* eps_out,
                                float* sig_out, int sms, cudaStream_t stream) {
  nsx_step_kernel<<<stream_grid(N * T * F, 256, sms), 256, 0, stream>>>(e, w4, b4, ws, bs, y, yT, gx, z, sched, n_steps, t, N,
                                                                        DH, T, F, out, eps_out, sig_out);
  return cudaGetLastError();
}

cudaError_t upd_launch_stg_gated_aggregate(const float* kqvs, const int* rowptr, const int* col, const float* bias,
                                           long long N, int V, int C, int relu, float* out, int sms, cudaStream_t stream) {
  // replicas of up to 768 nodes: q / v slab of the replica in shared memory (UPD_STG_AGG=gather forces the gather kernel)
  constexpr int CS = 32;
  static const bool gather_env = getenv("UPD_STG_AGG") && getenv("UPD_STG_AGG")[0] == 'g';
  const size_t smem = sizeof(float) * 2 * (size_t)V * CS;
  const long long reps = N / V;
  if (!gather_env && smem <= 192 * 1024 && reps <= 0x7fffffffLL && (C + CS - 1) / CS <= 65535) {
    auto kern = stg_gated_aggregate_smem_kernel<CS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)reps, (unsigned)((C + CS - 1) / CS));
    kern<<<grid, 256, smem, stream>>>(kqvs, rowptr, col, bias, V, C, relu, out);
    return cudaGetLastError();
  }
  stg_gated_aggregate_kernel<<<stream_grid(N * C, 256, sms), 256, 0, stream>>>(kqvs, rowptr, col, bias, N, V, C, relu, out);
  return cudaGetLastError();
}

cudaError_t upd_launch_stg_conv1d(const float* x, const float* w, const float* b, long long N, int CI, int CO, int Tin, int Tout,
                                   int K, int stride, int pad, int transposed, float* y, int sms, cudaStream_t stream) {
  const size_t smem = sizeof(float) * (size_t)CI * K * CO;
  if (smem > 48 * 1024) return cudaErrorInvalidValue;
  // ~8 CTAs per SM, every CTA stages the weights once and walks a block of rows
  long long ctas = (long long)sms * 8;
  if (ctas > N) ctas = N;
  const int rpc = (int)((N + ctas - 1) / ctas);
  if ((long long)rpc * Tout > 0x7fffffffLL) return cudaErrorInvalidValue;
  const unsigned grid = (unsigned)((N + rpc - 1) / rpc);
  if (CO >= 8)
    stg_conv1d_kernel<8><<<grid, 128, smem, stream>>>(x, w, b, N, CI, CO, Tin, Tout, K, stride, pad, transposed, rpc, y);
  else if (CO >= 2)
    stg_conv1d_kernel<4><<<grid, 128, smem, stream>>>(x, w, b, N, CI, CO, Tin, Tout, K, stride, pad, transposed, rpc, y);
  else
    stg_conv1d_kernel<1><<<grid, 128, smem, stream>>>(x, w, b, N, CI, CO, Tin, Tout, K, stride, pad, transposed, rpc, y);
  return cudaGetLastError();
}

cudaError_t upd_launch_stg_tcn_mma(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                                   const float* gamma, const float* beta, long long N, int CI, int C, int T, float* hn,
                                   void* a3, const float* wsc, float* sc_out, const float* x2, int CI1, int sms,
                                   cudaStream_t stream);

cudaError_t upd_launch_stg_tcn_ln(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                                  const float* gamma, const float* beta, long long N, int CI, int C, int T, float* hn,
                                  void* a3, const float* wsc, float* sc_out, const float* x2, int CI2, int sms,
                                  cudaStream_t stream) {
  if ((x2 == nullptr) != (CI2 == 0) || CI2 < 0 || CI2 >= CI || (reinterpret_cast<uintptr_t>(x2) & 15) != 0) return cudaErrorInvalidValue;
  int threads = ((T + 3) / 4 + 31) / 32 * 32;          // one segment of 4*threads positions at a time
  if (threads > 128) threads = 128;
  if ((T & 1) != 0 || T < 2 || CI < 1 || N > 0x7fffffffLL) return cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(hn) & 15) != 0 || (reinterpret_cast<uintptr_t>(a3) & 15) != 0) return cudaErrorInvalidValue;
  if (hn == nullptr && a3 == nullptr) return cudaErrorInvalidValue;
  const int seg = T <= 4 * threads ? ((T + 3) & ~3) : 4 * threads;
  // 16-channel blocks and segmented rows: x is read through L1 instead of being staged (measured: 16->16 ch x T=200
  // 0.98 -> 0.91 ms, 32->16 ch 1.77 -> 1.41 ms, 8 ch x T=2000 0.83 -> 0.75 ms; 8 ch x T=400 is faster staged, 0.46 vs
  // 0.54 ms).  UPD_TCN_XG=0/1 forces one path (measurement only).
  static const int xg_env = getenv("UPD_TCN_XG") ? atoi(getenv("UPD_TCN_XG")) : -1;
  const bool xg = (T & 3) == 0 && (xg_env >= 0 ? xg_env != 0 : (C == 16 || T > 4 * threads));
  const size_t smem = sizeof(float) * ((size_t)((xg ? 0 : CI) + C) * (seg + 4) + (size_t)CI * 3 * C + (size_t)C * 3 * C + (size_t)CI * C);
  if ((wsc == nullptr) != (sc_out == nullptr) || (reinterpret_cast<uintptr_t>(sc_out) & 15) != 0) return cudaErrorInvalidValue;
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) return cudaErrorInvalidValue;
  // Tensor-core kernel (csrc/stg_tcn_mma.cu) for rows of up to 512 positions, c_out >= 8 and 8 <= c_in <= 32; the FFMA kernel below keeps
  // the other shapes (narrow blocks, long segmented rows).  UPD_TCN_IMPL=ffma forces the FFMA kernel (measurement / tests).
  static const bool ffma_env = getenv("UPD_TCN_IMPL") && getenv("UPD_TCN_IMPL")[0] == 'f';
  if (!ffma_env) {
    cudaError_t e = upd_launch_stg_tcn_mma(x, w1, b1, w2, b2, gamma, beta, N, CI, C, T, hn, a3, wsc, sc_out, x2, CI - CI2, sms, stream);
    if (e != cudaErrorNotSupported) return e;
  }
  const int rpc = N >= 8192 ? 8 : 1;                   // rows per CTA: amortises the weight staging on big launches
  const unsigned grid = (unsigned)((N + rpc - 1) / rpc);
#define UPD_TCN_LAUNCH(CC, XX)                                                                                         \
  {                                                                                                                    \
    cudaError_t e = cudaFuncSetAttribute(stg_tcn_ln_kernel<CC, XX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return e;                                                                                    \
    stg_tcn_ln_kernel<CC, XX><<<grid, threads, smem, stream>>>(x, w1, b1, w2, b2, gamma, beta, N, CI, T, rpc, hn, (__half*)a3, wsc, sc_out, x2, CI - CI2); \
  }
#define UPD_TCN_CASE(CC)                                                                                               \
  case CC:                                                                                                             \
    if (xg) UPD_TCN_LAUNCH(CC, true) else UPD_TCN_LAUNCH(CC, false)                                                    \
    break;
  switch (C) {
    UPD_TCN_CASE(4)
    UPD_TCN_CASE(8)
    UPD_TCN_CASE(16)
    default: return cudaErrorInvalidValue;
  }
#undef UPD_TCN_LAUNCH
#undef UPD_TCN_CASE
  return cudaGetLastError();
}
