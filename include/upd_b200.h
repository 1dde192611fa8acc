/*
 * upd_b200.h -- C ABI of the B200-native uncertainty-inference hot path
 * (conditional reverse-diffusion sampling + MPV / gx reduction).
 *
 * The reference has no FFI layer: its boundary is the Python surface between
 * evaluation_and_analysis/diffusion_model_uncertainy.py and models/Diffusion_model/ (every file).
 * Each entry point below replaces one reference function at that surface; the reference
 * file:line it stands in for is cited per function.  INTEGRATION.md shows the ctypes stub a
 * reference maintainer would add.
 *
 * Conventions
 *   - Plain pointers and sizes only.  Every `*_dev` pointer is CUDA device memory owned by the
 *     caller; the library never allocates, frees, or keeps state between calls (re-entrant per
 *     stream; one process per GPU can run concurrently).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All work is
 *     enqueued asynchronously on it; nothing synchronises the device.
 *   - Return value: 0 = UPD_OK, otherwise a UPD_ERR_* code; upd_error_string() names it.  The
 *     Python host layer raises RuntimeError on non-zero.
 *   - All arithmetic is fp32 ("dtype f32").  Tensor-core contractions use an fp16 hi/lo split with
 *     three tcgen05 passes per product and fp32 accumulation (see DESIGN.md "Numerics").
 */
#ifndef UPD_B200_H
#define UPD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  UPD_OK = 0,
  UPD_ERR_BAD_ARG = 1,       /* NULL pointer, non-positive size, inconsistent sizes            */
  UPD_ERR_UNSUPPORTED = 2,   /* shape outside what the kernels are built for (see each call)   */
  UPD_ERR_CUDA = 3,          /* a CUDA runtime call failed; see upd_last_cuda_error()          */
  UPD_ERR_NO_DEVICE = 4      /* not an sm_100 device                                           */
};

/* Denoiser families sharing the fused sampler. */
enum {
  UPD_KIND_NSDIFF = 0,       /* models/Diffusion_model/NsDiff/denoise.py:23-51                 */
  UPD_KIND_TMDM = 1          /* models/Diffusion_model/TMDM/tmdm_model.py:23-64                */
};

/* Sampler implementations (same results within fp32 round-off). */
enum {
  UPD_IMPL_TCGEN05 = 0,      /* tcgen05/TMEM tensor-core sampler (default, the product path): the
                                warp-specialised kernel for F <= 2, the two-tile kernel for F = 3, 4;
                                a step count whose tables do not fit in shared memory (T > ~40) runs
                                on the FFMA kernel instead of failing                          */
  UPD_IMPL_SIMT = 1,         /* fp32 FFMA kernel (bring-up / cross-check of the tensor path)   */
  UPD_IMPL_TCGEN05_X2 = 2,   /* forced: two tiles per SM taking MUFU turns (sampler_tc.cu);
                                UPD_ERR_UNSUPPORTED when shared memory is exceeded (T > ~40)   */
  UPD_IMPL_TCGEN05_WS = 4    /* forced: warp-specialised kernel, 16 epilogue warps + 4 row warps
                                (sampler_ws.cu); UPD_ERR_UNSUPPORTED for F > 2                 */
};

const char* upd_error_string(int code);
/* cudaError_t of the last failing runtime call made by this library on the calling thread. */
int upd_last_cuda_error(void);
/* Library ABI version (bumped when a signature or the packed-weights layout changes). */
int upd_abi_version(void);

/* ------------------------------------------------------------------------------------------
 * Weights.  Host-side, plain fp32 row-major arrays exactly as they sit in the reference
 * state dict (keys `model.diffussion_model.*`):
 *   lin{1,2,3}_w [128, in] (in = 3F NsDiff / 2F TMDM for lin1, 128 for lin2/3), lin*_b [128],
 *   embed{1,2,3} [TE, 128] (TE = T for NsDiff, T+1 for TMDM), lin4_w [F,128], lin4_b [F],
 *   sigma_w [F,128] / sigma_b [F] (NsDiff only; NULL for TMDM),
 *   sched [n_sched, T]: NsDiff n_sched = 10 rows in the order p_sample_loop receives them
 *     (nsdiff_utils.py:271): alphas, one_minus_alphas_bar_sqrt, alphas_cumprod,
 *     alphas_cumprod_sum, alphas_cumprod_prev, alphas_cumprod_sum_prev, betas_tilde, betas_bar,
 *     betas_tilde_m_1, betas_bar_m_1 -- copied from the reference-built fp32 tables, never
 *     recomputed; TMDM n_sched = 2: alphas, one_minus_alphas_bar_sqrt (tmdm_diffusion_utils.py:107).
 * ------------------------------------------------------------------------------------------ */
typedef struct UpdDenoiserWeights {
  int kind;            /* UPD_KIND_*            */
  int F;               /* dataset_nf, 1..4      */
  int T;               /* diffusion_steps, 2..64 */
  const float* lin1_w; const float* lin1_b; const float* embed1;
  const float* lin2_w; const float* lin2_b; const float* embed2;
  const float* lin3_w; const float* lin3_b; const float* embed3;
  const float* lin4_w; const float* lin4_b;
  const float* sigma_w; const float* sigma_b;
  const float* sched;
} UpdDenoiserWeights;

/* Bytes of the packed blob for (kind, F, T); 0 if unsupported. */
size_t upd_denoiser_pack_bytes(int kind, int F, int T);
/* Pack host weights into `out_host` (capacity >= upd_denoiser_pack_bytes).  The caller uploads
 * the blob once per model load (replaces nothing in the reference: this is the B200 layout of
 * what utils/utils.py:660-689 load_diffusion_model materialises as nn.Parameters). */
int upd_denoiser_pack(const UpdDenoiserWeights* w, void* out_host, size_t capacity);

/* ------------------------------------------------------------------------------------------
 * upd_nsdiff_sample -- replaces p_sample_loop + the K//S chunk loop of evaluation_step
 *   (models/Diffusion_model/NsDiff/nsdiff_utils.py:271-284, :111-158, :209-239;
 *    NsDiff_model.py:227-257 / :461-487) for a batch of n_win windows of B rows each.
 *
 *   packed_dev   blob from upd_denoiser_pack (kind NSDIFF), on the device
 *   y0_hat_dev   [n_win*B, O, F] condition mean f(x); NULL = zeros (variants without f(x),
 *                NsDiff_model.py:446).  It is also the prior mean y_T_mean (:231 / :464).
 *   gx_dev       [n_win*B, O, F] condition variance g(x) (EPS already added where the
 *                reference adds it, NsDiff_model.py:450)
 *   K            samples to draw per row (= (n_z_samples // parallel_sample) * parallel_sample)
 *   S            parallel_sample: only fixes the layout of injected noise and the Philox-free
 *                reference ordering; results do not depend on it in Philox mode
 *   seed, window_base   Philox4x32-10 key and the global index of the first window of this
 *                batch: noise is keyed by (seed, global window, row, sample, position, feature,
 *                draw) so a sweep gives identical samples however it is split over launches/GPUs
 *   noise_dev    validation mode: NULL, or injected N(0,1) draws laid out as the reference
 *                consumes them, [n_win, K/S, T, B*S, O, F] (draw 0 = y_T, draw i = step t=T-i)
 *   out_dev      [n_win*B, K, O, F]  (the reference's cache element [B,O,F,K] is a permuted
 *                view of exactly this memory, NsDiff_model.py:259-262)
 *   impl         UPD_IMPL_*
 * Limits: F in 1..4, T in 2..64, O*F*4 bytes 16-byte aligned not required.
 * ------------------------------------------------------------------------------------------ */
int upd_nsdiff_sample(const void* packed_dev, const float* y0_hat_dev, const float* gx_dev,
                      int n_win, int B, int K, int S, int O, int F, int T,
                      uint64_t seed, uint64_t window_base, const float* noise_dev,
                      float* out_dev, int impl, void* stream);

/* upd_tmdm_sample -- replaces TMDM p_sample_loop + chunk loop
 *   (models/Diffusion_model/TMDM/tmdm_diffusion_utils.py:57-119, tmdm_adapter.py:132-151).
 *   y0_hat_dev [n_win*B, Lr, F] with Lr = label_len + pred_len rows per trajectory; unit-variance
 *   prior around it; out_dev [n_win*B, K, Lr, F] (caller slices the last pred_len positions).
 *   noise_dev layout [n_win, K/S, T, B*S, Lr, F]. */
int upd_tmdm_sample(const void* packed_dev, const float* y0_hat_dev,
                    int n_win, int B, int K, int S, int Lr, int F, int T,
                    uint64_t seed, uint64_t window_base, const float* noise_dev,
                    float* out_dev, int impl, void* stream);

/* ------------------------------------------------------------------------------------------
 * upd_mpv_reduce -- replaces summarize_pred_future_list / summarize_slbp_sensitivity /
 *   summarize_slbp_sampling_for_fig6 (diffusion_model_uncertainy.py:286-303, :529-541, :701-713):
 *   biased variance over the K trajectories (Welford, warp-shuffle merge), then means.
 *
 *   traj_dev     [n_win*B, K, O, F] as written by the samplers
 *   scale_dev    NULL, or [2,F] = (scaler_mean, scaler_std): trajectories are mapped to raw units
 *                x*std+mean before the statistics (_feature_inverse_transform, :267-283)
 *   var_dev      NULL or [n_win*B, O, F] per-position predictive variance
 *   mean_dev     NULL or [n_win*B, O, F] per-position predictive mean (prediction-error path, :542-549)
 *   mpv_dev      [n_win]   mean of var over (B,O,F)       -> `ews` of uncertainty_ews
 *   pmean_dev    [n_win]   mean of all samples            -> `pred_mean`
 *   mpv_f_dev    [n_win,F] mean of var over (B,O) per feature -> SLBP MPV picks [pred_dim]
 *   scratch_dev  >= upd_mpv_scratch_bytes(n_win, B, O, F) bytes
 * ------------------------------------------------------------------------------------------ */
size_t upd_mpv_scratch_bytes(int n_win, int B, int O, int F);
int upd_mpv_reduce(const float* traj_dev, const float* scale_dev, int n_win, int B, int K, int O, int F,
                   float* var_dev, float* mean_dev, float* mpv_dev, float* pmean_dev, float* mpv_f_dev,
                   void* scratch_dev, void* stream);

/* upd_gram_centered -- the contraction of `_slbp_intrinsic_dimension` (diffusion_model_uncertainy.py:686-698), batched over
 *   windows: traj_dev [n_win, K, D = O*F] (one SLBP cache element per window, as the samplers write it) ->
 *   gram_dev [n_win, K, K] (double) = C C^T / (K - 1), C = the K trajectories minus their mean.  It has the non-zero
 *   spectrum of the (O*F)^2 sample covariance the reference diagonalises; the caller counts eigenvalues up to 80 %.
 *   Limits: 2 <= K <= 128. */
int upd_gram_centered(const float* traj_dev, int n_win, int K, int D, double* gram_dev, void* stream);

/* upd_prediction_error -- `summarize_slbp_sensitivity`'s error term (diffusion_model_uncertainy.py:542-549), batched over
 *   windows: err_dev [n_win, F] = mean over the O positions of | mean_dev - target_dev |, both [n_win, O, F] (mean_dev =
 *   the per-position predictive means of upd_mpv_reduce, target_dev = the scaled future of each window).  F in 1..4. */
int upd_prediction_error(const float* mean_dev, const float* target_dev, int n_win, int O, int F, float* err_dev,
                         void* stream);

/* ------------------------------------------------------------------------------------------
 * upd_sigma_estimation -- replaces SigmaEstimation.forward, i.e. cond_pred_model_g, the whole
 *   "gx" uncertainty path (models/Diffusion_model/NsDiff/g_backbone.py:49-72, sigma.py:34-71).
 *   x_dev [rows, L, F] scaled windows; weights as in the state dict `cond_pred_model_g.mlp.*`
 *   (row-major fp32, device): w0 [H, L-R], b0 [H], ln1_w/ln1_b [F,H], w3 [H,H], b3 [H],
 *   ln2_w/ln2_b [F,H], w6 [O,H], b6 [O].  add_eps is added to the result (1e-7 where the
 *   reference adds EPS, else 0).  gx_dev [rows, O, F].  Limits: F in 1..4, O <= H.
 * ------------------------------------------------------------------------------------------ */
typedef struct UpdSigmaWeights {
  const float* w0; const float* b0; const float* ln1_w; const float* ln1_b;
  const float* w3; const float* b3; const float* ln2_w; const float* ln2_b;
  const float* w6; const float* b6;
} UpdSigmaWeights;
int upd_sigma_estimation(const UpdSigmaWeights* w_dev_ptrs, const float* x_dev, int rows, int L, int R,
                         int F, int H, int O, float add_eps, float* gx_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * DiffusionTS conditional sampler.  The x0-predicting transformer runs above this ABI
 * (library GEMMs + these kernels); the entry points below are the per-step algebra of
 * Diffusion_TS.fast_sample_infill (models/Diffusion_model/DiffusionTS/DiffusionTS.py:277-310) and the
 * Fourier seasonal head.  n = number of fp32 elements; every pointer is device memory.
 * ------------------------------------------------------------------------------------------ */

/* upd_dts_ddim_step -- replaces model_predictions' clamp / predict_noise_from_start and the DDIM mean
 *   (DiffusionTS.py:152-160, :294-303):  x_start = clamp(x0_raw,-1,1);
 *   pred_noise = (sqrt_recip_ac*img - x_start)/sqrt_recipm1_ac; pred_mean = x_start*sqrt_alpha_next + c*pred_noise;
 *   img_out = pred_mean + sigma*noise.  last != 0 (time_next < 0, :291-293): img_out = x_start.
 *   x_start_dev / pred_mean_dev may be NULL; noise_dev may be NULL when sigma == 0 (eta = 0). */
int upd_dts_ddim_step(const float* x0_raw_dev, const float* img_dev, long long n,
                      float sqrt_recip_ac, float sqrt_recipm1_ac, float sqrt_alpha_next, float c, float sigma,
                      const float* noise_dev, int last, float* x_start_dev, float* pred_mean_dev, float* img_out_dev,
                      void* stream);

/* upd_dts_adagrad_step -- replaces one iteration of langevin_fn's optimiser (DiffusionTS.py:384-401): a
 *   torch.optim.Adagrad created anew each iteration, i.e. p -= lr * g / (sqrt(g*g) + 1e-10), in place. */
int upd_dts_adagrad_step(float* p_dev, const float* grad_dev, long long n, float lr, void* stream);

/* upd_dts_infill -- replaces `sample[~mask] = refined[~mask]` (:404) followed by
 *   `img[mask] = q_sample(target, t)[mask]` (:305-306, q_sample :232-237) for the mask evaluation_step builds
 *   (first L_obs positions observed, DiffusionTS_model.py:47-54).  target_dev [rows, L_obs, F];
 *   noise_dev [rows, seq, F] = the full draw q_sample makes, or NULL: observed part = target (final
 *   overwrite, :308).  img_dev [rows, seq, F] is written, refined_dev [rows, seq, F] is read (may alias img). */
int upd_dts_infill(float* img_dev, const float* refined_dev, const float* target_dev, const float* noise_dev,
                   long long rows, int seq, int L_obs, int F, float sqrt_ac, float sqrt_one_minus_ac, void* stream);

/* upd_gauss_fill -- N(0,1) draws (Philox4x32-10 + Box-Muller) keyed by (seed, row_base + row, element, draw):
 *   production replacement of torch.randn / randn_like in the DiffusionTS and DiffSTG loops, independent of how
 *   rows are batched per launch or sharded over GPUs.  out_dev [rows, row_elems]. */
int upd_gauss_fill(float* out_dev, long long rows, long long row_elems, uint64_t seed, uint64_t row_base,
                   uint32_t draw, void* stream);

/* upd_dts_fourier_topk / _bwd -- replaces FourierLayer.forward (diffusionts_transformer.py:52-103) and its
 *   autograd backward.  spec_dev [rows, >= 2*NF, D] with row stride spec_row_stride floats: planes 0..NF-1 = Re,
 *   NF..2NF-1 = Im of rfft bins low..low+NF-1 along the sequence axis (the caller folds the DFT into the 1x1
 *   projection GEMM that precedes it).  top_k = int(log(NF)) <= 8 bins of largest |X| per (row, channel) are kept and
 *   resynthesised over seq positions; season_dev [rows, seq, D] is overwritten or accumulated into; idx_dev
 *   [rows, top_k, D] (int32) records the selection for the backward.  _bwd: gspec_dev must be zero-filled. */
int upd_dts_fourier_topk(const float* spec_dev, long long spec_row_stride, long long rows, int NF, int low, int seq,
                         int D, int top_k, int accumulate, float* season_dev, int* idx_dev, void* stream);
int upd_dts_fourier_topk_bwd(const float* gseason_dev, const int* idx_dev, long long gspec_row_stride, long long rows,
                             int NF, int low, int seq, int D, int top_k, float* gspec_dev, void* stream);

/* upd_dts_layernorm / _bwd -- replaces AdaLayerNorm (diffusionts_model_utils.py:187-202) and the blocks' nn.LayerNorm
 *   (diffusionts_transformer.py:215, 287) forward and input-gradient backward: y = LayerNorm(x) * gamma + beta over the
 *   last axis (eps 1e-5); gamma_dev / beta_dev [D] (AdaLN: 1 + scale[t] and shift[t], the step is shared by all rows).
 *   stats_dev [rows, 2] = (mean, rstd), written by the forward (may be NULL) and read by the backward, which returns
 *   dx only (the weights are constants during sampling).  D in {32,64,96,128,192,256,384,512,1024}.
 *   a3_dev (may be NULL): additionally emit the fp16 split operand [rows, 3D + 8] = [hi | lo | hi | 1 1 0..] that upd_gemm3
 *   consumes, so that a dense layer reading y needs no separate upd_fx_split pass; y_dev may then be NULL (the fp32 y is
 *   not written at all).  With a3_dev: D a multiple of 64 in {64,...,512,1024}. */
int upd_dts_layernorm(const float* x_dev, const float* gamma_dev, const float* beta_dev, long long rows, int D,
                      float* y_dev, float* stats_dev, void* a3_dev, void* stream);
int upd_dts_layernorm_bwd(const float* x_dev, const float* dy_dev, const float* gamma_dev, const float* stats_dev,
                          long long rows, int D, float* dx_dev, void* stream);

/* upd_dts_attention / _bwd -- replaces FullAttention / CrossAttention (diffusionts_transformer.py:126-203; head size 16)
 *   and their autograd backward (the refinement gradient of langevin_fn, DiffusionTS.py:384-399):
 *   o = softmax(scale * q k^T) v per (row r, head h), heads merged in o_dev [R*Lq, H*16].  q_dev: position (r*Lq + i) at
 *   q_dev + (r*Lq+i)*q_row_stride floats, head h at + h*16; k_dev / v_dev likewise over (r*S + j) with kv_row_stride, so
 *   Q|K|V may be one fused projection buffer.  lse_dev [R*H, Lq] (may be NULL in the forward) keeps the base-2
 *   log-sum-exp the backward needs; _bwd writes dq / dk / dv with the same addressing (dq_row_stride, dkv_row_stride).
 *   Runs on tcgen05 tensor cores (csrc/dts_attention_tc.cu: fp16 hi/lo operands, fp32 accumulation in TMEM; the backward
 *   scales each (row, head) tile of do_dev by a power of two, so cotangents of any magnitude keep fp32-grade accuracy)
 *   for S <= 224 (forward) / S, Lq <= 224 (backward); longer sequences run the fp32 FFMA kernels (csrc/dts_attention.cu;
 *   also with the environment variable UPD_DTS_ATTN_FFMA=1).
 *   a3_dev (forward, may be NULL): additionally emit the split operand [R*Lq, 3*H*16 + 8] of the out-projection GEMM.
 *   Limits: head_dim == 16; (S + Lq) * 136 bytes of shared memory on the FFMA path. */
int upd_dts_attention(const float* q_dev, long long q_row_stride, const float* k_dev, const float* v_dev,
                      long long kv_row_stride, int R, int H, int Lq, int S, int head_dim, float scale, float* o_dev,
                      float* lse_dev, void* a3_dev, void* stream);
int upd_dts_attention_bwd(const float* q_dev, long long q_row_stride, const float* k_dev, const float* v_dev,
                          long long kv_row_stride, int R, int H, int Lq, int S, int head_dim, float scale,
                          const float* o_dev, const float* lse_dev, const float* do_dev, float* dq_dev,
                          long long dq_row_stride, float* dk_dev, float* dv_dev, long long dkv_row_stride, void* stream);

/* ------------------------------------------------------------------------------------------
 * DiffSTG graph-conv sampler.
 * ------------------------------------------------------------------------------------------ */

/* upd_stg_posterior -- replaces gaussian_posterior (models/Diffusion_model/DiffSTG/graph_diffusion_model.py:46-73):
 *   out = a*(xt - b*pred) + c*w with w = z_dev (DDPM branch, t <= 1: a = 1/sqrt(alpha_t),
 *   b = (1-alpha_t)/sqrt(1-abar_t), c = sqrt(beta_tilde)) or w = pred (z_dev NULL, DDIM: a = sqrt(abar_target/abar_t),
 *   b = sqrt(1-abar_t), c = sqrt(1-abar_target)).  The caller forms a, b, c in float64 like the reference. */
int upd_stg_posterior(const float* xt_dev, const float* pred_dev, const float* z_dev, long long n,
                      float a, float b, float c, float* out_dev, void* stream);

/* upd_nsx_step -- NsDiff_spatial (models/Diffusion_model/NsDiff/NsDiff_model.py:695-790): replaces the two heads of the
 *   graph denoiser (models/Diffusion_model/NsDiff/ugnet.py:290-292: eps = lin4(e), sigma = softplus(sigma_lin(
 *   softplus(e)))) and the posterior step that consumes them (p_sample, models/Diffusion_model/NsDiff/nsdiff_utils.py:
 *   111-158; t == 0: p_sample_t_1to0, :209-239, no noise).  e_dev [N, DH, T] is UGnet's out block, channel-major;
 *   w4/ws [F, DH], b4/bs [F]; y/yT/gx/z/out [N, T, F]; sched_dev [10, n_steps] rows in the order of upd_nsdiff_sample.
 *   z_dev must be NULL exactly when t == 0.  y_dev NULL: heads only, written to eps_out_dev / sig_out_dev (either may
 *   be NULL otherwise). */
int upd_nsx_step(const float* e_dev, const float* w4_dev, const float* b4_dev, const float* ws_dev, const float* bs_dev,
                 const float* y_dev, const float* yT_dev, const float* gx_dev, const float* z_dev, const float* sched_dev,
                 int n_steps, int t, long long N, int DH, int T, int F, float* out_dev, float* eps_out_dev,
                 float* sig_out_dev, void* stream);

/* upd_stg_gated_aggregate -- replaces SpatialBlock's relu(ResGatedGraphConv(x, edge_index))
 *   (models/Diffusion_model/DiffSTG/ugnet.py:36-45; torch_geometric 2.5.3 layer, bias=True, root_weight=True) after
 *   its four projections, and duplicate_edge_index (graph_diffusion_model.py:77-84):
 *   kqvs_dev [N, 4C] = (key | query | value | skip) rows from ONE fused GEMM, N = n_rep * V nodes, replica r owns
 *   nodes r*V..r*V+V-1 and shares the CSR of the V-node graph: rowptr_dev [V+1], col_dev [E] = sources j of the
 *   edges j -> i, in edge order.  out[n,:] = act(sum_j sigmoid(k_n + q_j) * v_j + skip_n + bias). */
int upd_stg_gated_aggregate(const float* kqvs_dev, const int* rowptr_dev, const int* col_dev, const float* bias_dev,
                            long long N, int V, int C, int relu, float* out_dev, void* stream);

/* upd_stg_tcn_ln -- replaces the front half of ResidualBlock.forward (models/Diffusion_model/DiffSTG/ugnet.py:117-127):
 *   tcn1 (causal 3-tap conv + its 1x1 shortcut) + t_conv(time embedding), tcn2, LayerNorm([1,c]) in ONE pass.
 *   x_dev [N, CI, T] -> hn_dev [N, C, T].  w1_dev [C, CI, 3], w2_dev [C, C, 3]: the middle row of the reference's (3,3)
 *   kernels (the image height is 1) with the shortcut (or identity) folded into tap 2; b1_dev [C] = conv bias +
 *   shortcut bias + t_conv(emb(step)); b2_dev [C]; gamma/beta [C].  Exactly one of hn_dev (fp32) and a3_dev (the row as the
 *   fp16 split operand [N, 3*C*T+8] of the down-sampling GEMM, see the f(x) section) is written; the other is NULL.  wsc_dev [C, CI] / sc_dev [N, C, T] (both or neither): the block's 1x1 shortcut W_sc x
 *   (ugnet.py:129) evaluated in the same pass.  Limits: C in {4, 8, 16}, T even (any length: a row is walked in
 *   segments of 512 positions; T % 4 != 0 takes scalar global accesses).  Rows of T <= 512 positions with
 *   C >= 8 and 8 <= CI <= 32 run on the warp-MMA kernel (csrc/stg_tcn_mma.cu: fp16 hi/lo split operands, fp32 accumulate,
 *   ~1e-6 of the output rms from the fp32 FFMA kernel that keeps every other shape). */
int upd_stg_tcn_ln(const float* x_dev, const float* w1_dev, const float* b1_dev, const float* w2_dev, const float* b2_dev,
                   const float* gamma_dev, const float* beta_dev, long long N, int CI, int C, int T, float* hn_dev,
                   void* a3_dev, const float* wsc_dev, float* sc_dev, void* stream);

/* upd_stg_conv1d -- replaces the narrow convolutions of UGnet along the time axis: DownSample's Conv2d(c, c, (1,3), (1,2), (0,1))
 *   (models/Diffusion_model/DiffSTG/ugnet.py:152), UpSample's ConvTranspose2d(c, c, (1,4), (1,2), (0,1)) (:171) and the 1x1
 *   x_proj / out.0 projections (:245-246).  x_dev [N, CI, Tin] -> y_dev [N, CO, Tout] with
 *   Tout = (Tin + 2 pad - K) / stride + 1 (transposed == 0, w_dev [CO, CI, K]) or (Tin - 1) stride - 2 pad + K
 *   (transposed == 1, w_dev [CI, CO, K] as nn.ConvTranspose2d stores it); b_dev [CO] or NULL.  Limit: CI*K*CO <= 12288. */
int upd_stg_conv1d(const float* x_dev, const float* w_dev, const float* b_dev, long long N, int CI, int CO, int Tin, int K,
                   int stride, int pad, int transposed, float* y_dev, void* stream);

/* upd_stg_tcn_ln_cat -- upd_stg_tcn_ln on the channel concatenation cat(x_dev [N, CI1, T], x2_dev [N, CI2, T]) of the
 *   U-Net's up path (`x = torch.cat((x, s), dim=1)`, models/Diffusion_model/DiffSTG/ugnet.py:288-289), read from the two
 *   tensors in place; w1_dev [C, CI1+CI2, 3], wsc_dev [C, CI1+CI2].  Same limits. */
int upd_stg_tcn_ln_cat(const float* x_dev, int CI1, const float* x2_dev, int CI2, const float* w1_dev, const float* b1_dev,
                       const float* w2_dev, const float* b2_dev, const float* gamma_dev, const float* beta_dev, long long N,
                       int C, int T, float* hn_dev, void* a3_dev, const float* wsc_dev, float* sc_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * f(x) condition encoder (ns-Transformer; models/Diffusion_model/NsDiff/mu_backbone.py:53-183,
 * TMDM/tmdm_ns_transformer.py:40-174; blocks from torch-timeseries 0.1.10 -- parity unpinned).
 * Every dense layer runs as one fp16 tensor-core GEMM (fp32 accumulate) on the error-compensated operand
 *   A3(x) = [x_hi | x_lo | x_hi | 1 1 0 0 0 0 0 0]  (fp16, [rows, 3K+8]; hi = fp16(x), lo = fp16(x - hi))
 * against [W_hi | W_hi | W_lo | b_hi b_lo 0..] ; the two calls below produce A3 fused with what precedes the GEMM.
 * ------------------------------------------------------------------------------------------ */

/* upd_gemm3 -- the dense layer itself: out[M, n_out] (fp32) = A3[M, Kp] x W3[Nw, Kp]^T (+ addend[M, n_out] if not NULL),
 *   both operands fp16 row-major with Kp = 3K+8 as above (W3 rows padded to a multiple of 8, n_out <= Nw).  Replaces every
 *   nn.Linear / Conv1d(k=1) of the condition encoders (mu_backbone.py:150-183 via torch-timeseries' AttentionLayer /
 *   EncoderLayer / DecoderLayer, tmdm_ns_transformer.py:150-174), of the DiffusionTS transformer's forward
 *   (DiffusionTS/diffusionts_transformer.py:123-438) and the (1,T+1) / K|Q|V|skip contractions of the graph blocks
 *   (DiffSTG/ugnet.py:117-131, models/layer/gnn_conv.py:18-19).  Warp-specialised persistent tcgen05 kernel: TMA
 *   (SWIZZLE_128B tensor maps, out-of-bounds rows / K columns zero-filled), accumulators in TMEM, fp32 epilogue.
 *   Limits: Kp a multiple of 8, operands / output / addend 16-byte aligned; UPD_ERR_UNSUPPORTED otherwise. */
int upd_gemm3(const void* a3_dev, const void* w3_dev, long long M, int Nw, int n_out, int Kp, float* out_dev,
              const float* addend_dev, void* stream);

/* upd_fx_split -- A3(act(x)).  act: 0 none, 1 ReLU, 2 GELU (erf) = the feed-forward activation between conv1 and conv2
 *   (EncoderLayer/DecoderLayer of torch-timeseries as called at mu_backbone.py:70-104).  H > 1: x_dev is an attention
 *   output [B, H, L, K/H] and the row (b,l) gathers its heads (the `out.transpose(1,2).reshape(B, L, -1)` of
 *   AttentionLayer) -- rows = B*L.  K multiple of 4 (K/H too), K <= 2^20. */
int upd_fx_split(const float* x_dev, long long rows, int K, int H, int L, int act, void* a3_dev, void* stream);

/* upd_fx_add_ln_split -- y = LayerNorm_2(LayerNorm_1(x + res)) (eps 1e-5; res, the second norm, y_dev or a3_dev may be
 *   NULL): the `norm(x + sublayer(x))` of every encoder/decoder layer, with the stack's final norm folded into the
 *   last one; writes y [rows,K] fp32 and A3(y).  K in {32, 64, 96} or a multiple of 128, K <= 1024. */
int upd_fx_add_ln_split(const float* x_dev, const float* res_dev, const float* g1_dev, const float* b1_dev,
                        const float* g2_dev, const float* b2_dev, long long rows, int K, float* y_dev, void* a3_dev,
                        void* stream);

/* upd_fx_attention_hs16 -- the same de-stationary attention for head size 16 (TMDM's condition encoder: d_model 64, 4
 *   heads; tmdm_ns_transformer.py:53-91): fp32 FFMA kernel shared with DiffusionTS (dts_attention.cu), output o_dev
 *   [B*Lq, H*16] fp32 with the heads merged.  Same addressing, tau / delta / causal conventions as upd_fx_attention. */
int upd_fx_attention_hs16(const float* q_dev, long long q_row_stride, const float* k_dev, const float* v_dev,
                          long long kv_row_stride, const float* tau_dev, const float* delta_dev, int delta_pitch, int B, int H,
                          int Lq, int S, int causal, float scale, float* o_dev, void* stream);

/* upd_fx_embed_split -- DataEmbedding (circular token Conv1d(k=3, no bias) + positional table; torch-timeseries block used
 *   at mu_backbone.py:66-69): x_dev [rows/L, L, NF], w_dev [K, NF, 3], pe_dev [>= L, K] -> y_dev [rows, K] fp32 and its
 *   split operand a3_dev [rows, 3K+8].  K multiple of 4, K <= 1024. */
int upd_fx_embed_split(const float* x_dev, const float* w_dev, const float* pe_dev, long long rows, int L, int NF, int K,
                       float* y_dev, void* a3_dev, void* stream);

/* upd_fx_attention -- replaces DSAttention + the head merge of AttentionLayer (torch-timeseries 0.1.10, called from
 *   mu_backbone.py:70-104): out = softmax(scale * (tau_b * Q K^T + delta_b)) V on tcgen05 tensor cores (fp16 hi/lo
 *   operands, fp32 accumulation), written directly as the split operand A3 [B*Lq, 3*H*64+8] of the out-projection GEMM.
 *   q_dev: row (b*Lq + l) at q_dev + row*q_row_stride floats, head h at + h*64; k_dev / v_dev likewise with rows
 *   (b*S + s) and kv_row_stride (so Q|K|V may live in one fused projection buffer).  tau_dev [B] or NULL;
 *   delta_dev [B, delta_pitch] ALREADY multiplied by scale, or NULL; causal != 0 masks keys s > l (self-attention).
 *   Limits: head_dim == 64, S <= 192. */
int upd_fx_attention(const float* q_dev, long long q_row_stride, const float* k_dev, const float* v_dev,
                     long long kv_row_stride, const float* tau_dev, const float* delta_dev, int delta_pitch, int B, int H,
                     int Lq, int S, int head_dim, int causal, float scale, void* a3_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UPD_B200_H */
