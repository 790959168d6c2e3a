/* vlsa_b200 — C ABI of the B200 (sm_100a) language-guided patch-aggregation path of VLSA.
 *
 * The reference (liupei101/VLSA @ 915f37a) is pure Python/PyTorch and has no FFI; the boundary this
 * library replaces is the arithmetic inside
 *     model/vlsa.py:181-198       VLSA.forward
 *     model/deepmil.py:170-215    VLFAN.forward          (+ :16-67 logit_pooling / FeatMIL)
 *     utils/func.py:40-48         softmax output converter
 *     loss/loss_surv.py:144-169   SurvIFMLE.forward
 *     loss/loss_surv_ext.py:42-109 SurvEMD.forward
 * Each entry point names the reference lines it stands in for.  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - every call is asynchronous and ordered on `stream` (a cudaStream_t passed as void*);
 *   - the library never allocates device memory: scratch comes in through `workspace`
 *     (size from the matching *_workspace_bytes query) and must stay alive until the stream reaches
 *     the end of the call;
 *   - no global mutable state: calls on different streams are independent (the only process-wide data are
 *     once-initialised caches of per-device launch attributes);
 *   - return value: 0 on success, a negative VLSA_E* code for argument errors, a positive
 *     cudaError_t for CUDA failures (vlsa_error_string decodes both).
 *   - feature dim D is fixed at 512 (vlsa_img_encoder_dim_in, config/IFMLE/tcga_blca/cfg_vlsa_conch.yaml:47);
 *     1 <= P <= 16 text prototypes (num_query), 1 <= R <= 32 ordinal ranks (time_bins).
 */
#ifndef VLSA_B200_H_
#define VLSA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VLSA_D 512
#define VLSA_MAX_P 16
#define VLSA_MAX_R 32

#define VLSA_DTYPE_F32 0
#define VLSA_DTYPE_BF16 1
#define VLSA_DTYPE_SPLIT16 2 /* pre-split tile images of a device cohort, see vlsa_split16_pack (vlsa_agg_fwd / _partial_fwd / _bwd /
                              * _pooled_fwd only, together with VLSA_ROWS_RANGES) */
#define VLSA_DTYPE_MASK 0xff
/* Optional bits OR-ed into x_dtype of the vlsa_agg_* calls: force the streaming kernel of an fp32 pass for THIS call
 * (cross-checks in the parity tests; default = automatic: fp32 rows — register-staged tcgen05 kernel for P > 5, CUDA-core
 * kernel otherwise; bf16 rows — TMA-fed tcgen05 kernel agg_bf16_kernel, VLSA_KERNEL_SIMT selects the CUDA-core kernel).
 * All kernels compute the same function (model/deepmil.py:187-203). */
#define VLSA_KERNEL_SIMT 0x100   /* CUDA-core kernel (agg_simt_kernel) */
#define VLSA_KERNEL_TC 0x200     /* tcgen05 kernel (fp32 rows: agg_tc_kernel, bf16 rows: agg_bf16_kernel) */
/* Optional bit of x_dtype of the vlsa_agg_* calls: `cu_rows` holds 2 B entries (first row, one past the last row) per bag
 * instead of B + 1 offsets — the bags of the call lie anywhere inside X [total_rows, 512], in any order.  This is how a
 * step is drawn from a device-resident cohort (every patient of a split uploaded once; the reference re-uploads every
 * bag every epoch, runner/vlsa_handler.py:205) without a gather copy.  The plan (vlsa_agg_plan) only needs the sizes. */
#define VLSA_ROWS_RANGES 0x1000

#define VLSA_EINVAL (-1)      /* bad argument (null pointer, P/R/D out of range, ...) */
#define VLSA_EWORKSPACE (-2)  /* workspace too small */
#define VLSA_EUNSUPPORTED (-3)

#define VLSA_POOL_MEAN 0 /* 'logit_mean'  (deepmil.py:29-30) */
#define VLSA_POOL_TOPK 1 /* 'logit_topK' / 'logit_max' = top-1 (deepmil.py:24-28) */
#define VLSA_POOL_MAX 2  /* FeatMIL pooling 'max' over the patch FEATURES (deepmil.py:59-60); vlsa_feat_pool_fwd only */

int vlsa_version(void);
const char* vlsa_error_string(int code);
/* Device cohort stored as pre-split tile images (VLSA_DTYPE_SPLIT16).  Replaces, for bags that stay resident in HBM across
 * epochs, the per-epoch `.cuda()` of fp32 rows (runner/vlsa_handler.py:205, dataset/PatchWSI.py:197-215) AND the fp32 -> fp16
 * (hi, lo) split the tensor-core kernel would redo on every pass: rows are packed ONCE, at upload time, into records of 16 rows
 * — the shared-memory image of the tensor-core kernel's tile (8 slots x 2 row groups x (hi | lo) 128-byte-swizzled atoms) followed
 * by 16 x (1 / |x~|, 2^e) — `vlsa_split16_row_bytes()` = 2 056 bytes per row, the same 4 bytes per element as fp32.  A pass then
 * lands a tile with one bulk copy and converts nothing; results are bit-identical to the fp32-row tensor-core kernel's.
 * `image` holds records back to back; a bag starts at a record boundary: `first_row` (padded row space) % 16 == 0, and the
 * ranges of a plan (VLSA_ROWS_RANGES) are given in that padded row space.  X [n_rows, 512] fp32 on the device. */
size_t vlsa_split16_row_bytes(void);
int vlsa_split16_pack(const float* X, int64_t n_rows, void* image, int64_t first_row, void* stream);

/* Host-only helper: split the bags of one call into row chunks for the streaming kernels.
 * cu_rows_host[B+1] are the row offsets of the bags inside the packed X.  Writes chunk_start_host[B+1]
 * (first chunk id of every bag; chunk_start_host[B] = total chunks) and *chunk_rows_out (rows per
 * chunk, a multiple of the kernel's row tile).  sm_count <= 0 means "query the current device". */
int vlsa_agg_plan(const int64_t* cu_rows_host, int B, int sm_count, int* chunk_rows_out,
                  int32_t* chunk_start_host);

/* Scratch bytes needed by vlsa_agg_fwd / vlsa_agg_bwd for a plan with total_chunks chunks. */
size_t vlsa_agg_workspace_bytes(int total_chunks, int B, int P);

/* Fused forward of VLSA.forward (model/vlsa.py:181-198) on B packed bags:
 *   Qn = Q/|Q|; s = scale * Qn.(x/|x|); A = softmax_N(s); O = A@X      (deepmil.py:187-200)
 *   v = mean_P(O); f = W v + b                                         (deepmil.py:136,204)
 *   g = f/|f|; Tn = T/|T|; logits = (exp(logit_scale) g) @ Tn^T        (vlsa.py:185-192)
 *   IF = softmax(logits)                                               (utils/func.py:44)
 * X [total_rows, 512] holds the bags back to back (total_rows = cu_rows[B], passed by value because the TMA tensor map
 * of the tcgen05 kernel is built on the host) and is read exactly once.  out_ml / out_O / out_v / out_f are what vlsa_agg_bwd needs later
 * (out_O may be NULL when no backward follows).  out_if and out_Tn may be NULL.
 * Encoder-only use (VLFAN.forward alone, deepmil.py:170-215): T = NULL stops after f; out_g, out_logits,
 * out_if, out_Tn are then ignored.
 * q_prenorm = 0 everywhere except the gated query (deepmil.py:192-195), which passes the P difference rows
 * Qn_p - Qn_gate with q_prenorm = 1: they enter the scores as they are (see vlsa_agg_pooled_fwd). */
int vlsa_agg_fwd(const void* X, int x_dtype, int64_t total_rows, const int64_t* cu_rows, const int32_t* chunk_start, int B,
                 int chunk_rows, int total_chunks, const float* Q, int P, int q_prenorm, float coattn_scale, const float* W,
                 const float* bias, const float* T, int R, const float* logit_scale, void* workspace,
                 size_t workspace_bytes, float* out_v, float* out_f, float* out_g, float* out_logits,
                 float* out_if, float* out_ml, float* out_O, float* out_Tn, void* stream);

/* Only the streaming kernel of vlsa_agg_fwd (the one launch that reads X): writes the per-chunk
 * online-softmax partials into `workspace` and nothing else.  Exists so that the dominant kernel can be
 * timed in isolation (bench.py roofline) and so that callers can overlap the epilogue themselves. */
int vlsa_agg_partial_fwd(const void* X, int x_dtype, int64_t total_rows, const int64_t* cu_rows, const int32_t* chunk_start, int B,
                         int chunk_rows, int total_chunks, const float* Q, int P, float coattn_scale,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Backward of vlsa_agg_fwd w.r.t. (Q, W, bias, T, logit_scale) — what torch autograd computes for
 * model/vlsa.py:181-198 when `pred_loss.backward()` runs (runner/vlsa_handler.py:281).  X is data: no dX.
 * Second (and last) read of X: dQn_p = sum_n scale * A_pn (u_n - delta_p) x_n / |x_n| with
 * u_n = dv.x_n / P, delta_p = dv.O_p / P, A recomputed from the saved (max, sum) in `ml`.
 * v, f, g, logits, ml, O are the tensors vlsa_agg_fwd wrote; d_logits [B,R] is the incoming
 * gradient; d_g [B,D] (gradient w.r.t. the returned image features) and d_f [B,D] (gradient w.r.t. the
 * adapter output) may be NULL.  Encoder-only use (VLFAN.forward alone): T = NULL, then d_f is required and
 * f, g, logits, d_logits, dT, dlogit_scale are ignored.
 * Outputs are overwritten (not accumulated): dQ [P,D], dW [D,D], db [D], dT [R,D], dlogit_scale [1].
 * With q_prenorm = 1, dQ is the gradient w.r.t. the rows as passed (no normalisation Jacobian). */
int vlsa_agg_bwd(const void* X, int x_dtype, int64_t total_rows, const int64_t* cu_rows, const int32_t* chunk_start, int B,
                 int chunk_rows, int total_chunks, const float* Q, int P, int q_prenorm, float coattn_scale, const float* W,
                 const float* T, int R, const float* logit_scale, const float* v, const float* f, const float* g,
                 const float* logits, const float* ml, const float* O, const float* d_logits, const float* d_g,
                 const float* d_f, void* workspace, size_t workspace_bytes, float* dQ, float* dW, float* db,
                 float* dT, float* dlogit_scale, void* stream);

/* Pooled per-prototype features only: O[b][p] = softmax_n(scale * qdir_p . x_n / |x_n|) @ X_b  (model/deepmil.py:187-200),
 * for the VLFAN configurations whose tail is not "mean over P -> Linear": gated_query (deepmil.py:192-195),
 * query_pooling 'max' | 'weight' | 'attention' | 'gated_attention' (deepmil.py:133-150), pred_head 'Identity'
 * (deepmil.py:111-114).  The tail of those variants is P x 512 work and stays with the caller.
 * q_prenorm = 0: qdir_p = Q_p / |Q_p| as in vlsa_agg_fwd.  q_prenorm = 1: qdir_p = Q_p as given — the gated query
 * passes Qn_p - Qn_gate, the difference of two unit rows (A_[:, :-1] - A_[:, -1:] is linear in the query).
 * Same plan and workspace as vlsa_agg_fwd.  out_ml [B,P,2], out_O [B,P,D]. */
int vlsa_agg_pooled_fwd(const void* X, int x_dtype, int64_t total_rows, const int64_t* cu_rows, const int32_t* chunk_start, int B,
                        int chunk_rows, int total_chunks, const float* Q, int P, int q_prenorm, float coattn_scale,
                        void* workspace, size_t workspace_bytes, float* out_ml, float* out_O, void* stream);

/* Backward of vlsa_agg_pooled_fwd for an arbitrary gradient d_O [B,P,D] (one row per prototype):
 *   dS_pn = A_pn (d_O_p . x_n - d_O_p . O_p),  dqdir_p = scale * sum_n dS_pn x_n / |x_n|,
 * dQ [P,D] = dqdir pushed through the row normalisation (q_prenorm = 0) or dqdir itself (q_prenorm = 1), summed
 * over the B bags.  One pass over X with 2 P + 1 dot products per row (CUDA-core kernel for every P and dtype). */
int vlsa_agg_pooled_bwd(const void* X, int x_dtype, int64_t total_rows, const int64_t* cu_rows, const int32_t* chunk_start, int B,
                        int chunk_rows, int total_chunks, const float* Q, int P, int q_prenorm, float coattn_scale,
                        const float* ml, const float* O, const float* d_O, void* workspace, size_t workspace_bytes,
                        float* dQ, void* stream);

/* Gradient of vlsa_agg_pooled_fwd w.r.t. the patch rows themselves, for callers that train a module in front of the
 * aggregation — VLFAN's `feat_proj` (Linear + LayerNorm over every row, model/layers.py:65-82, deepmil.py:176-179):
 *   dx_n = sum_p ( A_pn d_O_p + (scale dS_pn / |x_n|) qdir_p ) - (sum_p dS_pn s_pn) x_n / |x_n|^2.
 * X [total_rows, D] fp32 (the projected rows), cu_rows [B+1] device, max_rows = the longest bag.  out_dX [total_rows, D].
 * One read of X, one write of dX. */
int vlsa_agg_pooled_bwd_dx(const float* X, const int64_t* cu_rows, int B, int64_t max_rows, const float* Q, int P,
                           int q_prenorm, float coattn_scale, const float* ml, const float* O, const float* d_O,
                           float* out_dX, void* stream);

/* Attention read-out of ONE bag (`ret_with_attn=True`, model/deepmil.py:206-213; utils/model_inference.py:104-113).
 * ml != NULL: A[p][n] = exp(scale * cos(Q_p, x_n) - ml[p][0]) / ml[p][1], the softmax over the N patches with the
 *   normalisers ml [P,2] from vlsa_agg_fwd (axis_softmax = 'V', the `cottn_score` of the forward);
 * ml == NULL: A[p][n] = softmax over the P prototypes of scale * cos(Q_p, x_n) (axis_softmax = 'L').
 * q_prenorm as in vlsa_agg_pooled_fwd (0 everywhere except the gated query).  out_A is [P,N]. */
int vlsa_attn_fwd(const void* X, int x_dtype, int64_t N, const float* Q, int P, int q_prenorm, float coattn_scale,
                  const float* ml, float* out_A, void* stream);

/* Interpretation path, "decoupled" text-image similarities per prototype (utils/model_inference.py:115-131):
 *   sim[b][p][r]  = cottn_score_p @ ((visual_adapter(X) / |f_b|) @ Tn_r)  =  (W O_bp + bias) . Tn_r / |f_b|
 * (attention rows sum to one, so the N-long contraction collapses onto the pooled O [B,P,D] that vlsa_agg_fwd
 * returns: no second pass over X), decoupled_imp = softmax over P of e^logit_scale sim, probs_2 = softmax over R of
 * e^logit_scale mean_P sim.  f [B,D] is the un-normalised adapter output of vlsa_agg_fwd.
 * out_sim, out_imp are [B,P,R]; out_probs is [B,R]. */
int vlsa_interp_fwd(const float* O, const float* W, const float* bias, const float* T, int R, const float* f,
                    const float* logit_scale, int B, int P, float* out_sim, float* out_imp, float* out_probs,
                    void* stream);

/* softmax -> w_ifmle * SurvIFMLE + w_emd * SurvEMD, value and gradient in one pass
 * (runner/vlsa_handler.py:241-258; loss/loss_surv.py:144-169 with alpha/eps; loss/loss_surv_ext.py:70-109
 * with p=2, raw distance, exp(logit_scale) detached).  t, e: int64 [B] (time bin, event indicator).
 * inv_norm = 1 / (number of samples the mean runs over; the GLOBAL batch when bags are sharded).
 * input_is_prob = 1: `logits` already holds the incidence (the reference loss modules' own signature,
 * loss_surv.py:144); the softmax is skipped and out_dlogits is the gradient w.r.t. that incidence.
 * out_loss [3] = (total, ifmle, emd); out_if [B,R] and out_dlogits [B,R] may be NULL;
 * out_per_sample [B,2] (ifmle_i, emd_i) is required (it is also the reduction scratch). */
int vlsa_surv_loss_fwd_bwd(const float* logits, const int64_t* t, const int64_t* e, int B, int R,
                           const float* logit_scale, float w_ifmle, float w_emd, float alpha, float eps,
                           float inv_norm, int input_is_prob, float* out_loss, float* out_if, float* out_dlogits,
                           float* out_per_sample, void* stream);

/* Zero-shot arm (model/vlsa.py:189-196 with FeatMIL identity, model/deepmil.py:16-37) for ONE bag:
 * per-patch logits exp(logit_scale) * cos(x_n, T_r), pooled over N per class (mean, or mean of the
 * top-min(k,N) values, k <= 64), preds = argmax_r (first maximal index).  out_logits [R], out_pred [1]. */
size_t vlsa_logit_pool_workspace_bytes(int64_t N, int R, int k);
int vlsa_logit_pool_fwd(const void* X, int x_dtype, int64_t N, const float* T, int R, const float* logit_scale,
                        int mode, int k, void* workspace, size_t workspace_bytes, float* out_logits,
                        int64_t* out_pred, void* stream);

/* Zero-shot arm, FeatMIL pooling 'mean' | 'max' (model/deepmil.py:57-60 + model/vlsa.py:188-192), ONE bag:
 *   f = mean | max over the N rows of X (column-wise); g = f/|f|; logits = (exp(logit_scale) g) @ (T/|T|)^T.
 * A one-row bag under any pooling is the same computation (mean of one row).  mode = VLSA_POOL_MEAN | VLSA_POOL_MAX.
 * out_f [512], out_g [512] (the returned image_features), out_logits [R], out_Tn [R,512] (may be NULL). */
size_t vlsa_feat_pool_workspace_bytes(void);
int vlsa_feat_pool_fwd(const void* X, int x_dtype, int64_t N, int mode, const float* T, int R, const float* logit_scale,
                       void* workspace, size_t workspace_bytes, float* out_f, float* out_g, float* out_logits,
                       float* out_Tn, void* stream);

/* image_features of the logit-pooling zero-shot modes: the N patches L2-normalised (F.normalize, model/vlsa.py:188-189).
 * out [N,512] fp32. */
int vlsa_row_normalize(const void* X, int x_dtype, int64_t N, float* out, void* stream);

/* One Adam step over the S trainable tensors of the path in one launch — `self.optimizer.step()` of the reference's
 * `_update_network` (runner/vlsa_handler.py:283) with the optimizer optim/optim_factory.py:25-37 builds for
 * cfg_vlsa_conch.yaml:111-118: torch.optim.Adam semantics (L2 weight decay added to the gradient, bias-corrected moments,
 * eps added to sqrt(v_hat)), amsgrad off.  Gradients and both moments are flat fp32 buffers of one layout (the all-reduce
 * bucket); `segments` is a DEVICE array of S records of vlsa_adam_segment_bytes() bytes:
 *     { float* param; int64 offset (floats, into the flat buffers); int64 n; float weight_decay; float lr; }
 * `step_count_in` [S] (device, float) holds the steps each tensor has taken; the counts after this step are written to
 * `step_count_out` [S], a DIFFERENT array (callers ping-pong two: no block can read a count another block already bumped, and
 * the step stays one launch); `flags` [S] (device, may be NULL = every tensor has a gradient): a tensor whose flag is 0
 * received no gradient this step and is skipped — moments, count and decay untouched, as torch.optim.Adam skips
 * `grad is None` — decided on the device, so an optimizer step needs no device -> host read.
 * max_n = the largest n of the segments. */
size_t vlsa_adam_segment_bytes(void);
int vlsa_adam_step(const void* segments, int S, int64_t max_n, const float* grads_flat, float* exp_avg, float* exp_avg_sq,
                   const float* step_count_in, float* step_count_out, const float* flags, float beta1, float beta2, float eps,
                   void* stream);

/* Whole path with HOST buffers (what a caller holding CPU tensors — the reference's DataLoader output,
 * dataset/PatchWSI.py:197-215 + runner/vlsa_handler.py:205,322-330 — would call): stage the packed bags
 * X_host [total_rows, D] (pinned memory for a truly asynchronous copy) to the device on `stream_copy`,
 * run vlsa_agg_fwd on `stream_compute` once the copy has landed, and copy the incidence [B,R] (and raw
 * logits if out_logits_host != NULL) back to the host on `stream_compute`.  Parameters (Q, W, bias, T,
 * logit_scale) are DEVICE pointers (they live on the GPU between calls).  Returns after enqueueing;
 * synchronise `stream_compute` before reading the outputs.  With two streams the call orders them both ways (the
 * copies wait for whatever `stream_compute` still does with `workspace`, the kernels wait for the copies), so a
 * workspace may be reused by the next call without a host synchronisation.  `workspace` is device memory of at least
 * vlsa_forward_host_workspace_bytes(total_rows, B, P, x_dtype) bytes. */
size_t vlsa_forward_host_workspace_bytes(int64_t total_rows, int B, int P, int x_dtype);
int vlsa_forward_host(const void* X_host, int x_dtype, const int64_t* cu_rows_host, int B, const float* Q, int P,
                      int q_prenorm, float coattn_scale, const float* W, const float* bias, const float* T, int R,
                      const float* logit_scale, void* workspace, size_t workspace_bytes, float* out_if_host,
                      float* out_logits_host, void* stream_compute, void* stream_copy);

#ifdef __cplusplus
}
#endif
#endif /* VLSA_B200_H_ */
