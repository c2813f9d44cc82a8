/*
 * b200clip C ABI -- the drop-in boundary of the B200-native CLIP head.
 *
 * The reference (cjycarrie/CLIP-FOR-DL) has no FFI: its boundary is a set of Python symbols in train.py /
 * disease_analysis.py.  The Python package `b200clip` re-exports those symbols
 * (same names, arguments and error behaviour) and lowers each of them onto the entry points below through
 * ctypes; every entry point cites the reference interface it replaces (paths relative to the reference root).
 *
 * Conventions
 *  - plain pointers + sizes; all pointers are DEVICE pointers unless a name ends in `_host`.
 *  - the caller owns every buffer including workspaces (`*_workspace_bytes` tells the size); kernels never allocate.
 *  - stream-ordered and re-entrant: `stream` is a cudaStream_t passed as void*; nothing synchronises the host.
 *  - return 0 on success, negative on failure (-1 invalid argument, -2 CUDA error, -3 workspace too small,
 *    -4 unsupported); b200clip_last_error_string() describes the last failure on this thread.  Nothing throws or
 *    aborts across the boundary and there is no CPU fallback.
 *  - bf16 tensors are row-major `__nv_bfloat16`, 16-byte aligned, leading dimension a multiple of 8 elements.
 */
#ifndef B200CLIP_H
#define B200CLIP_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

int b200clip_version(void);
const char* b200clip_last_error_string(void);
/* number of CUDA kernels this library has launched in this process (bench.py's gpu_launches) */
unsigned long long b200clip_launch_count(void);
/* debug: device buffer (1024 x int64) that receives per-role mbarrier wait-cycle counters of one cluster of the InfoNCE
 * backward kernel on the following launches; NULL switches it off (tools/nce_prof.py) */
void b200clip_debug_set_nce_prof(void* device_buf);

/* ---- generic tensor-core GEMM (tcgen05/TMEM, TMA) ------------------------------------------------------------
 * D[M,N] = A*B, bf16 operands, fp32 accumulate.  a_mn_major=0: a[M][K], =1: a[K][M];  b_mn_major=0: b[N][K]
 * (x W^T, nn.Linear 0426/train.py:78), =1: b[K][N].  epilogue: 0 store f32, 1 store bf16, 2 +bias,GELU -> (out0=p,
 * out1=gelu(p)) bf16, 3 +bias +resid(bf16) -> f32, 4 atomicAdd f32 (split_k > 1), 5 out0 bf16 = acc*gelu'(resid)+aux,
 * 6 relu(acc+bias) bf16. */
int b200clip_gemm_bf16(const void* a, const void* b, int a_mn_major, int b_mn_major, int M, int N, int K,
                       long long lda, long long ldb, int epilogue, float alpha, void* out0, long long ld0,
                       void* out1, long long ld1, const float* bias, const void* resid, long long ld_res,
                       const float* aux, long long ld_aux, int split_k, void* stream);

/* ---- a-L2: F.normalize(x, dim=-1), eps 1e-12 -- 0426/train.py:191-192, :971; 0426/disease_analysis.py:332 ----- */
int b200clip_l2norm_fwd(const void* x, int x_is_bf16, long long ldx, void* y_bf16, float* y_f32, float* inv_norm,
                        long long rows, int D, float eps, void* stream);
/* dx (+)= d/dx normalize(x) . dy  [+ addend * *addend_scale]   (addend [rows,D] f32 and its device scalar are optional);
 * dy may be given as dy_partials >= 1 partial sums, rows*D elements apart (b200clip_infonce_bwd's d_i splits) */
int b200clip_l2norm_bwd(const float* dy, int dy_partials, const void* x, int x_is_bf16, long long ldx, const float* inv_norm,
                        float* dx, int accumulate, long long rows, int D, float eps, const float* addend,
                        const float* addend_scale, void* stream);

/* ---- LayerNorm tail of the projection block -- nn.LayerNorm(512), 0426/train.py:82,95 ------------------------ */
int b200clip_layernorm_fwd(const float* z, const float* gamma, const float* beta, float* y_f32, void* yhat_bf16,
                           float* mean, float* rstd, float* inv_norm, long long rows, int D, float ln_eps, float l2_eps,
                           void* stream);
size_t b200clip_layernorm_bwd_workspace_bytes(long long rows, int D);
int b200clip_layernorm_bwd(const float* dy, const float* z, const float* mean, const float* rstd, const float* gamma,
                           float* dz_f32, void* dz_bf16, float* dgamma, float* dbeta, float* dz_colsum,
                           int accumulate_params, long long rows, int D, float drop_p, unsigned int drop_seed,
                           const unsigned int* drop_seed_dev, void* workspace, size_t workspace_bytes, void* stream);
/* LayerNorm backward with the backward of F.normalize(LayerNorm output) fused in front: dy = inv (g - yhat (yhat . g))
 * [+ addend * *addend_scale], g = sum of dyhat_partials partial sums (rows*D elements apart), yhat bf16 [rows,D]. */
int b200clip_layernorm_l2_bwd(const float* dyhat, int dyhat_partials, const void* yhat_bf16, const float* inv_norm,
                              float l2_eps, const float* addend, const float* addend_scale, const float* z, const float* mean,
                              const float* rstd, const float* gamma, float* dz_f32, void* dz_bf16, float* dgamma, float* dbeta,
                              float* dz_colsum, int accumulate_params, long long rows, int D, float drop_p,
                              unsigned int drop_seed, const unsigned int* drop_seed_dev, void* workspace,
                              size_t workspace_bytes, void* stream);
size_t b200clip_colsum_workspace_bytes(long long rows, int N);
int b200clip_colsum(const void* a, int a_is_bf16, long long lda, long long rows, int N, float* out, int accumulate,
                    void* workspace, size_t workspace_bytes, void* stream);
int b200clip_cast_f32_bf16(const float* in, void* out_bf16, long long n, void* stream);
/* scaled keep-mask of the fused dropout (nn.Dropout, 0426/train.py:81,93): keep ? 1/(1-p) : 0, a pure function of (seed,row,col) */
int b200clip_dropout_mask(float* out, long long rows, int cols, float p, unsigned int seed, void* stream);
/* The effective seed of every fused dropout is drop_seed + *drop_seed_dev (drop_seed_dev: optional DEVICE word).  A captured
 * CUDA graph freezes launch arguments; this one-thread kernel, captured at the head of the step graph, advances the device
 * word (seed <- seed * 1664525 + 1013904223) so that every replay trains with a fresh nn.Dropout mask (0426/train.py:81,93). */
int b200clip_dropout_seed_advance(unsigned int* seed_dev, void* stream);
int b200clip_sum_f32(const float* a, long long n, float* out, void* stream);

/* ---- a-P1 / a-P2: ImageProjection.forward 0426/train.py:84-96, TextProjection.forward :109-116 -----------------
 * Linear(E,D) -> GELU(erf) -> Linear(D,D) -> Dropout(drop_p; counter-based mask from drop_seed, 0 = off) -> +residual ->
 * LayerNorm [-> L2-normalised bf16].  Saved for backward (caller tensors): p, h (bf16), z (f32), mean, rstd; the backward
 * pass regenerates the dropout mask from (drop_p, drop_seed). */
int b200clip_proj_fwd(const void* x_bf16, long long B, int E, int D, const void* w1_bf16, const float* b1,
                      const void* w2_bf16, const float* b2, const float* gamma, const float* beta, float ln_eps,
                      float drop_p, unsigned int drop_seed, const unsigned int* drop_seed_dev, void* p_bf16, void* h_bf16,
                      float* z_f32, float* y_f32, void* yhat_bf16, float* mean, float* rstd, float* inv_norm, void* stream);
size_t b200clip_proj_bwd_workspace_bytes(long long B, int E, int D);
/* dy = gradient w.r.t. the LayerNorm output y.  Fused entry (dy == NULL): dyhat = gradient w.r.t. yhat = y/||y|| as
 * dyhat_partials partial sums, plus yhat, 1/||y|| and an optional addend (gradient that reaches y directly, times a device
 * scalar): the L2-norm backward then runs inside the LayerNorm-backward kernel. */
int b200clip_proj_bwd(const float* dy, const float* dyhat, int dyhat_partials, const void* yhat_bf16, const float* inv_norm,
                      const float* addend, const float* addend_scale, const void* x_bf16, long long B, int E, int D, const void* w1_bf16,
                      const void* w2_bf16, const float* gamma, const void* p_bf16, const void* h_bf16, const float* z_f32,
                      const float* mean, const float* rstd, float drop_p, unsigned int drop_seed,
                      const unsigned int* drop_seed_dev, float* dx_f32, void* dx_bf16,
                      float* dw1, float* db1, float* dw2,
                      float* db2, float* dgamma, float* dbeta, void* workspace, size_t workspace_bytes, void* stream);

/* ---- MultiViewFusion.forward(frontal_view, lateral_view) -- 0426/train.py:988-1000 --------------
 * x_bf16 = cat[frontal, lateral] as one [B, 2D] bf16 matrix (b200clip_cast_f32_bf16_2d writes the two halves);
 * h = dropout(relu(x W0^T + b0)) (bf16, saved for backward), y = h W3^T + b3.  Backward regenerates nothing: the saved h
 * carries the mask (h > 0 <=> positive and kept). */
int b200clip_cast_f32_bf16_2d(const float* in, long long ld_in, void* out_bf16, long long ld_out, long long rows, int cols,
                              void* stream);
int b200clip_fusion_fwd(const void* x_bf16, long long B, int D, const void* w0_bf16, const float* b0, const void* w3_bf16,
                        const float* b3, float drop_p, unsigned int drop_seed, void* h_bf16, float* y_f32, void* stream);
/* dx_f32 (or NULL) is [2][B][D]: d frontal, then d lateral, each contiguous (what autograd wants: no split copies). */
size_t b200clip_fusion_bwd_workspace_bytes(long long B, int D);
int b200clip_fusion_bwd(const float* dy, const void* x_bf16, long long B, int D, const void* w0_bf16, const void* w3_bf16,
                        const void* h_bf16, float drop_p, float* dx_f32, float* dw0, float* db0, float* dw3, float* db3,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ---- MultiModalAttention.forward(image_features, text_features) -- multimodal_attention/train.py:1069-1110
 * x_bf16 [B,D] image features, t [C,D] class text features (fp32, C <= 16, D <= 512); returns out [B,D] (enhanced features)
 * and w [B,C] (attention weights).  ip [B,D], tp [C,D], w, e_bf16 [B,D] are caller tensors saved for the backward pass;
 * the [B,C,D] tensor the reference expands is never formed. */
int b200clip_attention_fwd(const void* x_bf16, const float* t, long long B, int C, int D, const void* wi_bf16, const float* bi,
                           const float* wt, const float* bt, const float* wa, const float* ba, const void* wo_bf16,
                           const float* bo, float* ip, float* tp, float* w, void* e_bf16, float* out, void* stream);
size_t b200clip_attention_bwd_workspace_bytes(long long B, int C, int D);
int b200clip_attention_bwd(const float* d_out, const float* d_w, const void* x_bf16, const float* t, long long B, int C, int D,
                           const void* wi_bf16, const float* wt, const float* wa, const float* ba, const void* wo_bf16,
                           const float* ip, const float* tp, const float* w, const void* e_bf16, float* dx, float* dt,
                           float* dwi, float* dbi, float* dwt, float* dbt, float* dwa, float* dba, float* dwo, float* dbo,
                           void* workspace, size_t workspace_bytes, void* stream);

/* ---- a-N: contrastive_loss(image_features, text_features, temperature) -- 0426/train.py:154-176 ----------------
 * Inputs are L2-normalised bf16 rows.  Data-parallel form: i_hat = this rank's rows [b_loc, D] (global rows
 * row0 .. row0+b_loc), t_hat = all b_glob rows.  fwd_stats -> r (row sums, complete) and c_partial (this rank's
 * column sums; SUM-combine across ranks).  loss -> sums[3] = {sum log r, sum_{c_lo<=j<c_hi} log c_j, sum S_ii},
 * the half-inverse statistics the backward pass consumes and (if `loss` != NULL, single rank) the scalar loss.
 * bwd -> d_i [b_loc, D] and d_t_partial [b_glob, D] (reduce-scatter across ranks), both f32, scaled by the optional
 * device scalar *grad_scale. */
size_t b200clip_infonce_workspace_bytes(long long b_loc, long long b_glob);
int b200clip_infonce_fwd_stats(const void* i_hat, const void* t_hat, int D, long long b_loc, long long b_glob,
                               float temperature, float* r, float* c_partial, void* workspace, size_t workspace_bytes,
                               void* stream);
int b200clip_infonce_loss(const void* i_hat, const void* t_hat, int D, long long b_loc, long long b_glob, long long row0,
                          float temperature, const float* r, const float* c, long long c_lo, long long c_hi, float* rinvh,
                          float* cinvh, double* sums, float* loss, void* workspace, size_t workspace_bytes, void* stream);
/* rinvh = 0.5 / r, cinvh = 0.5 / c alone (one tiny kernel): everything the backward pass needs from the statistics, so that
 * b200clip_infonce_loss (then called with rinvh = cinvh = NULL) can run on another stream beside the backward kernel. */
int b200clip_infonce_inv_stats(const float* r, long long b_loc, const float* c, long long b_glob, float* rinvh, float* cinvh,
                               void* stream);
/* d_i is [d_i_splits][b_loc][D]: partial sums over column ranges (1 <= splits <= 8; b200clip_infonce_bwd_splits suggests
 * a count that balances the grid when b_loc << b_glob); their sum is the gradient.  D = 512 or 768.  directions: 1 = d_i only,
 * 2 = d_t_partial only, 3 = both in one launch (the data-parallel step launches 2 first so that the reduce-scatter of
 * d_t_partial overlaps 1). */
int b200clip_infonce_bwd_splits(long long b_loc, long long b_glob);
int b200clip_infonce_bwd(const void* i_hat, const void* t_hat, int D, long long b_loc, long long b_glob, long long row0,
                         float temperature, const float* rinvh, const float* cinvh, const float* grad_scale, float* d_i,
                         int d_i_splits, float* d_t_partial, int directions, void* stream);

/* ---- a-S: contrastive_clip_loss_function(text_projection, image_projection, temperature, mode) -- 0426/train.py:127-152
 * (soft targets softmax((I I^T + T T^T)/2 * tau), NOT detached; cross_entropy :118-125).  fp32 throughout (un-normalised inputs:
 * logits of +-10^3..10^4); the n x n matrices live in the caller's workspace, n <= 8192.  text, image: [n, D] f32.
 * _logits = mode "eval"; _fwd_bwd = mode "train": loss, and with d_text/d_image the gradients for upstream *grad_scale. */
size_t b200clip_softclip_workspace_bytes(long long n);
int b200clip_softclip_logits(const float* text, const float* image, long long n, int D, float temperature, float* logits,
                             void* stream);
int b200clip_softclip_fwd_bwd(const float* text, const float* image, long long n, int D, float temperature,
                              const float* grad_scale, float* loss, float* d_text, float* d_image, void* workspace,
                              size_t workspace_bytes, void* stream);

/* ---- a-B: multilabel_contrastive_loss(image_features, text_features, labels, temperature) -- 0426/train.py:178-230
 * label_sum: device scalar = sum(labels) over the GLOBAL batch; total_elems = B_glob * C.  status gets 1 when the
 * loss is NaN/Inf/>1000 (the reference's fallback condition, :224).  coef [B,C] = d loss / d (cos/tau) feeds
 * b200clip_skinny_outer for the text-side gradient. */
size_t b200clip_smallc_workspace_bytes(long long rows, int C, int D);
int b200clip_mlbce_fwd_bwd(const float* image_features, long long ldx, const float* text_features, const float* labels,
                           int label_cols, long long ld_labels, long long B, int C, int D, float temperature,
                           const float* label_sum, double total_elems, const float* grad_scale, float* d_image,
                           int d_image_accumulate, float* coef, float* x_inv_norm, double* sums, float* loss, int* status,
                           void* workspace, size_t workspace_bytes, void* stream);

/* ---- a-A: FC classification adapter nn.Linear(512,16) + nn.BCEWithLogitsLoss() -- NB02 c28:50-52, c29:23-25;
 * prediction sigmoid(z) > 0.5 -- NB02 c30:42-43 */
int b200clip_fc_bce_fwd_bwd(const float* x, long long ldx, const float* weight, const float* bias, const float* labels,
                            long long ld_labels, long long B, int C, int D, double total_elems, float threshold,
                            const float* grad_scale, float* d_x, int d_x_accumulate, float* coef, float* pred,
                            float* logits, double* sums, float* loss, void* workspace, size_t workspace_bytes,
                            void* stream);
/* out_w[C,D] (+)= *out_scale * coef^T (x * row_scale) ; out_b[C] (+)= *out_scale * colsum(coef)   (out_scale optional) */
int b200clip_skinny_outer(const float* coef, int C, const float* x, long long ldx, const float* row_scale, long long rows,
                          int D, float* out_w, float* out_b, int accumulate, const float* out_scale, void* workspace,
                          size_t workspace_bytes, void* stream);

/* a-B + a-A on warp-level tensor cores (head step, D = 512 or 768, 16 class texts + 16 FC rows): reads the L2-normalised bf16
 * features and 1/||y||; writes sums[3] as above, d_y [B,D] = d(text BCE + FC BCE)/dy for upstream gradient 1, the FC
 * coefficients times ||y|| in bf16 (coefn [B,16], feeds b200clip_skinny_outer_mma) and db_fc[16] = sum_rows dL/dz. */
size_t b200clip_bce_heads_mma_workspace_bytes(long long rows);
int b200clip_bce_heads_mma_fwd(const void* yhat_bf16, const float* inv_norm, long long B, int D, const float* class_text,
                               int c1, const float* fc_weight, const float* fc_bias, int c2, const float* labels,
                               int label_cols, long long ld_labels, float temperature, const float* label_sum,
                               double total_elems_text, double total_elems_fc, float* d_y, void* coefn_bf16, float* db_fc,
                               double* sums, void* workspace, size_t workspace_bytes, void* stream);
/* out_w[16,D] = *out_scale * coefn^T yhat ; db_out[16] = *out_scale * db_raw   (nn.Linear(512,16) weight/bias gradient) */
int b200clip_skinny_outer_mma(const void* coefn_bf16, const void* yhat_bf16, long long rows, int D, int C,
                              const float* out_scale, float* out_w, const float* db_raw, float* db_out, void* workspace,
                              size_t workspace_bytes, void* stream);

/* ---- multilabel_asymmetric_loss(logits, targets, gamma_pos=0, gamma_neg=4, clip=0.05, eps=1e-8, reduction='mean')
 * -- multimodal_attention/train.py:233-268.  reduction: 0 none, 1 mean, 2 sum. */
size_t b200clip_asl_workspace_bytes(void);
int b200clip_asl_fwd_bwd(const float* logits, const float* targets, long long n, float gamma_pos, float gamma_neg, float clip,
                         float eps, int reduction, const float* grad_scale, const float* grad_elem, float* loss_elem,
                         float* d_logits, double* sum, float* loss, void* workspace, size_t workspace_bytes, void* stream);

/* ---- torch.optim.AdamW for the head parameters: decoupled weight decay, bias-corrected moments.  `step`
 * is a device float counter: call b200clip_adamw_tick once per optimizer step, then b200clip_adamw_step per parameter tensor. */
int b200clip_adamw_tick(float* step, void* stream);
int b200clip_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                        float beta2, float eps, float weight_decay, const float* step, void* stream);

/* ---- zero-shot post-processing -- multimodal_attention/zero_shot_predict.py:66-213, L <= 32 labels.
 * dynamic_thresholds (:112-159): scores [N,L] = per-sample max over the two views of the sigmoid scores, labels [N,L] in {0,1}
 * (both f32) -> thresholds [L] (f64, device), optional best F1 per label.  merge_views (:183-213 on the per-view lists of
 * disease_analysis.py:361-413, top_k = None): prob_views [N,2,L] f32, view weights w0 = 1.0, w1 = 0.8 -> pred [N,L] u8. */
size_t b200clip_zs_thresholds_workspace_bytes(long long N, int L);
int b200clip_zs_dynamic_thresholds(const float* scores, const float* labels, long long N, int L, double* thresholds,
                                   double* best_f1, void* workspace, size_t workspace_bytes, void* stream);
int b200clip_zs_merge_views(const float* prob_views, const double* thresholds, long long N, int L, double w0, double w1,
                            uint8_t* pred, float* merged, void* stream);

/* loss of the fused head step from its six numerators (summed over ranks): sums6 = {sum_i log r_i, sum_j log c_j,
 * sum_i S_ii, text-BCE pos numerator, text-BCE neg numerator, FC-BCE sum}; parts3 = {InfoNCE, text BCE, FC BCE}. */
int b200clip_head_loss_finalize(const double* sums6, const float* label_sum, float temperature_nce, double b_glob,
                                double total_elems_text, double total_elems_fc, float* loss, float* parts3, int* status,
                                void* stream);

/* a-B + a-A fused for the head step: one pass over the image features serves both BCE heads (classes [0,c1) = class
 * texts, [c1,c1+c2) = FC adapter rows).  sums[3] = {text pos numerator, text neg numerator, FC BCE sum}. */
int b200clip_bce_heads_fwd_bwd(const float* image_features, long long ldx, const float* text_features, int c1,
                               const float* fc_weight, const float* fc_bias, int c2, const float* labels, int label_cols,
                               long long ld_labels, long long B, int D, float temperature, const float* label_sum,
                               double total_elems_text, double total_elems_fc, const float* grad_scale, float* d_image,
                               int d_image_accumulate, float* fc_coef, double* sums, float* loss_text, float* loss_fc,
                               int* status, void* workspace, size_t workspace_bytes, void* stream);

/* ---- a-M: predict_multilabel(image_features, text_features, threshold) -- 0426/train.py:869-886 ---------------- */
int b200clip_predict_multilabel(const float* image_features, long long ldx, const float* text_features, long long B,
                                int C, int D, float temperature, float threshold, float* pred, void* stream);

/* ---- a-Z: zero-shot scoring core of predict_zero_shot -- 0426/disease_analysis.py:329-356 (softmax top-k),
 * multimodal_attention/disease_analysis.py:345-413 (sigmoid(cos/0.5) >= thr), NB02 c41:27-32, c44:24-36, and the
 * 14 x (pos,neg) prompt shape.  thr_logit_host: HOST array of nlabels floats, a label passes when its score
 * (cos/tau, or l+ - l- in pair mode) is > (or >= if thr_inclusive) the value, i.e. logit(threshold). */
size_t b200clip_zeroshot_workspace_bytes(long long n);
/* D = 512 or 768 (the two widths BASELINE.json names; np <= 32 prompts).
 * workspace (optional, b200clip_zeroshot_workspace_bytes): rows whose decision margin is inside the guard band are listed
 * there and re-evaluated exactly by a second kernel; without it they are re-evaluated in place (same results, slower). */
int b200clip_zeroshot_score(const void* x_bf16, long long ldx, long long n, const void* prompts_bf16, int np, int D,
                            int pair_mode, int normalize_x, float temperature, const float* thr_logit_host,
                            int thr_inclusive, float guard, int topk, int value_mode, uint8_t* argmax, void* mask,
                            int mask_is_u32, uint8_t* topk_idx, float* topk_val, float* scores,
                            unsigned long long* guard_count, void* workspace, size_t workspace_bytes, void* stream);

/* contrastive_loss(image_features, text_features, temperature) -- 0426/train.py:154-176 -- for inputs that are NOT unit
 * vectors: fp32 logits, true row/column maxima, n <= 8192 (workspace: b200clip_softclip_workspace_bytes(n)).  The flash path
 * (b200clip_infonce_*) requires L2-normalised rows; b200clip_rows_unit_check sets *flag |= 1 when a row violates that. */
int b200clip_infonce_general_fwd_bwd(const float* image, const float* text, long long n, int D, float temperature,
                                     const float* grad_scale, float* loss, float* d_image, float* d_text, void* workspace,
                                     size_t workspace_bytes, void* stream);
/* the fixed shift m of the flash path: loss = m + (sums[0] + sums[1]) / (2 B) - sums[2] / B  (1/tau for tau >= 0.036) */
double b200clip_infonce_shift(float temperature);
int b200clip_rows_unit_check(const float* x, long long rows, int D, float tol, int* flag, void* stream);

/* ---- step edges --------------------------------------------------------------------------
 * calculate_multilabel_metrics(predictions, labels) -- 0426/train.py:251-302 -- and the in-loop accuracy counters of
 * train_epoch / validate (0426/train.py:437-447).  predictions [B,C] are probabilities (or any score compared with
 * `threshold`), labels [B,C] in {0,1}, C <= 32.  out (DEVICE, 7 + C doubles): sample_acc, label_acc, hamming_score,
 * exact_match, top1_acc (the reference's any-over-batch quirk, :277), top3_acc, f1_score, then per-class accuracy in %. */
size_t b200clip_multilabel_metrics_workspace_bytes(long long B);
int b200clip_multilabel_metrics(const float* predictions, long long ld_pred, const float* labels, long long ld_labels,
                                long long B, int C, float threshold, double* out, void* workspace,
                                size_t workspace_bytes, void* stream);
/* per-disease mean of L2-normalised prompt features (get_text_features_with_findings, 0426/disease_analysis.py:486-497);
 * prompts of disease d are rows offsets[d] .. offsets[d+1]-1 of prompt_features [*, D] (offsets: DEVICE int[n+1]). */
int b200clip_prompt_mean_pool(const float* prompt_features, const int* offsets, int num_diseases, int D, float eps,
                              int renormalize, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200CLIP_H */
