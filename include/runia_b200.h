/*
 * runia_b200.h -- C ABI of libruniab200.so: the B200 (sm_100a) implementation of RunIA-core's
 * post-hoc OoD scoring hot path.
 *
 * The reference (CEA-LIST/runia_core) is 100 % Python and has no FFI: every entry point below
 * replaces the NumPy / SciPy / scikit-learn / faiss / torch call sequence at the cited lines
 * (paths relative to the reference tree).  The Python classes in `runia_core_b200/` bind these
 * with ctypes (see INTEGRATION.md for the stub a maintainer of the reference would add).
 *
 * Conventions
 *  - All pointers are DEVICE pointers unless the name ends in `_host`.  All matrices are dense,
 *    row-major, contiguous.  The caller allocates inputs and outputs; nothing is retained.
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Calls are
 *    asynchronous with respect to the host and re-entrant per stream.
 *  - Return value: 0 on success; a negative RUNIA_E_* code for argument errors; a positive
 *    cudaError_t if a CUDA runtime call failed.  `runia_b200_last_error()` gives a message for
 *    the calling thread.  No C++ exception crosses the boundary.
 *  - "f32"/"f64" in a name is the element type of the streamed input; outputs are documented
 *    per call and match the dtype the reference returns.
 */
#ifndef RUNIA_B200_H
#define RUNIA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RUNIA_B200_ABI_VERSION 1

#define RUNIA_OK 0
#define RUNIA_E_BADARG (-1)      /* null pointer, non-positive size, k out of range ...        */
#define RUNIA_E_UNSUPPORTED (-2) /* shape outside what the kernels are built for               */
#define RUNIA_E_WORKSPACE (-3)   /* workspace too small (query the size with the *_workspace)  */

int runia_b200_abi_version(void);
const char *runia_b200_last_error(void);
/* Number of kernel launches issued through this library by the calling process (bench.py's
 * `gpu_launches`). */
int64_t runia_b200_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * (a1) MC-dropout latent-sample entropy -- evaluation/entropy.py:41-93 (`get_dl_h_z`) and :20-38
 * (`single_image_entropy_calculation`), i.e. entropy_estimators.continuous.get_h(x, k,
 * norm="max", min_dist=1e-5) per item (joint) and per (item, dimension).
 *   z      [n_items * n_mc, D] float32, item-major (rows i*n_mc .. i*n_mc+n_mc-1 = item i)
 *   h_z    [n_items, D] float64   per-dimension entropies            (entropy.py:73-92)
 *   h_mvn  [n_items]    float64   joint (Chebyshev) entropy          (entropy.py:67-71); may be NULL
 *   k      neighbours (entropy.py:66: 5 if n_mc > 5 else n_mc-1); 1 <= k < n_mc <= 128 (tuned kernels: n_mc <= 32)
 *   digamma_term = -psi(k) + psi(n_mc), computed by the host in float64.
 */
int runia_mcd_entropy_f32(const float *z, int64_t n_items, int n_mc, int D, int k, double min_dist,
                          double digamma_term, double *h_z, double *h_mvn, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Input staging: out[n, j] = (float)(in[n, j] - center[j])   (center may be NULL).
 * Used for float64 inputs (the reference keeps float64 end to end, e.g. postprocessors.py:241)
 * so that the subtraction of the fitted mean happens at input precision before the float32
 * contraction.  in_is_f64: 0 = float32 input, 1 = float64 input.
 */
int runia_center_cast(const void *in, int in_is_f64, int64_t N, int d, const double *center,
                      float *out, void *stream);

/* ---------------------------------------------------------------------------------------------
 * (a2) PCA projection -- dimensionality_reduction.py:75-87 (`apply_pca_transform` ->
 * sklearn PCA.transform):  Z = (X - mean) @ components^T, then / sqrt(explained_variance).
 *   X          [N, D0] float32
 *   mean       [D0]   float32  (subtracted in the prologue; may be NULL)
 *   components [d, D0] float32 (row j = component j, K-contiguous)
 *   inv_scale  [d]    float32  1/sqrt(explained_variance) (NULL when whiten=False)
 *   Z          [N, d] float32
 */
int runia_pca_transform_f32(const float *X, int64_t N, int D0, const float *mean,
                            const float *components, int d, const float *inv_scale, float *Z,
                            void *stream);

/* ---------------------------------------------------------------------------------------------
 * (a3) LaREM Mahalanobis -- inference/postprocessors.py:228-244 (`MDLatentSpace.postprocess`):
 *   out[n] = -(x_n - mu)^T P (x_n - mu).
 * The symmetric precision is passed factored, P = sum_j sign_j w_j w_j^T (host: float64 eigen-
 * decomposition, w_j = sqrt(|lambda_j|) v_j), so the score is a signed sum of squares of one
 * contraction:  out[n] = -sum_j sign_j ((x_n - mu) . w_j)^2  -- no cancellation in the sum.
 * (a7) ViM residual -- postprocessors.py:1082-1112: with Wt = NS^T, mu = u, sign = NULL and
 * mode = RUNIA_ROWNORM_VIM:  out[n] = -alpha * sqrt(sum_j ((x_n-u).NS_j)^2) + logsumexp(logits[n,:]).
 *   X     [N, d] float32;  mu [d] float32 (NULL = 0)
 *   Wt    [r, d] float32  (row j = w_j, K-contiguous);  sign [r] float32 (+1/-1/0; NULL = all +1)
 *   logits [N, C] float32 (ViM only), alpha (ViM only)
 *   out_f64 / out_f32: exactly one non-NULL; the reference returns float64 for MD, float32 for ViM.
 */
#define RUNIA_ROWNORM_MD 0
#define RUNIA_ROWNORM_VIM 1
int runia_rownorm_score_f32(const float *X, int64_t N, int d, const float *mu, const float *Wt, int r,
                            const float *sign, int mode, const float *logits, int C, float alpha,
                            double *out_f64, float *out_f32, void *stream);

/* Tensor-core (tcgen05, 3xTF32, FP32 accumulation in TMEM) versions of (a2) and (a3)/(a7).
 * runia_split_tf32: hi = tf32(x), lo = tf32(x - hi) -- the two operand planes of a fitted matrix
 *   (factored precision, PCA components, kNN / KDE bank), computed once at setup.
 * The streamed operand X is read as raw fp32 and split inside the kernel; products issued are
 *   X_lo*W_hi + X_hi*W_lo + X_hi*W_hi, so the result is FP32-faithful (~2^-21 relative per product).
 * Require K % 4 == 0 and 16-byte aligned pointers (RUNIA_E_UNSUPPORTED otherwise: use the FP32
 * SIMT entry points above).  Arguments as for the SIMT versions, with W / components replaced by
 * their two planes. */
int runia_split_tf32(const float *x, int64_t total, float *hi, float *lo, void *stream);
int runia_rownorm_score_tc(const float *X, int64_t N, int d, const float *mu, const float *Wt_hi,
                           const float *Wt_lo, int r, const float *sign, int mode, const float *logits, int C,
                           float alpha, double *out_f64, float *out_f32, void *stream);
int runia_pca_transform_tc(const float *X, int64_t N, int D0, const float *mean, const float *components_hi,
                           const float *components_lo, int d, const float *inv_scale, float *Z, void *stream);

/* ---------------------------------------------------------------------------------------------
 * (a6) Class-conditional Mahalanobis -- inference/funcs.py:69-102 (`mahalanobis_postprocess`) and
 * inference/postprocessors.py:320-357 (`cMDLatentSpace.postprocess`):
 *   out[n] = max_c -(x_n - mu_c)^T P (x_n - mu_c), classes without samples skipped.
 * With the same factorisation of P: y_n = (x_n - g) Wt^T, m_c = (mu_c - g) Wt^T (host, float64),
 *   out[n] = max_c -sum_j sign_j (y_nj - m_cj)^2.
 *   g [d] float32: any centre (the mean of the class means) -- keeps |y| small.
 *   Mc [C, r] float32; class_valid [C] int32 (0 = class had no training samples -> skipped)
 */
int runia_classcond_mahalanobis_f32(const float *X, int64_t N, int d, const float *g, const float *Wt,
                                    int r, const float *sign, const float *Mc, const int32_t *class_valid,
                                    int C, double *out_f64, float *out_f32, void *stream);

/* Tensor-core (3xTF32) version of (a6): Wt replaced by its two planes; C <= 16, d % 4 == 0, r % 4 == 0. */
int runia_classcond_mahalanobis_tc(const float *X, int64_t N, int d, const float *g, const float *Wt_hi,
                                   const float *Wt_lo, int r, const float *sign, const float *Mc,
                                   const int32_t *class_valid, int C, double *out_f64, float *out_f32, void *stream);

/* ---------------------------------------------------------------------------------------------
 * (a9) DDU / GMM log-density -- inference/postprocessors.py:490-491, 783-784 with the mixture of
 * inference/funcs.py:265-344:  out[n] = logsumexp_c log N(x_n; mu_c, Sigma_c).
 *   At [C * dpad, d] float32: block c holds rows of A_c = L_c^{-1} (Sigma_c = L_c L_c^T), zero-
 *        padded to dpad rows (dpad multiple of 128)
 *   off [C * dpad] float32: A_c mu_c (zero in the padding);  logconst [C] float32:
 *        -sum(log diag L_c) - d/2 log(2 pi)
 */
int runia_gmm_lse_f32(const float *X, int64_t N, int d, const float *At, const float *off, int dpad,
                      const float *logconst, int C, float *out, void *stream);
/* Tensor-core (3xTF32) version: At replaced by its two planes; d % 4 == 0. */
int runia_gmm_lse_tc(const float *X, int64_t N, int d, const float *At_hi, const float *At_lo, const float *off,
                     int dpad, const float *logconst, int C, float *out, void *stream);

/* ---------------------------------------------------------------------------------------------
 * (a5) kNN -- inference/postprocessors.py:385-423 (`KNNLatentSpace`), :825-883 (`KNN`),
 * inference/funcs.py:105-115 (`normalizer`), faiss.IndexFlatL2.search.
 *
 * runia_normalize_rows: out = float32(x / (||x||_2 + 1e-10)), norm accumulated in float64 in a
 *   fixed order (lane-interleaved partial sums + xor tree) so that the oracle reproduces it
 *   bit for bit.
 * runia_row_sqnorm_f32: out[n] = float32(sum_j x_nj^2) (float64 accumulation) -- the |b|^2 term
 *   of the distance expansion; computed once per bank at setup.
 * runia_knn_search_f32: for each query the k nearest bank rows under exact squared L2, total
 *   order (distance, index).  Candidates come from a fused FP32 distance-GEMM + per-row
 *   streaming top-KCAP filter (the Nq x Nb matrix never reaches HBM); candidates are re-ranked
 *   with exact float64 distances (fixed summation order); a row whose top-k cannot be PROVEN
 *   exact from the candidate bound (approximate distance of every non-candidate minus the
 *   rounding bound exceeds the exact k-th distance) is recomputed by an exhaustive exact pass.
 *   Outputs (any may be NULL):
 *     out_dist     [Nq, k] float32 ascending (FLT_MAX padding when k > Nb, like faiss)
 *     out_dist_f64 [Nq, k] float64 (the exact values before rounding; +inf padding)
 *     out_idx      [Nq, k] int64 (-1 padding); idx_offset is added to every index (bank shards)
 *     out_kth      [Nq]    float32 = out_dist[:, k-1]
 *   status [4] int32 (device): [0] rows that took the exhaustive pass; [1..3] reserved (always 0: the exhaustive
 *     pass handles any number of bank rows tying with the k-th neighbour by cutting its hit buffer back to the
 *     k smallest pairs whenever the next chunk of the bank could overflow it).
 *   The rounding bound of the certification scales with (|q|^2 + max_b |b|^2) / 2, so un-normalised rows
 *   (FlatL2Index used directly) are certified as strictly as the unit-norm rows of the postprocessors.
 *   Bn_tf32_hi / Bn_tf32_lo: optional pre-split planes of the bank (runia_split_tf32).  When both
 *     are given and d % 4 == 0 the candidate pass runs on the tcgen05 tensor cores (TMA-fed);
 *     NULL selects the FP32 SIMT pass.  The result is identical either way (exact re-rank).
 *   workspace: runia_knn_workspace_bytes(Nq, Nb, d, k) bytes of device memory.  1 <= k <= 1016.
 * runia_knn_search_ex_f32: the same with the tensor-core candidate filter's arithmetic chosen by the caller:
 *     filter_products = 1  one TF32 product (A_hi x B_hi; the raw query tile is the operand): a third of the tensor work;
 *                          the re-rank evaluates exactly every candidate within twice the product's rounding bound
 *                          (~3e-3 per unit of squared norm) of the k-th approximate distance and certifies against it;
 *                          rows with more than 4 x (k + 8 rounded up to a power of two) bank rows inside that band take
 *                          the exhaustive pass.  This is what runia_knn_search_f32 uses.
 *     filter_products = 3  the FP32-faithful 3xTF32 contraction (bound ~1e-4): for banks whose neighbourhoods are
 *                          denser than the single-product bound separates (the Python mirror measures that once per
 *                          bank: _ops.knn_filter_products).
 *   Same outputs, same exactness, same workspace size.
 * runia_topk_merge: merges R partial results ([R, Nq, k] float64 dist / int64 idx, each
 *   ascending) into the global top-k under the same total order -- the step after the NCCL
 *   all-gather when the bank is sharded across GPUs.  R <= 64.
 */
int runia_normalize_rows(const void *in, int in_is_f64, int64_t N, int d, float *out, void *stream);
int runia_row_sqnorm_f32(const float *X, int64_t N, int d, float *out, void *stream);
int64_t runia_knn_workspace_bytes(int64_t Nq, int64_t Nb, int d, int k);
int runia_knn_search_f32(const float *Qn, int64_t Nq, const float *Bn, const float *Bn_sqnorm,
                         const float *Bn_tf32_hi, const float *Bn_tf32_lo, int64_t Nb, int d, int k,
                         int64_t idx_offset, float *out_dist, double *out_dist_f64,
                         int64_t *out_idx, float *out_kth, int32_t *status, void *workspace,
                         int64_t workspace_bytes, void *stream);
int runia_knn_search_ex_f32(const float *Qn, int64_t Nq, const float *Bn, const float *Bn_sqnorm,
                            const float *Bn_tf32_hi, const float *Bn_tf32_lo, int64_t Nb, int d, int k,
                            int64_t idx_offset, float *out_dist, double *out_dist_f64,
                            int64_t *out_idx, float *out_kth, int32_t *status, void *workspace,
                            int64_t workspace_bytes, int filter_products, void *stream);
int runia_topk_merge(const double *part_dist, const int64_t *part_idx, int R, int64_t Nq, int k,
                     float *out_dist, int64_t *out_idx, float *out_kth, void *stream);

/* ---------------------------------------------------------------------------------------------
 * (a4) LaRED Gaussian KDE -- inference/postprocessors.py:109-128, 165-178
 * (sklearn KernelDensity(kernel="gaussian", bandwidth=h).score_samples):
 *   out[n] = logsumexp_i(-|q_n - b_i|^2 / (2 h^2)) - log(Nb_total) - d/2 log(2 pi h^2)
 * Fused distance-GEMM + online log-sum-exp.  For a bank shard pass partial outputs
 * (out_max, out_sum) [Nq] float32 (running max m and sum of exp(t - m)) and combine over ranks;
 * for a whole bank pass out_f64 and the total Nb.
 *   B_tf32_hi / B_tf32_lo: optional pre-split planes of the bank -> tcgen05 3xTF32 pass (as for kNN).
 *   workspace: runia_kde_workspace_bytes(Nq, Nb) bytes.
 */
int64_t runia_kde_workspace_bytes(int64_t Nq, int64_t Nb);
int runia_kde_lse_f32(const float *Q, int64_t Nq, const float *B, const float *B_tf32_hi, const float *B_tf32_lo,
                      int64_t Nb, int d, double bandwidth, int64_t Nb_total, double *out_f64, float *out_max, float *out_sum,
                      void *workspace, int64_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * (a8) Logit-space scores in one pass -- inference/postprocessors.py:519-551 (Energy),
 * :580-608 (MSP), :650-691 (GEN) + inference/funcs.py:347-375:
 *   energy[n] = logsumexp(l_n); msp[n] = max softmax(l_n);
 *   gen[n] = -sum_{top-M p} p^gamma (1-p)^gamma.     Any output may be NULL.  Any C, any M (M <= 0 or M >= C:
 *   all classes, like NumPy's [:, -M:]).
 * runia_gen_entropy_f32: `generalized_entropy(probs, gamma, M)` (funcs.py:347-375) on rows that already are
 *   probabilities -- no softmax, rows need not sum to one.
 */
int runia_logit_scores_f32(const float *logits, int64_t N, int C, float gamma, int M, float *energy,
                           float *msp, float *gen, void *stream);
int runia_gen_entropy_f32(const float *probs, int64_t N, int C, float gamma, int M, float *out, void *stream);

/* ---------------------------------------------------------------------------------------------
 * (a10) ReAct / DICE / DICE+ReAct -- inference/postprocessors.py:1444-1474, 1325-1354, 1591-1621,
 * inference/funcs.py:171-190:   out[n] = logsumexp_c( min(x_n, clip) . W_c + b_c )
 * W is the (masked, for DICE) final linear layer [C, d]; clip = +inf disables ReAct.  Any C and d (heads
 * that fit shared memory -- C <= 64, C*d*4 <= 200 KiB -- use the resident-weight kernels, others stream W).
 * ASH-S -- postprocessors.py:1192-1222 + funcs.py:230-261: keep the k_keep largest activations
 * of each row (ties with the k-th value: lowest indices), scale by exp(sum_all / sum_kept), then the same linear
 * layer + logsumexp.  runia_ash_linear_lse_f32 is the fused form for heads that fit shared memory;
 * runia_ash_prune_f32 writes the pruned and scaled rows [N, d] for any width, to be followed by
 * runia_clip_linear_lse_f32 / _tc with clip = +inf (any C).
 */
int runia_clip_linear_lse_f32(const float *X, int64_t N, int d, const float *W, const float *b, int C,
                              float clip, float *out, void *stream);
/* Tensor-core version for d % 4 == 0, d <= 4096, 16-byte aligned X: the same tcgen05 pipeline as the row
 * scorers (TMA-streamed rows, clip + 3xTF32 split by the converters, log-sum-exp straight from TMEM).
 * C <= 32: a 32-column panel, W_hi / W_lo = TF32 planes of W zero-padded to [32, d] (runia_split_tf32).
 * C > 32 (any): 256-column panels with an online log-sum-exp across them, W_hi / W_lo = planes of W [C, d],
 * b 16-byte aligned. */
int runia_clip_linear_lse_tc(const float *X, int64_t N, int d, const float *W_hi, const float *W_lo, const float *b,
                             int C, float clip, float *out, void *stream);
int runia_ash_linear_lse_f32(const float *X, int64_t N, int d, const float *W, const float *b, int C,
                             int k_keep, float *out, void *stream);
int runia_ash_prune_f32(const float *X, int64_t N, int d, int k_keep, float *out, void *stream);
/* Plain linear layer out[N, C] = X W^T + b (b may be NULL), FP32 SIMT contraction: RouteDICE.forward
 * (inference/funcs.py:171-190), which the reference materialises as an [N, C, d] product. */
int runia_linear_f32(const float *X, int64_t N, int d, const float *W, const float *b, int C, float *out, void *stream);

/* ---------------------------------------------------------------------------------------------
 * (f1) OoD detection metrics -- evaluation/metrics.py:37-100 (`get_auroc_results`: torchmetrics 1.8.2
 * binary auroc / roc / precision_recall_curve + sklearn.metrics.auc), InD = positive class.
 *   ind [n_ind], ood [n_ood]  scores, float32 (is_f64 = 0) or float64 (is_f64 = 1), device
 *   out4 [4] float64: auroc, fpr@95, aupr, number of ROC points (including the prepended origin)
 *   fpr_out / tpr_out [n_ind + n_ood + 1] float32 (optional): the ROC curve, one point per distinct
 *     score in descending order after (0, 0); only the first out4[3] entries are written
 *   workspace: runia_ood_metrics_workspace_bytes(n_ind, n_ood) bytes (keys, radix-sort double buffer,
 *     digit histograms, scan state).  n_ind + n_ood < 2^31.
 * Device-side LSD radix sort (8 bits per pass, stable) + one fused scan; the AUROC numerator is an exact
 * integer sum.
 */
int64_t runia_ood_metrics_workspace_bytes(int64_t n_ind, int64_t n_ood);
int runia_ood_metrics(const void *ind, int64_t n_ind, const void *ood, int64_t n_ood, int is_f64, double *out4,
                      float *fpr_out, float *tpr_out, void *workspace, int64_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * (f4) EigenScore -- llm_uncertainty/scores.py:49-66 (`eigen_score`): mean log singular value of
 * cov(E^T) + alpha I, E [n, d] float32 (n sampled generations x d hidden units), computed from the
 * n x n Gram matrix of the centred samples in float64 (Jacobi) instead of a d x d SVD.
 *   out [1] float64.  2 <= n <= 32, n <= d <= 25600.
 */
int runia_eigen_score_f32(const float *E, int n, int d, double alpha, double *out, void *stream);

/* ---------------------------------------------------------------------------------------------
 * (f4) Predictive entropy / mutual information of MC-dropout logits -- inference/funcs.py:430-465
 * (`get_predictive_uncertainty_score`).
 *   logits [n_items * n_mc, C] float32, item-major;  pred_h [n_items], mi [n_items] float32 (either may be NULL)
 *   pred_h = -sum_c mean_s(p) log mean_s(p), mi = pred_h - mean_s(-sum_c p log p), p = softmax per row.
 */
int runia_pred_uncertainty_f32(const float *logits, int64_t n_items, int n_mc, int C, float *pred_h, float *mi,
                               void *stream);

/* ---------------------------------------------------------------------------------------------
 * (f3) Spatial reduction of convolutional activation maps -- feature_extraction/utils.py:70-92
 * (`get_mean_or_fullmean_ls_sample`), the reducer that produces the rows `get_dl_h_z` consumes.
 *   x [P, H, W] float32 (P = batch * channels);  fullmean != 0: out [P];  fullmean == 0 ("mean"): out [P, H].
 */
int runia_spatial_mean_f32(const float *x, int64_t P, int H, int W, int fullmean, float *out, void *stream);

/* ---------------------------------------------------------------------------------------------
 * (f2) Ascending sort of float32 values (same LSD radix sort as the metrics): the order statistics
 * behind np.percentile(ind_train_data.flatten(), p) in ReAct / DICE+ReAct setup
 * (inference/postprocessors.py:1433, 1576).  NaNs sort last, -0.0 before +0.0.
 *   workspace: runia_sort_f32_workspace_bytes(n) bytes.  n < 2^31.
 */
int64_t runia_sort_f32_workspace_bytes(int64_t n);
int runia_sort_f32(const float *x, int64_t n, float *out_sorted, void *workspace, int64_t workspace_bytes,
                   void *stream);

/* ---------------------------------------------------------------------------------------------
 * (f2) setup() statistics: the mean / covariance halves of MDLatentSpace.setup
 * (inference/postprocessors.py:202-226), cMDLatentSpace.setup (:283-318) and mahalanobis_preprocess
 * (inference/funcs.py:33-66).
 *   runia_class_mean_f32: means[c] = X[labels == c].mean(0) for c < C, float32, accumulated row by row in
 *     ascending order like NumPy's axis-0 reduction (bit-identical); labels NULL (C must be 1): the column
 *     mean of all rows.  counts[c] (nullable) = rows of class c.  An empty class gives NaN like NumPy.
 *   runia_centered_gram_f64: G [d, d] = sum_i r_i r_i^T in float64 with r_i = f32(x_i - centers[labels_i])
 *     (centers NULL: r_i = x_i; labels NULL: every row uses centers[0]; rows whose label is outside [0, C)
 *     are skipped), colsum [d] (nullable) = sum_i r_i.  np.cov(R.T, bias=1) = (G - n a a^T) / n, a = colsum / n.
 *     Deterministic (fixed-order split reduction).  workspace: runia_centered_gram_workspace_bytes(N, d).
 *   runia_shifted_gram_f64: the same with ONE float64 centre u [d] (device): r_i = f64(x_i) - u, exact in float64 -- the
 *     rows ViM.setup hands EmpiricalCovariance(assume_centered=True) (inference/postprocessors.py:1060-1064:
 *     `ec.fit(train - self.u)`, u float64); covariance = G / N.  Same workspace.
 */
int runia_class_mean_f32(const float *X, const int32_t *labels, int64_t N, int d, int C, float *means,
                         int64_t *counts, void *stream);
size_t runia_centered_gram_workspace_bytes(int64_t N, int d);
int runia_centered_gram_f64(const float *X, const int32_t *labels, const float *centers, int64_t N, int d, int C,
                            double *G, double *colsum, void *workspace, size_t workspace_bytes, void *stream);
int runia_shifted_gram_f64(const float *X, const double *center, int64_t N, int d, double *G, double *colsum,
                           void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * (f2) float64 symmetric eigendecomposition and Cholesky factorisation for the setup() fits.
 *   runia_eigh_f64: A [n, n] symmetric (symmetrised as (A + A^T) / 2) -> evals [n] (unordered) and evecs [n, n] with
 *     ROW j = the unit eigenvector of evals[j]; one-sided Jacobi, deterministic, |A - V diag V^T| ~ 1e-14 |A|.  This is
 *     the decomposition behind scipy.linalg.pinvh (sklearn EmpiricalCovariance.precision_: inference/postprocessors.py:
 *     212-220, 296-314; inference/funcs.py:62-66) and behind PCA(svd_solver="covariance_eigh")
 *     (dimensionality_reduction.py:52-72).  The call synchronises `stream` once per sweep (the sweep count depends on
 *     the matrix); *sweeps_out (host, nullable) receives it.  workspace: runia_eigh_workspace_bytes(n).  n <= 8192.
 *   runia_cholesky_f64: L [batch, n, n] lower with A_b + jitter I = L_b L_b^T; fail[b] = 0, or 1 + the first column whose
 *     pivot is not positive (or, with rel_pivot > 0, below rel_pivot times its original diagonal entry: numerically
 *     singular at that resolution) -- the signal gmm_fit's jitter ladder reacts to (inference/funcs.py:296-342).
 *   runia_tril_inverse_f64: X_b = L_b^{-1} for a batch of lower-triangular factors (X != L): the whitening blocks the
 *     GMM scorer contracts against (DDU.setup, inference/postprocessors.py:751-760 -> MultivariateNormal.log_prob).
 */
size_t runia_eigh_workspace_bytes(int n);
int runia_eigh_f64(const double *A, int n, double *evals, double *evecs, void *workspace, size_t workspace_bytes,
                   int max_sweeps, int *sweeps_out, void *stream);
int runia_cholesky_f64(const double *A, int batch, int n, double jitter, double rel_pivot, double *L, int32_t *fail,
                       void *stream);
int runia_tril_inverse_f64(const double *L, int batch, int n, double *X, void *stream);

/* ---------------------------------------------------------------------------------------------
 * (f3) MC-DropBlock sampler fused with the spatial mean: MCSamplerModule.forward for layer_type "Conv"
 * (feature_extraction/abstract_classes.py:81-101 = n_mc x [DropBlock2D -> get_mean_or_fullmean_ls_sample
 * "fullmean", feature_extraction/utils.py:70-92]); DropBlock2D is dropblock==0.3.0 (published forward:
 * seed = rand(B,H,W) < drop_prob / bs^2; block_mask = 1 - max_pool2d(seed, bs, 1, bs/2) [even bs: cropped];
 * out = x * block_mask * numel / sum).
 *   x [B, C, H, W] float32; seed [n_mc, B, H, W] uint8 (non-zero = Bernoulli hit, drawn by the caller so that
 *   the RNG stream is torch's); out [B * n_mc, C] float32, row b * n_mc + m =
 *   sum over the cells kept by mask (m, b) of x[b, c] / number of kept cells   (normalised per image).
 *   n_mc <= 32 per call (more samples: one call per 32 seeds, as the Python mirror does).  workspace:
 *   runia_mc_dropblock_workspace_bytes(B, H, W, n_mc).
 */
size_t runia_mc_dropblock_workspace_bytes(int B, int H, int W, int n_mc);
int runia_mc_dropblock_mean_f32(const float *x, const uint8_t *seed, int B, int C, int H, int W, int n_mc,
                                int block_size, float *out, void *workspace, size_t workspace_bytes, void *stream);
/* The same sampler for layer_type "FC" / "RPN" (no spatial reduction): out [n_mc, B * C * H * W] =
 * x * block_mask(m) * (B * H * W) / sum(block_mask(m)), the mask normalised over the whole batch like DropBlock2D. */
int runia_mc_dropblock_apply_f32(const float *x, const uint8_t *seed, int B, int C, int H, int W, int n_mc,
                                 int block_size, float *out, void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * (f3) Object-level reducers -- feature_extraction/object_level.py:254-309 (`_reduce_features_to_rois`) and :312-366
 * (`_dropblock_rois_get_entropy`): torchvision.ops.roi_align(feat, [boxes], output_size, spatial_scale,
 * sampling_ratio, aligned) and the per-(box, channel) mean / unbiased std over the pooled bins.
 *   feat [B, C, H, W] float32; boxes [K, 4] float32 xyxy in image coordinates; batch_idx [K] int32 (NULL: image 0)
 *   sampling_ratio <= 0: adaptive ceil(roi_size / pooled_size) like torchvision; aligned: half-pixel shift
 *   runia_roi_align_f32: out [K, C, pooled_h, pooled_w]
 *   runia_roi_align_mean_f32: out_mean [K, C], out_std [K, C] (nullable), the maps are never materialised
 */
int runia_roi_align_f32(const float *feat, int B, int C, int H, int W, const float *boxes, const int32_t *batch_idx,
                        int64_t K, int pooled_h, int pooled_w, float spatial_scale, int sampling_ratio, int aligned,
                        float *out, void *stream);
int runia_roi_align_mean_f32(const float *feat, int B, int C, int H, int W, const float *boxes, const int32_t *batch_idx,
                             int64_t K, int pooled_h, int pooled_w, float spatial_scale, int sampling_ratio, int aligned,
                             float *out_mean, float *out_std, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Host -> device staging of PAGEABLE host memory (the NumPy arrays the reference's callers hand to postprocess(),
 * evaluation/metrics.py:322-340): pinned ring + parallel memcpy by a persistent worker pool + one cudaMemcpyAsync
 * per 4 MiB chunk on `stream`.  Returns once every chunk is enqueued; `src_host` may be reused immediately, the
 * device data is ready in stream order.  dst_dev is a DEVICE pointer, src_host a HOST pointer.
 */
int runia_stage_h2d(void *dst_dev, const void *src_host, int64_t bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Measurement aid (bench.py): TF32 tensor peak of this GPU.  Launches one CTA pair per TPC, each issuing
 * iters x 4 back-to-back tcgen05.mma.cta_group::2.kind::tf32 (256 x 256 x 8) on resident shared-memory tiles;
 * *flop_out (host pointer, nullable) receives the TF32 FLOP the launch issues.  Time it with CUDA events.
 */
int runia_tf32_peak_probe(int iters, double *flop_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RUNIA_B200_H */
