/* mvb.h -- C-ABI of libmvb.so: B200 (sm_100a) kernels for MVPy's cross-validated ridge hot path.
 *
 * The reference (FabulousFabs/MVPy) is pure Python and has no FFI; the boundary it offers is its
 * estimator API.  This header is the boundary the reference-side classes bind instead of calling
 * torch.linalg / ATen (see INTEGRATION.md for the ctypes stubs).  Each entry point names the
 * reference code it replaces (file:line under mvpy/).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (torch tensors), row-major, contiguous
 *     unless strides are given; nothing is allocated or freed inside the library;
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work, they never synchronise;
 *   - dtype: MVB_F32 or MVB_F64 (fp32 uses the tcgen05 split-TF32 contraction core, fp64 the
 *     CUDA-core path that the reference's fp64 test-suite needs);
 *   - return value: 0 = ok, <0 = error (MVB_E_*), message via mvb_last_error() (thread-local);
 *   - scratch memory comes from the caller: ask the matching *_workspace_bytes() first.
 */
#ifndef MVB_H_
#define MVB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVB_F32 0
#define MVB_F64 1

#define MVB_OK 0
#define MVB_E_ARG (-1)       /* bad argument                          */
#define MVB_E_WORKSPACE (-2) /* workspace too small                   */
#define MVB_E_CUDA (-3)      /* CUDA runtime error                    */
#define MVB_E_UNSUPPORTED (-4)

int mvb_version(void);
const char* mvb_last_error(void);
/* number of kernels launched by this library in the calling process (all threads) */
int64_t mvb_launch_count(void);

/* ---- measurement hook ----------------------------------------------------------------------------
 * Off by default.  After mvb_profile_reset() every tensor-core contraction launch is bracketed by CUDA
 * events on its own stream (at most 32768 timed launches per kernel class; the rest are only counted).
 * mvb_profile_read synchronises those events and reports, per class c < n_classes: timed milliseconds,
 * algorithmic flops (2*M*N*K) of the timed launches, flops of all launches, #timed and #all launches.
 * Used by bench.py for the roofline line; mvb_profile_disable() turns the hook off again.            */
void mvb_profile_reset(void);
void mvb_profile_disable(void);
/* CUDA-graph support (mvpy_b200/_graphs.py): the hook must be off while a call sequence is being stream-captured
 * (returns the previous state), and replays of a captured sequence report their kernel launches here.       */
int mvb_profile_set(int on);
void mvb_launch_count_add(int64_t n);
int mvb_profile_read(double* ms_timed, double* flops_timed, double* flops_all, int64_t* n_timed, int64_t* n_all,
                     int max_classes);
const char* mvb_profile_name(int cls);

/* ---- Scaler ------------------------------------------------------------------------------------
 * replaces _Scaler_torch.fit / transform / inverse_transform  (preprocessing/scaler.py:301-428).
 * X is viewed as [R][K]: R = product of the reduced dims (moved first), K = kept elements.
 * mean/var/scale have K entries; var is unbiased; scale = sqrt(var), 0 or NaN -> 1.            */
int mvb_scaler_fit(const void* X, int64_t R, int64_t K, int dtype, void* mean, void* var, void* scale, void* stream);
int mvb_scaler_apply(const void* X, int64_t R, int64_t K, int dtype, const void* mean, const void* scale,
                     int with_mean, int with_std, int inverse, void* out, void* stream);

/* ---- layout helpers ----------------------------------------------------------------------------
 * mvb_trf_pad: (N,C,T) -> (N,C,T+left+right) with replicated edge samples, i.e. the clamped lag
 * indexing of _TimeDelayed_torch._delay_matrices (estimators/timedelayed.py:415-417) turned into
 * a pure shift.  mvb_transpose: [R][C] -> [C][R] (dense 2-D designs become "series" rows).      */
int mvb_trf_pad(const void* X, int N, int C, int T, int left, int right, int dtype, void* Xp, void* stream);
int mvb_transpose(const void* X, int64_t R, int64_t C, int dtype, void* out, void* stream);
/* mvb_trf_shift: centring before multiplying (ridgecv.py:59-94 subtracts the column means from the design itself).
 * shift[c] = mean of feature c of the padded stimulus (N,C,Tp contiguous) over the selected trials; out = xp - shift[c].
 * The centred statistics of mvb_trf_stats are invariant under a per-column shift, so fitting on `out` and adding
 * shift back into the column means / intercept gives the same model without the n mu mu^T cancellation in fp32.     */
int mvb_trf_shift(const void* xp, int N, int C, int Tp, const int32_t* trial_idx, int n_trials, int dtype, void* shift, void* out,
                  void* stream);

/* ---- the implicit lag-expanded design -----------------------------------------------------------
 * Column j = (f, w) of row i = (s, t) is  xp[trial(s)*trial_stride + feat(f)*feat_stride + t + w]:
 * X_t of timedelayed.py:394-423 without ever materialising it.  n_lag = 1 describes a dense 2-D
 * design stored column-major ("series" layout).  trial_idx / feat_idx select training trials and
 * predictor subsets (NULL = all) -- what X[train] and index_select(X, dim, mask)
 * (crossvalidation/validator.py:92, model_selection/hierarchical.py:177-182) do by copying.     */
typedef struct {
  const void* xp;
  int64_t trial_stride, feat_stride;
  int n_time, n_lag;
  const int32_t* trial_idx; int n_trials;
  const int32_t* feat_idx;  int n_feat;
  int dtype;
} mvb_design;

/* targets: y[(trial*n_targets + f)*n_time + t]  (the (N,F,T) layout), same trial subset */

/* ---- generic separable-gather contraction  C[M,N] = sum_k A(m,k) * B(n,k) ------------------------
 * element (r,k) of an operand = base[roff[r] + koff[k]].  contig: 0 rows contiguous, 1 k contiguous,
 * 2 k contiguous + 16-byte aligned groups of 4 (dense row-major).  fp32 -> tcgen05 split-TF32.      */
typedef struct {
  const void* base; const int64_t* roff; const int64_t* koff; int rows; int contig;
} mvb_operand;
size_t mvb_contract_workspace_bytes(int M, int N, int K, int dtype);
int mvb_contract(const mvb_operand* A, const mvb_operand* B, int M, int N, int K, int dtype, void* C, int64_t ldc,
                 int symmetric, void* ws, size_t ws_bytes, void* stream);

/* ---- sufficient statistics of one training set ---------------------------------------------------
 * replaces the lag expansion + centring + X^T X / X^T y work of _RidgeCV_torch
 * (estimators/ridgecv.py:81-88,123 and the products inside :218-224): centred Gram G (p x p),
 * centred cross-moment XtY (p x F), column means mu (p), target means ybar (F).
 * fit_intercept = 0 leaves everything uncentred and returns zeros in mu / ybar.                  */
size_t mvb_trf_stats_workspace_bytes(const mvb_design* d, int n_targets);
int mvb_trf_stats(const mvb_design* d, const void* y, int n_targets, int fit_intercept,
                  void* G, void* XtY, void* mu, void* ybar, void* ws, size_t ws_bytes, void* stream);

/* mvb_gather_stats: statistics of a predictor subset (features feat_idx, all n_lag lags each) as the principal sub-block
 * of the full-predictor statistics of the same fold -- what lets hierarchical_score / shapley_score (hierarchical.py:150-207,
 * shapley.py:64-182: one full refit per predictor set) build X^T X once per fold.  G [p_full][p_full], XtY [p_full][F],
 * mu [p_full] -> G_sub [ps][ps], XtY_sub [ps][F], mu_sub [ps], ps = n_feat * n_lag.                                    */
int mvb_gather_stats(const void* G, const void* XtY, const void* mu, int64_t p_full, int n_targets, const int32_t* feat_idx, int n_feat,
                     int n_lag, int dtype, void* G_sub, void* XtY_sub, void* mu_sub, void* stream);

/* ---- batched symmetric eigensolver ---------------------------------------------------------------
 * replaces torch.linalg.svd / eigh in ridgecv.py:144,218.  A: batch x p x p symmetric (destroyed),
 * Vt: batch x p x p, row k = k-th eigenvector; lam: batch x p (unsorted, >= 0 enforced).        */
size_t mvb_eigh_workspace_bytes(int batch, int p, int dtype);
int mvb_eigh(void* A, void* Vt, void* lam, int batch, int p, int dtype, void* ws, size_t ws_bytes, void* stream);

/* ---- orthogonal reduction to block-tridiagonal form (fp32, tensor cores) ---------------------------
 * Large-p replacement for the decomposition in ridgecv.py:218: G (p x p, symmetric) is padded to
 * P = mvb_band_padded(p) (multiple of 128) and reduced to  Gp = Qt^T T Qt  with Qt (P x P) orthogonal
 * (rows = basis vectors) and T block tridiagonal: Tdiag[j], Tsub[j] = T_{j,j-1} ([P/128][128][128]).
 * mvb_ridge_solve uses it internally (block-Cholesky LOO sweep) when p >= 128 (fp32).                    */
int mvb_band_padded(int p);
size_t mvb_band_reduce_workspace_bytes(int p);
int mvb_band_reduce(const void* G, int p, void* Qt, void* Tdiag, void* Tsub, void* ws, size_t ws_bytes, void* stream);
/* diagnostics (synchronising): host_out[0] = residual blocks that needed the eigen-based orthonormalisation instead of
 * Cholesky QR, host_out[1] = blocks whose polish fell back from Newton-Schulz, during the last reduction in `ws`. */
int mvb_band_reduce_counters(const void* ws, int p, unsigned int* host_out);

/* ---- exact leave-one-out ridge over an alpha grid --------------------------------------------------
 * replaces _RidgeCV_torch.fit after the decomposition (ridgecv.py:223-301): per-row leverages,
 * LOO residuals, loss (A x F), alpha selection (first minimum; shared or per target), coefficients
 * coef (F x p), intercept (F).  G, XtY, mu, ybar come from mvb_trf_stats (G is destroyed).
 * alpha_out has 1 (shared) or F entries, best likewise (int32 grid index).  loss may be NULL.   */
size_t mvb_ridge_solve_workspace_bytes(const mvb_design* d, int n_targets, int n_alphas);
int mvb_ridge_solve(const mvb_design* d, const void* y, int n_targets, void* G, const void* XtY, const void* mu,
                    const void* ybar, const void* alphas, int n_alphas, int fit_intercept, int alpha_per_target,
                    void* coef, void* intercept, void* alpha_out, int32_t* best, void* loss,
                    void* ws, size_t ws_bytes, void* stream);

/* ---- batched small dense LOO ridge (mvb_dense_ridge_batched) -------------------------------------------------
 * replaces the per-slice estimator loop of _Sliding_torch.fit (estimators/sliding.py:575-622) around _RidgeCV_torch.fit
 * (estimators/ridgecv.py:96-303) -- per-timepoint decoding fits one small ridge per time slice -- and the same loop over any
 * stack of equally shaped dense problems.  Problem b of `batch`:  X_b[i][c] = X.base[b * batch_stride + s(i) * sample_stride +
 * c * col_stride], s(i) = sample_idx[i] (NULL = i), likewise Y_b[i][f]; a Sliding view of an (N, C, T) array has
 * sample_stride = C * T, col_stride = T, batch_stride = 1.  Outputs, per problem: coef [batch][F][p], intercept [batch][F],
 * alpha_out / best [batch][F] (alpha_per_target) or [batch], loss [batch][A][F] (may be NULL), gram [batch][p][p] = centred
 * X^T X (may be NULL; RidgeDecoder's patterns, ridgedecoder.py:267-285).  status[b] != 0: problem b met a pivot below
 * 1e-4 x its largest diagonal entry in some (G + alpha I) -- its outputs are not to be trusted and the caller refits it on the
 * general route (mvb_trf_stats + mvb_ridge_solve).  float32, 1 <= p <= mvb_dense_ridge_batched_max_cols(), n > p + 1.       */
typedef struct { const void* base; int64_t sample_stride, col_stride, batch_stride; } mvb_strided;
int mvb_dense_ridge_batched_max_cols(void);
size_t mvb_dense_ridge_batched_workspace_bytes(int batch, int n_samples, int n_cols, int n_targets, int n_alphas);
int mvb_dense_ridge_batched(const mvb_strided* X, const mvb_strided* Y, const int32_t* sample_idx, int n_samples, int n_cols,
                            int n_targets, int batch, const void* alphas, int n_alphas, int fit_intercept, int alpha_per_target,
                            int dtype, void* coef, void* intercept, void* alpha_out, int32_t* best, void* loss, void* gram,
                            int32_t* status, void* ws, size_t ws_bytes, void* stream);
/* the matching prediction, replaces the per-slice loop of _Sliding_torch.predict (estimators/sliding.py:770-831) over
 * _RidgeCV_torch.predict (ridgecv.py:305-323): yhat_b[i][f] = X_b[i] . coef[b][f] + intercept[b][f] (intercept may be NULL),
 * written at yhat[i * out_sample_stride + f * out_target_stride + b * out_batch_stride].                                   */
int mvb_dense_predict_batched(const mvb_strided* X, int n_samples, int n_cols, int batch, const void* coef, const void* intercept,
                              int n_targets, int dtype, void* yhat, int64_t out_sample_stride, int64_t out_target_stride,
                              int64_t out_batch_stride, void* stream);

/* ---- prediction -------------------------------------------------------------------------------------
 * replaces _TimeDelayed_torch.predict / _RidgeCV_torch.predict (timedelayed.py:522-533,
 * ridgecv.py:323): yhat[(s*F + f)*T + t] = sum_j X_t[(s,t), j] coef[f, j] + intercept[f].      */
size_t mvb_trf_predict_workspace_bytes(const mvb_design* d, int n_targets);
int mvb_trf_predict(const mvb_design* d, const void* coef, const void* intercept, int n_targets, void* yhat,
                    void* ws, size_t ws_bytes, void* stream);

/* ---- scoring ----------------------------------------------------------------------------------------
 * replaces math.r2 / math.pearsonr after metrics.score's reduce_ (math/r2.py:80-100,
 * math/pearsonr.py:55-56, metrics/score.py:17-46).  y, yhat viewed as [R][K] (R = reduced samples,
 * e.g. test trials; K = F*T kept cells); y_b has K entries or is NULL (mean of y).  Either output
 * may be NULL.  r2 is 0 where the denominator is <= 0; pearson is NaN on zero variance.          */
int mvb_score(const void* y, const void* yhat, const void* y_b, int64_t R, int64_t K, int dtype,
              void* r2_out, void* pearson_out, void* stream);

/* ---- ranks (widening, SURVEY 8f: spearmanr) --------------------------------------------------------------
 * replaces math.rank (math/rank.py:62-112) for the metric layout: x viewed as [R][K], average ranks (ties get the
 * mean of their positions, 1-based) along R for every one of the K cells.  spearmanr = mvb_score's pearson on ranks. */
int mvb_rank(const void* x, int64_t R, int64_t K, int dtype, void* out, void* stream);

/* ---- ReceptiveField pieces (widening) ----------------------------------------------------------
 * The per-epoch lagged Gram / cross-moment of _ReceptiveField_torch._compute_corrs (estimators/receptivefield.py:976-1118)
 * is mvb_trf_stats on a ZERO-padded stimulus with one trial selected (fit_intercept = 0): its rows are exactly the
 * epoch's samples, which is what the reference's edge correction (:779-807) removes from the FFT correlations.
 *
 * mvb_rf_assemble: G [n_trials][p][p] (kernel lag order) + the centred epochs Xc [n_trials][n_feat][n_time] -> cov
 *   [n_trials][p][p] as the reference lays cov_ out (feature-major, lag s_min.. ascending), reproducing its un-transposed
 *   lower blocks (:1097-1103) and its head correction for negative lags (:801-807).  With edge_correction = 0 pass the
 *   Gram of the extended rows (all rows any lagged sample reaches) instead.
 * mvb_sum_slabs: out = sum_i A[idx[i]] over slabs of slab_elems values -- X_XT[train].sum(0) (:1204-1206).
 * mvb_penalised_solve: for every alpha, (G + alpha * reg) x = rhs by LU with partial pivoting (torch.linalg.solve,
 *   :1224-1229; the matrix is not symmetric).  mats [n_alphas][p][p] is scratch, sol [n_alphas][p][n_rhs] the solutions,
 *   info[s] = 0 or (column + 1) without a non-zero pivot.                                                             */
int mvb_rf_assemble(const void* G, const void* Xc, int n_trials, int n_feat, int n_lag, int n_time, int s_min, int s_max,
                    int edge_correction, int dtype, void* cov, void* stream);
int mvb_sum_slabs(const void* A, int64_t slab_elems, const int32_t* idx, int n_idx, int dtype, void* out, void* stream);
int mvb_penalised_solve(const void* G, const void* reg, const void* alphas, int n_alphas, const void* rhs, int p, int n_rhs,
                        int dtype, void* mats, void* sol, int* info, void* stream);

/* ---- classification metrics (widening): same [R][K] layout as mvb_score -------------------------------
 * mvb_accuracy replaces math.accuracy (math/accuracy.py:13-33): fraction of equal entries along R per cell.
 * mvb_roc_auc replaces the rank formula of math.roc_auc (math/roc_auc.py:139-152): positive != 0 marks the positive class,
 * ties in score count one half, NaN where a cell has only one class.                                                   */
int mvb_accuracy(const void* y, const void* yhat, int64_t R, int64_t K, int dtype, void* out, void* stream);
int mvb_roc_auc(const void* positive, const void* score, int64_t R, int64_t K, int dtype, void* out, void* stream);

/* mvb_pairwise_coupling: multiclass probabilities of a one-versus-one classifier from its pairwise probabilities,
 * replaces _pairwise_coupling_torch + _compute_Qr_torch (estimators/classifier.py:527-653; call site :958-975).
 * pairwise [n_samples][K][K] (entry [j][k] = probability of class j in the (j, k) classifier), out [n_samples][K].
 * All samples iterate until the largest change over the batch is < tol, like the reference; K <= 16.
 * The repeated-index in-place updates of the reference (last write wins) are reproduced.                        */
size_t mvb_pairwise_coupling_workspace_bytes(int64_t n_samples, int n_classes, int max_iter);
int mvb_pairwise_coupling(const void* pairwise, int64_t n_samples, int n_classes, int max_iter, double eps, double tol, int dtype,
                          void* out, void* ws, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MVB_H_ */
