/*
 * barefoot_sm100.h - C ABI of libbarefoot_sm100.so, the B200 (sm_100a) implementation of
 * BAREFOOT's data-parallel inner loop (GP posterior -> reification/fusion -> acquisition ->
 * batch selection).
 *
 * The reference (/root/reference) has no FFI: its seam is a Python callable mapped over work
 * items (barefoot.py:1070-1073, 1271-1283, 2001-2013).  Each export below replaces the numerics
 * that one reference function runs on the CPU for ONE work item with a batched device call over
 * all hyper-parameter sets x all candidates; the reference site is cited per function.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless it says "host"; FP64 (double) / int64 / int32,
 *    row-major, contiguous.  The caller owns every buffer including workspace; the library never
 *    allocates or frees device memory and keeps no global state except a thread-local error text.
 *  - all launches are asynchronous on `stream` (a cudaStream_t passed as void*).
 *  - return value: 0 = ok, < 0 = bad argument (BF_ERR_*), > 0 = a cudaError_t.
 *  - "set" = one (hyper-parameter set[, fantasy point]) instance of a GP sharing the training
 *    inputs X; `set_hp` (optional, may be NULL = identity) maps a set to its row of inv_l2/amp.
 *  - np = n rounded up to a multiple of 32 (bf_padded_n).  Factor matrices are np x np row-major.
 */
#ifndef BAREFOOT_SM100_H
#define BAREFOOT_SM100_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BF_VERSION 100

/* covFunc: gpModel.py:29-36 */
#define BF_KERN_SE  0
#define BF_KERN_M32 1
#define BF_KERN_M52 2

/* acquisition mask bits for bf_gp_posterior (acquisitionFunc.py:98-214, util.py:655-665) */
#define BF_ACQ_EI     1
#define BF_ACQ_PI     2
#define BF_ACQ_UCB    4
#define BF_ACQ_GREEDY 8   /* also produces the top-2 of the mean that bf_kg_diag needs */
/* slots of best_val / best_idx rows written by bf_acq_finalize */
#define BF_SLOT_EI     0
#define BF_SLOT_PI     1
#define BF_SLOT_UCB    2
#define BF_SLOT_GREEDY 3
#define BF_SLOT_TOP2   4  /* value only: second largest mean (by position) */
#define BF_NSLOT_VAL   5
#define BF_NSLOT_IDX   4

#define BF_ERR_ARG      (-1)
#define BF_ERR_SIZE     (-2)   /* n, d or m outside what the kernels support */
#define BF_ERR_ALIGN    (-3)

#define BF_MAX_DIM      16     /* nDim supported by the cross-covariance generators */
#define BF_MAX_MODELS   8      /* m supported by bf_reification */

int         bf_version(void);
const char* bf_last_error(void);

/* ---- layout helpers (host side, pure arithmetic) ------------------------------------------ */
int    bf_padded_n(int n);                 /* np: n rounded up to 32 */
size_t bf_linv_packed_doubles(int n);      /* doubles per set of the packed L^-1 operand */
int    bf_posterior_tile(int n);           /* candidates per CTA chosen for this n (0: too large) */
int    bf_posterior_parts(int n, long long N, int S); /* partial-result slots (CTAs) per set of a
                                                          bf_gp_posterior call over S sets */

/* ---- (1)+(2) setup of a GP: gpModel.py:17-42 (george GP.compute) --------------------------- */
/* K[s] = amp * k(X, X; inv_l2) + diag((sqrt(yerr^2 + 1.25e-12))^2), written as an np x np
 * matrix padded with the identity.  noise is per set (noise_stride = n), shared (noise_stride =
 * 0, n values) or scalar (noise_n = 1). */
int bf_kernel_matrix(int kern, int n, int d, int S,
                     const double* X, const double* inv_l2, const double* amp,
                     const int32_t* set_hp,
                     const double* yerr, int yerr_n, long long yerr_stride,
                     double* K, void* stream);
/* Only the 32 x 32 blocks on and below the diagonal are written (all the factorisation reads; it
 * clears the blocks above the diagonal itself): half the kernel evaluations and half the writes. */
int bf_kernel_matrix_lower(int kern, int n, int d, int S,
                     const double* X, const double* inv_l2, const double* amp,
                     const int32_t* set_hp,
                     const double* yerr, int yerr_n, long long yerr_stride,
                     double* K, void* stream);

/* In-place lower Cholesky of S padded np x np matrices (george BasicSolver.compute ->
 * scipy.linalg.cholesky, reached from gpModel.py:41).  info[s] = 0 or 1 + index of the first
 * non-positive pivot (george would raise LinAlgError). */
int bf_cholesky_batched(int n, int S, double* K_inout_L, int32_t* info, void* stream);

/* The same factorisation, also storing the inverses of the 32 x 32 diagonal blocks of L:
 * dinv [S x (np / 32) x 32 x 32] row-major - the operands of the solve path (bf_gp_posterior_solve).
 * Available when bf_cholesky_dinv_supported(n) (padded size <= 736). */
int bf_cholesky_dinv_supported(int n);
int bf_cholesky_batched_dinv(int n, int S, double* K_inout_L, int32_t* info, double* dinv, void* stream);

/* L^-1 (row-major np x np into Linv, and the tensor-core operand layout into linv_packed) and
 * alpha = K^-1 y (george GP._compute_alpha -> cho_solve, reached from gpModel.py:47,53).
 * y is per set (y_stride = n) or shared (y_stride = 0). */
int bf_factor_finish(int n, int S, const double* L, const double* y, long long y_stride,
                     double* Linv, double* linv_packed, double* alpha, void* stream);
/* Kernels one bf_factor_finish call launches: 1 (one CTA per set) or, for a few sets, 3 (diagonal
 * block inverses; column groups of L^-1 spread over 4 np/32 CTAs per set; alpha), plus one that clears the
 * padding cells of the packed operand when n is not a multiple of 32.
 * linv_packed is the operand of bf_gp_posterior: the n x n inverse at the END of the padded index range
 * (element (i, j) at padded position (i + np - n, j + np - n)), zero in front. */
int bf_factor_finish_kernels(int n, int S, int with_alpha);

/* ---- (1)+(2)+(4) posterior + elementwise acquisitions: gpModel.py:50-54,
 * reificationFusion.py:189-206, acquisitionFunc.py:98-214 ------------------------------------ */
/* For every set s and candidate c: k* = amp k(x*_c, X); mu = k*.alpha; var = amp - |L^-1 k*|^2;
 * mu' = mu * scale[s] + shift[s]; var' = var * scale[s]^2 (scale/shift may be NULL = 1/0).
 * mu_out / var_out ([S x N]) are optional.  If acq_mask != 0 the per-tile partial maxima are
 * written to part_val [S x parts x 5] / part_idx [S x parts x 4], parts = bf_posterior_parts(n, N, S);
 * reduce with bf_acq_finalize(S, parts, ...).
 * spread_is_sqrt: PI/UCB receive sqrt(var') (ROM stage, util.py:451,519) instead of var'
 * (TM/batch stage, util.py:598-612).  EI always receives var' (util.py:323-326).
 * Xw [H x n x d] are the training inputs pre-scaled by the metric of every hyper-parameter row
 * (bf_prescale_train, once per GP); Xs are the plain candidates.
 * pair_ctr: optional scratch of S 32-bit words (any contents; cleared by the call on `stream`).  With it the
 * candidate tiles of a set are handed to the set's CTAs dynamically (the launch ends within one tile pair of a
 * perfect balance); NULL = static split.  Results are identical either way. */
int bf_prescale_train(int kern, int n, int d, int H, const double* X, const double* inv_l2,
                      double* Xw, void* stream);
int bf_gp_posterior(int kern, long long N, int n, int d, int S,
                    const double* Xs, const double* Xw,
                    const double* inv_l2, const double* amp, const int32_t* set_hp,
                    const double* linv_packed, const double* alpha,
                    const double* scale, const double* shift,
                    double* mu_out, double* var_out,
                    int acq_mask, int spread_is_sqrt, double curr_max, double xi, double kt,
                    double* part_val, long long* part_idx, uint32_t* pair_ctr, void* stream);

/* The same posterior for FEW candidates per set (the literal ROM stage: every work item of
 * barefoot.py:1001-1073 factorises its own fused GP and evaluates only sampleCount ~ 50 candidates,
 * util.py:243-262).  No L^-1 and no alpha: V = L^-1 [k* | y] by blocked forward substitution,
 * mu = v_c . t, var = amp - v_c . v_c (t = L^-1 y).  L [S x np x np] and dinv are the outputs of
 * bf_cholesky_batched_dinv; y [S x n] (y_stride = n) or shared (y_stride = 0).  Outputs and the
 * partial records as bf_gp_posterior, with parts = bf_posterior_solve_parts(n, N) CTAs per set of
 * bf_posterior_solve_cols(n) candidates each (0 = n too large for this path). */
int bf_posterior_solve_cols(int n);
int bf_posterior_solve_parts(int n, long long N);
int bf_gp_posterior_solve(int kern, long long N, int n, int d, int S,
                          const double* Xs, const double* X,
                          const double* inv_l2, const double* amp, const int32_t* set_hp,
                          const double* L, const double* dinv, const double* y, long long y_stride,
                          const double* scale, const double* shift,
                          double* mu_out, double* var_out,
                          int acq_mask, int spread_is_sqrt, double curr_max, double xi, double kt,
                          double* part_val, long long* part_idx, void* stream);

/* Reduce per-tile partials: best_val [S x 5], best_idx [S x 4]; first index wins ties
 * (np.where(v == max)[0][0], acquisitionFunc.py:135-138). */
int bf_acq_finalize(int S, int tiles, int acq_mask, const double* part_val,
                    const long long* part_idx, double* best_val, long long* best_idx,
                    void* stream);

/* ---- (4) knowledge gradient as the reference calls it, i.e. on np.diag(var):
 * acquisitionFunc.py:12-96 via util.py:259-262,618-621 ---------------------------------------- */
/* mu, var [S x N]; top [S x 5] / top_idx [S x 4] as written by bf_acq_finalize with
 * BF_ACQ_GREEDY (slot GREEDY = largest mean and its first index, slot TOP2 = second largest).
 * Candidates i < M are scored.  Output: kg_val[S], kg_idx[S] with the reference's floor (0, 0).
 * part_* are [S x tiles] scratch, tiles = bf_kg_tiles(N).  nu_out ([S x N], optional) receives
 * every log-KG value. */
int bf_kg_tiles(long long N);
/* top [S x 5] / top_idx [S x 4]: fills slots GREEDY (largest mean, first index) and TOP2 from
 * mu [S x N] for callers whose mean did not come out of bf_gp_posterior. */
int bf_top2(long long N, int S, const double* mu, double* top, long long* top_idx, void* stream);
int bf_kg_diag(long long N, long long M, int S, const double* mu, const double* var,
               const double* top, const long long* top_idx, double sn,
               double* nu_out, double* part_val, long long* part_idx,
               double* kg_val, long long* kg_idx, void* stream);

/* ---- full posterior covariance: gpModel.py:44-48 (george predict(..., return_cov=True)), called per
 * hyper-parameter set by the batch-only path (util.py:952) -------------------------------------- */
/* One set (pass that set's inv_l2 row, amp, Linv, alpha).  Ks, Vt: [Ap x np] scratch / output, Ap = A
 * rounded up to 32.  Ks = amp k(Xa, X) zero padded; Vt = Ks L^-T (row a = L^-1 k(X, xa));
 * mean[a] = Ks[a] . alpha (optional).  Linv is the row-major np x np L^-1 of bf_factor_finish. */
int bf_gp_solve_rows(int kern, long long A, int n, int d, const double* Xa, const double* X,
                     const double* inv_l2, double amp, const double* Linv, const double* alpha,
                     double* Ks, double* Vt, double* mean, void* stream);
/* C[a, b] = amp k(xa, xb) - Vta[a] . Vtb[b]  for a < A, b < B; C has leading dimension ldc. */
int bf_gp_cov(int kern, long long A, long long B, int n, int d, const double* Xa, const double* Xb,
              const double* inv_l2, double amp, const double* Vta, const double* Vtb, double* C,
              long long ldc, void* stream);

/* ---- (4) stand-alone elementwise acquisitions on given vectors: expected_improvement
 * (acquisitionFunc.py:130-138), probability_improvement (:172-176), upper_conf_bound (:210-214),
 * greedy first arg-max (util.py:655-665) ------------------------------------------------------- */
/* kind = exactly one BF_ACQ_* bit.  y, spread [S x N] (spread may be NULL for GREEDY).  out
 * ([S x N], optional) receives every value; best_val / best_idx [S] the maximum and its FIRST
 * index.  part_* are [S x tiles] scratch, tiles = bf_acq_tiles(N). */
int bf_acq_tiles(long long N);
int bf_acq_eval(int kind, long long N, int S, const double* y, const double* spread,
                double curr_max, double xi, double kt, double* out, double* part_val,
                long long* part_idx, double* best_val, long long* best_idx, void* stream);

/* Candidate-sharded stages over several GPUs: every rank holds the per-set maximum over ITS
 * candidates.  bf_records_pack writes val [S] and (idx + offset) [S] into one [2 x S] buffer for a single
 * all-gather; bf_records_reduce turns the gathered [world x 2 x S] buffer into the per-set maximum with numpy's
 * first-index tie rule over the concatenated candidates (lowest global index among equal values). */
int bf_records_pack(int S, const double* val, const long long* idx, long long offset, double* out, void* stream);
int bf_records_reduce(int world, int S, const double* gathered, double* best_val, long long* best_idx, void* stream);

/* ---- (4) knowledge gradient on a DENSE covariance (Frazier's sort / unique / upper-envelope
 * algorithm): acquisitionFunc.py:12-96 as the batch-only path calls it, util.py:1078-1081 ------- */
/* mu [S x N], sigma [S x N x N] (column i of sigma is read for candidate i, as the reference does),
 * candidates i < M scored.  nu_out [S x M] receives every log-KG value; kg_val / kg_idx [S] the
 * running maximum from the reference's floor (0, 0).  N <= BF_KG_FULL_MAX_N. */
/* N <= BF_KG_FULL_MAX_N: the lines of a candidate are sorted in shared memory (one CTA per candidate and set).
 * Larger N (up to BF_KG_FULL_BIG_MAX_N): the same algorithm on a per-CTA scratch slice in device memory with a
 * persistent grid over all (candidate, set) items; `workspace` must then hold bf_kg_full_workspace(N, M, S) bytes
 * (it may be NULL otherwise).  Either way ONE call covers all S sets. */
#define BF_KG_FULL_MAX_N 4096
#define BF_KG_FULL_BIG_MAX_N 65536
int bf_kg_full_max_n(void);
size_t bf_kg_full_workspace(int N, long long M, int S);
int bf_kg_full(int N, long long M, int S, const double* mu, const double* sigma, double sn,
               double* nu_out, double* kg_val, long long* kg_idx, void* workspace, void* stream);

/* ---- (3) reification: reificationFusion.py:26-57 ------------------------------------------- */
/* y, var: [m x P] (model-major).  mean_f, var_f: [P].  pinv semantics of numpy (rcond 1e-15). */
int bf_reification(int m, long long P, const double* y, const double* var,
                   double* mean_f, double* var_f, void* stream);

/* ---- (5) k-medoids primitives for kmedoids_max: util.py:38-49 (sklearn_extra KMedoids) ------ */
/* Every distance is sklearn's pairwise_distances(X, metric="euclidean") value, i.e. the EXPANDED form
 * sqrt(max(-2 x.y + |x|^2 + |y|^2, 0)) with a zero diagonal, operation for operation (bf_kmedoids.cu). */
/* rowsum[i] = sum_j |x_i - x_j| for rows [row0, row0 + rows) against all N rows. */
int bf_kmedoids_rowsum(long long N, int f, const double* X, long long row0, long long rows,
                       double* rowsum, void* stream);
/* labels[i] = first argmin_c |x_i - x_medoid[c]| */
int bf_kmedoids_assign(long long N, int f, int k, const double* X, const long long* medoid_idx,
                       int32_t* labels, void* stream);
/* perm: the rows sorted by (label, index); seg [k + 1]: segment offsets into perm.
 * cost[p] = sum over q in the same segment of |x_perm[p] - x_perm[q]| (in segment order). */
int bf_kmedoids_cluster_cost(long long N, int f, int k, const double* X, const long long* perm,
                             const long long* seg, double* cost, void* stream);
/* The same for the chunks (128 consecutive positions of one cluster, numbered cluster-major) with
 * chunk % nparts == part: the ranks of a multi-GPU job take one part each; positions of other parts
 * are left untouched, so zero-filled cost arrays combine by addition. */
int bf_kmedoids_cluster_cost_part(long long N, int f, int k, const double* X, const long long* perm,
                                  const long long* seg, int part, int nparts, double* cost, void* stream);
/* Stable counting sort of labels (np.where(labels == c)[0] for every c, util.py:42-45 and the
 * alternate loop of KMedoids): perm [N] = rows ordered by (label, row), seg [k + 1] = offsets.
 * workspace: bf_kmedoids_sort_workspace(N, k) bytes of device memory. */
size_t bf_kmedoids_sort_workspace(long long N, int k);
int bf_kmedoids_sort_labels(long long N, int k, const int32_t* labels, long long* perm, long long* seg,
                            void* workspace, void* stream);
/* Per cluster: first argmin of cost, and the alternate-method rule "move only if strictly
 * better than the current medoid".  medoid_idx is updated in place; changed[0] counts moves. */
int bf_kmedoids_update(long long N, int k, const long long* perm, const long long* seg,
                       const double* cost, long long* medoid_idx, int32_t* changed, void* stream);

/* The update when the cost chunks are split over nparts ranks: records [k x 3] = per cluster (smallest cost
 * among this part's positions, that position, cost at the current medoid's position if it is in this part),
 * +inf for "none".  The ranks all-gather their records ([nparts x k x 3]) and bf_kmedoids_update_final
 * applies the first-argmin / strictly-better rule; the result does not depend on nparts. */
int bf_kmedoids_update_partial(long long N, int k, const long long* perm, const long long* seg,
                               const double* cost, const long long* medoid_idx, int part, int nparts,
                               double* records, void* stream);
int bf_kmedoids_update_final(int k, int nparts, const double* records, const long long* perm,
                             long long* medoid_idx, int32_t* changed, void* stream);
/* One iteration of the alternate loop in one call: assign, sort_labels, cluster_cost_part, then update
 * (nparts == 1) or update_partial (nparts > 1, records required). */
int bf_kmedoids_iteration(long long N, int f, int k, const double* X, long long* medoid_idx,
                          int32_t* labels, long long* perm, long long* seg, void* workspace, int part,
                          int nparts, double* cost, double* records, int32_t* changed, void* stream);

/* ---- k-medoids on an explicit distance matrix: the vendored kMedoids(D, k, tmax), kmedoids.py:168-227 -------- */
/* D [n x ld] row-major.  assign: labels[i] = first argmin_c D[i, M[c]] (:208).  update: after
 * bf_kmedoids_sort_labels, cost[p] = mean of D[perm[p], perm[q]] over p's cluster (:213) and Mnew[c] = the
 * cluster's first minimum (:214-215, always moves); empty[0] counts empty clusters (zero it first). */
int bf_kmedoids_matrix_assign(long long n, int k, const double* D, long long ld, const long long* M,
                              int32_t* labels, void* stream);
int bf_kmedoids_matrix_update(long long n, int k, const double* D, long long ld, const long long* perm,
                              const long long* seg, double* cost, long long* Mnew, int32_t* empty, void* stream);

/* ---- clustering glue: barefoot.__kg_calc_clustering, barefoot.py:2102-2140 ---------------------- */
/* Records r < R carry key[r] in [0, G) (selected candidate x* and model folded into one key) and
 * value[r].  Per group, the reference's in-order scan: the first record sets (value, row), a later one
 * replaces them only if strictly greater (NaN never replaces; a NaN first record stays).
 * first_row[g] = first record of the group (R if empty), best_val[g] / best_row[g] = the kept pair
 * (0 / -1 if empty).  bad_key[0] is incremented for every key outside [0, G) (zero it first). */
int bf_group_best(long long R, long long G, const long long* key, const double* value,
                  long long* first_row, double* best_val, long long* best_row, int32_t* bad_key,
                  void* stream);

/* ---- small on-device helpers for the fused-GP targets: reificationFusion.py:147-159 -------- */
/* out[i] = a[i] * scale + shift */
int bf_affine(long long n, const double* a, double scale, double shift, double* out, void* stream);
/* total_var[i] = (err_mu[i] * err_std + err_mean)^2 + var[i] * model_std^2 (lines :148-152) */
int bf_total_variance(long long n, const double* err_mu, double err_std, double err_mean,
                      const double* var, double model_std, double* out, void* stream);
/* z-score with the population std (np.std), std 0 -> 1 (lines :154-159):
 * z = (f - mean) / std; zerr = |v / std^2|^0.5; stats[0] = mean, stats[1] = std. */
int bf_zscore_targets(int n, const double* f_mean, const double* f_var, double* z, double* zerr,
                      double* stats, void* stream);

/* ---- fused kernel matrix + Cholesky for many sets (gpModel.py:38-42 per work item) -------------------- */
/* L [S x np x np] <- chol(amp k(X, X) + diag(yerr^2 + 1.25e-12)) per set with the arguments of
 * bf_kernel_matrix, without ever storing K: a left-looking blocked factorisation that generates the kernel
 * tiles into its accumulators (bf_cholesky_ll.cu).  The 32 x 32 blocks strictly above the diagonal are left
 * untouched unless zero_upper is set.  info [S] as bf_cholesky_batched; dinv (optional) as
 * bf_cholesky_batched_dinv. */
int bf_cholesky_fused_supported(int n, int d);
/* The same left-looking factorisation on a buffer that already holds K in its lower blocks
 * (bf_kernel_matrix_lower): in place, like bf_cholesky_batched(_dinv). */
int bf_cholesky_ll_inplace(int n, int S, double* K_inout_L, int32_t* info, double* dinv, int zero_upper,
                           void* stream);
int bf_cholesky_fused(int kern, int n, int d, int S, const double* X, const double* inv_l2,
                      const double* amp, const int32_t* set_hp, const double* yerr, int yerr_n,
                      long long yerr_stride, double* L, int32_t* info, double* dinv, int zero_upper,
                      void* stream);

/* G independent z-scorings (f_mean, f_var, z, zerr: [G x n]; stats [G x 2]): bf_zscore_targets per group. */
int bf_zscore_targets_batched(int G, int n, const double* f_mean, const double* f_var, double* z,
                              double* zerr, double* stats, void* stream);
/* ROM-stage fantasy updates (util.py:243 -> reificationFusion.py:169-178 -> gpModel.py:56-64: append one point
 * and refactor) for K fantasy points of ONE model at once, as the bordered (rank-one) form of the appended
 * factor.  Inputs in the GP's normalised units from the UN-updated GP: mu_f, var_f [F] at the fused points;
 * mu_t, var_t [K] at the fantasy points; C [F x ldc] posterior covariance between them (bf_gp_cov); y_new [K]
 * the normalised fantasy targets (minus the GP's constant mean); noise_var = sigma_n^2 + 1.25e-12.
 * Outputs [K x F]: the de-normalised mean row and the total variance row (err_term [F] + var' model_std^2) that
 * reification consumes (reificationFusion.py:147-152). */
int bf_fantasy_rows(int F, int K, const double* mu_f, const double* var_f, const double* C, long long ldc,
                    const double* mu_t, const double* var_t, const double* y_new, double noise_var,
                    double gp_mean, double model_std, double model_mean, const double* err_term,
                    double* mean_rows, double* var_rows, void* stream);

/* ---- measurement aid: FP64 tensor (DMMA) issue rate, the roofline denominator bench.py quotes next to
 * cuBLAS DGEMM.  ctas CTAs of 8 warps x 12 accumulator chains x iters mma.sync.m8n8k4.f64; out receives
 * ctas * 256 doubles; bf_probe_dmma_flop = flop of one such launch. */
double bf_probe_dmma_flop(int ctas, int iters);
int bf_probe_dmma(int ctas, int iters, double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
