/*
 * bo_b200.h -- C ABI of libbo_b200.so: the sm_100a GP Bayesian-optimisation
 * inner loop that sits *under* pybo's pure-Python plugin surface.
 *
 * The reference (mwhoffman/pybo) has no FFI of its own: its policies call a
 * duck-typed `model` object (the absent `reggie` package).  Each entry point
 * below therefore cites the reference CALL SITE whose arithmetic it replaces.
 * All matrices are C-contiguous row-major float64.  Pointers are host pointers
 * unless the call's `flags` says BO_PTR_DEVICE.  Every function returns an int
 * status (BO_OK == 0) and never throws; bo_last_error() gives the message.
 * One host thread per handle; one CUDA stream per handle.
 */
#ifndef BO_B200_H
#define BO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes */
#define BO_OK 0
#define BO_ERR_CUDA 1      /* a CUDA runtime call or kernel failed            */
#define BO_ERR_NOT_PD 2    /* Cholesky hit a non-positive pivot (LinAlgError) */
#define BO_ERR_ARG 3       /* bad argument                                    */
#define BO_ERR_STATE 4     /* call made before the state it needs exists      */

/* kernel ids */
#define BO_KERNEL_SE 0         /* rho exp(-D/2)                               */
#define BO_KERNEL_MATERN52 1   /* rho (1 + r + r^2/3) exp(-r), r = sqrt(5 D)  */

/* acquisition ids */
#define BO_ACQ_MEAN 0   /* posterior mean            (recommenders.py:22-24)   */
#define BO_ACQ_EI 1     /* model.get_improvement     (policies/simple.py:25)   */
#define BO_ACQ_PI 2     /* model.get_tail            (policies/simple.py:39)   */
#define BO_ACQ_UCB 3    /* mu + sqrt(beta s2)        (policies/simple.py:62-72)*/

/* pointer flags */
#define BO_PTR_HOST 0
#define BO_PTR_DEVICE 1
#define BO_PTR_STAGED 2  /* candidates were generated into the handle by bo_candidates_sobol; Xc is ignored */

/* precision of the scoring contraction */
#define BO_PREC_F64 0      /* FP64 tensor-core (DMMA) path                     */
#define BO_PREC_OZAKI 1    /* error-bounded int8 slice emulation on tcgen05    */

typedef struct bo_ctx bo_ctx;

/* ---- lifetime ----------------------------------------------------------- */
int bo_create(int device, bo_ctx **out);
int bo_destroy(bo_ctx *ctx);
const char *bo_last_error(const bo_ctx *ctx);
/* cudaStream_t of the handle (as void*), so callers can time on it. */
void *bo_stream(bo_ctx *ctx);
int bo_sync(bo_ctx *ctx);
int bo_device_props(bo_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor,
                    size_t *l2_bytes, size_t *hbm_bytes);

/* ---- fit: what `model.add_data` implies (bayesopt.py:114,258,269) --------
 * For each of S hyper-samples s (S == 1: plain GP; S > 1: the MCMC mixture of
 * bayesopt.py:115):  K_s = k_s(X,X) + sn2_s I;  L_s = chol(K_s);
 * alpha_s = L_s^-1 (y - bias_s); also W_s = L_s^-1 (used by the scoring
 * contraction), beta_s = L_s^-T alpha_s and log|L_s|.
 * X: n x d, y: n, ell: S x d, rho/sn2/bias: S.  Returns BO_ERR_NOT_PD if any
 * factorisation fails (bo_fit_info gives sample and pivot). */
int bo_fit(bo_ctx *ctx, int kernel, int n, int d, int S, const double *X,
           const double *y, const double *ell, const double *rho,
           const double *sn2, const double *bias);
int bo_fit_shape(bo_ctx *ctx, int *kernel, int *n, int *d, int *S);
int bo_fit_info(bo_ctx *ctx, int *info /* S */);
/* log marginal likelihood of each hyper-sample (what every MCMC step costs) */
int bo_loglik(bo_ctx *ctx, double *out /* S */);
/* the same quantity WITHOUT building a scoring state (no W = L^-1, no transpose, no handle state touched):
 * Gram + Cholesky of the matrix bordered by the residual row, alpha = L^-1 (y - bias) read off the factor.
 * One likelihood evaluation of the slice sampler behind `MCMC(model, n=10, burn=100)` (bayesopt.py:108-115).
 * Returns BO_ERR_NOT_PD when a factorisation fails. */
int bo_loglik_fit(bo_ctx *ctx, int kernel, int n, int d, int S, const double *X, const double *y,
                  const double *ell, const double *rho, const double *sn2, const double *bias, double *out /* S */);
/* factor read-back, for parity tests: which = 0 L, 1 W=L^-1, 2 alpha, 3 beta */
int bo_get_factor(bo_ctx *ctx, int s, int which, double *out);

/* ---- device-side candidate grid (solvers/lbfgs.py:42-45 builds xgrid on the host; the TODO there asks
 * for a low-discrepancy grid).  Points [start, start + M) of the unscrambled Sobol sequence in the box
 * [lo, hi] (NULL: unit cube) from `bits` direction numbers per dimension (sv: d x bits, e.g. SciPy's
 * qmc.Sobol(d)._sv): x_i = XOR_{b in gray(i)} sv[k][b] / 2^bits -- bit-identical to SciPy / torch.
 * out == NULL: the grid stays in the handle and the next bo_score / bo_predict call passes
 * flags = BO_PTR_STAGED instead of a pointer (no host-to-device copy of candidates at all);
 * otherwise out receives the M x d points (host, or device with BO_PTR_DEVICE). */
int bo_candidates_sobol(bo_ctx *ctx, int d, int bits, const uint32_t *sv, const double *lo,
                        const double *hi, int64_t start, int64_t M, double *out, int flags);

/* ---- incremental refit: `model.add_data(x, y)` inside the loop (bayesopt.py:269)
 * Appends m observations (Xnew: m x d, ynew: m) to the fitted factor set without
 * refactorising: per point and hyper-sample l = W k, lam = sqrt(kss - |l|^2), the
 * new rows of L, W = L^-1 and W^T, alpha, beta and log|L| (two triangular
 * matrix-vector products, O(n^2)); the int8 slice planes get the new row too.
 * The handle has room for bo_fit_capacity() observations (n rounded up to 128);
 * beyond that, or after BO_ERR_NOT_PD, call bo_fit again. */
int bo_append(bo_ctx *ctx, int m, const double *Xnew, const double *ynew);
int bo_fit_capacity(bo_ctx *ctx, int *capacity);

/* ---- the hot call: `finit = f(xgrid, grad=False)` (solvers/lbfgs.py:50) ---
 * Scores M candidates Xc (M x d) with acquisition `acq`:
 *   param = target (EI, PI), beta (UCB), ignored (MEAN).
 * Fused on device: cross-kernel k(X, Xc) -> V = L^-1 k -> mu, s2 -> acquisition
 * (mixture-averaged over the S hyper-samples) -> (max, first argmax).
 * out_val (M) and out_grad (M x d) may be NULL.  best_val/best_idx may be NULL.
 * flags: BO_PTR_DEVICE if Xc/out_val/out_grad are device pointers. */
int bo_score(bo_ctx *ctx, int acq, double param, int64_t M, const double *Xc,
             int flags, double *out_val, double *out_grad, double *best_val,
             int64_t *best_idx);
/* The same pass with the incumbent left on the device: *record points at one 16-byte record
 * {int64 bits of the double value, int64 index + index_offset} in the handle, written on the handle's stream
 * -- nothing about the arg max is copied to the host.  A multi-GPU caller all-gathers the records of its ranks
 * (NCCL, on bo_stream()) and hands the gathered buffer to bo_incumbent_merge.  (solvers/lbfgs.py:50-51 across
 * candidate shards.) */
int bo_score_incumbent(bo_ctx *ctx, int acq, double param, int64_t M, const double *Xc, int flags,
                       double *out_val, int64_t index_offset, void **record);
/* records: count x k packed records (device); val/idx (k, host): per slot the maximum value and the lowest
 * index attaining it over the `count` contributions; NaN never wins.  One kernel + one read-back. */
int bo_incumbent_merge(bo_ctx *ctx, const void *records, int count, int k, double *val, int64_t *idx);
/* model.predict(X, grad) (simple.py:21,64; recommenders.py:22,24,34):
 * mu, s2 (M) and optionally dmu, ds2 (M x d); any output may be NULL. */
int bo_predict(bo_ctx *ctx, int64_t M, const double *Xc, int flags, double *mu,
               double *s2, double *dmu, double *ds2);
/* top-k of the values of the last bo_score call, descending, ties by lowest
 * index: replaces `argsort(finit)[::-1][:nbest]` (solvers/lbfgs.py:51).  NaN values never
 * rank; when fewer than k values are comparable the tail is returned as idx = -1, val = NaN. */
int bo_topk(bo_ctx *ctx, int k, int64_t *idx, double *val);
/* choose the precision path of the scoring contraction (default BO_PREC_F64).
 * For BO_PREC_OZAKI, 0 < tol < 2 is the target absolute error of the entries of
 * V = L^-1 k relative to sqrt(rho); the library picks the number of balanced base-256
 * int8 slices (digits in [-128, 127]) and whether the first dropped pair group is
 * accumulated as well from its error model.  tol >= 2 pins the level: slices = floor(tol)
 * (2..7), extra group iff the fractional part is >= 0.5 (e.g. 5.5).  Gradient requests and
 * batches of <= 64 points always run on the FP64 path. */
int bo_set_precision(bo_ctx *ctx, int prec, double tol);
/* FP64 rescue pass of the int8-slice path (on by default).  Every candidate carries an a-priori
 * bound on the error the slice truncation leaves in its acquisition value (bo_predict: in s2);
 * candidates whose bound exceeds tol * max(|value|, floor_rel * max|value|) are re-scored on the
 * FP64 path inside the same call, so the int8 path meets `tol` wherever the FP64 path does.
 * Defaults: tol = 5e-7 (2x inside the 1e-6 parity bar; the bound itself sits > 2.3x above the largest error the
 * calibration has seen, profiles/r2_oz_calib.txt), floor_rel = 1e-12.  A pass that has
 * to rescue more than a quarter of its candidates sends the following passes on this fit
 * straight to the FP64 path.
 * Tiers (tolerance-selected levels only; a level pinned with tol >= 2 runs exactly as pinned): a flagged list of
 * >= 4096 candidates is first re-scored on the int8 path one level up and only what that cannot certify goes to FP64.
 * Since whatever is not certified is re-scored, the level of the main pass only decides the speed: on passes of
 * >= 8 chunks of 32 768 candidates the first 4096 candidates run one half-level below the selected level, and if at
 * most 10 % of them are flagged there the rest of the pass does too. */
int bo_set_rescue(bo_ctx *ctx, int on, double tol, double floor_rel);
/* what the last bo_score / bo_predict call did: int8_path = 1 if it ran the int8-slice
 * contraction, how many of its `total` candidates the rescue pass re-scored in FP64 */
int bo_rescue_info(bo_ctx *ctx, int *int8_path, int64_t *flagged, int64_t *total);
/* the tiers of the last int8 pass.  Levels are coded 2 * slices + extra_group: level_first = candidate chunk 0,
 * level_rest = the other chunks, level_tier2 = the level the flagged list was re-scored at (0: straight to FP64);
 * first_flagged = candidates the main pass flagged, fp64_rescored = those that ended on the FP64 path
 * (= bo_rescue_info's `flagged`).  All zero after a pass on the FP64 path. */
int bo_tier_info(bo_ctx *ctx, int *level_first, int *level_rest, int *level_tier2, int64_t *first_flagged,
                 int64_t *fp64_rescored);
/* tuning knobs ("oz_cluster" leaves the results bit-identical, the tier knobs keep them within the rescue tolerance).
 * "oz_cluster": CTAs per thread-block cluster of the int8 scoring
 * contraction (1, 2 or 4; 0 = library default): the CTAs of a cluster work on the same candidate tile and adjacent row
 * blocks of W and fetch the K* slice tile once, by TMA multicast.  "oz_tiered" (0 / 1, default 1), "oz_tier_frac"
 * (default 0.10), "oz_tier_min" (default 4096): the tiers described at bo_set_rescue (results stay within the rescue
 * tolerance either way; which tier certified a candidate changes its last digits). */
int bo_set_option(bo_ctx *ctx, const char *key, double value);
/* the a-priori error model behind the rescue pass, for the level currently selected: on the int8 path
 * |s2 - s2_exact| <= errk[s] * sqrt(q rho_s), q = rho_s - s2, for hyper-sample s (errk: S values) */
int bo_ozaki_error_bound(bo_ctx *ctx, double *errk);
/* current path and, for BO_PREC_OZAKI after a scoring call, the level in use encoded as
 * 2 * slices + extra (extra = the digit pairs of group g = slices are accumulated too) */
int bo_precision_info(bo_ctx *ctx, int *prec, int *slices);

/* ---- Thompson: `model.sample_f(n, rng).get` (policies/simple.py:48) ------
 * ndraw weight-space posterior draws
 *   f_r(x) = bias_r + scale_r * sum_j cos(W_r[j] . x + b_r[j]) theta_r[j]
 * with m random Fourier features each.  The frequencies/phases may be shared
 * by all draws (nW == 1) or be per draw (nW == ndraw).
 * W: nW x m x d, b: nW x m, theta: ndraw x m, scale/bias: ndraw. */
int bo_thompson_set(bo_ctx *ctx, int ndraw, int nW, int m, int d, const double *W,
                    const double *b, const double *theta, const double *scale,
                    const double *bias);
/* Build the draws on the device -- the work `model.sample_f(n, rng)` does before its `.get` is evaluated
 * (policies/simple.py:48): for a GP with data X (n x d), y (n), signal variance rho, noise sn2 and constant
 * mean `bias`, and m random Fourier features phi(x) = sqrt(2 rho / m) cos(W x + b),
 *   theta_r = A^-1 Phi^T (y - bias) + sqrt(sn2) L^-T eps_r,   A = Phi^T Phi + sn2 I = L L^T,  Phi = phi(X),
 * i.e. theta_r ~ N(A^-1 Phi^T r, sn2 A^-1), then installs them as bo_thompson_set would (scale_r = sqrt(2 rho / m),
 * bias_r = bias).  W: nW x m x d, b: nW x m (nW == 1: one basis shared by all draws; nW == ndraw: one basis per
 * draw, i.e. ndraw independent sample_f calls), noise: ndraw x m standard normals from the caller's random stream.
 * theta_out (ndraw x m, host) may be NULL.  BO_ERR_NOT_PD if a feature system fails to factor. */
int bo_thompson_build(bo_ctx *ctx, int n, int d, const double *X, const double *y, double rho, double sn2,
                      double bias, int ndraw, int nW, int m, const double *W, const double *b,
                      const double *noise, double *theta_out);
/* out (ndraw x M) / out_grad (ndraw x M x d) may be NULL; best_* (ndraw, host)
 * may be NULL.  flags: BO_PTR_DEVICE if Xc/out/out_grad are device pointers. */
int bo_thompson_eval(bo_ctx *ctx, int64_t M, const double *Xc, int flags,
                     double *out, double *out_grad, double *best_val,
                     int64_t *best_idx);

/* per-draw arg max over M candidates left on the device as ndraw packed records (see bo_score_incumbent) */
int bo_thompson_incumbents(bo_ctx *ctx, int64_t M, const double *Xc, int flags, int64_t index_offset,
                           void **records, int *ndraw);

/* ---- stand-alone pieces (metric "Cholesky GB/s", parity tests) ----------- */
/* In-place lower Cholesky of `batch` n x n matrices (only the lower triangle
 * is read; the strict upper triangle of the result is zeroed). info[b] = 0 or
 * 1-based failing pivot.  Replaces scipy.linalg.cholesky (LAPACK dpotrf). */
int bo_cholesky(bo_ctx *ctx, int n, int batch, double *A, int flags, int *info);
/* K = k(X,X) + sn2 I  (n x n), the Gram matrix `add_data` builds. */
int bo_gram(bo_ctx *ctx, int kernel, int n, int d, const double *X,
            const double *ell, double rho, double sn2, double *K, int flags);

/* ---- live per-kernel timing (CUDA events on the handle's stream) --------- */
int bo_profile_enable(bo_ctx *ctx, int on);
int bo_profile_reset(bo_ctx *ctx);
/* number of distinct kernels seen; then name/launch-count/total ms of each */
int bo_profile_count(bo_ctx *ctx, int *count);
int bo_profile_get(bo_ctx *ctx, int i, char *name, int name_cap, int64_t *launches,
                   double *total_ms);
/* total kernel launches issued by this handle since creation */
int bo_launch_count(bo_ctx *ctx, int64_t *launches);
/* measured FP64 roof of this device: kind 0 = tensor-core DMMA m8n8k4,
 * kind 1 = DFMA, both register resident on every SM; result in TFLOP/s. */
int bo_microbench(bo_ctx *ctx, int kind, int iters, double *tflops);
/* self-test hook of the int8-slice path: runs it on hyper-sample 0 for the first mc
 * candidates and returns mu, s2, the raw int32 group accumulators of candidate tile 0
 * ([np/64][S + extra][128][64]) and the slice planes, for exact comparison on the host. */
int bo_ozaki_debug(bo_ctx *ctx, int S, int extra, int mc, const double *Xc, double *mu, double *s2,
                   int32_t *acc, int8_t *wslices, int8_t *kslices, double *rowscale);

#ifdef __cplusplus
}
#endif
#endif /* BO_B200_H */
