// common.cuh -- handle layout, error plumbing and the live event profiler of
// libbo_b200.so.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/bo_b200.h"

#define BO_PAD 128           // every factor dimension is padded to a multiple of this
#define BO_NB 64             // Cholesky panel width
#define BO_MAX_D 32          // largest supported input dimension

static inline int bo_round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline int64_t bo_round_up64(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

struct bo_prof_entry {
    std::string name;
    int64_t launches = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
    double total_ms = 0.0;
};

struct bo_thompson_state {
    int ndraw = 0, nW = 0, m = 0, d = 0;
    double *dBestVal = nullptr;
    int64_t *dBestIdx = nullptr;
    double *W = nullptr, *b = nullptr, *theta = nullptr, *scale = nullptr, *bias = nullptr;
    double *thetaT = nullptr;   // (m padded to 16) x (ndraw padded to 256), shared-basis contraction operand
    int ndp = 0;
    // int8-slice path of the shared-basis batch (ozaki.cu): host copies, slice planes of Theta, padded basis
    std::vector<double> h_theta, h_scale, h_bias, h_W, h_b;
    int8_t *ozTheta = nullptr, *ozPhi = nullptr;
    double *ozRowScale = nullptr, *ozRowBias = nullptr, *ozWp = nullptr, *ozBp = nullptr;
    size_t ozTheta_capacity = 0, ozPhi_capacity = 0, ozRow_capacity = 0, ozWp_capacity = 0;
    int oz_mp = 0, oz_ndp = 0, oz_dpad = 0, oz_S = 0, oz_extra = 0, oz_mp_hint = 0;
    bool oz_ready = false;
    double *build = nullptr;            // bo_thompson_build scratch (features, feature system, its factor and inverse)
    size_t build_capacity = 0;
};

struct bo_chol_graph {
    double *A = nullptr, *dinv = nullptr;
    int *info = nullptr, *flags = nullptr;
    int np = 0, batch = 0;
    int64_t launches = 0;
    uint64_t stamp = 0;
    cudaGraphExec_t exec = nullptr;
};

#define BO_CHOL_MAX_LANES 8
#define BO_OZ_ERRK_SLOTS 4

struct bo_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;       // low-priority side stream (operand slicing overlaps the contraction)
    cudaEvent_t ev_sliced[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
    cudaStream_t chol_lane_main[BO_CHOL_MAX_LANES - 1] = {}, chol_lane_side[BO_CHOL_MAX_LANES - 1] = {};   // extra stream pairs of the batched Cholesky (lazy)
    cudaEvent_t chol_lane_ev[BO_CHOL_MAX_LANES - 1][5] = {}, chol_lane_fork = nullptr;
    char err[512] = {0};
    int sm_count = 0;
    cudaDeviceProp prop;

    // ---- fit state --------------------------------------------------------
    bool fitted = false;
    int kernel = 0, n = 0, np = 0, d = 0, dp = 0, S = 0;
    double *dX = nullptr;       // n x d raw observations
    double *dXs = nullptr;      // S x np x dp : X / ell_s, zero padded
    double *dY = nullptr;       // n
    double *dInvEll = nullptr;  // S x dp (zero padded)
    double *dRho = nullptr, *dSn2 = nullptr, *dBias = nullptr;   // S each
    double *dL = nullptr;       // S x np x np   Cholesky factor (lower)
    double *dW = nullptr;       // S x np x np   W = L^-1 (lower)
    double *dWT = nullptr;      // S x np x np   W^T (upper) -- gradient path
    double *dDinv = nullptr;    // S x (np/64) x 64 x 64 inverted diagonal blocks
    double *dAlpha = nullptr;   // S x np   L^-1 (y - bias)
    double *dBeta = nullptr;    // S x np   L^-T alpha
    double *dLogdet = nullptr;  // S        sum log diag L
    int *dInfo = nullptr;       // S
    double *dAppend = nullptr;  // bo_append scratch: kvec[np], lvec[np], per-sample scalars
    int *dAppendInfo = nullptr;
    size_t append_capacity = 0, appendinfo_capacity = 0;
    double *dCholDinv = nullptr;   // stand-alone bo_cholesky scratch
    int *dCholInfo = nullptr;
    int *dCholFlags = nullptr;     // per (matrix, panel) 'diagonal block factored' flags of the fused step kernel
    size_t cholflags_capacity = 0;
    std::vector<bo_chol_graph> chol_graphs;    // captured factorisation patterns (linalg.cu bo_linalg_cholesky)
    std::vector<bo_chol_graph> chol_seen;      // patterns requested once so far (captured when they come back)
    uint64_t chol_graph_clock = 0;
    int chol_group = 0;                        // panels per trailing update of the launch-per-step Cholesky (0: default)
    int chol_flow_grid = 0;                    // co-resident CTAs of the persistent factorisation kernel (0: unavailable)
    size_t choldinv_capacity = 0, cholinfo_capacity = 0;
    std::vector<double> h_rho, h_sn2, h_bias, h_ell;
    std::vector<int> h_info;
    size_t fit_capacity = 0;    // S*np*np currently allocated
    size_t small_capacity[11] = {0};
    size_t info_capacity = 0;

    // ---- scoring scratch ---------------------------------------------------
    int64_t chunk = 0;          // candidates per chunk (multiple of 128)
    double *dKs = nullptr;      // np x chunk cross-kernel tile
    double *dV = nullptr;       // np x chunk (gradient path)
    double *dU = nullptr;       // np x chunk (gradient path)
    double *dQpart = nullptr;   // (np/128) x chunk
    double *dPpart = nullptr;   // (np/128) x chunk
    double *dMuS = nullptr;     // S x chunk
    double *dS2S = nullptr;     // S x chunk
    double *dDmuS = nullptr;    // S x chunk x d
    double *dDs2S = nullptr;    // S x chunk x d
    double *dGpart = nullptr;   // gradient partial sums
    size_t ks_capacity = 0, mom_capacity = 0, gm_capacity = 0;
    size_t v_capacity = 0, u_capacity = 0, gpart_capacity = 0, dmus_capacity = 0, ds2s_capacity = 0;   // gradient scratch, each its own size
    double *dXc = nullptr;      // staged candidates (host path, or generated by bo_candidates_sobol)
    size_t xc_capacity = 0;
    int64_t staged_M = 0;       // > 0: dXc holds a generated grid of staged_M x staged_d points (BO_PTR_STAGED)
    int staged_d = 0;
    double *dSobol = nullptr;   // box bounds + direction numbers of the generator
    size_t sobol_capacity = 0;
    double *dVal = nullptr;     // M values of the last score (device resident)
    size_t val_capacity = 0;
    int64_t last_M = 0;
    bool last_val_valid = false;
    const double *last_val_ptr = nullptr;
    double *dGradOut = nullptr;
    size_t gradout_capacity = 0;
    double *dPredict = nullptr;  // bo_predict host-path staging (mu, s2, dmu, ds2)
    size_t predict_capacity = 0;
    double *dLoglik = nullptr;   // bo_loglik / bo_loglik_fit output staging
    size_t loglik_capacity = 0;
    double *dLLK = nullptr, *dLLSmall = nullptr;   // bo_loglik_fit: augmented Gram / factor, small inputs
    size_t llk_capacity = 0, llsmall_capacity = 0;
    int64_t *dIncumbent = nullptr;   // packed {value bits, global index} records of the last pass (cross-rank exchange)
    size_t incumbent_capacity = 0;
    double *dMerged = nullptr;       // bo_incumbent_merge result staging
    size_t merged_capacity = 0;
    bool best_valid = false;
    int64_t *rec_ptr = nullptr;      // when set, the final arg-max kernel of a scoring pass also writes the packed record
    int64_t rec_offset = 0;
    double *dBlkVal = nullptr;  // per-block argmax staging
    int64_t *dBlkIdx = nullptr;
    size_t blk_capacity = 0;

    int prec = BO_PREC_F64;
    double prec_tol = 1e-9;
    // int8-slice (Ozaki) path state
    bool oz_ready = false;
    int oz_slices = 0;
    bool oz_extra = false;          // also accumulate the digit pairs of group g = S
    int8_t *dWs = nullptr;        // S_hyper x slices x np x np  slice planes of W
    int8_t *dKss = nullptr;       // slices x chunk x np          slice planes of K*^T
    double *dRowScale = nullptr;  // S_hyper x np   2^(e_i - 12) rho
    int *dRowExp = nullptr;       // S_hyper x np (+ S_hyper maxima)
    size_t ws_capacity = 0, kss_capacity = 0, kss_stride = 0, rowscale_capacity = 0, rowexp_capacity = 0;
    std::vector<int> h_emax;
    double *dXsHalfSq = nullptr;               // S_hyper x np   |xs_j|^2 / 2
    size_t halfsq_capacity = 0;
    double *dOzQ = nullptr;                     // (np/64) x chunk partial |v|^2
    size_t ozpart_capacity = 0;
    double *dOzMu = nullptr;                    // per slice buffer: (blocks) x chunk partials of kappa . beta
    size_t ozmu_capacity = 0, ozmu_stride = 0;
    int oz_mu_slot = 0, oz_mu_rows[2] = {0, 0};
    // FP64 rescue pass of the int8-slice path (score.cu run_oz): per-candidate a-priori error bound -> flag list
    // -> compact FP64 re-score -> scatter
    int oz_cluster = 0;                 // CTAs per cluster of the scoring contraction (0: library default)
    bool oz_rescue = true;
    double oz_rescue_tol = 5e-7;        // flag when the bound exceeds tol * max(|value|, floor * max |value|): 2x inside the 1e-6 bar
    double oz_rescue_floor = 1e-12;
    double oz_demote_frac = 0.25;       // a pass that rescues more than this fraction sends later passes to FP64
    bool oz_demoted = false;
    int64_t oz_last_total = 0, oz_last_flagged = 0;      // flagged = candidates re-scored on the FP64 path
    // tiers of the last int8 pass: levels are 2 S + extra; first = candidate chunk 0, rest = the other chunks,
    // tier2 = the level the flagged candidates were re-scored at before FP64 (0: went straight to FP64)
    int oz_last_first = 0, oz_last_rest = 0, oz_last_tier2 = 0;
    int64_t oz_last_first_flagged = 0;
    bool oz_tiered = true;              // main pass one half-level below the selected one when few candidates need more
    double oz_tier_frac = 0.10;         // ... i.e. when at most this fraction of the first chunk is flagged there
    int64_t oz_tier_min = 4096;         // flagged lists shorter than this go straight to FP64
    int oz_last_path = 0;               // 1: the last scoring pass ran the int8-slice contraction
    double *dErrEst = nullptr, *dErrK = nullptr, *dRescue = nullptr;
    int *dFlagList = nullptr;
    double *dErrEst2 = nullptr, *dRescue2 = nullptr;      // second tier's own scratch (the first tier's stays live)
    int *dFlagList2 = nullptr;
    size_t errest2_capacity = 0, rescue2_capacity = 0, flaglist2_capacity = 0;
    unsigned long long *dFlagBits = nullptr;
    size_t errest_capacity = 0, errk_capacity = 0, rescue_capacity = 0, flaglist_capacity = 0, flagbits_capacity = 0;
    std::vector<double> h_errk;

    bo_thompson_state th;

    // ---- profiler -----------------------------------------------------------
    bool prof_on = false;
    std::vector<bo_prof_entry> prof;
    int64_t launches = 0;
};

int bo_set_err(bo_ctx *ctx, int code, const char *fmt, ...);

#define BO_CUDA(ctx, call)                                                          \
    do {                                                                            \
        cudaError_t e__ = (call);                                                   \
        if (e__ != cudaSuccess)                                                     \
            return bo_set_err(ctx, BO_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, \
                              #call, cudaGetErrorString(e__));                      \
    } while (0)

#define BO_TRY(expr)                \
    do {                            \
        int rc__ = (expr);          \
        if (rc__ != BO_OK) return rc__; \
    } while (0)

// Ensure a device buffer holds at least `count` elements of T.
template <typename T>
static inline int bo_reserve(bo_ctx *ctx, T **ptr, size_t *cap, size_t count) {
    if (*cap >= count && *ptr) return BO_OK;
    if (*ptr) BO_CUDA(ctx, cudaFree(*ptr));
    *ptr = nullptr;
    *cap = 0;
    BO_CUDA(ctx, cudaMalloc((void **)ptr, count * sizeof(T)));
    *cap = count;
    return BO_OK;
}

// RAII marker around one kernel launch: counts it and, when the profiler is on,
// brackets it with events on the handle's stream.
struct bo_launch_scope {
    bo_ctx *ctx;
    cudaEvent_t a = nullptr, b = nullptr;
    int slot = -1;
    cudaStream_t st;
    bo_launch_scope(bo_ctx *c, const char *name, cudaStream_t stream = nullptr) : ctx(c), st(stream ? stream : c->stream) {
        ctx->launches++;
        if (!ctx->prof_on) return;
        for (size_t i = 0; i < ctx->prof.size(); ++i)
            if (ctx->prof[i].name == name) slot = (int)i;
        if (slot < 0) {
            ctx->prof.emplace_back();
            ctx->prof.back().name = name;
            slot = (int)ctx->prof.size() - 1;
        }
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a, st);
    }
    ~bo_launch_scope() {
        if (slot < 0) return;
        cudaEventRecord(b, st);
        ctx->prof[slot].launches++;
        ctx->prof[slot].pending.emplace_back(a, b);
    }
};

#define BO_LAUNCH(ctx, name) bo_launch_scope scope__(ctx, name)
#define BO_LAUNCH_ON(ctx, name, stream) bo_launch_scope scope__(ctx, name, stream)

#define BO_CHECK_LAUNCH(ctx)                                                     \
    do {                                                                         \
        cudaError_t e__ = cudaGetLastError();                                    \
        if (e__ != cudaSuccess)                                                  \
            return bo_set_err(ctx, BO_ERR_CUDA, "%s:%d launch -> %s", __FILE__,  \
                              __LINE__, cudaGetErrorString(e__));                \
    } while (0)

// ---- stage entry points implemented across the .cu files -------------------
int bo_linalg_gram(bo_ctx *ctx, int kernel, int n, int np, int dp, int S, const double *dXs,
                   const double *rho_dev, const double *sn2_dev, double *K, int mirror, const double *aug,
                   const double *ann);
int bo_linalg_loglik_aug(bo_ctx *ctx, const double *L, int n, int np, int S, double *out);
int bo_linalg_cholesky(bo_ctx *ctx, int np, int batch, double *A, double *dinv, int *dInfo);
int bo_linalg_trtri(bo_ctx *ctx, int np, int batch, const double *L, const double *dinv,
                    double *W, double *tmp);
int bo_linalg_transpose(bo_ctx *ctx, int np, int batch, const double *A, double *AT);
int bo_linalg_finish_fit(bo_ctx *ctx);
int bo_linalg_init(bo_ctx *ctx);
void bo_linalg_drop_graphs(bo_ctx *ctx);
int bo_score_init(bo_ctx *ctx);
int bo_ozaki_init(bo_ctx *ctx);
int bo_thompson_init(bo_ctx *ctx);
int bo_thompson_build_init(bo_ctx *ctx);
int bo_ozaki_prepare(bo_ctx *ctx, int S);
int bo_ozaki_append_row(bo_ctx *ctx, int row);
int bo_ozaki_choose_slices(bo_ctx *ctx, double tol);
int bo_ozaki_error_scale(bo_ctx *ctx, int S, bool extra, int slot);
int bo_ozaki_slice(bo_ctx *ctx, int s, int S, const double *dXc, int64_t c0, int mc, int mcp, int buf,
                   cudaStream_t stream);
int bo_ozaki_contract(bo_ctx *ctx, int s, int S, bool extra, int mcp, int buf, double *mu, double *s2, int32_t *dbg);
int bo_ozaki_reserve(bo_ctx *ctx, int S, int mcp_max, int nbuf);
bool bo_thompson_ozaki_usable(bo_ctx *ctx, int64_t M);
int64_t bo_thompson_ozaki_blocks(int64_t M);
int bo_thompson_ozaki_run(bo_ctx *ctx, int64_t M, const double *dXc, double *dOut, double *blkval, int64_t *blkidx,
                          int64_t blk_ld);

// one scoring / prediction pass over M device-resident candidates
struct ScoreRequest {
    int mode = 0, acq = 0;     // mode 0: acquisition, 1: predict moments
    double param = 0.0;
    int64_t M = 0;
    const double *dXc = nullptr;
    double *dVal = nullptr, *dGrad = nullptr;
    double *dMu = nullptr, *dS2 = nullptr, *dDmu = nullptr, *dDs2 = nullptr;
    bool want_best = false;
};
int bo_score_run(bo_ctx *ctx, const ScoreRequest &rq);
int bo_topk_run(bo_ctx *ctx, const double *vals, int64_t M, int k, double *h_val, int64_t *h_idx);
int bo_thompson_run(bo_ctx *ctx, int64_t M, const double *dXc, double *dOut, double *dGrad,
                    double *dBestVal, int64_t *dBestIdx);
