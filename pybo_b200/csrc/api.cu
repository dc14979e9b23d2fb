// api.cu -- the extern "C" boundary declared in include/bo_b200.h.
#include <math.h>
#include <stdarg.h>

#include <algorithm>

#include "common.cuh"

int bo_set_err(bo_ctx *ctx, int code, const char *fmt, ...) {
    if (ctx) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

#define BO_ENTER(ctx)                                              \
    do {                                                           \
        if (!(ctx)) return BO_ERR_ARG;                             \
        BO_CUDA(ctx, cudaSetDevice((ctx)->device));                \
    } while (0)

static int padded_dim(int d) {
    int dp = 2;
    while (dp < d) dp *= 2;
    return dp;
}

// ---------------------------------------------------------------------------
// small utility kernels
// ---------------------------------------------------------------------------
__global__ void pad_identity_kernel(double *A, int n, int np) {
    // rows/cols >= n of each np x np matrix := identity
    const int64_t boff = (int64_t)blockIdx.z * np * np;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= np || j >= np) return;
    if (i >= n || j >= n) A[boff + (int64_t)i * np + j] = (i == j) ? 1.0 : 0.0;
}

__global__ void zero_upper_kernel(double *A, int n, int64_t ld, int64_t stride) {
    const int64_t boff = (int64_t)blockIdx.z * stride;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i < n && j < n && j > i) A[boff + (int64_t)i * ld + j] = 0.0;
}

// Xs[s][i][k] = X[i][k] / ell_s[k] for i < n, k < d; zero elsewhere (rows up to np, columns up to dp)
__global__ void scale_x_kernel(const double *__restrict__ X, const double *__restrict__ invell, int n, int np, int d,
                               int dp, int S, double *__restrict__ Xs) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)S * np * dp) return;
    const int k = (int)(e % dp);
    const int i = (int)((e / dp) % np);
    const int s = (int)(e / ((int64_t)dp * np));
    Xs[e] = (i < n && k < d) ? X[(int64_t)i * d + k] * invell[(int64_t)s * dp + k] : 0.0;
}

__global__ void loglik_kernel(const double *alpha, const double *logdet, int n, int np, double *out) {
    __shared__ double red[32];
    const int s = blockIdx.x;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double a = alpha[(int64_t)s * np + i];
        acc = fma(a, a, acc);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) out[s] = -0.5 * v - logdet[s] - 0.5 * n * 1.8378770664093453;   // log(2 pi)
    }
}

// ---------------------------------------------------------------------------
// lifetime
// ---------------------------------------------------------------------------
extern "C" int bo_create(int device, bo_ctx **out) {
    if (!out) return BO_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return BO_ERR_CUDA;
    bo_ctx *ctx = new bo_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&ctx->prop, device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return BO_ERR_CUDA;
    }
    {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);       // lo = least priority (largest number)
        cudaStreamDestroy(ctx->stream);
        bool ok = cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, hi) == cudaSuccess &&
                  cudaStreamCreateWithPriority(&ctx->stream2, cudaStreamNonBlocking, lo) == cudaSuccess;
        for (int i = 0; i < 2 && ok; ++i)
            ok = cudaEventCreateWithFlags(&ctx->ev_sliced[i], cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&ctx->ev_consumed[i], cudaEventDisableTiming) == cudaSuccess;
        if (!ok) {
            delete ctx;
            return BO_ERR_CUDA;
        }
    }
    ctx->sm_count = ctx->prop.multiProcessorCount;
    int rc = bo_linalg_init(ctx);
    if (rc == BO_OK) rc = bo_score_init(ctx);
    if (rc == BO_OK) rc = bo_ozaki_init(ctx);
    if (rc == BO_OK) rc = bo_thompson_init(ctx);
    if (rc == BO_OK) rc = bo_thompson_build_init(ctx);
    if (rc != BO_OK) {
        fprintf(stderr, "bo_create: %s\n", ctx->err);
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return rc;
    }
    *out = ctx;
    return BO_OK;
}

static void free_all(bo_ctx *ctx) {
    double **ptrs[] = {&ctx->dX, &ctx->dXs, &ctx->dY, &ctx->dInvEll, &ctx->dRho, &ctx->dSn2, &ctx->dBias,
                       &ctx->dL, &ctx->dW, &ctx->dWT, &ctx->dDinv, &ctx->dAlpha, &ctx->dBeta, &ctx->dLogdet,
                       &ctx->dKs, &ctx->dV, &ctx->dU, &ctx->dQpart, &ctx->dPpart, &ctx->dMuS,
                       &ctx->dS2S, &ctx->dDmuS, &ctx->dDs2S, &ctx->dGpart, &ctx->dXc, &ctx->dVal,
                       &ctx->dGradOut, &ctx->dBlkVal, &ctx->th.W, &ctx->th.b, &ctx->th.theta,
                       &ctx->th.scale, &ctx->th.bias, &ctx->th.dBestVal, &ctx->th.thetaT, &ctx->dOzQ, &ctx->dXsHalfSq, &ctx->dCholDinv, &ctx->dOzMu, &ctx->dAppend, &ctx->dSobol,
                       &ctx->dMerged, &ctx->dPredict, &ctx->dLoglik, &ctx->dErrEst, &ctx->dErrK, &ctx->dRescue, &ctx->dErrEst2, &ctx->dRescue2, &ctx->dLLK, &ctx->dLLSmall,
                       &ctx->th.build};
    for (auto p : ptrs)
        if (*p) { cudaFree(*p); *p = nullptr; }
    if (ctx->dInfo) { cudaFree(ctx->dInfo); ctx->dInfo = nullptr; }
    if (ctx->dFlagList) { cudaFree(ctx->dFlagList); ctx->dFlagList = nullptr; }
    if (ctx->dFlagList2) { cudaFree(ctx->dFlagList2); ctx->dFlagList2 = nullptr; }
    if (ctx->dIncumbent) { cudaFree(ctx->dIncumbent); ctx->dIncumbent = nullptr; }
    if (ctx->dFlagBits) { cudaFree(ctx->dFlagBits); ctx->dFlagBits = nullptr; }
    if (ctx->dAppendInfo) { cudaFree(ctx->dAppendInfo); ctx->dAppendInfo = nullptr; }
    if (ctx->dWs) { cudaFree(ctx->dWs); ctx->dWs = nullptr; }
    if (ctx->dCholInfo) { cudaFree(ctx->dCholInfo); ctx->dCholInfo = nullptr; }
    if (ctx->dCholFlags) { cudaFree(ctx->dCholFlags); ctx->dCholFlags = nullptr; }
    if (ctx->dKss) { cudaFree(ctx->dKss); ctx->dKss = nullptr; }
    if (ctx->dRowScale) { cudaFree(ctx->dRowScale); ctx->dRowScale = nullptr; }
    if (ctx->dRowExp) { cudaFree(ctx->dRowExp); ctx->dRowExp = nullptr; }
    if (ctx->dBlkIdx) { cudaFree(ctx->dBlkIdx); ctx->dBlkIdx = nullptr; }
    if (ctx->th.dBestIdx) { cudaFree(ctx->th.dBestIdx); ctx->th.dBestIdx = nullptr; }
    if (ctx->th.ozTheta) { cudaFree(ctx->th.ozTheta); ctx->th.ozTheta = nullptr; }
    if (ctx->th.ozPhi) { cudaFree(ctx->th.ozPhi); ctx->th.ozPhi = nullptr; }
    if (ctx->th.ozRowScale) { cudaFree(ctx->th.ozRowScale); ctx->th.ozRowScale = nullptr; }
    if (ctx->th.ozWp) { cudaFree(ctx->th.ozWp); ctx->th.ozWp = nullptr; }
}

static void prof_drain(bo_ctx *ctx) {
    for (auto &e : ctx->prof) {
        for (auto &pr : e.pending) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) e.total_ms += ms;
            cudaEventDestroy(pr.first);
            cudaEventDestroy(pr.second);
        }
        e.pending.clear();
    }
}

extern "C" int bo_destroy(bo_ctx *ctx) {
    if (!ctx) return BO_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->stream2) cudaStreamSynchronize(ctx->stream2);
    prof_drain(ctx);
    bo_linalg_drop_graphs(ctx);
    free_all(ctx);
    for (int i = 0; i < 2; ++i) {
        if (ctx->ev_sliced[i]) cudaEventDestroy(ctx->ev_sliced[i]);
        if (ctx->ev_consumed[i]) cudaEventDestroy(ctx->ev_consumed[i]);
    }
    for (int l = 0; l < BO_CHOL_MAX_LANES - 1; ++l) {
        for (auto &e : ctx->chol_lane_ev[l])
            if (e) cudaEventDestroy(e);
        if (ctx->chol_lane_main[l]) cudaStreamDestroy(ctx->chol_lane_main[l]);
        if (ctx->chol_lane_side[l]) cudaStreamDestroy(ctx->chol_lane_side[l]);
    }
    if (ctx->chol_lane_fork) cudaEventDestroy(ctx->chol_lane_fork);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return BO_OK;
}

extern "C" const char *bo_last_error(const bo_ctx *ctx) { return ctx ? ctx->err : "null handle"; }
extern "C" void *bo_stream(bo_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

extern "C" int bo_sync(bo_ctx *ctx) {
    BO_ENTER(ctx);
    BO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BO_OK;
}

extern "C" int bo_device_props(bo_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor, size_t *l2_bytes,
                               size_t *hbm_bytes) {
    if (!ctx) return BO_ERR_ARG;
    if (sm_count) *sm_count = ctx->prop.multiProcessorCount;
    if (cc_major) *cc_major = ctx->prop.major;
    if (cc_minor) *cc_minor = ctx->prop.minor;
    if (l2_bytes) *l2_bytes = (size_t)ctx->prop.l2CacheSize;
    if (hbm_bytes) *hbm_bytes = ctx->prop.totalGlobalMem;
    return BO_OK;
}

// ---------------------------------------------------------------------------
// fit
// ---------------------------------------------------------------------------
template <typename T>
static int realloc_dev(bo_ctx *ctx, T **p, size_t count) {
    if (*p) { BO_CUDA(ctx, cudaFree(*p)); *p = nullptr; }
    BO_CUDA(ctx, cudaMalloc((void **)p, count * sizeof(T)));
    return BO_OK;
}

extern "C" int bo_fit(bo_ctx *ctx, int kernel, int n, int d, int S, const double *X, const double *y,
                      const double *ell, const double *rho, const double *sn2, const double *bias) {
    BO_ENTER(ctx);
    if (kernel != BO_KERNEL_SE && kernel != BO_KERNEL_MATERN52)
        return bo_set_err(ctx, BO_ERR_ARG, "unknown kernel id %d", kernel);
    if (n < 1 || d < 1 || d > BO_MAX_D || S < 1 || !X || !y || !ell || !rho || !sn2 || !bias)
        return bo_set_err(ctx, BO_ERR_ARG, "bo_fit: bad shape n=%d d=%d S=%d (d <= %d)", n, d, S, BO_MAX_D);
    for (int s = 0; s < S; ++s) {
        if (!(rho[s] > 0.0) || !(sn2[s] >= 0.0)) return bo_set_err(ctx, BO_ERR_ARG, "bo_fit: rho must be > 0, sn2 >= 0");
        for (int k = 0; k < d; ++k)
            if (!(ell[s * d + k] > 0.0)) return bo_set_err(ctx, BO_ERR_ARG, "bo_fit: ell must be > 0");
    }
    ctx->fitted = false;
    ctx->oz_ready = false;
    ctx->oz_demoted = false;
    ctx->last_val_valid = false;
    const int np = bo_round_up(n, BO_PAD), dp = padded_dim(d), nblk64 = np / 64;
    const size_t mat = (size_t)S * np * np;
    if (mat > ctx->fit_capacity) {
        BO_TRY(realloc_dev(ctx, &ctx->dL, mat));
        BO_TRY(realloc_dev(ctx, &ctx->dW, mat));
        BO_TRY(realloc_dev(ctx, &ctx->dWT, mat));
        ctx->fit_capacity = mat;
    }
    // small per-fit buffers grow on demand and are reused (bo_fit runs hundreds of times in the
    // hyper-parameter sampler)
    {
        struct { double **p; size_t need; } small[] = {
            {&ctx->dX, (size_t)np * d}, {&ctx->dY, (size_t)np}, {&ctx->dXs, (size_t)S * np * dp},
            {&ctx->dInvEll, (size_t)S * dp}, {&ctx->dRho, (size_t)S}, {&ctx->dSn2, (size_t)S},
            {&ctx->dBias, (size_t)S}, {&ctx->dDinv, (size_t)S * nblk64 * 4096}, {&ctx->dAlpha, (size_t)S * np},
            {&ctx->dBeta, (size_t)S * np}, {&ctx->dLogdet, (size_t)S}};
        for (size_t i = 0; i < sizeof(small) / sizeof(small[0]); ++i) {
            if (ctx->small_capacity[i] >= small[i].need && *small[i].p) continue;
            BO_TRY(realloc_dev(ctx, small[i].p, small[i].need));
            ctx->small_capacity[i] = small[i].need;
        }
        if (ctx->info_capacity < (size_t)S || !ctx->dInfo) {
            BO_TRY(realloc_dev(ctx, &ctx->dInfo, (size_t)S));
            ctx->info_capacity = (size_t)S;
        }
    }
    ctx->kernel = kernel; ctx->n = n; ctx->np = np; ctx->d = d; ctx->dp = dp; ctx->S = S;
    ctx->h_rho.assign(rho, rho + S);
    ctx->h_sn2.assign(sn2, sn2 + S);
    ctx->h_bias.assign(bias, bias + S);
    ctx->h_ell.assign(ell, ell + (size_t)S * d);
    ctx->h_info.assign(S, 0);

    // inverse lengthscales (zero padded); the scaled, zero-padded coordinates Xs[s][i][k] = X[i][k] / ell_s[k] are
    // built on the device.  The host buffers are pageable: cudaMemcpyAsync returns once they are staged, so no
    // synchronisation is needed before they go out of scope.
    std::vector<double> inv((size_t)S * dp, 0.0);
    for (int s = 0; s < S; ++s)
        for (int k = 0; k < d; ++k) inv[(size_t)s * dp + k] = 1.0 / ell[s * d + k];
    cudaStream_t st = ctx->stream;
    BO_CUDA(ctx, cudaMemcpyAsync(ctx->dX, X, sizeof(double) * n * d, cudaMemcpyHostToDevice, st));
    BO_CUDA(ctx, cudaMemcpyAsync(ctx->dY, y, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    BO_CUDA(ctx, cudaMemcpyAsync(ctx->dInvEll, inv.data(), sizeof(double) * inv.size(), cudaMemcpyHostToDevice, st));
    BO_CUDA(ctx, cudaMemcpyAsync(ctx->dRho, rho, sizeof(double) * S, cudaMemcpyHostToDevice, st));
    BO_CUDA(ctx, cudaMemcpyAsync(ctx->dSn2, sn2, sizeof(double) * S, cudaMemcpyHostToDevice, st));
    BO_CUDA(ctx, cudaMemcpyAsync(ctx->dBias, bias, sizeof(double) * S, cudaMemcpyHostToDevice, st));
    {
        BO_LAUNCH(ctx, "scale_x_kernel");
        const int64_t tot = (int64_t)S * np * dp;
        scale_x_kernel<<<(int)((tot + 255) / 256), 256, 0, st>>>(ctx->dX, ctx->dInvEll, n, np, d, dp, S, ctx->dXs);
        BO_CHECK_LAUNCH(ctx);
    }

    BO_TRY(bo_linalg_gram(ctx, kernel, n, np, dp, S, ctx->dXs, ctx->dRho, ctx->dSn2, ctx->dL, 0, nullptr, nullptr));
    BO_TRY(bo_linalg_cholesky(ctx, np, S, ctx->dL, ctx->dDinv, ctx->dInfo));
    BO_CUDA(ctx, cudaMemcpyAsync(ctx->h_info.data(), ctx->dInfo, sizeof(int) * S, cudaMemcpyDeviceToHost, st));
    BO_CUDA(ctx, cudaStreamSynchronize(st));
    for (int s = 0; s < S; ++s)
        if (ctx->h_info[s] != 0)
            return bo_set_err(ctx, BO_ERR_NOT_PD, "Cholesky failed: hyper-sample %d, leading minor %d is not positive definite",
                              s, ctx->h_info[s]);
    BO_TRY(bo_linalg_trtri(ctx, np, S, ctx->dL, ctx->dDinv, ctx->dW, ctx->dWT /* scratch */));
    BO_TRY(bo_linalg_transpose(ctx, np, S, ctx->dW, ctx->dWT));
    BO_TRY(bo_linalg_finish_fit(ctx));          // (stream ordered: the first call that reads results synchronises)

    int64_t chunk = ((int64_t)1 << 25) / np / 128 * 128;
    ctx->chunk = std::min<int64_t>(16384, std::max<int64_t>(1024, chunk));
    ctx->fitted = true;
    return BO_OK;
}

extern "C" int bo_fit_shape(bo_ctx *ctx, int *kernel, int *n, int *d, int *S) {
    if (!ctx) return BO_ERR_ARG;
    if (!ctx->fitted) return bo_set_err(ctx, BO_ERR_STATE, "not fitted");
    if (kernel) *kernel = ctx->kernel;
    if (n) *n = ctx->n;
    if (d) *d = ctx->d;
    if (S) *S = ctx->S;
    return BO_OK;
}

extern "C" int bo_fit_info(bo_ctx *ctx, int *info) {
    if (!ctx || !info) return BO_ERR_ARG;
    for (size_t s = 0; s < ctx->h_info.size(); ++s) info[s] = ctx->h_info[s];
    return BO_OK;
}

extern "C" int bo_loglik(bo_ctx *ctx, double *out) {
    BO_ENTER(ctx);
    if (!ctx->fitted) return bo_set_err(ctx, BO_ERR_STATE, "bo_loglik before bo_fit");
    BO_TRY(bo_reserve(ctx, &ctx->dLoglik, &ctx->loglik_capacity, (size_t)ctx->S));
    {
        BO_LAUNCH(ctx, "loglik_kernel");
        loglik_kernel<<<ctx->S, 256, 0, ctx->stream>>>(ctx->dAlpha, ctx->dLogdet, ctx->n, ctx->np, ctx->dLoglik);
        BO_CHECK_LAUNCH(ctx);
    }
    BO_CUDA(ctx, cudaMemcpyAsync(out, ctx->dLoglik, sizeof(double) * ctx->S, cudaMemcpyDeviceToHost, ctx->stream));
    BO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BO_OK;
}

// Log marginal likelihood of S hyper-samples WITHOUT building a scoring state: Gram + Cholesky only.  The
// residual r_s = y - bias_s rides through the factorisation as one extra row of the padded matrix,
//     [[K_s, .], [r_s^T, c]] = [[L, 0], [alpha^T, .]] [[L, 0], [alpha^T, .]]^T,   alpha = L^-1 r_s,
// (c = 1 + 2 |r|^2 / sn2 >= 1 + |alpha|^2 keeps the last pivot positive), so alpha needs no triangular solve,
// no W = L^-1, no transpose.  One likelihood evaluation of the hyper-parameter sampler (bayesopt.py:108-115).
extern "C" int bo_loglik_fit(bo_ctx *ctx, int kernel, int n, int d, int S, const double *X, const double *y,
                             const double *ell, const double *rho, const double *sn2, const double *bias, double *out) {
    BO_ENTER(ctx);
    if (kernel != BO_KERNEL_SE && kernel != BO_KERNEL_MATERN52)
        return bo_set_err(ctx, BO_ERR_ARG, "unknown kernel id %d", kernel);
    if (n < 1 || d < 1 || d > BO_MAX_D || S < 1 || !X || !y || !ell || !rho || !sn2 || !bias || !out)
        return bo_set_err(ctx, BO_ERR_ARG, "bo_loglik_fit: bad shape n=%d d=%d S=%d (d <= %d)", n, d, S, BO_MAX_D);
    for (int s = 0; s < S; ++s) {
        if (!(rho[s] > 0.0) || !(sn2[s] >= 0.0)) return bo_set_err(ctx, BO_ERR_ARG, "bo_loglik_fit: rho must be > 0, sn2 >= 0");
        for (int k = 0; k < d; ++k)
            if (!(ell[s * d + k] > 0.0)) return bo_set_err(ctx, BO_ERR_ARG, "bo_loglik_fit: ell must be > 0");
    }
    const int np = bo_round_up(n + 1, 64), dp = padded_dim(d), nblk = np / 64;
    // small inputs in one host block -> one copy: [X n*d | inv S*dp | rho S | sn2 S | ann S | aug S*np]
    const size_t oX = 0, oInv = oX + (size_t)n * d, oRho = oInv + (size_t)S * dp, oSn = oRho + S, oAnn = oSn + S,
                 oAug = oAnn + S, oXs = oAug + (size_t)S * np, total = oXs + (size_t)S * np * dp;
    std::vector<double> h(oXs, 0.0);
    std::copy(X, X + (size_t)n * d, h.begin() + oX);
    for (int s = 0; s < S; ++s) {
        for (int k = 0; k < d; ++k) h[oInv + (size_t)s * dp + k] = 1.0 / ell[s * d + k];
        h[oRho + s] = rho[s];
        h[oSn + s] = sn2[s];
        double rr = 0.0;
        for (int i = 0; i < n; ++i) {
            const double r = y[i] - bias[s];
            h[oAug + (size_t)s * np + i] = r;
            rr += r * r;
        }
        const double c = 1.0 + 2.0 * rr / (sn2[s] > 1e-300 ? sn2[s] : 1e-300);
        h[oAnn + s] = c < 1e300 ? c : 1e300;
    }
    BO_TRY(bo_reserve(ctx, &ctx->dLLSmall, &ctx->llsmall_capacity, total));
    BO_TRY(bo_reserve(ctx, &ctx->dLLK, &ctx->llk_capacity, (size_t)S * np * np));
    BO_TRY(bo_reserve(ctx, &ctx->dCholDinv, &ctx->choldinv_capacity, (size_t)S * nblk * 4096));
    BO_TRY(bo_reserve(ctx, &ctx->dCholInfo, &ctx->cholinfo_capacity, (size_t)S));
    BO_TRY(bo_reserve(ctx, &ctx->dLoglik, &ctx->loglik_capacity, (size_t)S));
    cudaStream_t st = ctx->stream;
    double *ds = ctx->dLLSmall;
    BO_CUDA(ctx, cudaMemcpyAsync(ds, h.data(), sizeof(double) * oXs, cudaMemcpyHostToDevice, st));
    {
        BO_LAUNCH(ctx, "scale_x_kernel");
        const int64_t tot = (int64_t)S * np * dp;
        scale_x_kernel<<<(int)((tot + 255) / 256), 256, 0, st>>>(ds + oX, ds + oInv, n, np, d, dp, S, ds + oXs);
        BO_CHECK_LAUNCH(ctx);
    }
    BO_TRY(bo_linalg_gram(ctx, kernel, n, np, dp, S, ds + oXs, ds + oRho, ds + oSn, ctx->dLLK, 0, ds + oAug, ds + oAnn));
    BO_TRY(bo_linalg_cholesky(ctx, np, S, ctx->dLLK, ctx->dCholDinv, ctx->dCholInfo));
    BO_TRY(bo_linalg_loglik_aug(ctx, ctx->dLLK, n, np, S, ctx->dLoglik));
    std::vector<int> info(S, 0);
    BO_CUDA(ctx, cudaMemcpyAsync(info.data(), ctx->dCholInfo, sizeof(int) * S, cudaMemcpyDeviceToHost, st));
    BO_CUDA(ctx, cudaMemcpyAsync(out, ctx->dLoglik, sizeof(double) * S, cudaMemcpyDeviceToHost, st));
    BO_CUDA(ctx, cudaStreamSynchronize(st));
    for (int s = 0; s < S; ++s)
        if (info[s] != 0 && info[s] <= n)
            return bo_set_err(ctx, BO_ERR_NOT_PD, "bo_loglik_fit: hyper-sample %d, leading minor %d is not positive definite", s, info[s]);
    return BO_OK;
}

extern "C" int bo_get_factor(bo_ctx *ctx, int s, int which, double *out) {
    BO_ENTER(ctx);
    if (!ctx->fitted) return bo_set_err(ctx, BO_ERR_STATE, "bo_get_factor before bo_fit");
    if (s < 0 || s >= ctx->S || !out) return bo_set_err(ctx, BO_ERR_ARG, "bo_get_factor: bad sample index");
    const int n = ctx->n, np = ctx->np;
    if (which == 0 || which == 1) {
        const double *src = (which == 0 ? ctx->dL : ctx->dW) + (size_t)s * np * np;
        BO_CUDA(ctx, cudaMemcpy2DAsync(out, sizeof(double) * n, src, sizeof(double) * np, sizeof(double) * n, n,
                                       cudaMemcpyDeviceToHost, ctx->stream));
        BO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < n; ++i)
            for (int j = i + 1; j < n; ++j) out[(size_t)i * n + j] = 0.0;
    } else if (which == 2 || which == 3) {
        const double *src = (which == 2 ? ctx->dAlpha : ctx->dBeta) + (size_t)s * np;
        BO_CUDA(ctx, cudaMemcpyAsync(out, src, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
        BO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    } else {
        return bo_set_err(ctx, BO_ERR_ARG, "bo_get_factor: which must be 0..3");
    }
    return BO_OK;
}

// ---------------------------------------------------------------------------
// scoring / prediction
// ---------------------------------------------------------------------------
// ---------------------------------------------------------------------------
// device-side candidate grid: points [start, start + M) of the unscrambled Sobol sequence,
// x_i = XOR over the set bits b of gray(i) = i ^ (i >> 1) of the direction numbers sv[k][b], scaled to the
// box.  One thread per coordinate (coalesced row-major M x d output).
// ---------------------------------------------------------------------------
__global__ void sobol_kernel(int d, int bits, const uint32_t *__restrict__ sv, const double *__restrict__ lohi,
                             int64_t start, int64_t M, double *__restrict__ out) {
    __shared__ uint32_t ssv[BO_MAX_D * 32];
    __shared__ double slo[BO_MAX_D], sw[BO_MAX_D];
    for (int e = threadIdx.x; e < d * bits; e += blockDim.x) ssv[e] = sv[e];
    for (int k = threadIdx.x; k < d; k += blockDim.x) {
        slo[k] = lohi[k];
        sw[k] = lohi[d + k] - lohi[k];
    }
    __syncthreads();
    const double unit = 1.0 / (double)(1ull << bits);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < M * d; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / d;
        const int k = (int)(e - i * d);
        const uint64_t idx = (uint64_t)(start + i);
        uint32_t g = (uint32_t)(idx ^ (idx >> 1)), x = 0;
        while (g) {
            const int b = __ffs(g) - 1;
            x ^= ssv[k * bits + b];
            g &= g - 1;
        }
        out[e] = fma((double)x * unit, sw[k], slo[k]);
    }
}

extern "C" int bo_candidates_sobol(bo_ctx *ctx, int d, int bits, const uint32_t *sv, const double *lo, const double *hi,
                                   int64_t start, int64_t M, double *out, int flags) {
    BO_ENTER(ctx);
    if (d < 1 || d > BO_MAX_D || bits < 1 || bits > 32 || !sv || M <= 0 || start < 0 ||
        (uint64_t)(start + M) > (1ull << bits))
        return bo_set_err(ctx, BO_ERR_ARG, "bo_candidates_sobol: bad arguments (d=%d bits=%d start=%lld M=%lld)", d, bits,
                          (long long)start, (long long)M);
    const bool dev = (flags & BO_PTR_DEVICE) && out;
    double *dst = dev ? out : nullptr;
    if (!dst) {
        BO_TRY(bo_reserve(ctx, &ctx->dXc, &ctx->xc_capacity, (size_t)M * d));
        dst = ctx->dXc;
    }
    BO_TRY(bo_reserve(ctx, &ctx->dSobol, &ctx->sobol_capacity, (size_t)BO_MAX_D * 32 / 2 + 2 * BO_MAX_D));
    std::vector<double> lohi(2 * d);
    for (int k = 0; k < d; ++k) {
        lohi[k] = lo ? lo[k] : 0.0;
        lohi[d + k] = hi ? hi[k] : 1.0;
    }
    double *dLoHi = ctx->dSobol;
    uint32_t *dSv = reinterpret_cast<uint32_t *>(ctx->dSobol + 2 * BO_MAX_D);
    cudaStream_t st = ctx->stream;
    BO_CUDA(ctx, cudaMemcpyAsync(dLoHi, lohi.data(), sizeof(double) * 2 * d, cudaMemcpyHostToDevice, st));
    BO_CUDA(ctx, cudaMemcpyAsync(dSv, sv, sizeof(uint32_t) * d * bits, cudaMemcpyHostToDevice, st));
    BO_CUDA(ctx, cudaStreamSynchronize(st));        // `lohi` is a host temporary
    {
        BO_LAUNCH(ctx, "sobol_kernel");
        const int64_t want = (M * d + 255) / 256;
        const int grid = (int)(want < (int64_t)ctx->sm_count * 16 ? want : (int64_t)ctx->sm_count * 16);
        sobol_kernel<<<grid, 256, 0, st>>>(d, bits, dSv, dLoHi, start, M, dst);
        BO_CHECK_LAUNCH(ctx);
    }
    ctx->staged_M = dev ? 0 : M;
    ctx->staged_d = dev ? 0 : d;
    if (out && !dev) {
        BO_CUDA(ctx, cudaMemcpyAsync(out, dst, sizeof(double) * M * d, cudaMemcpyDeviceToHost, st));
        BO_CUDA(ctx, cudaStreamSynchronize(st));
    }
    return BO_OK;
}

static int stage_candidates(bo_ctx *ctx, int64_t M, const double *Xc, int flags, const double **dXc) {
    if (flags & BO_PTR_STAGED) {       // grid generated into the handle by bo_candidates_sobol
        if (ctx->staged_M != M || ctx->staged_d != ctx->d || !ctx->dXc)
            return bo_set_err(ctx, BO_ERR_STATE, "BO_PTR_STAGED: the handle holds %lld staged candidates of dimension %d, asked for %lld x %d",
                              (long long)ctx->staged_M, ctx->staged_d, (long long)M, ctx->d);
        *dXc = ctx->dXc;
        return BO_OK;
    }
    ctx->staged_M = 0;
    if (flags & BO_PTR_DEVICE) {
        *dXc = Xc;
        return BO_OK;
    }
    BO_TRY(bo_reserve(ctx, &ctx->dXc, &ctx->xc_capacity, (size_t)M * ctx->d));
    BO_CUDA(ctx, cudaMemcpyAsync(ctx->dXc, Xc, sizeof(double) * M * ctx->d, cudaMemcpyHostToDevice, ctx->stream));
    *dXc = ctx->dXc;
    return BO_OK;
}

// what == 0: plain bo_score; what == 1: bo_score_incumbent (arg max stays on the device, no read-back)
static int score_impl(bo_ctx *ctx, int acq, double param, int64_t M, const double *Xc, int flags, double *out_val,
                      double *out_grad, double *best_val, int64_t *best_idx, bool device_best) {
    if (!ctx->fitted) return bo_set_err(ctx, BO_ERR_STATE, "bo_score before bo_fit");
    if (acq < BO_ACQ_MEAN || acq > BO_ACQ_UCB) return bo_set_err(ctx, BO_ERR_ARG, "unknown acquisition id %d", acq);
    if (M <= 0 || (!Xc && !(flags & BO_PTR_STAGED))) return bo_set_err(ctx, BO_ERR_ARG, "bo_score: need M > 0 candidates");
    const bool dev = flags & BO_PTR_DEVICE;
    ScoreRequest rq;
    rq.mode = 0; rq.acq = acq; rq.param = param; rq.M = M;
    BO_TRY(stage_candidates(ctx, M, Xc, flags, &rq.dXc));
    if (dev && out_val) {
        rq.dVal = out_val;
    } else {
        BO_TRY(bo_reserve(ctx, &ctx->dVal, &ctx->val_capacity, (size_t)M));
        rq.dVal = ctx->dVal;
    }
    if (out_grad) {
        if (dev) rq.dGrad = out_grad;
        else {
            BO_TRY(bo_reserve(ctx, &ctx->dGradOut, &ctx->gradout_capacity, (size_t)M * ctx->d));
            rq.dGrad = ctx->dGradOut;
        }
    }
    const bool fetch_best = (best_val != nullptr) || (best_idx != nullptr);
    rq.want_best = fetch_best || device_best;
    ctx->best_valid = false;
    BO_TRY(bo_score_run(ctx, rq));
    ctx->last_val_ptr = rq.dVal;
    ctx->last_M = M;
    ctx->last_val_valid = true;
    ctx->best_valid = rq.want_best;
    cudaStream_t st = ctx->stream;
    if (!dev && out_val)
        BO_CUDA(ctx, cudaMemcpyAsync(out_val, rq.dVal, sizeof(double) * M, cudaMemcpyDeviceToHost, st));
    if (!dev && out_grad)
        BO_CUDA(ctx, cudaMemcpyAsync(out_grad, rq.dGrad, sizeof(double) * M * ctx->d, cudaMemcpyDeviceToHost, st));
    double bv = 0.0;
    int64_t bi = -1;
    if (fetch_best) {
        BO_CUDA(ctx, cudaMemcpyAsync(&bv, ctx->dBlkVal + ctx->blk_capacity - 1, sizeof(double), cudaMemcpyDeviceToHost, st));
        BO_CUDA(ctx, cudaMemcpyAsync(&bi, ctx->dBlkIdx + ctx->blk_capacity - 1, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    }
    // host candidates / outputs are only borrowed for the duration of the call
    if ((!dev && !(flags & BO_PTR_STAGED)) || (!dev && (out_val || out_grad)) || fetch_best) BO_CUDA(ctx, cudaStreamSynchronize(st));
    if (best_val) *best_val = bv;
    if (best_idx) *best_idx = (bi == INT64_MAX) ? 0 : bi;
    return BO_OK;
}

extern "C" int bo_score(bo_ctx *ctx, int acq, double param, int64_t M, const double *Xc, int flags,
                        double *out_val, double *out_grad, double *best_val, int64_t *best_idx) {
    BO_ENTER(ctx);
    return score_impl(ctx, acq, param, M, Xc, flags, out_val, out_grad, best_val, best_idx, false);
}

// ---------------------------------------------------------------------------
// incumbents that stay on the device: packed 16-byte records {value bits, global index} for the
// cross-rank exchange (the collective follows the scoring kernels on the same stream)
// ---------------------------------------------------------------------------
__global__ void incumbent_pack_kernel(const double *__restrict__ val, const int64_t *__restrict__ idx, int k,
                                      int64_t offset, int64_t *__restrict__ rec) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    const double v = val[i];
    const int64_t ix = idx[i];
    const bool none = (ix == INT64_MAX) || (v != v);
    rec[2 * i] = __double_as_longlong(none ? -INFINITY : v);
    rec[2 * i + 1] = none ? INT64_MAX : ix + offset;
}

// recs: [count][k] records; out[k]: max value, lowest index among equal values (NaN never wins)
__global__ void incumbent_merge_kernel(const int64_t *__restrict__ recs, int count, int k, double *__restrict__ oval,
                                       int64_t *__restrict__ oidx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    double bv = -INFINITY;
    int64_t bi = INT64_MAX;
    for (int r = 0; r < count; ++r) {
        const double v = __longlong_as_double(recs[((int64_t)r * k + i) * 2]);
        const int64_t ix = recs[((int64_t)r * k + i) * 2 + 1];
        if (v != v) continue;
        if (v > bv || (v == bv && ix < bi)) { bv = v; bi = ix; }
    }
    oval[i] = bv;
    oidx[i] = bi;
}

extern "C" int bo_score_incumbent(bo_ctx *ctx, int acq, double param, int64_t M, const double *Xc, int flags,
                                  double *out_val, int64_t index_offset, void **record) {
    BO_ENTER(ctx);
    if (!record) return BO_ERR_ARG;
    // the record is written by the pass's own final arg-max kernel (fused: no extra launch between the scoring
    // kernels and the collective that follows on the stream)
    BO_TRY(bo_reserve(ctx, &ctx->dIncumbent, &ctx->incumbent_capacity, (size_t)2));
    ctx->rec_ptr = ctx->dIncumbent;
    ctx->rec_offset = index_offset;
    const int rc = score_impl(ctx, acq, param, M, Xc, flags, out_val, nullptr, nullptr, nullptr, true);
    ctx->rec_ptr = nullptr;
    ctx->rec_offset = 0;
    if (rc != BO_OK) return rc;
    *record = ctx->dIncumbent;
    return BO_OK;
}

extern "C" int bo_thompson_incumbents(bo_ctx *ctx, int64_t M, const double *Xc, int flags, int64_t index_offset,
                                      void **records, int *ndraw) {
    BO_ENTER(ctx);
    bo_thompson_state &th = ctx->th;
    if (!records) return BO_ERR_ARG;
    if (th.ndraw == 0) return bo_set_err(ctx, BO_ERR_STATE, "bo_thompson_incumbents before bo_thompson_set / bo_thompson_build");
    if (M <= 0 || !Xc) return bo_set_err(ctx, BO_ERR_ARG, "bo_thompson_incumbents: need M > 0 points");
    const double *dXc = Xc;
    if (!(flags & BO_PTR_DEVICE)) {
        BO_TRY(bo_reserve(ctx, &ctx->dXc, &ctx->xc_capacity, (size_t)M * th.d));
        BO_CUDA(ctx, cudaMemcpyAsync(ctx->dXc, Xc, sizeof(double) * M * th.d, cudaMemcpyHostToDevice, ctx->stream));
        dXc = ctx->dXc;
        ctx->staged_M = 0;
    }
    BO_TRY(bo_thompson_run(ctx, M, dXc, nullptr, nullptr, th.dBestVal, th.dBestIdx));
    ctx->last_val_valid = false;
    BO_TRY(bo_reserve(ctx, &ctx->dIncumbent, &ctx->incumbent_capacity, (size_t)2 * th.ndraw));
    {
        BO_LAUNCH(ctx, "incumbent_pack_kernel");
        incumbent_pack_kernel<<<(th.ndraw + 127) / 128, 128, 0, ctx->stream>>>(th.dBestVal, th.dBestIdx, th.ndraw, index_offset,
                                                                              ctx->dIncumbent);
        BO_CHECK_LAUNCH(ctx);
    }
    *records = ctx->dIncumbent;
    if (ndraw) *ndraw = th.ndraw;
    return BO_OK;
}

extern "C" int bo_incumbent_merge(bo_ctx *ctx, const void *records, int count, int k, double *val, int64_t *idx) {
    BO_ENTER(ctx);
    if (!records || count < 1 || k < 1 || !val || !idx) return bo_set_err(ctx, BO_ERR_ARG, "bo_incumbent_merge: bad arguments");
    BO_TRY(bo_reserve(ctx, &ctx->dMerged, &ctx->merged_capacity, (size_t)2 * k));
    double *ov = ctx->dMerged;
    int64_t *oi = reinterpret_cast<int64_t *>(ctx->dMerged + k);
    {
        BO_LAUNCH(ctx, "incumbent_merge_kernel");
        incumbent_merge_kernel<<<(k + 127) / 128, 128, 0, ctx->stream>>>(static_cast<const int64_t *>(records), count, k, ov, oi);
        BO_CHECK_LAUNCH(ctx);
    }
    BO_CUDA(ctx, cudaMemcpyAsync(val, ov, sizeof(double) * k, cudaMemcpyDeviceToHost, ctx->stream));
    BO_CUDA(ctx, cudaMemcpyAsync(idx, oi, sizeof(int64_t) * k, cudaMemcpyDeviceToHost, ctx->stream));
    BO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BO_OK;
}

extern "C" int bo_predict(bo_ctx *ctx, int64_t M, const double *Xc, int flags, double *mu, double *s2,
                          double *dmu, double *ds2) {
    BO_ENTER(ctx);
    if (!ctx->fitted) return bo_set_err(ctx, BO_ERR_STATE, "bo_predict before bo_fit");
    if (M <= 0 || (!Xc && !(flags & BO_PTR_STAGED))) return bo_set_err(ctx, BO_ERR_ARG, "bo_predict: need M > 0 points");
    const bool dev = flags & BO_PTR_DEVICE;
    const int d = ctx->d;
    ScoreRequest rq;
    rq.mode = 1; rq.M = M;
    BO_TRY(stage_candidates(ctx, M, Xc, flags, &rq.dXc));
    if (dev) {
        rq.dMu = mu; rq.dS2 = s2; rq.dDmu = dmu; rq.dDs2 = ds2;
    } else {
        // staging lives in the handle (no malloc / free, hence no device-wide sync, per L-BFGS callback)
        const bool g = dmu || ds2;
        BO_TRY(bo_reserve(ctx, &ctx->dPredict, &ctx->predict_capacity, (size_t)M * (2 + (g ? 2 * (size_t)d : 0))));
        double *scratch = ctx->dPredict;
        rq.dMu = scratch;
        rq.dS2 = scratch + M;
        if (g) {
            rq.dDmu = scratch + 2 * M;
            rq.dDs2 = scratch + 2 * M + M * d;
        }
    }
    int rc = bo_score_run(ctx, rq);
    cudaError_t e = cudaSuccess;
    if (rc == BO_OK && !dev) {
        cudaStream_t st = ctx->stream;
        if (mu) e = cudaMemcpyAsync(mu, rq.dMu, sizeof(double) * M, cudaMemcpyDeviceToHost, st);
        if (s2 && e == cudaSuccess) e = cudaMemcpyAsync(s2, rq.dS2, sizeof(double) * M, cudaMemcpyDeviceToHost, st);
        if (dmu && e == cudaSuccess) e = cudaMemcpyAsync(dmu, rq.dDmu, sizeof(double) * M * d, cudaMemcpyDeviceToHost, st);
        if (ds2 && e == cudaSuccess) e = cudaMemcpyAsync(ds2, rq.dDs2, sizeof(double) * M * d, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    if (rc != BO_OK) return rc;
    if (e != cudaSuccess) return bo_set_err(ctx, BO_ERR_CUDA, "bo_predict: %s", cudaGetErrorString(e));
    return BO_OK;
}

extern "C" int bo_topk(bo_ctx *ctx, int k, int64_t *idx, double *val) {
    BO_ENTER(ctx);
    if (!ctx->last_val_valid) return bo_set_err(ctx, BO_ERR_STATE, "bo_topk before bo_score");
    if (k < 1 || k > 4096 || !idx) return bo_set_err(ctx, BO_ERR_ARG, "bo_topk: need 1 <= k <= 4096");
    if (k > ctx->last_M) k = (int)ctx->last_M;
    std::vector<double> hv(k);
    BO_TRY(bo_topk_run(ctx, ctx->last_val_ptr, ctx->last_M, k, hv.data(), idx));
    // fewer than k finite (non-NaN) values: the tail is marked idx = -1, val = NaN (callers truncate there)
    for (int j = 0; j < k; ++j)
        if (idx[j] == INT64_MAX || idx[j] < 0) { idx[j] = -1; hv[j] = NAN; }
    if (val) std::copy(hv.begin(), hv.end(), val);
    return BO_OK;
}

extern "C" int bo_set_precision(bo_ctx *ctx, int prec, double tol) {
    if (!ctx) return BO_ERR_ARG;
    if (prec != BO_PREC_F64 && prec != BO_PREC_OZAKI) return bo_set_err(ctx, BO_ERR_ARG, "unknown precision path %d", prec);
    if (prec == BO_PREC_OZAKI && !(tol > 0.0)) return bo_set_err(ctx, BO_ERR_ARG, "bo_set_precision: tol must be > 0");
    if (ctx->prec != prec || ctx->prec_tol != tol) ctx->oz_demoted = false;
    ctx->prec = prec;
    ctx->prec_tol = tol;         // (the W slice planes stay valid: bo_ozaki_prepare re-slices only when the level changes)
    return BO_OK;
}

extern "C" int bo_set_rescue(bo_ctx *ctx, int on, double tol, double floor_rel) {
    if (!ctx) return BO_ERR_ARG;
    if (on && (!(tol > 0.0) || !(floor_rel >= 0.0))) return bo_set_err(ctx, BO_ERR_ARG, "bo_set_rescue: tol must be > 0, floor >= 0");
    ctx->oz_rescue = on != 0;
    if (on) { ctx->oz_rescue_tol = tol; ctx->oz_rescue_floor = floor_rel; }
    ctx->oz_demoted = false;
    return BO_OK;
}

extern "C" int bo_set_option(bo_ctx *ctx, const char *key, double value) {
    if (!ctx || !key) return BO_ERR_ARG;
    if (strcmp(key, "oz_cluster") == 0) {
        const int v = (int)value;
        if (v != 0 && v != 1 && v != 2 && v != 4) return bo_set_err(ctx, BO_ERR_ARG, "oz_cluster must be 0 (default), 1, 2 or 4");
        ctx->oz_cluster = v;
        return BO_OK;
    }
    if (strcmp(key, "oz_tiered") == 0) {
        ctx->oz_tiered = value != 0.0;
        return BO_OK;
    }
    if (strcmp(key, "oz_tier_frac") == 0) {
        if (!(value >= 0.0 && value <= 1.0)) return bo_set_err(ctx, BO_ERR_ARG, "oz_tier_frac must lie in [0, 1]");
        ctx->oz_tier_frac = value;
        return BO_OK;
    }
    if (strcmp(key, "oz_tier_min") == 0) {
        if (!(value >= 1.0)) return bo_set_err(ctx, BO_ERR_ARG, "oz_tier_min must be >= 1");
        ctx->oz_tier_min = (int64_t)value;
        return BO_OK;
    }
    return bo_set_err(ctx, BO_ERR_ARG, "bo_set_option: unknown key '%s'", key);
}

extern "C" int bo_ozaki_error_bound(bo_ctx *ctx, double *errk) {
    BO_ENTER(ctx);
    if (!errk) return BO_ERR_ARG;
    if (!ctx->fitted || ctx->prec != BO_PREC_OZAKI) return bo_set_err(ctx, BO_ERR_STATE, "bo_ozaki_error_bound: fit and select BO_PREC_OZAKI first");
    const int S = bo_ozaki_choose_slices(ctx, ctx->prec_tol);
    if (S < 1) return BO_ERR_CUDA;
    BO_TRY(bo_ozaki_prepare(ctx, S));
    BO_TRY(bo_ozaki_error_scale(ctx, S, ctx->oz_extra, 0));
    for (int s = 0; s < ctx->S; ++s) errk[s] = ctx->h_errk[s];
    return BO_OK;
}

extern "C" int bo_rescue_info(bo_ctx *ctx, int *int8_path, int64_t *flagged, int64_t *total) {
    if (!ctx) return BO_ERR_ARG;
    if (int8_path) *int8_path = ctx->oz_last_path;
    if (flagged) *flagged = ctx->oz_last_flagged;
    if (total) *total = ctx->oz_last_total;
    return BO_OK;
}

extern "C" int bo_precision_info(bo_ctx *ctx, int *prec, int *slices) {
    if (!ctx) return BO_ERR_ARG;
    if (prec) *prec = ctx->prec;
    // (the level most candidates of the last pass ran at; before the first pass, the level the tolerance selects)
    if (slices) *slices = (ctx->prec == BO_PREC_OZAKI) ? (ctx->oz_last_rest ? ctx->oz_last_rest : ctx->oz_slices * 2 + (ctx->oz_extra ? 1 : 0)) : 0;
    return BO_OK;
}

extern "C" int bo_tier_info(bo_ctx *ctx, int *level_first, int *level_rest, int *level_tier2, int64_t *first_flagged,
                            int64_t *fp64_rescored) {
    if (!ctx) return BO_ERR_ARG;
    if (level_first) *level_first = ctx->oz_last_first;
    if (level_rest) *level_rest = ctx->oz_last_rest;
    if (level_tier2) *level_tier2 = ctx->oz_last_tier2;
    if (first_flagged) *first_flagged = ctx->oz_last_first_flagged;
    if (fp64_rescored) *fp64_rescored = ctx->oz_last_flagged;
    return BO_OK;
}

// ---------------------------------------------------------------------------
// Thompson
// ---------------------------------------------------------------------------
extern "C" int bo_thompson_set(bo_ctx *ctx, int ndraw, int nW, int m, int d, const double *W, const double *b,
                               const double *theta, const double *scale, const double *bias) {
    BO_ENTER(ctx);
    if (ndraw < 1 || m < 1 || d < 1 || d > BO_MAX_D || (nW != 1 && nW != ndraw) || !W || !b || !theta || !scale || !bias)
        return bo_set_err(ctx, BO_ERR_ARG, "bo_thompson_set: bad shape ndraw=%d nW=%d m=%d d=%d", ndraw, nW, m, d);
    bo_thompson_state &th = ctx->th;
    BO_TRY(realloc_dev(ctx, &th.W, (size_t)nW * m * d));
    BO_TRY(realloc_dev(ctx, &th.b, (size_t)nW * m));
    BO_TRY(realloc_dev(ctx, &th.theta, (size_t)ndraw * m));
    BO_TRY(realloc_dev(ctx, &th.scale, (size_t)ndraw));
    BO_TRY(realloc_dev(ctx, &th.bias, (size_t)ndraw));
    BO_TRY(realloc_dev(ctx, &th.dBestVal, (size_t)ndraw));
    BO_TRY(realloc_dev(ctx, &th.dBestIdx, (size_t)ndraw));
    cudaStream_t st = ctx->stream;
    BO_CUDA(ctx, cudaMemcpyAsync(th.W, W, sizeof(double) * nW * m * d, cudaMemcpyHostToDevice, st));
    BO_CUDA(ctx, cudaMemcpyAsync(th.b, b, sizeof(double) * nW * m, cudaMemcpyHostToDevice, st));
    BO_CUDA(ctx, cudaMemcpyAsync(th.theta, theta, sizeof(double) * ndraw * m, cudaMemcpyHostToDevice, st));
    BO_CUDA(ctx, cudaMemcpyAsync(th.scale, scale, sizeof(double) * ndraw, cudaMemcpyHostToDevice, st));
    BO_CUDA(ctx, cudaMemcpyAsync(th.bias, bias, sizeof(double) * ndraw, cudaMemcpyHostToDevice, st));
    BO_CUDA(ctx, cudaStreamSynchronize(st));
    th.ndraw = ndraw; th.nW = nW; th.m = m; th.d = d;
    th.oz_ready = false;
    th.h_theta.clear();
    if (nW == 1) {      // host copies for the int8-slice path (Theta is sliced on the host)
        th.h_theta.assign(theta, theta + (size_t)ndraw * m);
        th.h_scale.assign(scale, scale + ndraw);
        th.h_bias.assign(bias, bias + ndraw);
        th.h_W.assign(W, W + (size_t)m * d);
        th.h_b.assign(b, b + m);
    }
    // transposed, zero-padded theta for the shared-basis tensor-core path
    if (th.thetaT) { cudaFree(th.thetaT); th.thetaT = nullptr; }
    if (nW == 1) {
        const int mp = bo_round_up(m, 16), ndp = bo_round_up(ndraw, 256);
        std::vector<double> tt((size_t)mp * ndp, 0.0);
        for (int r = 0; r < ndraw; ++r)
            for (int j = 0; j < m; ++j) tt[(size_t)j * ndp + r] = theta[(size_t)r * m + j];
        BO_CUDA(ctx, cudaMalloc(&th.thetaT, sizeof(double) * tt.size()));
        BO_CUDA(ctx, cudaMemcpy(th.thetaT, tt.data(), sizeof(double) * tt.size(), cudaMemcpyHostToDevice));
        th.ndp = ndp;
    }
    return BO_OK;
}

extern "C" int bo_thompson_eval(bo_ctx *ctx, int64_t M, const double *Xc, int flags, double *out,
                                double *out_grad, double *best_val, int64_t *best_idx) {
    BO_ENTER(ctx);
    bo_thompson_state &th = ctx->th;
    if (th.ndraw == 0) return bo_set_err(ctx, BO_ERR_STATE, "bo_thompson_eval before bo_thompson_set");
    if (M <= 0 || !Xc) return bo_set_err(ctx, BO_ERR_ARG, "bo_thompson_eval: need M > 0 points");
    const bool dev = flags & BO_PTR_DEVICE;
    const double *dXc = Xc;
    if (!dev) {
        BO_TRY(bo_reserve(ctx, &ctx->dXc, &ctx->xc_capacity, (size_t)M * th.d));
        BO_CUDA(ctx, cudaMemcpyAsync(ctx->dXc, Xc, sizeof(double) * M * th.d, cudaMemcpyHostToDevice, ctx->stream));
        dXc = ctx->dXc;
    }
    double *dOut = out, *dGrad = out_grad;
    if (!dev && out) {
        BO_TRY(bo_reserve(ctx, &ctx->dVal, &ctx->val_capacity, (size_t)M * th.ndraw));
        dOut = ctx->dVal;
    }
    if (!dev && out_grad) {
        BO_TRY(bo_reserve(ctx, &ctx->dGradOut, &ctx->gradout_capacity, (size_t)M * th.ndraw * th.d));
        dGrad = ctx->dGradOut;
    }
    const bool want_best = best_val || best_idx;
    BO_TRY(bo_thompson_run(ctx, M, dXc, dOut, dGrad, want_best ? th.dBestVal : nullptr,
                           want_best ? th.dBestIdx : nullptr));
    ctx->last_val_valid = false;
    cudaStream_t st = ctx->stream;
    if (!dev && out) BO_CUDA(ctx, cudaMemcpyAsync(out, dOut, sizeof(double) * M * th.ndraw, cudaMemcpyDeviceToHost, st));
    if (!dev && out_grad)
        BO_CUDA(ctx, cudaMemcpyAsync(out_grad, dGrad, sizeof(double) * M * th.ndraw * th.d, cudaMemcpyDeviceToHost, st));
    std::vector<double> bv(th.ndraw);
    std::vector<int64_t> bi(th.ndraw);
    if (want_best) {
        BO_CUDA(ctx, cudaMemcpyAsync(bv.data(), th.dBestVal, sizeof(double) * th.ndraw, cudaMemcpyDeviceToHost, st));
        BO_CUDA(ctx, cudaMemcpyAsync(bi.data(), th.dBestIdx, sizeof(int64_t) * th.ndraw, cudaMemcpyDeviceToHost, st));
    }
    if (!dev || want_best) BO_CUDA(ctx, cudaStreamSynchronize(st));
    if (best_val) std::copy(bv.begin(), bv.end(), best_val);
    if (best_idx)
        for (int r = 0; r < th.ndraw; ++r) best_idx[r] = (bi[r] == INT64_MAX) ? 0 : bi[r];
    return BO_OK;
}

// ---------------------------------------------------------------------------
// stand-alone Cholesky / Gram
// ---------------------------------------------------------------------------
extern "C" int bo_cholesky(bo_ctx *ctx, int n, int batch, double *A, int flags, int *info) {
    BO_ENTER(ctx);
    if (n < 1 || batch < 1 || !A) return bo_set_err(ctx, BO_ERR_ARG, "bo_cholesky: bad shape");
    const bool dev = flags & BO_PTR_DEVICE;
    const int np = bo_round_up(n, 64), nblk = np / 64;
    cudaStream_t st = ctx->stream;
    double *dA = nullptr, *dinv = nullptr;
    int *dInfo = nullptr;
    const bool inplace = dev && (np == n);
    // scratch (inverted diagonal blocks, info) persists in the handle: no malloc per call
    BO_TRY(bo_reserve(ctx, &ctx->dCholDinv, &ctx->choldinv_capacity, (size_t)batch * nblk * 4096));
    BO_TRY(bo_reserve(ctx, &ctx->dCholInfo, &ctx->cholinfo_capacity, (size_t)batch));
    dinv = ctx->dCholDinv;
    dInfo = ctx->dCholInfo;
    cudaError_t e = cudaSuccess;
    if (!inplace) e = cudaMalloc(&dA, sizeof(double) * (size_t)batch * np * np);
    if (e != cudaSuccess) return bo_set_err(ctx, BO_ERR_CUDA, "bo_cholesky: %s", cudaGetErrorString(e));
    int rc = BO_OK;
    if (inplace) {
        dA = A;
    } else {
        for (int b = 0; b < batch && e == cudaSuccess; ++b)
            e = cudaMemcpy2DAsync(dA + (size_t)b * np * np, sizeof(double) * np, A + (size_t)b * n * n,
                                  sizeof(double) * n, sizeof(double) * n, n,
                                  dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess && np != n) {
            ctx->launches++;
            pad_identity_kernel<<<dim3((np + 31) / 32, (np + 7) / 8, batch), dim3(32, 8), 0, st>>>(dA, n, np);
        }
    }
    if (e == cudaSuccess) rc = bo_linalg_cholesky(ctx, np, batch, dA, dinv, dInfo);
    if (e == cudaSuccess && rc == BO_OK) {
        ctx->launches++;
        zero_upper_kernel<<<dim3((np + 31) / 32, (np + 7) / 8, batch), dim3(32, 8), 0, st>>>(dA, np, np, (int64_t)np * np);
        if (!inplace)
            for (int b = 0; b < batch && e == cudaSuccess; ++b)
                e = cudaMemcpy2DAsync(A + (size_t)b * n * n, sizeof(double) * n, dA + (size_t)b * np * np,
                                      sizeof(double) * np, sizeof(double) * n, n,
                                      dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st);
        std::vector<int> hinfo(batch, 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(hinfo.data(), dInfo, sizeof(int) * batch, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        bool bad = false;
        for (int b = 0; b < batch; ++b) {
            if (info) info[b] = hinfo[b];
            bad = bad || hinfo[b] != 0;
        }
        if (e == cudaSuccess && bad) rc = bo_set_err(ctx, BO_ERR_NOT_PD, "bo_cholesky: matrix is not positive definite");
    }
    cudaStreamSynchronize(st);
    if (!inplace) cudaFree(dA);
    if (e != cudaSuccess) return bo_set_err(ctx, BO_ERR_CUDA, "bo_cholesky: %s", cudaGetErrorString(e));
    return rc;
}

extern "C" int bo_gram(bo_ctx *ctx, int kernel, int n, int d, const double *X, const double *ell, double rho,
                       double sn2, double *K, int flags) {
    BO_ENTER(ctx);
    if (n < 1 || d < 1 || d > BO_MAX_D || !X || !ell || !K) return bo_set_err(ctx, BO_ERR_ARG, "bo_gram: bad shape");
    if (kernel != BO_KERNEL_SE && kernel != BO_KERNEL_MATERN52) return bo_set_err(ctx, BO_ERR_ARG, "unknown kernel id");
    const bool dev = flags & BO_PTR_DEVICE;
    const int np = bo_round_up(n, BO_PAD), dp = padded_dim(d);
    std::vector<double> xs((size_t)np * dp, 0.0);
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < d; ++k) xs[(size_t)i * dp + k] = X[(size_t)i * d + k] / ell[k];
    double *dXs = nullptr, *dK = nullptr, *dHyp = nullptr;
    double hyp[2] = {rho, sn2};
    cudaStream_t st = ctx->stream;
    cudaError_t e = cudaMalloc(&dXs, sizeof(double) * xs.size());
    if (e == cudaSuccess) e = cudaMalloc(&dK, sizeof(double) * (size_t)np * np);
    if (e == cudaSuccess) e = cudaMalloc(&dHyp, sizeof(double) * 2);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dXs, xs.data(), sizeof(double) * xs.size(), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dHyp, hyp, sizeof(hyp), cudaMemcpyHostToDevice, st);
    int rc = BO_OK;
    if (e == cudaSuccess) rc = bo_linalg_gram(ctx, kernel, n, np, dp, 1, dXs, dHyp, dHyp + 1, dK, 1, nullptr, nullptr);
    if (e == cudaSuccess && rc == BO_OK)
        e = cudaMemcpy2DAsync(K, sizeof(double) * n, dK, sizeof(double) * np, sizeof(double) * n, n,
                              dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(dXs); cudaFree(dK); cudaFree(dHyp);
    if (e != cudaSuccess) return bo_set_err(ctx, BO_ERR_CUDA, "bo_gram: %s", cudaGetErrorString(e));
    return rc;
}

// ---------------------------------------------------------------------------
// profiler
// ---------------------------------------------------------------------------
extern "C" int bo_profile_enable(bo_ctx *ctx, int on) {
    if (!ctx) return BO_ERR_ARG;
    ctx->prof_on = on != 0;
    return BO_OK;
}

extern "C" int bo_profile_reset(bo_ctx *ctx) {
    BO_ENTER(ctx);
    BO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    prof_drain(ctx);
    ctx->prof.clear();
    return BO_OK;
}

extern "C" int bo_profile_count(bo_ctx *ctx, int *count) {
    BO_ENTER(ctx);
    BO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    prof_drain(ctx);
    if (count) *count = (int)ctx->prof.size();
    return BO_OK;
}

extern "C" int bo_profile_get(bo_ctx *ctx, int i, char *name, int name_cap, int64_t *launches, double *total_ms) {
    if (!ctx || i < 0 || i >= (int)ctx->prof.size()) return BO_ERR_ARG;
    if (name && name_cap > 0) {
        strncpy(name, ctx->prof[i].name.c_str(), name_cap - 1);
        name[name_cap - 1] = 0;
    }
    if (launches) *launches = ctx->prof[i].launches;
    if (total_ms) *total_ms = ctx->prof[i].total_ms;
    return BO_OK;
}

extern "C" int bo_launch_count(bo_ctx *ctx, int64_t *launches) {
    if (!ctx || !launches) return BO_ERR_ARG;
    *launches = ctx->launches;
    return BO_OK;
}
