// thompson_build.cu -- construction of weight-space posterior draws on the device: what
// `model.sample_f(n, rng)` (reference policies/simple.py:48) does once per Thompson iteration before
// its `.get` is evaluated.  For a basis of m random Fourier features phi(x) = scale cos(W x + b),
// scale = sqrt(2 rho / m):
//     Phi = phi(X)                      (n x m)      th_features_kernel   (stored transposed, m x n)
//     A   = Phi^T Phi + sn2 I           (m x m)      DMMA GEMM (dgemm_kernel, NT)
//     z   = Phi^T (y - bias)                         th_gemv_kernel
//     L   = chol(A), Wm = L^-1                       bo_linalg_cholesky / bo_linalg_trtri (linalg.cu)
//     u   = Wm z                                     th_lower_mv_kernel
//     theta_r = Wm^T (u + sqrt(sn2) eps_r)           th_theta_kernel
// i.e. theta_r ~ N(A^-1 Phi^T r, sn2 A^-1).  Batched over nW bases: nW == 1 (all ndraw draws share one
// basis: BASELINE config 4 as one dense contraction) or nW == ndraw (one basis per draw = ndraw
// independent `sample_f` calls).  The random numbers (W, b, eps) come from the caller's NumPy stream,
// so a draw is the same function whether it was built here or by the oracle.
#include <math.h>

#include <algorithm>

#include "common.cuh"
#include "dgemm.cuh"

typedef DTile<128, 128, 64, 32, 4, true> TB128NT;

__device__ __forceinline__ double tb_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// PhiT[b][j][i] = scale cos(W_b[j] . x_i + b_b[j]) for j < m, i < n; 0 on the padding.  grid (npk / 128, mp, nW)
__global__ void __launch_bounds__(128)
th_features_kernel(int n, int npk, int d, int m, int mp, const double *__restrict__ X, const double *__restrict__ W,
                   const double *__restrict__ b, double scale, double *__restrict__ PhiT) {
    __shared__ double w[BO_MAX_D + 1];
    const int j = blockIdx.y, bb = blockIdx.z;
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (j < m) {
        if (threadIdx.x < d) w[threadIdx.x] = W[((int64_t)bb * m + j) * d + threadIdx.x];
        if (threadIdx.x == d) w[d] = b[(int64_t)bb * m + j];
    }
    __syncthreads();
    double v = 0.0;
    if (j < m && i < n) {
        double a = w[d];
        for (int k = 0; k < d; ++k) a = fma(w[k], X[(int64_t)i * d + k], a);
        v = scale * cos(a);
    }
    PhiT[((int64_t)bb * mp + j) * npk + i] = v;
}

// A[b][j][j] += sn2 (j < m), := 1 on the padded diagonal (rows of zeros otherwise)
__global__ void th_diag_kernel(int m, int mp, double sn2, double *__restrict__ A) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= mp) return;
    double *a = A + (int64_t)blockIdx.y * mp * mp + (int64_t)j * (mp + 1);
    *a = (j < m) ? *a + sn2 : 1.0;
}

// z[b][j] = sum_i PhiT[b][j][i] (y_i - bias): one warp per row
__global__ void __launch_bounds__(256)
th_gemv_kernel(int n, int npk, int mp, const double *__restrict__ PhiT, const double *__restrict__ y, double bias,
               double *__restrict__ z) {
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= mp) return;
    const double *row = PhiT + ((int64_t)blockIdx.y * mp + j) * npk;
    double acc = 0.0;
    for (int i = lane; i < n; i += 32) acc = fma(row[i], y[i] - bias, acc);
    acc = tb_warp_sum(acc);
    if (lane == 0) z[(int64_t)blockIdx.y * mp + j] = acc;
}

// u[b][i] = sum_{j <= i} Wm[b][i][j] z[b][j]
__global__ void __launch_bounds__(256)
th_lower_mv_kernel(int mp, const double *__restrict__ Wm, const double *__restrict__ z, double *__restrict__ u) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= mp) return;
    const double *row = Wm + ((int64_t)blockIdx.y * mp + i) * mp;
    const double *zz = z + (int64_t)blockIdx.y * mp;
    double acc = 0.0;
    for (int j = lane; j <= i; j += 32) acc = fma(row[j], zz[j], acc);
    acc = tb_warp_sum(acc);
    if (lane == 0) u[(int64_t)blockIdx.y * mp + i] = acc;
}

// theta[r][j] = sum_{i >= j} WmT[b][j][i] (u[b][i] + sq eps[r][i]) for the R draws r = b R + r0 .. of basis b;
// one warp per row j, eight draws per pass.  grid (mp / 8, ceil(R / 8), nW)
__global__ void __launch_bounds__(256)
th_theta_kernel(int m, int mp, int R, double sq, const double *__restrict__ WmT, const double *__restrict__ u,
                const double *__restrict__ eps /* ndraw x mp */, double *__restrict__ theta /* ndraw x m */) {
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int bb = blockIdx.z, r0 = blockIdx.y * 8;
    if (j >= m) return;
    const double *row = WmT + ((int64_t)bb * mp + j) * mp;
    const double *uu = u + (int64_t)bb * mp;
    const int nr = (R - r0) < 8 ? (R - r0) : 8;
    const double *e0 = eps + ((int64_t)bb * R + r0) * mp;
    double acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.0;
    for (int i = (j & ~31) + lane; i < mp; i += 32) {
        if (i < j) continue;
        const double w = row[i], ui = uu[i];
#pragma unroll
        for (int c = 0; c < 8; ++c)
            if (c < nr) acc[c] = fma(w, fma(sq, e0[(int64_t)c * mp + i], ui), acc[c]);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        acc[c] = tb_warp_sum(acc[c]);
        if (lane == 0 && c < nr) theta[((int64_t)bb * R + r0 + c) * m + j] = acc[c];
    }
}

int bo_thompson_build_init(bo_ctx *ctx) {
    BO_CUDA(ctx, cudaFuncSetAttribute(dgemm_kernel<TB128NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, TB128NT::SMEM_BYTES));
    return BO_OK;
}

extern "C" int bo_thompson_build(bo_ctx *ctx, int n, int d, const double *X, const double *y, double rho, double sn2,
                                 double bias, int ndraw, int nW, int m, const double *W, const double *b,
                                 const double *noise, double *theta_out) {
    if (!ctx) return BO_ERR_ARG;
    BO_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n < 1 || d < 1 || d > BO_MAX_D || m < 1 || ndraw < 1 || (nW != 1 && nW != ndraw) || !X || !y || !W || !b || !noise ||
        !(rho > 0.0) || !(sn2 > 0.0))
        return bo_set_err(ctx, BO_ERR_ARG, "bo_thompson_build: bad arguments (n=%d d=%d m=%d ndraw=%d nW=%d; sn2 must be > 0)",
                          n, d, m, ndraw, nW);
    const int mp = bo_round_up(m, 128), npk = bo_round_up(n, 128), R = ndraw / nW, nblk = mp / 64;
    const double scale = sqrt(2.0 * rho / m), sq = sqrt(sn2);
    const size_t nPhi = (size_t)nW * mp * npk, nMat = (size_t)nW * mp * mp;
    const size_t nIn = (size_t)n * d + n + (size_t)nW * m * d + (size_t)nW * m + (size_t)ndraw * mp;
    const size_t total = nPhi + 3 * nMat + 2 * (size_t)nW * mp + nIn + (size_t)ndraw * m + (size_t)nW * nblk * 4096;
    BO_TRY(bo_reserve(ctx, &ctx->th.build, &ctx->th.build_capacity, total));
    BO_TRY(bo_reserve(ctx, &ctx->dCholInfo, &ctx->cholinfo_capacity, (size_t)nW));
    double *PhiT = ctx->th.build, *A = PhiT + nPhi, *Wm = A + nMat, *T = Wm + nMat, *z = T + nMat, *u = z + (size_t)nW * mp;
    double *dX = u + (size_t)nW * mp, *dY = dX + (size_t)n * d, *dW = dY + n, *dB = dW + (size_t)nW * m * d;
    double *dEps = dB + (size_t)nW * m, *dTheta = dEps + (size_t)ndraw * mp, *dinv = dTheta + (size_t)ndraw * m;
    cudaStream_t st = ctx->stream;
    BO_CUDA(ctx, cudaMemcpyAsync(dX, X, sizeof(double) * n * d, cudaMemcpyHostToDevice, st));
    BO_CUDA(ctx, cudaMemcpyAsync(dY, y, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    BO_CUDA(ctx, cudaMemcpyAsync(dW, W, sizeof(double) * nW * m * d, cudaMemcpyHostToDevice, st));
    BO_CUDA(ctx, cudaMemcpyAsync(dB, b, sizeof(double) * nW * m, cudaMemcpyHostToDevice, st));
    BO_CUDA(ctx, cudaMemsetAsync(dEps, 0, sizeof(double) * ndraw * mp, st));
    BO_CUDA(ctx, cudaMemcpy2DAsync(dEps, sizeof(double) * mp, noise, sizeof(double) * m, sizeof(double) * m, ndraw,
                                   cudaMemcpyHostToDevice, st));
    {
        BO_LAUNCH(ctx, "th_features_kernel");
        th_features_kernel<<<dim3(npk / 128, mp, nW), 128, 0, st>>>(n, npk, d, m, mp, dX, dW, dB, scale, PhiT);
        BO_CHECK_LAUNCH(ctx);
    }
    {   // A = PhiT PhiT^T
        DGemmParams p = {};
        p.A = PhiT; p.lda = npk; p.B = PhiT; p.ldb = npk; p.C = A; p.ldc = mp;
        p.strideA = p.strideB = (int64_t)mp * npk; p.strideC = (int64_t)mp * mp;
        p.inner = nW; p.tiles_m = p.tiles_n = mp / 128; p.K = npk; p.krule = KR_FULL; p.alpha = 1.0; p.beta = 0.0;
        BO_LAUNCH(ctx, "th_syrk_kernel");
        dgemm_kernel<TB128NT><<<dim3(p.tiles_m * p.tiles_n, 1, nW), TB128NT::NTHREADS, TB128NT::SMEM_BYTES, st>>>(p);
        BO_CHECK_LAUNCH(ctx);
    }
    {
        BO_LAUNCH(ctx, "th_diag_kernel");
        th_diag_kernel<<<dim3((mp + 127) / 128, nW), 128, 0, st>>>(m, mp, sn2, A);
        BO_CHECK_LAUNCH(ctx);
    }
    {
        BO_LAUNCH(ctx, "th_gemv_kernel");
        th_gemv_kernel<<<dim3(mp / 8, nW), 256, 0, st>>>(n, npk, mp, PhiT, dY, bias, z);
        BO_CHECK_LAUNCH(ctx);
    }
    BO_TRY(bo_linalg_cholesky(ctx, mp, nW, A, dinv, ctx->dCholInfo));
    std::vector<int> info(nW, 0);
    BO_CUDA(ctx, cudaMemcpyAsync(info.data(), ctx->dCholInfo, sizeof(int) * nW, cudaMemcpyDeviceToHost, st));
    BO_CUDA(ctx, cudaStreamSynchronize(st));
    for (int i = 0; i < nW; ++i)
        if (info[i] != 0)
            return bo_set_err(ctx, BO_ERR_NOT_PD, "bo_thompson_build: feature system %d is not positive definite (pivot %d)", i, info[i]);
    BO_TRY(bo_linalg_trtri(ctx, mp, nW, A, dinv, Wm, T));
    BO_TRY(bo_linalg_transpose(ctx, mp, nW, Wm, T));
    {
        BO_LAUNCH(ctx, "th_lower_mv_kernel");
        th_lower_mv_kernel<<<dim3(mp / 8, nW), 256, 0, st>>>(mp, Wm, z, u);
        BO_CHECK_LAUNCH(ctx);
    }
    {
        BO_LAUNCH(ctx, "th_theta_kernel");
        th_theta_kernel<<<dim3(mp / 8, (R + 7) / 8, nW), 256, 0, st>>>(m, mp, R, sq, T, u, dEps, dTheta);
        BO_CHECK_LAUNCH(ctx);
    }
    // hand the draws to the evaluation state (bo_thompson_eval) exactly as bo_thompson_set would
    std::vector<double> theta((size_t)ndraw * m), sc(ndraw, scale), bs(ndraw, bias);
    BO_CUDA(ctx, cudaMemcpyAsync(theta.data(), dTheta, sizeof(double) * ndraw * m, cudaMemcpyDeviceToHost, st));
    BO_CUDA(ctx, cudaStreamSynchronize(st));
    if (theta_out) std::copy(theta.begin(), theta.end(), theta_out);
    return bo_thompson_set(ctx, ndraw, nW, m, d, W, b, theta.data(), sc.data(), bs.data());
}
