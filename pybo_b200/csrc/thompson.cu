// thompson.cu -- evaluation of weight-space posterior draws
//   f_r(x) = bias_r + scale_r * sum_j cos(W_r[j] . x + b_r[j]) theta_r[j]
// i.e. what `model.sample_f(n, rng).get` returns (reference policies/simple.py:48),
// batched over ndraw draws and M candidates, with a per-draw (max, first argmax).
// The spectral frequencies may be shared by all draws (nW == 1) or per draw.
#include <math.h>

#include "common.cuh"
#include "dgemm.cuh"

#define TH_TILE 64   // features staged per shared-memory tile

__device__ __forceinline__ bool th_better(double v, int64_t i, double bv, int64_t bi) {
    return (v > bv) || (v == bv && i < bi);
}

// grid: (ceil(M/128), ndraw); one thread per candidate.
__global__ void __launch_bounds__(128)
thompson_kernel(int m, int d, int nW, const double *__restrict__ W, const double *__restrict__ b,
                const double *__restrict__ theta, const double *__restrict__ scale,
                const double *__restrict__ bias, int64_t M, const double *__restrict__ Xc,
                double *__restrict__ out, double *__restrict__ grad, double *__restrict__ blkval,
                int64_t *__restrict__ blkidx) {
    __shared__ double sW[TH_TILE][BO_MAX_D + 1];
    __shared__ double sb[TH_TILE], st[TH_TILE];
    const int r = blockIdx.y;
    const int wr = (nW == 1) ? 0 : r;
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const bool live = i < M;
    double x[BO_MAX_D], gacc[BO_MAX_D];
#pragma unroll
    for (int k = 0; k < BO_MAX_D; ++k) {
        x[k] = (live && k < d) ? Xc[i * d + k] : 0.0;
        gacc[k] = 0.0;
    }
    double acc = 0.0;
    for (int j0 = 0; j0 < m; j0 += TH_TILE) {
        __syncthreads();
        for (int e = threadIdx.x; e < TH_TILE * d; e += 128) {
            int jj = e / d, k = e % d;
            sW[jj][k] = (j0 + jj < m) ? W[((int64_t)wr * m + j0 + jj) * d + k] : 0.0;
        }
        for (int e = threadIdx.x; e < TH_TILE; e += 128) {
            sb[e] = (j0 + e < m) ? b[(int64_t)wr * m + j0 + e] : 0.0;
            st[e] = (j0 + e < m) ? theta[(int64_t)r * m + j0 + e] : 0.0;
        }
        __syncthreads();
        const int lim = (m - j0) < TH_TILE ? (m - j0) : TH_TILE;
        for (int jj = 0; jj < lim; ++jj) {
            double a = sb[jj];
            for (int k = 0; k < d; ++k) a = fma(sW[jj][k], x[k], a);
            double sn, cs;
            sincos(a, &sn, &cs);
            acc = fma(cs, st[jj], acc);
            if (grad != nullptr) {
                const double w = -sn * st[jj];
                for (int k = 0; k < d; ++k) gacc[k] = fma(w, sW[jj][k], gacc[k]);
            }
        }
    }
    const double sc = scale[r];
    const double val = bias[r] + sc * acc;
    if (live) {
        if (out) out[(int64_t)r * M + i] = val;
        if (grad)
            for (int k = 0; k < d; ++k) grad[((int64_t)r * M + i) * d + k] = sc * gacc[k];
    }
    if (blkval == nullptr) return;
    __shared__ double sv[4];
    __shared__ int64_t si[4];
    double bv = live ? val : -INFINITY;
    int64_t bi = live ? i : INT64_MAX;
    if (bv != bv) { bv = -INFINITY; bi = INT64_MAX; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (th_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 4; ++w)
            if (th_better(sv[w], si[w], bv, bi)) { bv = sv[w]; bi = si[w]; }
        blkval[(int64_t)r * gridDim.x + blockIdx.x] = bv;
        blkidx[(int64_t)r * gridDim.x + blockIdx.x] = bi;
    }
}

// one block per draw
__global__ void __launch_bounds__(256)
thompson_best_kernel(const double *__restrict__ blkval, const int64_t *__restrict__ blkidx, int nb,
                     double *__restrict__ bestval, int64_t *__restrict__ bestidx) {
    __shared__ double sv[8];
    __shared__ int64_t si[8];
    const int r = blockIdx.x;
    double bv = -INFINITY;
    int64_t bi = INT64_MAX;
    for (int i = threadIdx.x; i < nb; i += 256) {
        double v = blkval[(int64_t)r * nb + i];
        int64_t ix = blkidx[(int64_t)r * nb + i];
        if (th_better(v, ix, bv, bi)) { bv = v; bi = ix; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (th_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w)
            if (th_better(sv[w], si[w], bv, bi)) { bv = sv[w]; bi = si[w]; }
        bestval[r] = bv;
        bestidx[r] = bi;
    }
}

// ---------------------------------------------------------------------------
// Shared-basis batch (nW == 1): F = Phi Theta with Phi[i][j] = cos(w_j . x_i + b_j) generated on
// the fly as the A operand of an FP64 tensor-core (DMMA) contraction over the features.
// Tile: 64 candidates x 256 draws, BK = 16 features; 8 warps as 2 (candidates) x 4 (draws).
// ---------------------------------------------------------------------------
#define TG_BM 64
#define TG_BN 256
#define TG_BK 16
#define TG_SMEM_BYTES ((TG_BM * (TG_BK + 4) + 2 * TG_BK * (TG_BN + 4) + TG_BK * BO_MAX_D + TG_BK + 4 * TG_BN) * 8)
__global__ void __launch_bounds__(256, 1)
thompson_gemm_kernel(int m, int d, int ndraw, int ndp, const double *__restrict__ W, const double *__restrict__ b,
                     const double *__restrict__ thetaT /* mp x ndp */, const double *__restrict__ scale,
                     const double *__restrict__ bias, int64_t M, const double *__restrict__ Xc,
                     double *__restrict__ out, double *__restrict__ blkval, int64_t *__restrict__ blkidx) {
    extern __shared__ __align__(16) double tg_smem[];
    double (*As)[TG_BK + 4] = reinterpret_cast<double (*)[TG_BK + 4]>(tg_smem);
    double (*Bs)[TG_BK][TG_BN + 4] = reinterpret_cast<double (*)[TG_BK][TG_BN + 4]>(tg_smem + TG_BM * (TG_BK + 4));
    double (*sW)[BO_MAX_D] = reinterpret_cast<double (*)[BO_MAX_D]>(tg_smem + TG_BM * (TG_BK + 4) + 2 * TG_BK * (TG_BN + 4));
    double *sb = tg_smem + TG_BM * (TG_BK + 4) + 2 * TG_BK * (TG_BN + 4) + TG_BK * BO_MAX_D;
    double (*rv)[TG_BN] = reinterpret_cast<double (*)[TG_BN]>(sb + TG_BK);
    int64_t (*ri)[TG_BN] = reinterpret_cast<int64_t (*)[TG_BN]>(sb + TG_BK + 2 * TG_BN);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 2, wn = warp & 3, g = lane >> 2, t = lane & 3;
    const int64_t i0 = (int64_t)blockIdx.x * TG_BM;
    const int n0 = blockIdx.y * TG_BN;
    // feature generation: thread -> candidate (tid & 63), features (tid >> 6) * 4 .. + 3
    const int gc = tid & 63, gj = (tid >> 6) * 4;
    double x[BO_MAX_D];
    {
        const int64_t gi = i0 + gc;
#pragma unroll
        for (int k = 0; k < BO_MAX_D; ++k) x[k] = (gi < M && k < d) ? Xc[gi * d + k] : 0.0;
    }
    double acc[4][8][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    const int nk = (m + TG_BK - 1) / TG_BK;
    auto load_B = [&](int kt, int buf) {
        const double *src = thetaT + (int64_t)kt * TG_BK * ndp + n0;
        for (int c = tid; c < TG_BK * (TG_BN / 2); c += 256) {
            const int r = c / (TG_BN / 2), cc = (c % (TG_BN / 2)) * 2;
            cp_async16(&Bs[buf][r][cc], src + (int64_t)r * ndp + cc);
        }
        cp_async_commit();
    };
    load_B(0, 0);
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        __syncthreads();                                   // previous tile fully consumed
        for (int e = tid; e < TG_BK * d; e += 256) {
            const int jj = e / d, k = e % d, j = kt * TG_BK + jj;
            sW[jj][k] = (j < m) ? W[(int64_t)j * d + k] : 0.0;
        }
        if (tid < TG_BK) sb[tid] = (kt * TG_BK + tid < m) ? b[kt * TG_BK + tid] : 0.0;
        if (kt + 1 < nk) load_B(kt + 1, buf ^ 1);
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int jj = gj + q;
            double a = sb[jj];
            for (int k = 0; k < d; ++k) a = fma(sW[jj][k], x[k], a);
            As[gc][jj] = (kt * TG_BK + jj < m) ? cos(a) : 0.0;
        }
        if (kt + 1 < nk) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TG_BK; kk += 4) {
            double af[4], bf[8];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) af[mi] = As[wm * 32 + mi * 8 + g][kk + t];
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) bf[ni] = Bs[buf][kk + t][wn * 64 + ni * 8 + g];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 8; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
        }
    }
    // epilogue: scale / bias, store, per-draw (max, first arg max) over this block's candidates
#pragma unroll
    for (int ni = 0; ni < 8; ++ni)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int r = n0 + wn * 64 + ni * 8 + 2 * t + e;
            double bv = -INFINITY;
            int64_t bi = INT64_MAX;
            if (r < ndraw) {
                const double sc = scale[r], bs = bias[r];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) {
                    const int64_t i = i0 + wm * 32 + mi * 8 + g;
                    const double v = bs + sc * acc[mi][ni][e];
                    if (i < M) {
                        if (out) out[(int64_t)r * M + i] = v;
                        if (th_better(v, i, bv, bi)) { bv = v; bi = i; }
                    }
                }
            }
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (th_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
            }
            if (g == 0) {
                rv[wm][wn * 64 + ni * 8 + 2 * t + e] = bv;
                ri[wm][wn * 64 + ni * 8 + 2 * t + e] = bi;
            }
        }
    __syncthreads();
    if (blkval != nullptr && tid < TG_BN && n0 + tid < ndraw) {
        double bv = rv[0][tid];
        int64_t bi = ri[0][tid];
        if (th_better(rv[1][tid], ri[1][tid], bv, bi)) { bv = rv[1][tid]; bi = ri[1][tid]; }
        blkval[(int64_t)(n0 + tid) * gridDim.x + blockIdx.x] = bv;
        blkidx[(int64_t)(n0 + tid) * gridDim.x + blockIdx.x] = bi;
    }
}

int bo_thompson_run(bo_ctx *ctx, int64_t M, const double *dXc, double *dOut, double *dGrad,
                    double *dBestVal, int64_t *dBestIdx) {
    bo_thompson_state &th = ctx->th;
    if (th.ndraw == 0) return bo_set_err(ctx, BO_ERR_STATE, "bo_thompson_eval before bo_thompson_set");
    const bool gemm_path = (th.nW == 1) && (dGrad == nullptr) && th.thetaT != nullptr;
    // shared basis, values only, int8 path selected: tcgen05 contraction of sliced cosine features (ozaki.cu)
    const bool oz_path = gemm_path && bo_thompson_ozaki_usable(ctx, M);
    const int nb = oz_path ? (int)bo_thompson_ozaki_blocks(M)
                           : (gemm_path ? (int)((M + TG_BM - 1) / TG_BM) : (int)((M + 127) / 128));
    double *blkval = nullptr;
    int64_t *blkidx = nullptr;
    if (dBestVal != nullptr) {
        size_t need = (size_t)nb * th.ndraw + 8;
        if (ctx->blk_capacity < need || !ctx->dBlkVal) {
            size_t c1 = ctx->blk_capacity, c2 = ctx->blk_capacity;
            BO_TRY(bo_reserve(ctx, &ctx->dBlkVal, &c1, need));
            BO_TRY(bo_reserve(ctx, &ctx->dBlkIdx, &c2, need));
            ctx->blk_capacity = need;
        }
        blkval = ctx->dBlkVal;
        blkidx = ctx->dBlkIdx;
    }
    if (oz_path) {
        BO_TRY(bo_thompson_ozaki_run(ctx, M, dXc, dOut, blkval, blkidx, nb));
    } else if (gemm_path) {
        BO_LAUNCH(ctx, "thompson_gemm_kernel");
        thompson_gemm_kernel<<<dim3(nb, th.ndp / TG_BN), 256, TG_SMEM_BYTES, ctx->stream>>>(
            th.m, th.d, th.ndraw, th.ndp, th.W, th.b, th.thetaT, th.scale, th.bias, M, dXc, dOut, blkval, blkidx);
        BO_CHECK_LAUNCH(ctx);
    } else {
        BO_LAUNCH(ctx, "thompson_kernel");
        thompson_kernel<<<dim3(nb, th.ndraw), 128, 0, ctx->stream>>>(
            th.m, th.d, th.nW, th.W, th.b, th.theta, th.scale, th.bias, M, dXc, dOut, dGrad, blkval, blkidx);
        BO_CHECK_LAUNCH(ctx);
    }
    if (dBestVal != nullptr) {
        BO_LAUNCH(ctx, "thompson_best_kernel");
        thompson_best_kernel<<<th.ndraw, 256, 0, ctx->stream>>>(blkval, blkidx, nb, dBestVal, dBestIdx);
        BO_CHECK_LAUNCH(ctx);
    }
    return BO_OK;
}

int bo_thompson_init(bo_ctx *ctx) {
    BO_CUDA(ctx, cudaFuncSetAttribute(thompson_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM_BYTES));
    return BO_OK;
}
