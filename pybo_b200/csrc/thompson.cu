// thompson.cu -- evaluation of weight-space posterior draws
//   f_r(x) = bias_r + scale_r * sum_j cos(W_r[j] . x + b_r[j]) theta_r[j]
// i.e. what `model.sample_f(n, rng).get` returns (reference policies/simple.py:48),
// batched over ndraw draws and M candidates, with a per-draw (max, first argmax).
// The spectral frequencies may be shared by all draws (nW == 1) or per draw.
#include <math.h>

#include "common.cuh"

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

int bo_thompson_run(bo_ctx *ctx, int64_t M, const double *dXc, double *dOut, double *dGrad,
                    double *dBestVal, int64_t *dBestIdx) {
    bo_thompson_state &th = ctx->th;
    if (th.ndraw == 0) return bo_set_err(ctx, BO_ERR_STATE, "bo_thompson_eval before bo_thompson_set");
    const int nb = (int)((M + 127) / 128);
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
    {
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
