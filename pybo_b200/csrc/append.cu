// append.cu -- incremental refit: one observation appended to a fitted factor set in O(n^2)
// instead of the O(n^3) refactorisation `model.add_data` (reference bayesopt.py:269) implies.
//
// With K' = [[K, k], [k^T, kss]] (kss = rho + sn2) the bordered factors are
//     L' = [[L, 0], [l^T, lam]],        l = L^-1 k = W k,     lam = sqrt(kss - |l|^2)
//     W' = [[W, 0], [w^T, 1/lam]],      w = -(W^T l) / lam
//     alpha' = [alpha, a],              a = (y - bias - l . alpha) / lam
//     beta'  = W'^T alpha' = [beta + a w, a / lam],            log|L'| = log|L| + log lam
// i.e. two triangular matrix-vector products that stream W and W^T once each (HBM-bound:
// n^2/2 * 8 B per product), a handful of dot products and two row writes.  The padded layout
// (factor dimension np = multiple of 128, identity on the padded diagonal) already holds room
// for the new row, so nothing is re-laid out until n reaches np.
#include <math.h>

#include "common.cuh"

#define AP_ROWS 8            // rows (warps) per block in the matrix-vector kernels

__device__ __forceinline__ double ap_warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double ap_kernel_value(int kernel, double D, double rho) {
    if (kernel == BO_KERNEL_SE) return rho * exp(-0.5 * D);
    const double r = sqrt(5.0 * D);
    return rho * (1.0 + r + r * r * (1.0 / 3.0)) * exp(-r);
}

// kvec[j] = k(x_j, x_new) for j < n, 0 on [n, np); also stores the new scaled row of Xs.
// (`raw` = {x[d], y} of the new observation, appended to the raw copies dX / dY by the first sample's launch)
__global__ void append_kvec_kernel(int kernel, int n, int np, int dp, double rho, double *__restrict__ Xs,
                                   const double *__restrict__ xnew, double *__restrict__ kvec, const double *__restrict__ raw,
                                   int d, double *__restrict__ dX, double *__restrict__ dY) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (raw != nullptr && blockIdx.x == 0 && threadIdx.x <= d) {
        if (threadIdx.x < d) dX[(int64_t)n * d + threadIdx.x] = raw[threadIdx.x];
        else dY[n] = raw[d];
    }
    if (j >= np) return;
    double v = 0.0;
    if (j < n) {
        double D = 0.0;
        for (int k = 0; k < dp; ++k) {
            const double t = Xs[(int64_t)j * dp + k] - xnew[k];
            D = fma(t, t, D);
        }
        v = ap_kernel_value(kernel, D, rho);
    }
    kvec[j] = v;
    if (j == n)
        for (int k = 0; k < dp; ++k) Xs[(int64_t)n * dp + k] = xnew[k];
}

// Dot product of one triangular row with a vector: elements [lo, hi) of `row` (lo, hi arbitrary; the
// 16-byte loads start at the even index below lo and are masked at both ends), four independent
// 16-byte loads per lane in flight.
__device__ __forceinline__ double ap_row_dot(const double *__restrict__ row, const double *__restrict__ vec, int lo, int hi,
                                             int lane) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    const int beg = lo & ~1;
    int j = beg + 2 * lane;
    // 256 elements per round: four independent 16-byte loads of the matrix row per lane in flight (a warp streams its
    // two rows alone, so the bytes it keeps in flight set the bandwidth it reaches)
    for (; j + 192 < hi; j += 256) {
        const double2 w0 = *reinterpret_cast<const double2 *>(row + j);
        const double2 w1 = *reinterpret_cast<const double2 *>(row + j + 64);
        const double2 w2 = *reinterpret_cast<const double2 *>(row + j + 128);
        const double2 w3 = *reinterpret_cast<const double2 *>(row + j + 192);
        const double2 v0 = *reinterpret_cast<const double2 *>(vec + j);
        const double2 v1 = *reinterpret_cast<const double2 *>(vec + j + 64);
        const double2 v2 = *reinterpret_cast<const double2 *>(vec + j + 128);
        const double2 v3 = *reinterpret_cast<const double2 *>(vec + j + 192);
        a0 = fma(j >= lo ? w0.x : 0.0, v0.x, a0);
        a1 = fma(w0.y, v0.y, a1);
        a2 = fma(w1.x, v1.x, a2);
        a3 = fma(w1.y, v1.y, a3);
        a0 = fma(w2.x, v2.x, a0);
        a1 = fma(w2.y, v2.y, a1);
        a2 = fma(w3.x, v3.x, a2);
        a3 = fma(j + 193 < hi ? w3.y : 0.0, v3.y, a3);
    }
    for (; j < hi; j += 64) {
        const double2 w0 = *reinterpret_cast<const double2 *>(row + j);
        const double2 v0 = *reinterpret_cast<const double2 *>(vec + j);
        a0 = fma(j >= lo ? w0.x : 0.0, v0.x, a0);
        a1 = fma(j + 1 < hi ? w0.y : 0.0, v0.y, a1);
    }
    return ap_warp_sum((a0 + a1) + (a2 + a3));
}

// The same dot product shared by the AP_ROWS warps of a block: warp w takes the 256-element rounds w, w + AP_ROWS, ...
// of the row, so a 4096-element row is two dependent load rounds per warp instead of sixteen (one warp streaming a
// row alone is bound by the latency of its own load rounds: 29 us per product at n = 4096, a third of the HBM roof).
// Returns the block-wide sum in every thread of warp 0 (red: AP_ROWS doubles of shared memory).
__device__ __forceinline__ double ap_row_dot_block(const double *__restrict__ row, const double *__restrict__ vec, int lo, int hi,
                                                   double *red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int beg = lo & ~1;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    for (int j0 = beg + 256 * w; j0 < hi; j0 += 256 * AP_ROWS) {
        const int j = j0 + 2 * lane;
        double2 wv[4], vv[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int jt = j + 64 * t;
            const bool in = jt < hi;                           // (rows are padded to an even length >= hi)
            wv[t] = in ? *reinterpret_cast<const double2 *>(row + jt) : make_double2(0.0, 0.0);
            vv[t] = in ? *reinterpret_cast<const double2 *>(vec + jt) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int jt = j + 64 * t;
            const double x = (jt >= lo) ? wv[t].x : 0.0;
            const double y = (jt + 1 < hi) ? wv[t].y : 0.0;
            if (t & 1) { a2 = fma(x, vv[t].x, a2); a3 = fma(y, vv[t].y, a3); }
            else { a0 = fma(x, vv[t].x, a0); a1 = fma(y, vv[t].y, a1); }
        }
    }
    const double part = ap_warp_sum((a0 + a1) + (a2 + a3));
    __syncthreads();                                           // red is reused between calls
    if (lane == 0) red[w] = part;
    __syncthreads();
    double tot = 0.0;
    if (w == 0) {
#pragma unroll
        for (int i = 0; i < AP_ROWS; ++i) tot += red[i];       // fixed order
    }
    return tot;
}

// lvec[i] = sum_{j <= i} W[i][j] kvec[j] for i < n (coalesced along the rows), lvec[i] = 0 on [n, np).
// A warp takes rows r and n - 1 - r, so every warp streams n + 1 elements whatever r is.
__global__ void __launch_bounds__(32 * AP_ROWS)
append_wk_kernel(const double *__restrict__ W, const double *__restrict__ kvec, int n, int np, double *__restrict__ lvec) {
    __shared__ double red[AP_ROWS];
    const int r = blockIdx.x;                                  // one block per row pair (r, n - 1 - r): n + 1 elements
    const int half = (n + 1) >> 1;
    if (r < half) {
        const int r2 = n - 1 - r;
        const double d0 = ap_row_dot_block(W + (int64_t)r * np, kvec, 0, r + 1, red);
        if (threadIdx.x == 0) lvec[r] = d0;
        if (r2 != r) {
            const double d1 = ap_row_dot_block(W + (int64_t)r2 * np, kvec, 0, r2 + 1, red);
            if (threadIdx.x == 0) lvec[r2] = d1;
        }
    }
    // zero the padding once (last block)
    if (blockIdx.x == gridDim.x - 1)
        for (int i = n + threadIdx.x; i < np; i += blockDim.x) lvec[i] = 0.0;
}

// One block: lam, a, the new row of L, alpha[n], log|L|.  scal = {lam, a, 1/lam}; info = n + 1 when
// the bordered matrix is not positive definite (nothing is written in that case).
__global__ void append_pivot_kernel(const double *__restrict__ lvec, double *__restrict__ alpha, double *__restrict__ L,
                                    double *__restrict__ logdet, int n, int np, double kss, double resid,
                                    double *__restrict__ scal, int *__restrict__ info) {
    __shared__ double red[2][32];
    __shared__ double sh_lam;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double q = 0.0, la = 0.0;
    for (int i = tid; i < n; i += blockDim.x) {
        const double l = lvec[i];
        q = fma(l, l, q);
        la = fma(l, alpha[i], la);
    }
    q = ap_warp_sum(q);
    la = ap_warp_sum(la);
    if (lane == 0) { red[0][wid] = q; red[1][wid] = la; }
    __syncthreads();
    if (wid == 0) {
        q = lane < (int)(blockDim.x >> 5) ? red[0][lane] : 0.0;
        la = lane < (int)(blockDim.x >> 5) ? red[1][lane] : 0.0;
        q = ap_warp_sum(q);
        la = ap_warp_sum(la);
        if (lane == 0) {
            const double piv = kss - q;
            if (!(piv > 0.0)) {
                *info = n + 1;
                sh_lam = 0.0;
            } else {
                const double lam = sqrt(piv);
                const double a = (resid - la) / lam;
                scal[0] = lam; scal[1] = a; scal[2] = 1.0 / lam;
                alpha[n] = a;
                *logdet += log(lam);
                *info = 0;
                sh_lam = lam;
            }
        }
    }
    __syncthreads();
    const double lam = sh_lam;
    if (lam == 0.0) return;
    double *row = L + (int64_t)n * np;
    for (int j = tid; j < n; j += blockDim.x) row[j] = lvec[j];
    if (tid == 0) row[n] = lam;
}

// w_j = -(sum_{i = j}^{n-1} WT[j][i] l_i) / lam for j < n (rows of W^T, paired j / n - 1 - j per warp),
// written to row n of W and column n of W^T; beta_j += a w_j.  The last block closes the diagonal.
__global__ void __launch_bounds__(32 * AP_ROWS)
append_wrow_kernel(double *__restrict__ W, double *__restrict__ WT, const double *__restrict__ lvec,
                   const double *__restrict__ scal, const int *__restrict__ info, double *__restrict__ beta, int n, int np) {
    __shared__ double red[AP_ROWS];
    if (*info != 0) return;
    const int r = blockIdx.x;                                  // one block per row pair (r, n - 1 - r)
    const double lam_inv = scal[2], a = scal[1];
    const int half = (n + 1) >> 1;
    if (r >= half) {                                           // the extra block closes the diagonal
        if (threadIdx.x == 0) {
            W[(int64_t)n * np + n] = lam_inv;
            WT[(int64_t)n * np + n] = lam_inv;
            beta[n] = a * lam_inv;
        }
        return;
    }
#pragma unroll 1
    for (int t = 0; t < 2; ++t) {
        const int j = t == 0 ? r : n - 1 - r;
        if (t == 1 && j == r) break;
        const double acc = ap_row_dot_block(WT + (int64_t)j * np, lvec, j, n, red);
        if (threadIdx.x == 0) {
            const double w = -acc * lam_inv;
            W[(int64_t)n * np + j] = w;
            WT[(int64_t)j * np + n] = w;
            beta[j] = fma(a, w, beta[j]);
        }
    }
}

int bo_ozaki_append_row(bo_ctx *ctx, int row);

extern "C" int bo_fit_capacity(bo_ctx *ctx, int *capacity) {
    if (!ctx || !capacity) return BO_ERR_ARG;
    if (!ctx->fitted) return bo_set_err(ctx, BO_ERR_STATE, "not fitted");
    *capacity = ctx->np;
    return BO_OK;
}

extern "C" int bo_append(bo_ctx *ctx, int m, const double *Xnew, const double *ynew) {
    if (!ctx) return BO_ERR_ARG;
    BO_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->fitted) return bo_set_err(ctx, BO_ERR_STATE, "bo_append before bo_fit");
    if (m < 1 || !Xnew || !ynew) return bo_set_err(ctx, BO_ERR_ARG, "bo_append: bad arguments");
    if (ctx->n + m > ctx->np)
        return bo_set_err(ctx, BO_ERR_STATE, "bo_append: %d + %d observations exceed the padded capacity %d; refit", ctx->n, m, ctx->np);
    const int np = ctx->np, dp = ctx->dp, d = ctx->d, S = ctx->S;
    cudaStream_t st = ctx->stream;
    // scratch: kvec[np], lvec[np], S x (xnew[dp] + scal[4]), raw {x[d], y}, info[S]
    const size_t per = (size_t)dp + 4;
    BO_TRY(bo_reserve(ctx, &ctx->dAppend, &ctx->append_capacity, (size_t)2 * np + S * per + d + 1));
    BO_TRY(bo_reserve(ctx, &ctx->dAppendInfo, &ctx->appendinfo_capacity, (size_t)S + S));
    double *kvec = ctx->dAppend, *lvec = kvec + np, *small = lvec + np;
    std::vector<double> h_small((size_t)S * per + d + 1, 0.0);
    std::vector<int> h_info(2 * S, 0);
    ctx->last_val_valid = false;
    for (int p = 0; p < m; ++p) {
        const int n = ctx->n;
        const double *x = Xnew + (size_t)p * d;
        for (int s = 0; s < S; ++s)
            for (int k = 0; k < d; ++k) h_small[s * per + k] = x[k] / ctx->h_ell[(size_t)s * d + k];
        for (int k = 0; k < d; ++k) h_small[(size_t)S * per + k] = x[k];
        h_small[(size_t)S * per + d] = ynew[p];
        BO_CUDA(ctx, cudaMemcpyAsync(small, h_small.data(), sizeof(double) * h_small.size(), cudaMemcpyHostToDevice, st));
        for (int s = 0; s < S; ++s) {
            const size_t mo = (size_t)s * np * np;
            double *xs = ctx->dXs + (size_t)s * np * dp, *sm = small + s * per;
            {
                BO_LAUNCH(ctx, "append_kvec_kernel");
                append_kvec_kernel<<<(np + 255) / 256, 256, 0, st>>>(ctx->kernel, n, np, dp, ctx->h_rho[s], xs, sm, kvec,
                                                                    s == 0 ? small + (size_t)S * per : nullptr, d, ctx->dX, ctx->dY);
                BO_CHECK_LAUNCH(ctx);
            }
            {
                BO_LAUNCH(ctx, "append_wk_kernel");
                append_wk_kernel<<<(n + 1) / 2 + 1, 32 * AP_ROWS, 0, st>>>(ctx->dW + mo, kvec, n, np, lvec);
                BO_CHECK_LAUNCH(ctx);
            }
            {
                BO_LAUNCH(ctx, "append_pivot_kernel");
                append_pivot_kernel<<<1, 1024, 0, st>>>(lvec, ctx->dAlpha + (size_t)s * np, ctx->dL + mo, ctx->dLogdet + s, n, np,
                                                        ctx->h_rho[s] + ctx->h_sn2[s], ynew[p] - ctx->h_bias[s], sm + dp,
                                                        ctx->dAppendInfo + s);
                BO_CHECK_LAUNCH(ctx);
            }
            {
                BO_LAUNCH(ctx, "append_wrow_kernel");
                append_wrow_kernel<<<(n + 1) / 2 + 1, 32 * AP_ROWS, 0, st>>>(
                    ctx->dW + mo, ctx->dWT + mo, lvec, sm + dp, ctx->dAppendInfo + s, ctx->dBeta + (size_t)s * np, n, np);
                BO_CHECK_LAUNCH(ctx);
            }
        }
        BO_CUDA(ctx, cudaMemcpyAsync(h_info.data(), ctx->dAppendInfo, sizeof(int) * S, cudaMemcpyDeviceToHost, st));
        BO_CUDA(ctx, cudaStreamSynchronize(st));
        for (int s = 0; s < S; ++s)
            if (h_info[s] != 0) {
                // samples before s already hold the new row; the handle is no longer consistent
                ctx->fitted = false;
                ctx->oz_ready = false;
                return bo_set_err(ctx, BO_ERR_NOT_PD, "append: hyper-sample %d, bordered matrix of order %d is not positive definite; refit",
                                  s, n + 1);
            }
        ctx->n = n + 1;
        if (ctx->oz_ready) BO_TRY(bo_ozaki_append_row(ctx, n));
    }
    return BO_OK;
}
