// score.cu -- the hot call `finit = f(xgrid, grad=False)` (reference
// solvers/lbfgs.py:50) and its gradient form (lbfgs.py:56-58), i.e. what
// model.predict / get_improvement / get_tail do for a batch of candidates
// (policies/simple.py:21,25,39,64):
//
//   K* = k(X, Xc)            kstar_kernel      (SIMT fp64, exp / Matern epilogue)
//   V  = W K*  (W = L^-1)    score_gemm_kernel (FP64 DMMA, lower-triangular K range,
//                                               fused column reductions |v|^2, v.alpha)
//   mu, s2 per hyper-sample  moments_kernel
//   U  = W^T V, dmu, ds2     dgemm_kernel + grad_partial_kernel (gradient path only)
//   acquisition + arg max    acq_kernel / argmax_final_kernel
//
// Candidates are processed in chunks so K* (np x chunk) stays a bounded scratch.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "dgemm.cuh"

typedef DTile<128, 128, 64, 32, 4, false> TS;   // scoring tile, B = K* stored [k][n]
#define GRAD_SLICES 16
#define S2_FLOOR 1e-300
#define INV_SQRT_2PI 0.3989422804014326779

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------
// K*[j][m] = k(x_j, xc_m); rows j >= n and columns m >= mc are written as 0.
// ---------------------------------------------------------------------------
#define KS_ROWS 32
template <int DP>
__global__ void __launch_bounds__(128)
kstar_kernel(int kernel, int n, int d, const double *__restrict__ Xs, const double *__restrict__ invell,
             double rho, const double *__restrict__ Xc, int64_t c0, int mc, int mcp,
             double *__restrict__ Ks) {
    __shared__ double xs[KS_ROWS][DP];
    const int tid = threadIdx.x;
    const int m = blockIdx.x * 128 + tid;
    const int j0 = blockIdx.y * KS_ROWS;
    for (int e = tid; e < KS_ROWS * DP; e += 128) xs[e / DP][e % DP] = Xs[(int64_t)j0 * DP + e];
    double xc[DP];
    const bool live = m < mc;
#pragma unroll
    for (int k = 0; k < DP; ++k) xc[k] = (live && k < d) ? Xc[(c0 + m) * d + k] * invell[k] : 0.0;
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < KS_ROWS; ++r) {
        double D = 0.0;
#pragma unroll
        for (int k = 0; k < DP; ++k) {
            double t = xc[k] - xs[r][k];
            D = fma(t, t, D);
        }
        double v;
        if (kernel == BO_KERNEL_SE) {
            v = rho * exp(-0.5 * D);
        } else {
            double rr = sqrt(5.0 * D);
            v = rho * (1.0 + rr + rr * rr * (1.0 / 3.0)) * exp(-rr);
        }
        if (!live || (j0 + r) >= n) v = 0.0;
        Ks[(int64_t)(j0 + r) * mcp + m] = v;
    }
}

// ---------------------------------------------------------------------------
// V = W K* on 128 x 128 tiles, k restricted to the lower triangle of W, with the
// column reductions q = sum_rows v^2 and p = sum_rows v alpha fused in.
// Tiles are issued heaviest (largest row block) first.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(TS::NTHREADS, 1)
score_gemm_kernel(const double *__restrict__ W, int np, const double *__restrict__ Ks, int mcp,
                  const double *__restrict__ alpha, double *__restrict__ qpart,
                  double *__restrict__ ppart, double *__restrict__ Vout) {
    extern __shared__ __align__(16) double smem[];
    __shared__ double red[2][2][128];
    typedef TS T;
    const int ctiles = mcp / 128, nblk = np / 128;
    const int rb = nblk - 1 - (int)(blockIdx.x / ctiles);
    const int ct = blockIdx.x % ctiles;
    const double *A = W + (int64_t)rb * 128 * np;
    const double *B = Ks + (int64_t)ct * 128;
    double acc[T::MI][T::NI][2];
#pragma unroll
    for (int mi = 0; mi < T::MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < T::NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    T::mainloop(acc, A, np, B, mcp, 0, (rb + 1) * 128, smem);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp / T::WARPS_N, wn = warp % T::WARPS_N;
    const int g = lane >> 2, t = lane & 3;
    const double *al = alpha + rb * 128 + wm * T::WM;
    double qs[T::NI][2], ps[T::NI][2];
#pragma unroll
    for (int ni = 0; ni < T::NI; ++ni) qs[ni][0] = qs[ni][1] = ps[ni][0] = ps[ni][1] = 0.0;
#pragma unroll
    for (int mi = 0; mi < T::MI; ++mi) {
        const double a_r = al[mi * 8 + g];
#pragma unroll
        for (int ni = 0; ni < T::NI; ++ni)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const double v = acc[mi][ni][e];
                qs[ni][e] = fma(v, v, qs[ni][e]);
                ps[ni][e] = fma(v, a_r, ps[ni][e]);
            }
    }
#pragma unroll
    for (int ni = 0; ni < T::NI; ++ni)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                qs[ni][e] += __shfl_xor_sync(0xffffffffu, qs[ni][e], o);
                ps[ni][e] += __shfl_xor_sync(0xffffffffu, ps[ni][e], o);
            }
            if (g == 0) {
                const int c = wn * T::WN + ni * 8 + 2 * t + e;
                red[wm][0][c] = qs[ni][e];
                red[wm][1][c] = ps[ni][e];
            }
        }
    __syncthreads();
    if (tid < 128) {
        const int64_t o = (int64_t)rb * mcp + (int64_t)ct * 128 + tid;
        qpart[o] = red[0][0][tid] + red[1][0][tid];
        ppart[o] = red[0][1][tid] + red[1][1][tid];
    }
    if (Vout != nullptr) {
        double *C = Vout + (int64_t)rb * 128 * mcp + (int64_t)ct * 128;
#pragma unroll
        for (int mi = 0; mi < T::MI; ++mi)
#pragma unroll
            for (int ni = 0; ni < T::NI; ++ni) {
                int r = wm * T::WM + mi * 8 + g, c = wn * T::WN + ni * 8 + 2 * t;
                *reinterpret_cast<double2 *>(&C[(int64_t)r * mcp + c]) =
                    make_double2(acc[mi][ni][0], acc[mi][ni][1]);
            }
    }
}

__global__ void moments_kernel(int nblk, int mcp, const double *__restrict__ qpart,
                               const double *__restrict__ ppart, double rho, double bias,
                               double *__restrict__ mu, double *__restrict__ s2) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= mcp) return;
    double q = 0.0, p = 0.0;
    for (int i = 0; i < nblk; ++i) {
        q += qpart[(int64_t)i * mcp + m];
        p += ppart[(int64_t)i * mcp + m];
    }
    mu[m] = bias + p;
    s2[m] = rho - q;
}

// ---------------------------------------------------------------------------
// Small batches (the L-BFGS callback shape f(x[None], grad=True), reference lbfgs.py:56-58):
// a tiled GEMM would walk the whole k range serially in one CTA, so M <= 16 candidates use
// bandwidth-bound triangular GEMVs instead: one warp per row of W (or W^T), MC right-hand sides.
//   UPPER == false: out[row][m] = sum_{j <= row} Mx[row][j] rhs[j][m]     (V = W K*)
//   UPPER == true : out[row][m] = sum_{j >= row} Mx[row][j] rhs[j][m]     (U = W^T V)
// ---------------------------------------------------------------------------
template <int MC, bool UPPER>
__global__ void __launch_bounds__(256)
trimv_small_kernel(const double *__restrict__ Mx, int np, const double *__restrict__ rhs, int ld,
                   double *__restrict__ out) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= np) return;
    const double *mrow = Mx + (int64_t)row * np;
    double acc[MC];
#pragma unroll
    for (int m = 0; m < MC; ++m) acc[m] = 0.0;
    const int jbeg = UPPER ? (row & ~31) : 0;
    const int jend = UPPER ? np : row + 1;
    for (int j = jbeg + lane; j < jend; j += 32) {
        if (UPPER && j < row) continue;
        const double w = mrow[j];
#pragma unroll
        for (int m = 0; m < MC; ++m) acc[m] = fma(w, rhs[(int64_t)j * ld + m], acc[m]);
    }
#pragma unroll
    for (int m = 0; m < MC; ++m) {
        acc[m] = warp_sum_d(acc[m]);
        if (lane == 0) out[(int64_t)row * ld + m] = acc[m];
    }
}

// Same products for 5..16 right-hand sides (the batched multi-start refinement, solvers.py
// `batched_lbfgs`): re-reading MC values of rhs per matrix element from L2 made the call slower
// than the tiled GEMM, so the block stages 128 rows of rhs in shared memory per step and the
// eight warps (eight consecutive rows) share them; the matrix is still read exactly once.
template <int MC, bool UPPER>
__global__ void __launch_bounds__(256)
trimv_multi_kernel(const double *__restrict__ Mx, int np, const double *__restrict__ rhs, int ld,
                   double *__restrict__ out) {
    __shared__ double sr[128][MC + 1];
    const int tid = threadIdx.x, lane = tid & 31;
    const int r0 = blockIdx.x * 8, row = r0 + (tid >> 5);
    const double *mrow = Mx + (int64_t)row * np;
    double acc[MC];
#pragma unroll
    for (int m = 0; m < MC; ++m) acc[m] = 0.0;
    const int jb = UPPER ? (r0 & ~127) : 0;
    const int je = UPPER ? np : r0 + 8;
    for (int j0 = jb; j0 < je; j0 += 128) {
        for (int e = tid; e < 128 * MC; e += 256) sr[e / MC][e % MC] = rhs[(int64_t)(j0 + e / MC) * ld + (e % MC)];
        __syncthreads();
#pragma unroll
        for (int jj = lane; jj < 128; jj += 32) {
            const int j = j0 + jj;
            const bool in = UPPER ? (j >= row) : (j <= row);
            const double w = in ? mrow[j] : 0.0;
#pragma unroll
            for (int m = 0; m < MC; ++m) acc[m] = fma(w, sr[jj][m], acc[m]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int m = 0; m < MC; ++m) {
        acc[m] = warp_sum_d(acc[m]);
        if (lane == 0) out[(int64_t)row * ld + m] = acc[m];
    }
}

// mu[m] = bias + sum_i V[i][m] alpha[i], s2[m] = rho - sum_i V[i][m]^2 ; one block per candidate
__global__ void __launch_bounds__(256)
small_moments_kernel(const double *__restrict__ V, int np, int ld, const double *__restrict__ alpha, double rho,
                     double bias, double *__restrict__ mu, double *__restrict__ s2) {
    __shared__ double rq[8], rp[8];
    const int m = blockIdx.x;
    double q = 0.0, p = 0.0;
    for (int i = threadIdx.x; i < np; i += 256) {
        const double v = V[(int64_t)i * ld + m];
        q = fma(v, v, q);
        p = fma(v, alpha[i], p);
    }
    q = warp_sum_d(q);
    p = warp_sum_d(p);
    if ((threadIdx.x & 31) == 0) { rq[threadIdx.x >> 5] = q; rp[threadIdx.x >> 5] = p; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { q += rq[w]; p += rp[w]; }
        mu[m] = bias + p;
        s2[m] = rho - q;
    }
}

// ---------------------------------------------------------------------------
// gradient partial sums: one warp per candidate, lanes stride the observations.
//   gm_k = sum_j g_j beta_j (xc_k - x_jk),  gs_k = sum_j g_j U_jm (xc_k - x_jk)
// (scaled coordinates), g = -dk/dD * 2:  SE: k,  Matern-5/2: rho 5/3 (1+r) e^-r.
// ---------------------------------------------------------------------------
template <int DP>
__global__ void __launch_bounds__(256)
grad_partial_kernel(int kernel, int n, int np, int d, const double *__restrict__ Xs,
                    const double *__restrict__ invell, double rho, const double *__restrict__ Xc,
                    int64_t c0, int mc, int mcp, const double *__restrict__ U,
                    const double *__restrict__ beta, double *__restrict__ gpart) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m = blockIdx.x * 8 + warp;
    if (m >= mc) return;
    const int rps = np / GRAD_SLICES, j0 = blockIdx.y * rps;
    double xc[DP], gm[DP], gs[DP];
#pragma unroll
    for (int k = 0; k < DP; ++k) {
        xc[k] = (k < d) ? Xc[(c0 + m) * d + k] * invell[k] : 0.0;
        gm[k] = gs[k] = 0.0;
    }
    for (int j = j0 + lane; j < j0 + rps && j < n; j += 32) {
        double diff[DP], D = 0.0;
#pragma unroll
        for (int k = 0; k < DP; ++k) {
            diff[k] = xc[k] - Xs[(int64_t)j * DP + k];
            D = fma(diff[k], diff[k], D);
        }
        double gk;
        if (kernel == BO_KERNEL_SE) {
            gk = rho * exp(-0.5 * D);
        } else {
            double rr = sqrt(5.0 * D);
            gk = rho * (5.0 / 3.0) * (1.0 + rr) * exp(-rr);
        }
        const double gb = gk * beta[j], gu = gk * U[(int64_t)j * mcp + m];
#pragma unroll
        for (int k = 0; k < DP; ++k) {
            gm[k] = fma(gb, diff[k], gm[k]);
            gs[k] = fma(gu, diff[k], gs[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < DP; ++k) {
        gm[k] = warp_sum_d(gm[k]);
        gs[k] = warp_sum_d(gs[k]);
    }
    if (lane == 0) {
        double *o = gpart + ((int64_t)blockIdx.y * mcp + m) * 2 * DP;
#pragma unroll
        for (int k = 0; k < DP; ++k) {
            o[k] = gm[k];
            o[DP + k] = gs[k];
        }
    }
}

__global__ void grad_finish_kernel(int dp, int d, int mc, int mcp, const double *__restrict__ gpart,
                                   const double *__restrict__ invell, double *__restrict__ dmu,
                                   double *__restrict__ ds2) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= mc * d) return;
    const int m = e / d, k = e % d;
    double a = 0.0, b = 0.0;
    for (int sl = 0; sl < GRAD_SLICES; ++sl) {
        const double *o = gpart + ((int64_t)sl * mcp + m) * 2 * dp;
        a += o[k];
        b += o[dp + k];
    }
    dmu[e] = -invell[k] * a;
    ds2[e] = 2.0 * invell[k] * b;
}

// ---------------------------------------------------------------------------
// acquisition epilogue over the S hyper-samples + per-block (max, first argmax)
// ---------------------------------------------------------------------------
struct AcqParams {
    int mode;          // 0: acquisition value (+grad), 1: predict moments (+grad)
    int acq;
    double param;
    int S, d, mc, mcp;
    int64_t c0;
    const double *muS, *s2S, *dmuS, *ds2S;   // S x mcp, S x mcp x d
    double *out_val, *out_grad;              // global arrays indexed by c0 + m
    double *out_mu, *out_s2, *out_dmu, *out_ds2;
    double *blkval;
    int64_t *blkidx;
    int64_t blk0;
    // int8-slice path only: a-priori bound on the absolute error of s2_s is  errK[s] * sqrt(q_s rho_s),
    // q_s = rho_s - s2_s (ozaki.cu, bo_ozaki_error_scale); errest[c0 + m] receives the first-order bound on the
    // absolute error of this candidate's output (mode 0: the acquisition value, mode 1: s2), +inf where the
    // first-order estimate itself is not trustworthy (s2_s < 16 x its own error bound)
    const double *rhoS, *errK;
    double *errest;
};

__device__ __forceinline__ bool better(double v, int64_t i, double bv, int64_t bi) {
    return (v > bv) || (v == bv && i < bi);
}

__global__ void __launch_bounds__(256) acq_kernel(AcqParams p) {
    const int m = blockIdx.x * 256 + threadIdx.x;
    const bool live = m < p.mc;
    double val = -INFINITY;
    if (live) {
        const int S = p.S, d = p.d;
        const double invS = 1.0 / S;
        const bool want_grad = (p.mode == 0) ? (p.out_grad != nullptr) : (p.out_dmu != nullptr || p.out_ds2 != nullptr);
        const int64_t gi = p.c0 + m;
        double eacc = 0.0;                  // error bound accumulator (int8-slice path)
        if (p.mode == 0 && (p.acq == BO_ACQ_EI || p.acq == BO_ACQ_PI)) {
            // mean over hyper-samples of the per-sample EI / PI at the common target
            double acc = 0.0;
            for (int s = 0; s < S; ++s) {
                const double mu = p.muS[(int64_t)s * p.mcp + m];
                const double s2r = p.s2S[(int64_t)s * p.mcp + m];
                const double s2 = fmax(s2r, S2_FLOOR);
                const double sd = sqrt(s2), dl = mu - p.param, z = dl / sd;
                const double cdf = 0.5 * erfc(-z * M_SQRT1_2);
                const double pdf = INV_SQRT_2PI * exp(-0.5 * z * z);
                acc += (p.acq == BO_ACQ_EI) ? (dl * cdf + sd * pdf) : cdf;
                if (p.errest) {
                    // |dEI/ds2| = pdf / (2 sd),  |dPI/ds2| = pdf |z| / (2 s2)
                    const double E = p.errK[s] * sqrt(fmax(p.rhoS[s] - s2r, 0.0) * p.rhoS[s]);
                    const double sens = (p.acq == BO_ACQ_EI) ? 0.5 * pdf / sd : 0.5 * pdf * fabs(z) / s2;
                    eacc += (s2r < 16.0 * E) ? INFINITY : sens * E;
                }
            }
            val = acc * invS;
            if (want_grad) {
                for (int k = 0; k < d; ++k) {
                    double ga = 0.0;
                    for (int s = 0; s < S; ++s) {
                        const double mu = p.muS[(int64_t)s * p.mcp + m];
                        const double s2 = fmax(p.s2S[(int64_t)s * p.mcp + m], S2_FLOOR);
                        const double sd = sqrt(s2), z = (mu - p.param) / sd;
                        const double cdf = 0.5 * erfc(-z * M_SQRT1_2);
                        const double pdf = INV_SQRT_2PI * exp(-0.5 * z * z);
                        const double dm = p.dmuS[((int64_t)s * p.mcp + m) * d + k];
                        const double dv = p.ds2S[((int64_t)s * p.mcp + m) * d + k];
                        ga += (p.acq == BO_ACQ_EI) ? (cdf * dm + (0.5 * pdf / sd) * dv)
                                                   : ((pdf / sd) * (dm - (0.5 * z / sd) * dv));
                    }
                    p.out_grad[gi * d + k] = ga * invS;
                }
            }
        } else {
            // mixture moments: mu = mean mu_s, s2 = mean (s2_s + (mu_s - mu)^2)
            double mub = 0.0;
            for (int s = 0; s < S; ++s) mub += p.muS[(int64_t)s * p.mcp + m];
            mub *= invS;
            double s2b = 0.0;
            for (int s = 0; s < S; ++s) {
                const double dm = p.muS[(int64_t)s * p.mcp + m] - mub;
                const double s2r = p.s2S[(int64_t)s * p.mcp + m];
                s2b += s2r + dm * dm;
                if (p.errest) {
                    const double E = p.errK[s] * sqrt(fmax(p.rhoS[s] - s2r, 0.0) * p.rhoS[s]);
                    eacc += (p.mode == 0 && p.acq == BO_ACQ_UCB && s2r < 16.0 * E) ? INFINITY : E;
                }
            }
            s2b *= invS;
            // rho - |v|^2 can round below zero next to an observation (sn2 ~ 1e-6 rho): the variance handed back
            // is clamped at 0 and the UCB square roots see a tiny positive floor instead of NaN / Inf
            const double s2c = fmax(s2b, S2_FLOOR);
            if (p.mode == 1) {
                if (p.out_mu) p.out_mu[gi] = mub;
                if (p.out_s2) p.out_s2[gi] = fmax(s2b, 0.0);
                val = mub;
            } else if (p.acq == BO_ACQ_MEAN) {
                val = mub;
                eacc = 0.0;                                  // the mean never goes through the int8 contraction
            } else {   // UCB, reference policies/simple.py:72
                val = mub + sqrt(p.param * s2c);
                eacc *= 0.5 * sqrt(p.param / s2c);           // |dUCB/ds2b| x error of s2b (invS applied below)
            }
            if (want_grad) {
                for (int k = 0; k < d; ++k) {
                    double dmb = 0.0;
                    for (int s = 0; s < S; ++s) dmb += p.dmuS[((int64_t)s * p.mcp + m) * d + k];
                    dmb *= invS;
                    double dsb = 0.0;
                    for (int s = 0; s < S; ++s) {
                        const double dm = p.muS[(int64_t)s * p.mcp + m] - mub;
                        dsb += p.ds2S[((int64_t)s * p.mcp + m) * d + k] +
                               2.0 * dm * (p.dmuS[((int64_t)s * p.mcp + m) * d + k] - dmb);
                    }
                    dsb *= invS;
                    if (p.mode == 1) {
                        if (p.out_dmu) p.out_dmu[gi * d + k] = dmb;
                        if (p.out_ds2) p.out_ds2[gi * d + k] = dsb;
                    } else if (p.acq == BO_ACQ_MEAN) {
                        p.out_grad[gi * d + k] = dmb;
                    } else {   // simple.py:69-70
                        p.out_grad[gi * d + k] = dmb + 0.5 * sqrt(p.param / s2c) * dsb;
                    }
                }
            }
        }
        if (p.mode == 0 && p.out_val) p.out_val[gi] = val;
        if (p.errest) p.errest[gi] = eacc * invS;
    }
    if (p.blkval == nullptr) return;
    // block arg max, ties to the lowest index, NaN never wins
    __shared__ double sv[8];
    __shared__ int64_t si[8];
    double bv = live ? val : -INFINITY;
    int64_t bi = live ? (p.c0 + m) : INT64_MAX;
    if (bv != bv) { bv = -INFINITY; bi = INT64_MAX; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w)
            if (better(sv[w], si[w], bv, bi)) { bv = sv[w]; bi = si[w]; }
        p.blkval[p.blk0 + blockIdx.x] = bv;
        p.blkidx[p.blk0 + blockIdx.x] = bi;
    }
}

// single block: reduce (val, idx) pairs; writes result[0] = val, ridx[0] = idx
__global__ void __launch_bounds__(1024)
argmax_final_kernel(const double *__restrict__ bval, const int64_t *__restrict__ bidx, int64_t nb,
                    double *__restrict__ rval, int64_t *__restrict__ ridx, int64_t *__restrict__ rec, int64_t rec_offset) {
    __shared__ double sv[32];
    __shared__ int64_t si[32];
    double bv = -INFINITY;
    int64_t bi = INT64_MAX;
    for (int64_t i = threadIdx.x; i < nb; i += 1024)
        if (better(bval[i], bidx[i], bv, bi)) { bv = bval[i]; bi = bidx[i]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 32; ++w)
            if (better(sv[w], si[w], bv, bi)) { bv = sv[w]; bi = si[w]; }
        rval[0] = bv;
        ridx[0] = bi;
        // cross-rank exchange: the packed {value bits, global index} record is written by the same thread, so the
        // collective can follow this kernel directly on the stream (bo_score_incumbent)
        if (rec != nullptr) {
            const bool none = (bi == INT64_MAX) || (bv != bv);
            rec[0] = __double_as_longlong(none ? -INFINITY : bv);
            rec[1] = none ? INT64_MAX : bi + rec_offset;
        }
    }
}

// top-k pass: arg max over entries strictly after (prev_val, prev_idx) in the
// (value descending, index ascending) order.
__global__ void __launch_bounds__(256)
topk_pass_kernel(const double *__restrict__ vals, int64_t M, const double *__restrict__ prev_val,
                 const int64_t *__restrict__ prev_idx, int first, double *__restrict__ bval,
                 int64_t *__restrict__ bidx) {
    __shared__ double sv[8];
    __shared__ int64_t si[8];
    const double pv = first ? INFINITY : prev_val[0];
    const int64_t pi = first ? -1 : prev_idx[0];
    double bv = -INFINITY;
    int64_t bi = INT64_MAX;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < M; i += (int64_t)gridDim.x * 256) {
        const double v = vals[i];
        if (v != v) continue;
        const bool eligible = first || (v < pv) || (v == pv && i > pi);
        if (eligible && better(v, i, bv, bi)) { bv = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w)
            if (better(sv[w], si[w], bv, bi)) { bv = sv[w]; bi = si[w]; }
        bval[blockIdx.x] = bv;
        bidx[blockIdx.x] = bi;
    }
}

// ---------------------------------------------------------------------------
// rescue pass of the int8-slice path (kernels): candidates whose a-priori error bound exceeds the
// acquisition tolerance are compacted, re-scored on the FP64 path and scattered back.
// ---------------------------------------------------------------------------
// flagged <=> !(errest <= tol * max(|x|, floor)), floor = floor_abs + floor_rel * |*gmax| (NaN flags too)
__global__ void oz_flag_kernel(int64_t M, const double *__restrict__ x, const double *__restrict__ errest, double tol,
                               double floor_abs, double floor_rel, const double *__restrict__ gmax,
                               int *__restrict__ list, int *__restrict__ count) {
    const double gm = gmax ? fabs(gmax[0]) : 0.0;
    const double floor = floor_abs + ((gm == gm && gm < INFINITY) ? floor_rel * gm : 0.0);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x) {
        const double e = errest[i];
        if (e == 0.0) continue;
        const double thr = tol * fmax(fabs(x[i]), floor);
        if (!(e <= thr)) list[atomicAdd(count, 1)] = (int)i;
    }
}

// the flagged list is sorted so the compact batch (and with it every FP64 tile) does not depend on the
// order in which the atomics landed: bitonic sort in one block for small lists, else a rank-by-count pass
__global__ void oz_sort_small_kernel(int *__restrict__ list, int count) {
    extern __shared__ int sl[];
    int np2 = 1;
    while (np2 < count) np2 <<= 1;
    for (int i = threadIdx.x; i < np2; i += blockDim.x) sl[i] = i < count ? list[i] : 0x7fffffff;
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < np2; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const int a = sl[i], b = sl[l];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { sl[i] = b; sl[l] = a; }
                }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < count; i += blockDim.x) list[i] = sl[i];
}

// large lists: mark flags in a bitmap, then compact in index order (block prefix over 64-bit words)
__global__ void oz_mark_kernel(const int *__restrict__ list, int count, unsigned long long *__restrict__ bits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) atomicOr(bits + (list[i] >> 6), 1ull << (list[i] & 63));
}
__global__ void oz_wordcount_kernel(const unsigned long long *__restrict__ bits, int nwords, int *__restrict__ blocksum) {
    __shared__ int red[32];
    const int w = blockIdx.x * 1024 + threadIdx.x;
    int c = w < nwords ? __popcll(bits[w]) : 0;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x < 32) {
        int v = red[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) blocksum[blockIdx.x] = v;
    }
}
__global__ void oz_compact_kernel(const unsigned long long *__restrict__ bits, int nwords, const int *__restrict__ blocksum,
                                  int *__restrict__ list) {
    __shared__ int wsum[32];
    __shared__ int base_s;
    const int w = blockIdx.x * 1024 + threadIdx.x;
    const unsigned long long word = w < nwords ? bits[w] : 0ull;
    const int c = __popcll(word);
    if (threadIdx.x == 0) {
        int b = 0;
        for (int i = 0; i < (int)blockIdx.x; ++i) b += blocksum[i];
        base_s = b;
    }
    // inclusive scan of c over the block
    int v = c;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    if (lane == 31) wsum[wid] = v;
    __syncthreads();
    if (wid == 0) {
        int t = wsum[lane];
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, t, o);
            if (lane >= o) t += u;
        }
        wsum[lane] = t;
    }
    __syncthreads();
    int pos = base_s + v - c + (wid ? wsum[wid - 1] : 0);
    unsigned long long rem = word;
    while (rem) {
        const int b = __ffsll((long long)rem) - 1;
        list[pos++] = w * 64 + b;
        rem &= rem - 1;
    }
}

__global__ void oz_gather_rows_kernel(const double *__restrict__ Xc, int d, const int *__restrict__ list, int count,
                                      double *__restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)count * d) return;
    const int i = (int)(e / d), k = (int)(e - (int64_t)i * d);
    out[e] = Xc[(int64_t)list[i] * d + k];
}

__global__ void oz_scatter_kernel(const int *__restrict__ list, int count, const double *__restrict__ a,
                                  double *__restrict__ outa, const double *__restrict__ b, double *__restrict__ outb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int g = list[i];
    if (outa) outa[g] = a[i];
    if (outb) outb[g] = b[i];
}

// ---------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------
int bo_score_init(bo_ctx *ctx) {
    BO_CUDA(ctx, cudaFuncSetAttribute(score_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TS::SMEM_BYTES));
    BO_CUDA(ctx, cudaFuncSetAttribute(dgemm_kernel<TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS::SMEM_BYTES));
    return BO_OK;
}

template <int DP>
static int launch_kstar(bo_ctx *ctx, int s, const double *dXc, int64_t c0, int mc, int mcp) {
    BO_LAUNCH(ctx, "kstar_kernel");
    kstar_kernel<DP><<<dim3(mcp / 128, ctx->np / KS_ROWS), 128, 0, ctx->stream>>>(
        ctx->kernel, ctx->n, ctx->d, ctx->dXs + (int64_t)s * ctx->np * ctx->dp, ctx->dInvEll + (int64_t)s * ctx->dp,
        ctx->h_rho[s], dXc, c0, mc, mcp, ctx->dKs);
    BO_CHECK_LAUNCH(ctx);
    return BO_OK;
}

template <int DP>
static int launch_grad_partial(bo_ctx *ctx, int s, const double *dXc, int64_t c0, int mc, int mcp) {
    BO_LAUNCH(ctx, "grad_partial_kernel");
    grad_partial_kernel<DP><<<dim3((mc + 7) / 8, GRAD_SLICES), 256, 0, ctx->stream>>>(
        ctx->kernel, ctx->n, ctx->np, ctx->d, ctx->dXs + (int64_t)s * ctx->np * ctx->dp,
        ctx->dInvEll + (int64_t)s * ctx->dp, ctx->h_rho[s], dXc, c0, mc, mcp, ctx->dU,
        ctx->dBeta + (int64_t)s * ctx->np, ctx->dGpart);
    BO_CHECK_LAUNCH(ctx);
    return BO_OK;
}

#define DISPATCH_DP(ctx, fn, ...)                                             \
    do {                                                                      \
        switch ((ctx)->dp) {                                                  \
            case 2: BO_TRY(fn<2>(__VA_ARGS__)); break;                        \
            case 4: BO_TRY(fn<4>(__VA_ARGS__)); break;                        \
            case 8: BO_TRY(fn<8>(__VA_ARGS__)); break;                        \
            case 16: BO_TRY(fn<16>(__VA_ARGS__)); break;                      \
            case 32: BO_TRY(fn<32>(__VA_ARGS__)); break;                      \
            default: return bo_set_err(ctx, BO_ERR_ARG, "unsupported padded dimension %d", (ctx)->dp); \
        }                                                                     \
    } while (0)

// gradient / small-batch scratch: every buffer has its own capacity (a refit with another n, S or d changes
// their sizes independently of np * cap)
static int reserve_v(bo_ctx *ctx, size_t n_v) {
    BO_TRY(bo_reserve(ctx, &ctx->dV, &ctx->v_capacity, n_v));
    return BO_OK;
}
static int reserve_grad(bo_ctx *ctx, size_t n_v, size_t cap, int S, int d, int dp) {
    BO_TRY(bo_reserve(ctx, &ctx->dV, &ctx->v_capacity, n_v));
    BO_TRY(bo_reserve(ctx, &ctx->dU, &ctx->u_capacity, n_v));
    BO_TRY(bo_reserve(ctx, &ctx->dGpart, &ctx->gpart_capacity, (size_t)GRAD_SLICES * cap * 2 * dp));
    BO_TRY(bo_reserve(ctx, &ctx->dDmuS, &ctx->dmus_capacity, (size_t)S * cap * d));
    BO_TRY(bo_reserve(ctx, &ctx->dDs2S, &ctx->ds2s_capacity, (size_t)S * cap * d));
    return BO_OK;
}

static int reserve_moments(bo_ctx *ctx, int nblk, int64_t cap, int S) {
    const size_t need = (size_t)nblk * cap;
    if (ctx->mom_capacity < need || !ctx->dQpart) {
        size_t c1 = ctx->mom_capacity, c2 = ctx->mom_capacity;
        BO_TRY(bo_reserve(ctx, &ctx->dQpart, &c1, need));
        BO_TRY(bo_reserve(ctx, &ctx->dPpart, &c2, need));
        ctx->mom_capacity = need;
    }
    const size_t needm = (size_t)S * cap;
    if (ctx->gm_capacity < needm || !ctx->dMuS) {
        size_t c1 = ctx->gm_capacity, c2 = ctx->gm_capacity;
        BO_TRY(bo_reserve(ctx, &ctx->dMuS, &c1, needm));
        BO_TRY(bo_reserve(ctx, &ctx->dS2S, &c2, needm));
        ctx->gm_capacity = needm;
    }
    return BO_OK;
}

static int reserve_blocks(bo_ctx *ctx, size_t need) {
    if (ctx->blk_capacity < need || !ctx->dBlkVal) {
        size_t c1 = ctx->blk_capacity, c2 = ctx->blk_capacity;
        BO_TRY(bo_reserve(ctx, &ctx->dBlkVal, &c1, need));
        BO_TRY(bo_reserve(ctx, &ctx->dBlkIdx, &c2, need));
        ctx->blk_capacity = need;
    }
    return BO_OK;
}

#define ARGMAX_PASS_BLOCKS 592

// (max, first arg max) of a device array into the tail record dBlkVal/dBlkIdx[blk_capacity - 1]
static int final_argmax_over(bo_ctx *ctx, const double *vals, int64_t M) {
    const int nb = (int)((M + 255) / 256 < ARGMAX_PASS_BLOCKS ? (M + 255) / 256 : ARGMAX_PASS_BLOCKS);
    {
        BO_LAUNCH(ctx, "topk_pass_kernel");
        topk_pass_kernel<<<nb, 256, 0, ctx->stream>>>(vals, M, nullptr, nullptr, 1, ctx->dBlkVal, ctx->dBlkIdx);
        BO_CHECK_LAUNCH(ctx);
    }
    BO_LAUNCH(ctx, "argmax_final_kernel");
    argmax_final_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->dBlkVal, ctx->dBlkIdx, nb, ctx->dBlkVal + ctx->blk_capacity - 1,
                                                    ctx->dBlkIdx + ctx->blk_capacity - 1, ctx->rec_ptr, ctx->rec_offset);
    BO_CHECK_LAUNCH(ctx);
    return BO_OK;
}

// FP64 path: K* (SIMT) -> V = W K* (DMMA) -> moments -> [gradients] -> acquisition, chunk by chunk
static int run_fp64(bo_ctx *ctx, const ScoreRequest &rq) {
    const int np = ctx->np, S = ctx->S, d = ctx->d, dp = ctx->dp, nblk = np / 128;
    const bool grad = (rq.mode == 0) ? (rq.dGrad != nullptr) : (rq.dDmu != nullptr || rq.dDs2 != nullptr);
    const int64_t M = rq.M;
    const int64_t chunk = ctx->chunk;
    const int64_t cap = bo_round_up64(M < chunk ? M : chunk, 128);
    BO_TRY(bo_reserve(ctx, &ctx->dKs, &ctx->ks_capacity, (size_t)np * cap));
    BO_TRY(reserve_moments(ctx, nblk, cap, S));
    if (grad) BO_TRY(reserve_grad(ctx, (size_t)np * cap, (size_t)cap, S, d, dp));
    else if (M <= 16) BO_TRY(reserve_v(ctx, (size_t)np * cap));
    const int64_t nchunks = (M + chunk - 1) / chunk;
    const int64_t nblocks_total = nchunks * ((chunk + 255) / 256);
    if (rq.want_best) BO_TRY(reserve_blocks(ctx, (size_t)(nblocks_total > ARGMAX_PASS_BLOCKS ? nblocks_total : ARGMAX_PASS_BLOCKS) + 8));

    int64_t blk0 = 0;
    for (int64_t c0 = 0; c0 < M; c0 += chunk) {
        const int mc = (int)((M - c0) < chunk ? (M - c0) : chunk);
        const int mcp = bo_round_up(mc, 128);
        for (int s = 0; s < S; ++s) {
            DISPATCH_DP(ctx, launch_kstar, ctx, s, rq.dXc, c0, mc, mcp);
            if (mc <= 16) {
                // small batch: triangular GEMVs (V, and U = W^T V for gradients) instead of tiled GEMMs
                const int MC = mc <= 1 ? 1 : (mc <= 2 ? 2 : (mc <= 4 ? 4 : (mc <= 8 ? 8 : 16)));
                const double *Wm = ctx->dW + (int64_t)s * np * np, *WTm = ctx->dWT + (int64_t)s * np * np;
                {
                    BO_LAUNCH(ctx, "trimv_small_kernel");
                    switch (MC) {
                        case 1: trimv_small_kernel<1, false><<<np / 8, 256, 0, ctx->stream>>>(Wm, np, ctx->dKs, mcp, ctx->dV); break;
                        case 2: trimv_small_kernel<2, false><<<np / 8, 256, 0, ctx->stream>>>(Wm, np, ctx->dKs, mcp, ctx->dV); break;
                        case 4: trimv_small_kernel<4, false><<<np / 8, 256, 0, ctx->stream>>>(Wm, np, ctx->dKs, mcp, ctx->dV); break;
                        case 8: trimv_multi_kernel<8, false><<<np / 8, 256, 0, ctx->stream>>>(Wm, np, ctx->dKs, mcp, ctx->dV); break;
                        default: trimv_multi_kernel<16, false><<<np / 8, 256, 0, ctx->stream>>>(Wm, np, ctx->dKs, mcp, ctx->dV); break;
                    }
                    BO_CHECK_LAUNCH(ctx);
                }
                {
                    BO_LAUNCH(ctx, "small_moments_kernel");
                    small_moments_kernel<<<mc, 256, 0, ctx->stream>>>(ctx->dV, np, mcp, ctx->dAlpha + (int64_t)s * np,
                                                                    ctx->h_rho[s], ctx->h_bias[s],
                                                                    ctx->dMuS + (int64_t)s * mcp, ctx->dS2S + (int64_t)s * mcp);
                    BO_CHECK_LAUNCH(ctx);
                }
                if (grad) {
                    {
                        BO_LAUNCH(ctx, "trimv_small_kernel");
                        switch (MC) {
                            case 1: trimv_small_kernel<1, true><<<np / 8, 256, 0, ctx->stream>>>(WTm, np, ctx->dV, mcp, ctx->dU); break;
                            case 2: trimv_small_kernel<2, true><<<np / 8, 256, 0, ctx->stream>>>(WTm, np, ctx->dV, mcp, ctx->dU); break;
                            case 4: trimv_small_kernel<4, true><<<np / 8, 256, 0, ctx->stream>>>(WTm, np, ctx->dV, mcp, ctx->dU); break;
                            case 8: trimv_multi_kernel<8, true><<<np / 8, 256, 0, ctx->stream>>>(WTm, np, ctx->dV, mcp, ctx->dU); break;
                            default: trimv_multi_kernel<16, true><<<np / 8, 256, 0, ctx->stream>>>(WTm, np, ctx->dV, mcp, ctx->dU); break;
                        }
                        BO_CHECK_LAUNCH(ctx);
                    }
                    DISPATCH_DP(ctx, launch_grad_partial, ctx, s, rq.dXc, c0, mc, mcp);
                    {
                        BO_LAUNCH(ctx, "grad_finish_kernel");
                        grad_finish_kernel<<<(mc * d + 255) / 256, 256, 0, ctx->stream>>>(
                            dp, d, mc, mcp, ctx->dGpart, ctx->dInvEll + (int64_t)s * dp,
                            ctx->dDmuS + (int64_t)s * mcp * d, ctx->dDs2S + (int64_t)s * mcp * d);
                        BO_CHECK_LAUNCH(ctx);
                    }
                }
                continue;
            }
            {
                BO_LAUNCH(ctx, "score_gemm_kernel");
                score_gemm_kernel<<<nblk * (mcp / 128), TS::NTHREADS, TS::SMEM_BYTES, ctx->stream>>>(
                    ctx->dW + (int64_t)s * np * np, np, ctx->dKs, mcp, ctx->dAlpha + (int64_t)s * np,
                    ctx->dQpart, ctx->dPpart, grad ? ctx->dV : nullptr);
                BO_CHECK_LAUNCH(ctx);
            }
            {
                BO_LAUNCH(ctx, "moments_kernel");
                moments_kernel<<<(mcp + 255) / 256, 256, 0, ctx->stream>>>(
                    nblk, mcp, ctx->dQpart, ctx->dPpart, ctx->h_rho[s], ctx->h_bias[s],
                    ctx->dMuS + (int64_t)s * mcp, ctx->dS2S + (int64_t)s * mcp);
                BO_CHECK_LAUNCH(ctx);
            }
            if (grad) {
                {   // U = W^T V   (W^T upper triangular: k >= row block)
                    DGemmParams p = {};
                    p.A = ctx->dWT + (int64_t)s * np * np; p.lda = np;
                    p.B = ctx->dV; p.ldb = mcp;
                    p.C = ctx->dU; p.ldc = mcp;
                    p.inner = 1; p.tiles_m = nblk; p.tiles_n = mcp / 128; p.K = np;
                    p.krule = KR_A_UPPER; p.alpha = 1.0; p.beta = 0.0;
                    BO_LAUNCH(ctx, "grad_gemm_kernel");
                    dgemm_kernel<TS><<<dim3(p.tiles_m * p.tiles_n, 1, 1), TS::NTHREADS, TS::SMEM_BYTES, ctx->stream>>>(p);
                    BO_CHECK_LAUNCH(ctx);
                }
                DISPATCH_DP(ctx, launch_grad_partial, ctx, s, rq.dXc, c0, mc, mcp);
                {
                    BO_LAUNCH(ctx, "grad_finish_kernel");
                    grad_finish_kernel<<<(mc * d + 255) / 256, 256, 0, ctx->stream>>>(
                        dp, d, mc, mcp, ctx->dGpart, ctx->dInvEll + (int64_t)s * dp,
                        ctx->dDmuS + (int64_t)s * mcp * d, ctx->dDs2S + (int64_t)s * mcp * d);
                    BO_CHECK_LAUNCH(ctx);
                }
            }
        }
        AcqParams ap = {};
        ap.mode = rq.mode; ap.acq = rq.acq; ap.param = rq.param;
        ap.S = S; ap.d = d; ap.mc = mc; ap.mcp = mcp; ap.c0 = c0;
        ap.muS = ctx->dMuS; ap.s2S = ctx->dS2S; ap.dmuS = ctx->dDmuS; ap.ds2S = ctx->dDs2S;
        ap.out_val = rq.dVal; ap.out_grad = rq.dGrad;
        ap.out_mu = rq.dMu; ap.out_s2 = rq.dS2; ap.out_dmu = rq.dDmu; ap.out_ds2 = rq.dDs2;
        ap.blkval = rq.want_best ? ctx->dBlkVal : nullptr;
        ap.blkidx = rq.want_best ? ctx->dBlkIdx : nullptr;
        ap.blk0 = blk0;
        const int nb = (mc + 255) / 256;
        {
            BO_LAUNCH(ctx, "acq_kernel");
            acq_kernel<<<nb, 256, 0, ctx->stream>>>(ap);
            BO_CHECK_LAUNCH(ctx);
        }
        blk0 += nb;
    }
    if (rq.want_best) {
        BO_LAUNCH(ctx, "argmax_final_kernel");
        argmax_final_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->dBlkVal, ctx->dBlkIdx, blk0,
                                                        ctx->dBlkVal + ctx->blk_capacity - 1,
                                                        ctx->dBlkIdx + ctx->blk_capacity - 1, ctx->rec_ptr, ctx->rec_offset);
        BO_CHECK_LAUNCH(ctx);
    }
    return BO_OK;
}

// int8-slice path: slices -> tcgen05 contraction -> moments -> acquisition + error bound, chunk by chunk;
// then the FP64 rescue of the candidates whose bound exceeds the tolerance.
// Precision levels of the int8-slice path are coded 2 S + extra (S slices per operand, `extra`: the digit pairs of
// group g = S are accumulated too); one step up divides the error bound by 8 (extra on) or 32 (S + 1, extra off).
static inline int lv_S(int L) { return L >> 1; }
static inline bool lv_extra(int L) { return (L & 1) != 0; }

#define OZ_TIER_MIN_CHUNKS 8
#define OZ_PILOT_CANDIDATES 4096     // length of the pilot chunk of a tiered pass

struct OzLevels {
    int first;      // level of candidate chunk 0
    int rest;       // level of the other chunks; if first < rest the pass may lower it to `first` after chunk 0 (see run_oz)
    int tier2;      // level the flagged candidates are re-scored at before the FP64 path takes what is left (0: none) ...
    int sel;        // ... unless the whole pass ran below the level the tolerance selects: then that level is the next tier
};

static int64_t oz_chunk_candidates() {
    static const int chunk_tiles = getenv("BO_OZ_CHUNK_TILES") ? atoi(getenv("BO_OZ_CHUNK_TILES")) : 256;
    return (int64_t)(chunk_tiles > 0 ? chunk_tiles : 256) * 128;
}

// One pass of the int8-slice path over rq's candidates, then the repair of what it cannot certify:
//   1. every chunk of 32 768 candidates: slice K*, contract on tcgen05, moments, acquisition + error bound
//   2. flag the candidates whose bound exceeds the rescue tolerance, compact them (sorted)
//   3. depth 0, long list, a higher level available: re-score the list on this same path at lv.tier2 (depth 1), which
//      hands what it still cannot certify to the FP64 path; otherwise the FP64 path takes the list directly
//   4. scatter back
// Because every candidate that is not certified at the rescue tolerance (5e-7) is re-scored, the level of step 1 only decides the SPEED.
// With lv.first < lv.rest (bo_score_run: one half-level below the level the tolerance selects) chunk 0 runs at the
// lower level and its flagged fraction decides the level of the other chunks: <= oz_tier_frac -> they run at the
// lower level too (2 of 15 digit-pair products saved at the headline shape), else at lv.rest.
static int run_oz(bo_ctx *ctx, const ScoreRequest &rq, OzLevels lv, int depth, const double *gmax_outer, int64_t *fp64_count) {
    const int np = ctx->np, S = ctx->S, d = ctx->d;
    const int64_t M = rq.M;
    const int64_t chunk = oz_chunk_candidates();
    const int64_t cap = bo_round_up64(M < chunk ? M : chunk, 128);
    BO_TRY(reserve_moments(ctx, np / 128, cap, S));
    // what the rescue pass looks at: mode 0 -> the acquisition value (not for the mean, which never goes through
    // the int8 contraction); mode 1 -> s2
    const bool rescue = ctx->oz_rescue && ((rq.mode == 0) ? (rq.acq != BO_ACQ_MEAN) : (rq.dS2 != nullptr));
    const bool own_max = rescue && rq.mode == 0 && !gmax_outer;       // the flag floor needs max |value|
    const bool need_best = rq.want_best || own_max;
    if (rescue && M > (int64_t)0x7fffffff)
        return bo_set_err(ctx, BO_ERR_ARG, "int8 path: at most 2^31 - 1 candidates per call (the rescue list holds 32-bit indices)");
    if (!rescue) lv.first = lv.rest;
    // candidate chunks: [0, first_size), then steps of `chunk`; the pilot chunk of a tiered pass is short (a pilot that
    // fails has its flagged candidates re-scored one level up: 4096 of them at most, and 32 candidate tiles x np / 64
    // row blocks still fill the machine)
    const int64_t first_size = (lv.first < lv.rest && OZ_PILOT_CANDIDATES < chunk) ? OZ_PILOT_CANDIDATES : chunk;
    const int64_t nchunk = M <= first_size ? 1 : 1 + (M - first_size + chunk - 1) / chunk;
    const int64_t nblocks_total = nchunk * ((chunk + 255) / 256);
    if (need_best) BO_TRY(reserve_blocks(ctx, (size_t)(nblocks_total > ARGMAX_PASS_BLOCKS ? nblocks_total : ARGMAX_PASS_BLOCKS) + 8));
    double **errest_p = depth ? &ctx->dErrEst2 : &ctx->dErrEst;
    int **flag_p = depth ? &ctx->dFlagList2 : &ctx->dFlagList;
    double **rescue_p = depth ? &ctx->dRescue2 : &ctx->dRescue;
    const int slot_first = 2 * depth, slot_rest = 2 * depth + 1;
    double floor_abs = 0.0;
    if (rescue) {
        BO_TRY(bo_reserve(ctx, errest_p, depth ? &ctx->errest2_capacity : &ctx->errest_capacity, (size_t)M));
        BO_TRY(bo_reserve(ctx, flag_p, depth ? &ctx->flaglist2_capacity : &ctx->flaglist_capacity, (size_t)M + 1));
        BO_TRY(bo_ozaki_error_scale(ctx, lv_S(lv.first), lv_extra(lv.first), slot_first));
        BO_TRY(bo_ozaki_error_scale(ctx, lv_S(lv.rest), lv_extra(lv.rest), slot_rest));
        if (rq.mode == 1) {                                // s2: floor 1e-9 * mean rho (the parity metric's floor)
            for (int i = 0; i < S; ++i) floor_abs += ctx->h_rho[i];
            floor_abs *= 1e-9 / S;
        }
    }
    double *errest = rescue ? *errest_p : nullptr;
    int *flaglist = rescue ? *flag_p : nullptr;
    int *count_dev = rescue ? flaglist + M : nullptr;
    const double *xflag = rq.mode == 0 ? rq.dVal : rq.dS2;
    const double floor_rel = rq.mode == 0 ? ctx->oz_rescue_floor : 0.0;

    // Software pipeline over work items (chunk, hyper-sample): the operand slicer of item w+1
    // runs on the low-priority side stream while the tcgen05 contraction of item w runs on
    // the main stream (different pipes: FP64/ALU vs tensor).  Two slice buffers.
    const int64_t nitems = nchunk * S;
    const int S_buf = lv_S(lv.first) > lv_S(lv.rest) ? lv_S(lv.first) : lv_S(lv.rest);
    BO_TRY(bo_ozaki_reserve(ctx, S_buf, (int)cap, nitems > 1 ? 2 : 1));
    auto item = [&](int64_t w, int64_t &c0, int &mc, int &mcp, int &s) {
        const int64_t ci = w / S;
        c0 = ci == 0 ? 0 : first_size + (ci - 1) * chunk;
        s = (int)(w % S);
        const int64_t len = ci == 0 ? first_size : chunk;
        mc = (int)((M - c0) < len ? (M - c0) : len);
        mcp = bo_round_up(mc, 128);
    };
    auto level_of = [&](int64_t w) { return (w / S) == 0 ? lv.first : lv.rest; };
    int64_t c0; int mc, mcp, s;
    // Overlap is opt-in (BO_OZ_OVERLAP=1): on a power-capped B200 the two kernels only slow each
    // other down (measured: contraction 2.77 -> 3.56 ms per chunk), so by default both stages
    // run back to back on the main stream.
    static const bool overlap = getenv("BO_OZ_OVERLAP") && atoi(getenv("BO_OZ_OVERLAP")) != 0;
    cudaStream_t side = overlap ? ctx->stream2 : ctx->stream;
    // candidates staged on the main stream must be visible to the side stream
    BO_CUDA(ctx, cudaEventRecord(ctx->ev_consumed[0], ctx->stream));
    BO_CUDA(ctx, cudaStreamWaitEvent(side, ctx->ev_consumed[0], 0));
    item(0, c0, mc, mcp, s);
    BO_TRY(bo_ozaki_slice(ctx, s, lv_S(lv.first), rq.dXc, c0, mc, mcp, 0, side));
    BO_CUDA(ctx, cudaEventRecord(ctx->ev_sliced[0], side));
    int64_t blk0 = 0;
    int64_t pilot_count = -1;
    for (int64_t w = 0; w < nitems; ++w) {
        const int buf = (int)(w & 1);
        const int L = level_of(w);
        item(w, c0, mc, mcp, s);
        BO_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_sliced[buf], 0));
        BO_TRY(bo_ozaki_contract(ctx, s, lv_S(L), lv_extra(L), mcp, buf, ctx->dMuS + (int64_t)s * mcp, ctx->dS2S + (int64_t)s * mcp, nullptr));
        BO_CUDA(ctx, cudaEventRecord(ctx->ev_consumed[buf], ctx->stream));
        if (s == S - 1) {
            AcqParams ap = {};
            ap.mode = rq.mode; ap.acq = rq.acq; ap.param = rq.param;
            ap.S = S; ap.d = d; ap.mc = mc; ap.mcp = mcp; ap.c0 = c0;
            ap.muS = ctx->dMuS; ap.s2S = ctx->dS2S; ap.dmuS = nullptr; ap.ds2S = nullptr;
            ap.out_val = rq.dVal; ap.out_grad = nullptr;
            ap.out_mu = rq.dMu; ap.out_s2 = rq.dS2; ap.out_dmu = nullptr; ap.out_ds2 = nullptr;
            ap.blkval = need_best ? ctx->dBlkVal : nullptr;
            ap.blkidx = need_best ? ctx->dBlkIdx : nullptr;
            ap.blk0 = blk0;
            if (rescue) {
                ap.rhoS = ctx->dRho;
                ap.errK = ctx->dErrK + (size_t)((w / S) == 0 ? slot_first : slot_rest) * S;
                ap.errest = errest;
            }
            const int nb = (mc + 255) / 256;
            {
                BO_LAUNCH(ctx, "acq_kernel");
                acq_kernel<<<nb, 256, 0, ctx->stream>>>(ap);
                BO_CHECK_LAUNCH(ctx);
            }
            blk0 += nb;
            if (w / S == 0 && lv.first < lv.rest) {
                // pilot: how much of chunk 0 does the lower level leave to the tiers above it?  (floor from the
                // chunk's own maximum, never above the global one: the estimate errs towards the higher level)
                if (nchunk < OZ_TIER_MIN_CHUNKS) {
                    pilot_count = mc;       // short passes: a failed pilot costs too much (bo_score_run does not ask for it)
                } else {
                    double *pmax = nullptr;
                    if (rq.mode == 0) {
                        pmax = gmax_outer ? const_cast<double *>(gmax_outer) : ctx->dBlkVal + ctx->blk_capacity - 2;
                        if (!gmax_outer) {
                            BO_LAUNCH(ctx, "argmax_final_kernel");
                            argmax_final_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->dBlkVal, ctx->dBlkIdx, blk0, pmax,
                                                                            ctx->dBlkIdx + ctx->blk_capacity - 2, nullptr, 0);
                            BO_CHECK_LAUNCH(ctx);
                        }
                    }
                    BO_CUDA(ctx, cudaMemsetAsync(count_dev, 0, sizeof(int), ctx->stream));
                    {
                        BO_LAUNCH(ctx, "oz_flag_kernel");
                        oz_flag_kernel<<<(mc + 255) / 256, 256, 0, ctx->stream>>>(mc, xflag, errest, ctx->oz_rescue_tol, floor_abs, floor_rel,
                                                                                 pmax, flaglist, count_dev);
                        BO_CHECK_LAUNCH(ctx);
                    }
                    int cnt = 0;
                    BO_CUDA(ctx, cudaMemcpyAsync(&cnt, count_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
                    BO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                    pilot_count = cnt;
                }
                if ((double)pilot_count <= ctx->oz_tier_frac * (double)mc) {
                    lv.rest = lv.first;
                    BO_TRY(bo_ozaki_error_scale(ctx, lv_S(lv.rest), lv_extra(lv.rest), slot_rest));   // bound of the level that runs
                }
            }
        }
        if (w + 1 < nitems) {
            int64_t c1; int mc1, mcp1, s1;
            item(w + 1, c1, mc1, mcp1, s1);
            const int nbuf = (int)((w + 1) & 1);
            if (w >= 1) BO_CUDA(ctx, cudaStreamWaitEvent(side, ctx->ev_consumed[nbuf], 0));
            BO_TRY(bo_ozaki_slice(ctx, s1, lv_S(level_of(w + 1)), rq.dXc, c1, mc1, mcp1, nbuf, side));
            BO_CUDA(ctx, cudaEventRecord(ctx->ev_sliced[nbuf], side));
        }
    }
    double *gmax = ctx->dBlkVal ? ctx->dBlkVal + ctx->blk_capacity - 1 : nullptr;
    if (need_best) {
        BO_LAUNCH(ctx, "argmax_final_kernel");
        argmax_final_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->dBlkVal, ctx->dBlkIdx, blk0, gmax,
                                                        ctx->dBlkIdx + ctx->blk_capacity - 1, ctx->rec_ptr, ctx->rec_offset);
        BO_CHECK_LAUNCH(ctx);
    }
    if (gmax_outer) gmax = const_cast<double *>(gmax_outer);
    if (fp64_count) *fp64_count = 0;
    if (depth == 0) {
        ctx->oz_last_total = M;
        ctx->oz_last_flagged = 0;
        ctx->oz_last_first = lv.first; ctx->oz_last_rest = lv.rest; ctx->oz_last_tier2 = 0;
        ctx->oz_last_first_flagged = 0;
    }
    if (!rescue) return BO_OK;

    // ---- rescue: flag, compact (sorted), re-score (higher level, then FP64), scatter ----
    BO_CUDA(ctx, cudaMemsetAsync(count_dev, 0, sizeof(int), ctx->stream));
    {
        BO_LAUNCH(ctx, "oz_flag_kernel");
        const int nb = (int)((M + 255) / 256 < 1184 ? (M + 255) / 256 : 1184);
        oz_flag_kernel<<<nb, 256, 0, ctx->stream>>>(M, xflag, errest, ctx->oz_rescue_tol, floor_abs, floor_rel,
                                                    rq.mode == 0 ? gmax : nullptr, flaglist, count_dev);
        BO_CHECK_LAUNCH(ctx);
    }
    int count = 0;
    BO_CUDA(ctx, cudaMemcpyAsync(&count, count_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    BO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (depth == 0) ctx->oz_last_first_flagged = count;
    if (count == 0) return BO_OK;
    if (count <= 8192) {
        int np2 = 1;
        while (np2 < count) np2 <<= 1;
        BO_LAUNCH(ctx, "oz_sort_small_kernel");
        oz_sort_small_kernel<<<1, 1024, np2 * sizeof(int), ctx->stream>>>(flaglist, count);
        BO_CHECK_LAUNCH(ctx);
    } else {
        const int nwords = (int)((M + 63) / 64), nwb = (nwords + 1023) / 1024;
        BO_TRY(bo_reserve(ctx, &ctx->dFlagBits, &ctx->flagbits_capacity, (size_t)nwords + (size_t)(nwb + 1) / 2 + 1));
        int *blocksum = reinterpret_cast<int *>(ctx->dFlagBits + nwords);
        BO_CUDA(ctx, cudaMemsetAsync(ctx->dFlagBits, 0, sizeof(unsigned long long) * nwords, ctx->stream));
        ctx->launches += 3;
        oz_mark_kernel<<<(count + 255) / 256, 256, 0, ctx->stream>>>(flaglist, count, ctx->dFlagBits);
        oz_wordcount_kernel<<<nwb, 1024, 0, ctx->stream>>>(ctx->dFlagBits, nwords, blocksum);
        oz_compact_kernel<<<nwb, 1024, 0, ctx->stream>>>(ctx->dFlagBits, nwords, blocksum, flaglist);
        BO_CHECK_LAUNCH(ctx);
    }
    const size_t per = (size_t)d + 2;
    BO_TRY(bo_reserve(ctx, rescue_p, depth ? &ctx->rescue2_capacity : &ctx->rescue_capacity, (size_t)count * per));
    double *xr = *rescue_p, *ra = xr + (size_t)count * d, *rb = ra + count;
    {
        BO_LAUNCH(ctx, "oz_gather_rows_kernel");
        oz_gather_rows_kernel<<<(int)(((int64_t)count * d + 255) / 256), 256, 0, ctx->stream>>>(rq.dXc, d, flaglist, count, xr);
        BO_CHECK_LAUNCH(ctx);
    }
    ScoreRequest r2;
    r2.mode = rq.mode; r2.acq = rq.acq; r2.param = rq.param; r2.M = count; r2.dXc = xr;
    if (rq.mode == 0) r2.dVal = ra;
    else { r2.dMu = ra; r2.dS2 = rb; }
    r2.want_best = false;
    int64_t to_fp64 = count;
    const int top = lv.first > lv.rest ? lv.first : lv.rest;
    const int next = (lv.tier2 > 0 && top < lv.sel) ? lv.sel : lv.tier2;
    // (a list longer than the demotion threshold is a fit whose variance has collapsed over the candidate set: one more
    //  digit rarely certifies it, it goes to FP64 directly and the fit is demoted below)
    if (depth == 0 && next > top && count >= ctx->oz_tier_min && (double)count <= ctx->oz_demote_frac * (double)M) {
        OzLevels l2 = {next, next, 0, next};
        BO_TRY(run_oz(ctx, r2, l2, 1, rq.mode == 0 ? gmax : nullptr, &to_fp64));
        ctx->oz_last_tier2 = next;
    } else {
        BO_TRY(run_fp64(ctx, r2));
    }
    if (fp64_count) *fp64_count = to_fp64;
    {
        BO_LAUNCH(ctx, "oz_scatter_kernel");
        if (rq.mode == 0)
            oz_scatter_kernel<<<(count + 255) / 256, 256, 0, ctx->stream>>>(flaglist, count, ra, rq.dVal, nullptr, nullptr);
        else
            oz_scatter_kernel<<<(count + 255) / 256, 256, 0, ctx->stream>>>(flaglist, count, ra, rq.dMu, rb, rq.dS2);
        BO_CHECK_LAUNCH(ctx);
    }
    if (depth != 0) return BO_OK;
    ctx->oz_last_flagged = to_fp64;
    if (rq.want_best) BO_TRY(final_argmax_over(ctx, rq.dVal, M));
    // a fit where the int8 path hands most of its work on (posterior variance collapsed over the whole candidate
    // set, e.g. n = 1024 in d = 4) is an FP64 problem: later passes go there directly.  Judged on what the selected
    // level (lv.rest, or everything when the pass ran at one level) left over.
    {
        double frac;
        if (pilot_count >= 0 && lv.rest != lv.first && M > first_size)
            frac = (double)(count - (pilot_count < count ? pilot_count : count)) / (double)(M - first_size);
        else if (pilot_count >= 0 && lv.rest == lv.first && lv.tier2 > 0)
            frac = (double)to_fp64 / (double)M;         // whole pass below the selected level: judge by what reached FP64
        else
            frac = (double)count / (double)M;
        if (M >= 4096 && frac > ctx->oz_demote_frac) ctx->oz_demoted = true;
    }
    return BO_OK;
}

int bo_score_run(bo_ctx *ctx, const ScoreRequest &rq) {
    if (!ctx->fitted) return bo_set_err(ctx, BO_ERR_STATE, "bo_score/bo_predict before bo_fit");
    const bool grad = (rq.mode == 0) ? (rq.dGrad != nullptr) : (rq.dDmu != nullptr || rq.dDs2 != nullptr);
    if (rq.M <= 0) return bo_set_err(ctx, BO_ERR_ARG, "M must be positive");
    // int8-slice path for value-only passes when selected; gradients stay on the FP64 path
    // (tiny batches stay on the exact FP64 GEMV/GEMM path: nothing to gain from the tensor pipe there)
    const bool oz = (ctx->prec == BO_PREC_OZAKI) && !grad && rq.M > 64 && !ctx->oz_demoted;
    ctx->oz_last_path = oz ? 1 : 0;
    if (!oz) {
        ctx->oz_last_total = rq.M;
        ctx->oz_last_flagged = 0;
        ctx->oz_last_first = ctx->oz_last_rest = ctx->oz_last_tier2 = 0;
        ctx->oz_last_first_flagged = 0;
        return run_fp64(ctx, rq);
    }
    const int oz_S = bo_ozaki_choose_slices(ctx, ctx->prec_tol);
    if (oz_S < 1) return BO_ERR_CUDA;
    const int L_sel = 2 * oz_S + (ctx->oz_extra ? 1 : 0);
    OzLevels lv = {L_sel, L_sel, 0, L_sel};
    // a pinned level (tol >= 2: tests, calibration) runs exactly as pinned: one level, flagged candidates to FP64
    const bool pinned = ctx->prec_tol >= 2.0;
    if (!pinned && ctx->oz_rescue && ctx->oz_tiered) {
        const int L2 = L_sel + 1;                                   // one step up for the flagged list
        if (lv_S(L2) <= 7 && (int64_t)ctx->np * lv_S(L2) < (1 << 17)) lv.tier2 = L2;
        // (a pilot that fails has its 4096 candidates re-scored one level up: < 2 % of a pass of this length)
        if (L_sel - 1 >= 6 && rq.M >= OZ_TIER_MIN_CHUNKS * oz_chunk_candidates() && lv.tier2 > 0) lv.first = L_sel - 1;
    }
    BO_TRY(bo_ozaki_prepare(ctx, lv.tier2 > 0 ? lv_S(lv.tier2) : oz_S));
    return run_oz(ctx, rq, lv, 0, nullptr, nullptr);
}


// top-k of a device-resident value array; results land in dBlkVal/dBlkIdx tail
int bo_topk_run(bo_ctx *ctx, const double *vals, int64_t M, int k, double *h_val, int64_t *h_idx) {
    const int nb = (int)((M + 255) / 256 < 592 ? (M + 255) / 256 : 592);
    size_t need = (size_t)nb + 2 * (size_t)k + 8;
    if (ctx->blk_capacity < need || !ctx->dBlkVal) {
        size_t c1 = ctx->blk_capacity, c2 = ctx->blk_capacity;
        BO_TRY(bo_reserve(ctx, &ctx->dBlkVal, &c1, need));
        BO_TRY(bo_reserve(ctx, &ctx->dBlkIdx, &c2, need));
        ctx->blk_capacity = need;
    }
    double *rv = ctx->dBlkVal + nb;
    int64_t *ri = ctx->dBlkIdx + nb;
    for (int j = 0; j < k; ++j) {
        {
            BO_LAUNCH(ctx, "topk_pass_kernel");
            topk_pass_kernel<<<nb, 256, 0, ctx->stream>>>(vals, M, j ? rv + j - 1 : rv, j ? ri + j - 1 : ri,
                                                         j == 0, ctx->dBlkVal, ctx->dBlkIdx);
            BO_CHECK_LAUNCH(ctx);
        }
        {
            BO_LAUNCH(ctx, "argmax_final_kernel");
            argmax_final_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->dBlkVal, ctx->dBlkIdx, nb, rv + j, ri + j, nullptr, 0);
            BO_CHECK_LAUNCH(ctx);
        }
    }
    BO_CUDA(ctx, cudaMemcpyAsync(h_val, rv, sizeof(double) * k, cudaMemcpyDeviceToHost, ctx->stream));
    BO_CUDA(ctx, cudaMemcpyAsync(h_idx, ri, sizeof(int64_t) * k, cudaMemcpyDeviceToHost, ctx->stream));
    BO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BO_OK;
}
