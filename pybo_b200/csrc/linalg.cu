// linalg.cu -- the fit side of the hot path (what `model.add_data` implies,
// reference bayesopt.py:114,258,269): Gram matrix K = k(X,X) + sn2 I, blocked
// right-looking Cholesky (64-wide panels, FP64 DMMA trailing update), blocked
// recursive triangular inverse W = L^-1, and alpha / beta / log-det.
#include <stdlib.h>

#include "common.cuh"
#include "dgemm.cuh"

typedef DTile<64, 64, 32, 32, 3, true> T64NT;
typedef DTile<64, 64, 32, 32, 3, false> T64NN;
typedef DTile<128, 128, 64, 32, 4, true> T128NT;

// ---------------------------------------------------------------------------
// Gram matrix.  Xs is already divided by ell (zero padded to dp columns).
// Padded rows/cols get the identity so the factor stays well defined.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double kern_from_sqdist(int kernel, double D, double rho) {
    if (kernel == BO_KERNEL_SE) return rho * exp(-0.5 * D);
    double r = sqrt(5.0 * D);
    return rho * (1.0 + r + r * r * (1.0 / 3.0)) * exp(-r);
}

// One CTA = one 64 x 64 tile of the LOWER triangle (tile row ti >= tile column tj; upper tiles are produced by
// mirroring, never recomputed).  The 2 x 64 scaled points are staged in shared memory k-major, so a thread reads its
// four row points as broadcasts and its four column points as two 16-byte loads; thread (ty, tx) owns rows
// ty + 16 a and columns 4 tx + b and stores 32 contiguous bytes per row (a warp covers two full 512-byte row
// segments).  With `mirror` the tile also goes through a padded shared-memory transpose and is written to the upper
// triangle with the same store pattern.  `aug` (optional, S x np): row `n` of the padded matrix receives
// [aug[0..n), ann] instead of the identity -- the right-hand side rides through the factorisation (bo_loglik_fit).
template <int DP>
__global__ void __launch_bounds__(256)
gram_tile_kernel(int kernel, int n, int np, const double *__restrict__ Xs, const double *__restrict__ rho,
                 const double *__restrict__ sn2, double *__restrict__ K, int mirror, const double *__restrict__ aug,
                 const double *__restrict__ ann) {
    // the point stage (2 x DP x 64 doubles) and the transpose tile (64 x 65) share one buffer: the tile is written
    // only after every thread has finished reading the points
    constexpr int STAGE = 2 * DP * 64, TILE = 64 * 65;
    __shared__ __align__(16) double buf[STAGE > TILE ? STAGE : TILE];
    double (*xi)[64] = reinterpret_cast<double (*)[64]>(buf);
    double (*xj)[64] = reinterpret_cast<double (*)[64]>(buf + DP * 64);
    double (*tile)[65] = reinterpret_cast<double (*)[65]>(buf);
    const int s = blockIdx.z;
    int x = blockIdx.x;
    int ti = (int)((sqrt(8.0 * x + 1.0) - 1.0) * 0.5);
    while ((ti + 1) * (ti + 2) / 2 <= x) ++ti;
    while (ti * (ti + 1) / 2 > x) --ti;
    const int tj = x - ti * (ti + 1) / 2;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const double *Xb = Xs + (int64_t)s * np * DP;
    for (int e = tid; e < 64 * DP; e += 256) {
        xi[e % DP][e / DP] = Xb[(int64_t)ti * 64 * DP + e];
        xj[e % DP][e / DP] = Xb[(int64_t)tj * 64 * DP + e];
    }
    __syncthreads();
    double D[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) D[a][b] = 0.0;
#pragma unroll
    for (int k = 0; k < DP; ++k) {
        double pi[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) pi[a] = xi[k][ty + 16 * a];
        const double2 p01 = *reinterpret_cast<const double2 *>(&xj[k][4 * tx]);
        const double2 p23 = *reinterpret_cast<const double2 *>(&xj[k][4 * tx + 2]);
        const double pj[4] = {p01.x, p01.y, p23.x, p23.y};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const double t = pi[a] - pj[b];
                D[a][b] = fma(t, t, D[a][b]);
            }
    }
    const double rh = rho[s], sn = sn2[s];
    double *Kb = K + (int64_t)s * np * np;
    const bool offdiag = ti != tj;
    if (mirror && offdiag) __syncthreads();                    // points consumed: the buffer becomes the tile
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int i = ti * 64 + ty + 16 * a;
        double v[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int j = tj * 64 + 4 * tx + b;
            if (i < n && j < n) {
                v[b] = kern_from_sqdist(kernel, D[a][b], rh);
                if (i == j) v[b] += sn;
            } else if (aug != nullptr && i == n && j <= n) {
                v[b] = (j < n) ? aug[(int64_t)s * np + j] : ann[s];
            } else {
                v[b] = (i == j) ? 1.0 : 0.0;
            }
        }
        double *dst = Kb + (int64_t)i * np + tj * 64 + 4 * tx;
        *reinterpret_cast<double2 *>(dst) = make_double2(v[0], v[1]);
        *reinterpret_cast<double2 *>(dst + 2) = make_double2(v[2], v[3]);
        if (mirror && offdiag) {
#pragma unroll
            for (int b = 0; b < 4; ++b) tile[ty + 16 * a][4 * tx + b] = v[b];
        }
    }
    if (mirror && offdiag) {
        __syncthreads();
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int r = ty + 16 * a;                         // row of the mirrored tile = column of the computed one
            double *dst = Kb + (int64_t)(tj * 64 + r) * np + ti * 64 + 4 * tx;
            *reinterpret_cast<double2 *>(dst) = make_double2(tile[4 * tx][r], tile[4 * tx + 1][r]);
            *reinterpret_cast<double2 *>(dst + 2) = make_double2(tile[4 * tx + 2][r], tile[4 * tx + 3][r]);
        }
    }
}

// K_s = k_s(X, X) + sn2_s I on the lower triangle of S padded np x np matrices (np % 64 == 0); `mirror` also fills
// the upper triangle (the factorisation never reads it).
int bo_linalg_gram(bo_ctx *ctx, int kernel, int n, int np, int dp, int S, const double *dXs,
                   const double *rho_dev, const double *sn2_dev, double *K, int mirror, const double *aug,
                   const double *ann) {
    const int T = np / 64;
    dim3 grd(T * (T + 1) / 2, 1, S);
    BO_LAUNCH(ctx, "gram_kernel");
    switch (dp) {
#define GRAM_RUN(DPV) case DPV: gram_tile_kernel<DPV><<<grd, 256, 0, ctx->stream>>>(kernel, n, np, dXs, rho_dev, sn2_dev, K, mirror, aug, ann); break
        GRAM_RUN(2); GRAM_RUN(4); GRAM_RUN(8); GRAM_RUN(16); GRAM_RUN(32);
#undef GRAM_RUN
        default: return bo_set_err(ctx, BO_ERR_ARG, "unsupported padded dimension %d", dp);
    }
    BO_CHECK_LAUNCH(ctx);
    return BO_OK;
}

// log marginal likelihood from an augmented factor: row n of L holds alpha^T = (L^-1 r)^T
//   ll = -1/2 |alpha|^2 - sum_{i<n} log L_ii - n/2 log(2 pi)
__global__ void loglik_aug_kernel(const double *__restrict__ L, int n, int np, double *__restrict__ out) {
    __shared__ double red[32];
    const double *Ls = L + (int64_t)blockIdx.x * np * np;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double a = Ls[(int64_t)n * np + i];
        acc += -0.5 * a * a - log(Ls[(int64_t)i * (np + 1)]);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) out[blockIdx.x] = v - 0.5 * n * 1.8378770664093453;   // log(2 pi)
    }
}

int bo_linalg_loglik_aug(bo_ctx *ctx, const double *L, int n, int np, int S, double *out) {
    BO_LAUNCH(ctx, "loglik_aug_kernel");
    loglik_aug_kernel<<<S, 256, 0, ctx->stream>>>(L, n, np, out);
    BO_CHECK_LAUNCH(ctx);
    return BO_OK;
}

// ---------------------------------------------------------------------------
// 64 x 64 diagonal block: factor in shared memory (one barrier per column) and
// invert the factor (needed by the panel solve and by W = L^-1).
// ---------------------------------------------------------------------------
typedef DTile<64, 64, 32, 16, 3, true> T64NT8;       // 256-thread variant for the fused panel step

// Register-resident factorisation of a 64 x 64 diagonal block by one CTA of 256 threads, followed by the
// inversion of the factor (needed by the panel solve and by W = L^-1).
//
// Factor: thread (ty, tx) owns a[ty + 16 ai][tx + 16 b] in registers for the whole factorisation; per
// column only the pivot column is broadcast through double-buffered shared memory (one barrier per
// column).  Columns stay unscaled (a[i][j] = L[i][j] sqrt(d_j)) so the multiplier is m_ij = a[i][j] / d_j
// with 1 / d_j from MUFU.RCP64H + two Newton steps (the column loop is the serial chain of the whole
// Cholesky: tools/latency.py measures 67 cycles for the smem -> barrier -> smem round trip, 75-80 for
// rsqrt / __drcp_rn, 9 per dependent DFMA); the columns are scaled by 1 / sqrt(d_j) at the end.
//
// Inverse: NOT carried through the column loop (that doubled its instruction count and the loop is
// issue-bound) but built afterwards in shared memory by recursive doubling over block sizes 1, 2, .. 32:
//     X = [[X11, 0], [-X22 L21 X11, X22]]
// all 32 / b pairs of a level in parallel, two barriers per level, ~87 k FMA in total.
struct PotrfSmem {
    double colbuf[2][64];
    double dj[64];              // pivots d_j
    double rs[64];              // 1 / sqrt(d_j) = 1 / L_jj
    int bad;
};
#define PF_LD 65                // leading dimension of the 64 x 64 shared-memory tiles (conflict-free columns)
#define PF_TILE (64 * PF_LD)

__device__ __forceinline__ double pf_rcp(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double e = fma(-d, y, 1.0);
    y = fma(y, e, y);
    e = fma(-d, y, 1.0);
    return fma(y, e, y);
}

// One level of the recursive-doubling inverse: for each of the 32 / B pairs of B x B diagonal blocks
// (rows r0 .. r0 + 2B), T = L21 X11 then X21 = -X22 T.  A thread owns column ci of one pair and R = max(1, B / 8)
// rows of it, so the R dot products share the loads of the X11 / T column and run as independent chains.
template <int B>
__device__ __forceinline__ void pf_invert_level(const double *Ls, double *Xs, double *Ts) {
    constexpr int R = B >= 8 ? B / 8 : 1;                 // rows per thread
    constexpr int NT = 32 * B / R;                        // active threads (32 B outputs per level)
    constexpr int RS = B / R;                             // row stride between a thread's rows
    const int tid = threadIdx.x;
    const int ci = tid & (B - 1);
    const int rr = (tid / B) % RS;                        // first row of this thread inside the block
    const int q = tid / (B * RS);                         // pair
    const int r0 = q * 2 * B;
    const bool on = tid < NT;
    if (on) {
        double acc[R];
#pragma unroll
        for (int m = 0; m < R; ++m) acc[m] = 0.0;
        const double *x11 = Xs + r0 * PF_LD + r0 + ci;
        const double *l21 = Ls + (r0 + B + rr) * PF_LD + r0;
        // (X11 is lower triangular, so terms t < ci are zero: fixed bounds keep the loop fully unrolled and
        //  every load independent of the lane)
#pragma unroll
        for (int t = 0; t < B; ++t) {
            const double x = x11[t * PF_LD];
#pragma unroll
            for (int m = 0; m < R; ++m) acc[m] = fma(l21[m * RS * PF_LD + t], x, acc[m]);
        }
#pragma unroll
        for (int m = 0; m < R; ++m) Ts[(r0 + B + rr + m * RS) * 33 + ci] = acc[m];
    }
    __syncthreads();
    if (on) {
        double acc[R];
#pragma unroll
        for (int m = 0; m < R; ++m) acc[m] = 0.0;
        const double *tt = Ts + (r0 + B) * 33 + ci;
        const double *x22 = Xs + (r0 + B + rr) * PF_LD + r0 + B;
        // X22 is lower triangular: beyond the diagonal its entries are zero
#pragma unroll
        for (int t = 0; t < B; ++t) {
            const double tv = tt[t * 33];
#pragma unroll
            for (int m = 0; m < R; ++m) acc[m] = fma(x22[m * RS * PF_LD + t], tv, acc[m]);
        }
#pragma unroll
        for (int m = 0; m < R; ++m) Xs[(r0 + B + rr + m * RS) * PF_LD + r0 + ci] = -acc[m];
    }
    __syncthreads();
}

// `ra` holds the (already updated) block on entry; L goes to Ab (global) and Ls (shared), the inverse to Db.
// Ls, Xs: PF_TILE doubles each; Ts: 64 x 33 doubles.
__device__ __forceinline__ void potrf64_regs(double (&ra)[4][4], PotrfSmem &sm, double *Ls, double *Xs, double *Ts,
                                             double *Ab, int ld, double *Db, int *info, int kblk) {
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    if (tid == 0) sm.bad = 0;
#pragma unroll
    for (int ai = 0; ai < 4; ++ai)
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if (tx + 16 * b > ty + 16 * ai) ra[ai][b] = 0.0;
    int p = 0;
    const bool tx_le_ty = tx <= ty;
    // The column loop is unrolled over the 16-column block jb so that which register tiles take
    // part in a step (rows below the pivot, columns right of it) is known at compile time;
    // only the comparisons against jl inside the pivot's own tile remain at run time.
#pragma unroll
    for (int jb = 0; jb < 4; ++jb) {
#pragma unroll 1
        for (int jl = 0; jl < 16; ++jl) {
            const int j = 16 * jb + jl;
            if (tx == jl) {            // owners of column j of a
#pragma unroll
                for (int ai = jb; ai < 4; ++ai) sm.colbuf[p][ty + 16 * ai] = ra[ai][jb];
            }
            __syncthreads();
            const double d = sm.colbuf[p][j];
            const double r2 = pf_rcp(d);
            if (tid == 0) {
                sm.dj[j] = d;
                if (!(d > 0.0) && sm.bad == 0) sm.bad = j + 1;
            }
            const bool tx_gt_jl = tx > jl, ty_gt_jl = ty > jl;
            double cc[4];
#pragma unroll
            for (int b = jb; b < 4; ++b) cc[b] = sm.colbuf[p][tx + 16 * b];
#pragma unroll
            for (int ai = jb; ai < 4; ++ai) {
                const bool row_on = (ai > jb) || ty_gt_jl;          // i > j
                if (!row_on) continue;
                const double mij = sm.colbuf[p][ty + 16 * ai] * r2;
#pragma unroll
                for (int b = jb; b <= ai; ++b) {                    // trailing update: j < k <= i
                    const bool k_gt_j = (b > jb) || tx_gt_jl;
                    const bool k_le_i = (b < ai) || tx_le_ty;
                    if (k_gt_j && k_le_i) ra[ai][b] = fma(-mij, cc[b], ra[ai][b]);
                }
            }
            p ^= 1;
        }
    }
    __syncthreads();
    if (tid < 64) sm.rs[tid] = rsqrt(sm.dj[tid]);
    __syncthreads();
    // L = a diag(rs): to global, and to shared memory for the inversion; X starts as diag(1 / L_ii)
#pragma unroll
    for (int ai = 0; ai < 4; ++ai)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int i = ty + 16 * ai, k = tx + 16 * b;
            const double l = (k <= i) ? ra[ai][b] * sm.rs[k] : 0.0;      // diag: d / sqrt(d)
            Ab[(int64_t)i * ld + k] = l;
            Ls[i * PF_LD + k] = l;
            Xs[i * PF_LD + k] = (i == k) ? sm.rs[i] : 0.0;
        }
    __syncthreads();
    pf_invert_level<1>(Ls, Xs, Ts);
    pf_invert_level<2>(Ls, Xs, Ts);
    pf_invert_level<4>(Ls, Xs, Ts);
    pf_invert_level<8>(Ls, Xs, Ts);
    pf_invert_level<16>(Ls, Xs, Ts);
    pf_invert_level<32>(Ls, Xs, Ts);
    for (int e = tid; e < 4096; e += 256) Db[e] = Xs[(e >> 6) * PF_LD + (e & 63)];
    if (tid == 0 && sm.bad != 0) atomicCAS(info, 0, kblk * 64 + sm.bad);
}

// One step of the blocked factorisation in ONE launch (the serial chain of the algorithm):
//   CTA 0      : D = A_cc - L_{c,c-1} L_{c,c-1}^T (panel c-1, if any), factor D, invert it, raise flag[c];
//   CTA x >= 1 : R = A_ic - L_{i,c-1} L_{c,c-1}^T for row tile i = c + x (DMMA), then -- once the flag is
//                up -- L_ic = R Dinv_c^T.
// The column update of the tiles below the diagonal overlaps the factorisation of the diagonal block;
// CTA 0 never waits and is dispatched before the CTAs that wait for it, so the spin cannot deadlock.
// mode 0: fused (above); mode 1: diagonal CTA only (grid.x = 1); mode 2: panel CTAs only, launched after mode 1 in
// stream order (grid.x = nblk - c - 1, no flag to wait for, the smaller GEMM-only shared-memory footprint).  The split
// form is used when many matrices are factored at once: CTAs that spin on the flag while the diagonal block is being
// factored hold shared memory that the trailing updates of the other matrices could be using.
__global__ void __launch_bounds__(256)
chol_step_kernel(double *A, int ld, int64_t strideA, int c, double *dinv, int64_t strideD, int *info, int *flags,
                 int nblk, int mode, int kprev) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x;
    double *Az = A + blockIdx.z * strideA;
    double *Db = dinv + blockIdx.z * strideD + (int64_t)c * 4096;
    int *flag = flags + blockIdx.z * nblk + c;
    // kprev: width of the finished columns directly left of block column c that still have to be applied to it
    // (0, 64 or 128: trailing updates run per PAIR of panels, see chol_enqueue)
    const bool has_prev = kprev > 0;
    const int64_t cprev = (int64_t)c * 64 - kprev;
    const int bx = (mode == 2) ? (int)blockIdx.x + 1 : (int)blockIdx.x;
    if (bx == 0) {
        PotrfSmem &sm = *reinterpret_cast<PotrfSmem *>(smem);
        double *P = smem + (sizeof(PotrfSmem) + 7) / 8;           // [64][65] copy of L_{c,c-1}, later the factor
        double *Xs = P + PF_TILE, *Ts = Xs + PF_TILE;
        const int tx = tid & 15, ty = tid >> 4;
        double *Ab = Az + (int64_t)c * 64 * (ld + 1);
        double ra[4][4];
#pragma unroll
        for (int ai = 0; ai < 4; ++ai)
#pragma unroll
            for (int b = 0; b < 4; ++b) ra[ai][b] = Ab[(int64_t)(ty + 16 * ai) * ld + tx + 16 * b];
        if (has_prev) {
            // D = A_cc - L_{c,c-1} L_{c,c-1}^T on the FP64 tensor cores (half the shared-memory traffic of a DFMA
            // register tile), then through shared memory from the MMA fragment layout into the (ty, tx) layout
            typedef T64NT8 T;
            const double *Lp = Az + (int64_t)c * 64 * ld + cprev;
            const int warp = tid >> 5, lane = tid & 31;
            const int wm = warp / T::WARPS_N, wn = warp % T::WARPS_N, g = lane >> 2, t = lane & 3;
            double acc[T::MI][T::NI][2];
#pragma unroll
            for (int mi = 0; mi < T::MI; ++mi)
#pragma unroll
                for (int ni = 0; ni < T::NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
            T::mainloop(acc, Lp, ld, Lp, ld, 0, kprev, smem);        // ends with a barrier: smem is free again
#pragma unroll
            for (int mi = 0; mi < T::MI; ++mi)
#pragma unroll
                for (int ni = 0; ni < T::NI; ++ni) {
                    const int r = wm * T::WM + mi * 8 + g, cc = wn * T::WN + ni * 8 + 2 * t;
                    Xs[r * PF_LD + cc] = acc[mi][ni][0];
                    Xs[r * PF_LD + cc + 1] = acc[mi][ni][1];
                }
            __syncthreads();
#pragma unroll
            for (int ai = 0; ai < 4; ++ai)
#pragma unroll
                for (int b = 0; b < 4; ++b) ra[ai][b] -= Xs[(ty + 16 * ai) * PF_LD + tx + 16 * b];
            __syncthreads();                                      // Xs is rebuilt by the inversion below
        }
        potrf64_regs(ra, sm, P, Xs, Ts, Ab, ld, Db, info + blockIdx.z, c);
        __threadfence();
        __syncthreads();
        if (tid == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(1) : "memory");
        return;
    }
    typedef T64NT8 T;
    const int i = c + bx;
    double *C = Az + (int64_t)i * 64 * ld + (int64_t)c * 64;
    const int warp = tid >> 5, lane = tid & 31;
    const int wm = warp / T::WARPS_N, wn = warp % T::WARPS_N;
    const int g = lane >> 2, t = lane & 3;
    double acc[T::MI][T::NI][2];
    if (has_prev) {
#pragma unroll
        for (int mi = 0; mi < T::MI; ++mi)
#pragma unroll
            for (int ni = 0; ni < T::NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
        const double *Li = Az + (int64_t)i * 64 * ld + cprev;
        const double *Lc = Az + (int64_t)c * 64 * ld + cprev;
        T::mainloop(acc, Li, ld, Lc, ld, 0, kprev, smem);
#pragma unroll
        for (int mi = 0; mi < T::MI; ++mi)
#pragma unroll
            for (int ni = 0; ni < T::NI; ++ni) {
                const int r = wm * T::WM + mi * 8 + g, cc = wn * T::WN + ni * 8 + 2 * t;
                double2 *dst = reinterpret_cast<double2 *>(&C[(int64_t)r * ld + cc]);
                double2 v = *dst;
                v.x -= acc[mi][ni][0];
                v.y -= acc[mi][ni][1];
                *dst = v;
            }
        __threadfence();             // the tile is re-read below through cp.async (L2)
        __syncthreads();
    }
    if (mode == 0) {
        if (tid == 0) {
            int v;
            do {
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
            } while (v == 0);
        }
        __syncthreads();
    }

#pragma unroll
    for (int mi = 0; mi < T::MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < T::NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    T::mainloop(acc, C, ld, Db, 64, 0, 64, smem);
#pragma unroll
    for (int mi = 0; mi < T::MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < T::NI; ++ni) {
            const int r = wm * T::WM + mi * 8 + g, cc = wn * T::WN + ni * 8 + 2 * t;
            *reinterpret_cast<double2 *>(&C[(int64_t)r * ld + cc]) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
        }
}

// ---------------------------------------------------------------------------
// Single-matrix factorisation as ONE persistent cooperative kernel (dataflow over 64 x 64 tiles).
//
// The launch-per-step pipeline above spends a third of every step on kernel boundaries (drain of the slowest panel
// CTA, launch, cross-stream event) and synchronises whole kernels where single tiles depend on each other.  Here every
// CTA is resident for the whole factorisation and tiles are handed over through release / acquire flags in global
// memory:
//   solved[r][c]   (c <  r)  tile L_rc is final;   solved[c][c]  the diagonal block c is factored and Dinv_c stored
//   applied[i][j]            number of panel PAIRS (2q, 2q+1) the update workers have applied to tile (i, j)
// Row CTA r (blockIdx.x = r < nblk) owns block row r of the panel chain: for c = 0 .. r-1 it applies panel c-1 to its
// tile (r, c), multiplies by Dinv_c^T as soon as that exists, and at c = r factors the diagonal block -- the same
// arithmetic, in the same order, as chol_step_kernel.  While it waits for Dinv_c at an odd c it already applies panel
// c-1 to its tile in column c+1, so that every tile meets its own step with only ONE panel (K = 64) left to apply.
// All other CTAs -- and the row CTAs once their row is done -- are update workers: they draw trailing-update tasks
// (pair p, tile (i, j), i >= j >= 2p + 3:  C_ij -= L_i,pair L_j,pair^T, K = 128) from one atomic counter, in an order
// (pair, column, row) that is a topological order of the dependency graph and serves the columns the chain needs next
// first.  Cooperative launch guarantees that every CTA is resident, so a waiting CTA always waits for a running one.
// ---------------------------------------------------------------------------
struct CholFlowParams {
    double *A;
    int ld, nblk;
    double *dinv;
    int *info;
    int *solved, *applied, *next;
};

__device__ __forceinline__ int flow_ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void flow_wait(const int *flag, int v) {
    if (threadIdx.x == 0) {
        while (flow_ld_acquire(flag) < v) __nanosleep(64);
    }
    __syncthreads();
}
// every thread's global writes first (fence), then one release store
__device__ __forceinline__ void flow_signal(int *flag, int v) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(v) : "memory");
}

// SUB: C -= A B^T (C is read into the accumulators, negated, while the operand stages are in flight); else C = A B^T.
// A: 64 rows (lda), B: 64 rows stored [n][k] (ldb); K a multiple of 16.  Ends with every thread's stores issued.
template <bool SUB>
__device__ __forceinline__ void flow_gemm(double *C, int ldc, const double *Ap, int lda, const double *Bp, int ldb, int K,
                                          double *smem) {
    typedef T64NT8 T;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp / T::WARPS_N, wn = warp % T::WARPS_N;
    const int g = lane >> 2, t = lane & 3;
    double acc[T::MI][T::NI][2];
#pragma unroll
    for (int mi = 0; mi < T::MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < T::NI; ++ni) {
            if (SUB) {
                const int r = wm * T::WM + mi * 8 + g, c = wn * T::WN + ni * 8 + 2 * t;
                // (L2 load: the tile may have been written by another SM since this SM last cached it)
                const double2 v = __ldcg(reinterpret_cast<const double2 *>(&C[(int64_t)r * ldc + c]));
                acc[mi][ni][0] = -v.x;
                acc[mi][ni][1] = -v.y;
            } else {
                acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
            }
        }
    T::mainloop(acc, Ap, lda, Bp, ldb, 0, K, smem);
    const double sg = SUB ? -1.0 : 1.0;
#pragma unroll
    for (int mi = 0; mi < T::MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < T::NI; ++ni) {
            const int r = wm * T::WM + mi * 8 + g, c = wn * T::WN + ni * 8 + 2 * t;
            *reinterpret_cast<double2 *>(&C[(int64_t)r * ldc + c]) = make_double2(sg * acc[mi][ni][0], sg * acc[mi][ni][1]);
        }
}

__device__ __forceinline__ int flow_napp(int c) { return c >= 3 ? (c - 3) / 2 + 1 : 0; }

__global__ void __launch_bounds__(256)
chol_flow_kernel(CholFlowParams p) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_task[4];
    const int tid = threadIdx.x, nblk = p.nblk, ld = p.ld;
    double *Az = p.A;
    if ((int)blockIdx.x < nblk) {
        // ===================== row CTA: block row r of the panel chain =====================
        const int r = blockIdx.x;
        double *row = Az + (int64_t)r * 64 * ld;
        for (int c = 0; c < r; ++c) {
            double *Trc = row + (int64_t)c * 64;
            flow_wait(p.applied + (int64_t)r * nblk + c, flow_napp(c));
            if (c >= 1) {                                   // panel c-1 onto this tile
                flow_wait(p.solved + (int64_t)c * nblk + (c - 1), 1);
                flow_gemm<true>(Trc, ld, row + (int64_t)(c - 1) * 64, ld, Az + (int64_t)c * 64 * ld + (int64_t)(c - 1) * 64, ld, 64, smem);
            }
            if ((c & 1) && c + 1 <= r) {                    // idle until Dinv_c exists: panel c-1 onto the tile in column c+1
                flow_wait(p.applied + (int64_t)r * nblk + c + 1, flow_napp(c + 1));
                if (c + 1 < r) flow_wait(p.solved + (int64_t)(c + 1) * nblk + (c - 1), 1);
                flow_gemm<true>(row + (int64_t)(c + 1) * 64, ld, row + (int64_t)(c - 1) * 64, ld,
                                Az + (int64_t)(c + 1) * 64 * ld + (int64_t)(c - 1) * 64, ld, 64, smem);
            }
            __threadfence();                                // the tile is re-read below through cp.async (L2)
            flow_wait(p.solved + (int64_t)c * nblk + c, 1); // Dinv_c (also the barrier behind the fence)
            flow_gemm<false>(Trc, ld, Trc, ld, p.dinv + (int64_t)c * 4096, 64, 64, smem);
            flow_signal(p.solved + (int64_t)r * nblk + c, 1);
        }
        {   // diagonal block r
            PotrfSmem &sm = *reinterpret_cast<PotrfSmem *>(smem);
            double *P = smem + (sizeof(PotrfSmem) + 7) / 8;
            double *Xs = P + PF_TILE, *Ts = Xs + PF_TILE;
            const int tx = tid & 15, ty = tid >> 4;
            double *Ab = row + (int64_t)r * 64;
            flow_wait(p.applied + (int64_t)r * nblk + r, flow_napp(r));
            double ra[4][4];
#pragma unroll
            for (int ai = 0; ai < 4; ++ai)
#pragma unroll
                for (int b = 0; b < 4; ++b) ra[ai][b] = __ldcg(&Ab[(int64_t)(ty + 16 * ai) * ld + tx + 16 * b]);
            if (r >= 1) {
                typedef T64NT8 T;
                const double *Lp = row + (int64_t)(r - 1) * 64;
                const int warp = tid >> 5, lane = tid & 31;
                const int wm = warp / T::WARPS_N, wn = warp % T::WARPS_N, g = lane >> 2, t = lane & 3;
                double acc[T::MI][T::NI][2];
#pragma unroll
                for (int mi = 0; mi < T::MI; ++mi)
#pragma unroll
                    for (int ni = 0; ni < T::NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
                T::mainloop(acc, Lp, ld, Lp, ld, 0, 64, smem);           // ends with a barrier: smem is free again
#pragma unroll
                for (int mi = 0; mi < T::MI; ++mi)
#pragma unroll
                    for (int ni = 0; ni < T::NI; ++ni) {
                        const int rr = wm * T::WM + mi * 8 + g, cc = wn * T::WN + ni * 8 + 2 * t;
                        Xs[rr * PF_LD + cc] = acc[mi][ni][0];
                        Xs[rr * PF_LD + cc + 1] = acc[mi][ni][1];
                    }
                __syncthreads();
#pragma unroll
                for (int ai = 0; ai < 4; ++ai)
#pragma unroll
                    for (int b = 0; b < 4; ++b) ra[ai][b] -= Xs[(ty + 16 * ai) * PF_LD + tx + 16 * b];
                __syncthreads();
            }
            potrf64_regs(ra, sm, P, Xs, Ts, Ab, ld, p.dinv + (int64_t)r * 4096, p.info, r);
            flow_signal(p.solved + (int64_t)r * nblk + r, 1);
        }
    }
    // ===================== update worker =====================
    for (;;) {
        __syncthreads();                                    // s_task of the previous round fully consumed
        if (tid == 0) {
            const int t = atomicAdd(p.next, 1);
            // task t -> (pair q, tile (i, j)): pairs in order, inside a pair column by column
            int q = 0, rem = t, T = nblk - 3;
            while (T > 0 && rem >= T * (T + 1) / 2) { rem -= T * (T + 1) / 2; T -= 2; ++q; }
            if (T <= 0) {
                s_task[0] = -1;
            } else {
                int jj = 0;
                while (rem >= T - jj) { rem -= T - jj; ++jj; }
                s_task[0] = q;
                s_task[1] = 2 * q + 3 + jj + rem;           // i
                s_task[2] = 2 * q + 3 + jj;                 // j
            }
        }
        __syncthreads();
        const int q = s_task[0], i = s_task[1], j = s_task[2];
        if (q < 0) break;
        flow_wait(p.solved + (int64_t)i * nblk + 2 * q + 1, 1);          // row i of the pair (its even panel came first)
        if (j != i) flow_wait(p.solved + (int64_t)j * nblk + 2 * q + 1, 1);
        flow_wait(p.applied + (int64_t)i * nblk + j, q);                 // earlier pairs on this tile
        flow_gemm<true>(Az + (int64_t)i * 64 * ld + (int64_t)j * 64, ld, Az + (int64_t)i * 64 * ld + (int64_t)(2 * q) * 64, ld,
                        Az + (int64_t)j * 64 * ld + (int64_t)(2 * q) * 64, ld, 128, smem);
        flow_signal(p.applied + (int64_t)i * nblk + j, q + 1);
    }
}

#define CHOL_DIAG_SMEM ((int)(sizeof(PotrfSmem) + 8 + (2 * PF_TILE + 64 * 33) * 8))
#define CHOL_STEP_SMEM (T64NT8::SMEM_BYTES > CHOL_DIAG_SMEM ? T64NT8::SMEM_BYTES : CHOL_DIAG_SMEM)


// Trailing update, one 64 x 64 tile of the lower triangle per CTA: C -= P_m P_n^T (K = 64 or 128 panel columns).  The accumulators start as
// -C: the tile is fetched from global memory straight into registers while the operand stages are still in flight, and
// the result is written back negated, so the read of C is off the tail of the kernel (a launch of this kernel is
// latency-bound in the second half of the factorisation, where the triangle holds fewer tiles than the chip has SMs).
template <class T>
__global__ void __launch_bounds__(T::NTHREADS)
syrk_tri_kernel(double *A22, const double *P, int ld, int64_t strideA, int K) {
    extern __shared__ __align__(16) double smem[];
    // linear index -> (tm >= tn)
    int x = blockIdx.x;
    int tm = (int)((sqrt(8.0 * x + 1.0) - 1.0) * 0.5);
    while ((tm + 1) * (tm + 2) / 2 <= x) ++tm;
    while (tm * (tm + 1) / 2 > x) --tm;
    int tn = x - tm * (tm + 1) / 2;
    const double *Pb = P + blockIdx.z * strideA;
    double *Cb = A22 + blockIdx.z * strideA;
    const double *Ap = Pb + (int64_t)tm * T::BM * ld;
    const double *Bp = Pb + (int64_t)tn * T::BN * ld;
    double *C = Cb + (int64_t)tm * T::BM * ld + (int64_t)tn * T::BN;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp / T::WARPS_N, wn = warp % T::WARPS_N;
    const int g = lane >> 2, t = lane & 3;
    double acc[T::MI][T::NI][2];
#pragma unroll
    for (int mi = 0; mi < T::MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < T::NI; ++ni) {
            const int r = wm * T::WM + mi * 8 + g, c = wn * T::WN + ni * 8 + 2 * t;
            const double2 v = *reinterpret_cast<const double2 *>(&C[(int64_t)r * ld + c]);
            acc[mi][ni][0] = -v.x;
            acc[mi][ni][1] = -v.y;
        }
    T::mainloop(acc, Ap, ld, Bp, ld, 0, K, smem);
#pragma unroll
    for (int mi = 0; mi < T::MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < T::NI; ++ni) {
            int r = wm * T::WM + mi * 8 + g, c = wn * T::WN + ni * 8 + 2 * t;
            *reinterpret_cast<double2 *>(&C[(int64_t)r * ld + c]) = make_double2(-acc[mi][ni][0], -acc[mi][ni][1]);
        }
}

template <class K>
static int set_smem(bo_ctx *ctx, K kernel, int bytes) {
    BO_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    return BO_OK;
}

// In-place lower Cholesky of `batch` padded np x np matrices (np % 64 == 0).
// Right-looking with a one-panel lookahead: step c (chol_step_kernel, main high-priority stream)
// applies panel c-1 to block column c, factors the diagonal block and solves the panel in one
// launch, while the rest of the trailing update of panel c-1 (columns > c) runs on the side stream.
struct chol_lane {                 // one (main, side) stream pair and its hand-over events
    cudaStream_t main, side;
    cudaEvent_t sliced[2], consumed[2];
};

static int chol_enqueue_lane(bo_ctx *ctx, const chol_lane &ln, bool split, int np, int batch, double *A, double *dinv,
                             int *dInfo, int *flags) {
    const int nblk = np / BO_NB;
    const int64_t strideA = (int64_t)np * np, strideD = (int64_t)nblk * 4096;
    cudaStream_t main = ln.main, side = ln.side;
    auto step = [&](int c, int kprev) -> int {
        if (!split) {
            BO_LAUNCH_ON(ctx, "chol_step_kernel", main);
            chol_step_kernel<<<dim3(nblk - c, 1, batch), 256, CHOL_STEP_SMEM, main>>>(A, np, strideA, c, dinv, strideD, dInfo,
                                                                                      flags, nblk, 0, kprev);
            BO_CHECK_LAUNCH(ctx);
            return BO_OK;
        }
        {
            BO_LAUNCH_ON(ctx, "chol_diag_kernel", main);
            chol_step_kernel<<<dim3(1, 1, batch), 256, CHOL_STEP_SMEM, main>>>(A, np, strideA, c, dinv, strideD, dInfo,
                                                                               flags, nblk, 1, kprev);
            BO_CHECK_LAUNCH(ctx);
        }
        if (nblk - c - 1 > 0) {
            BO_LAUNCH_ON(ctx, "chol_panel_kernel", main);
            chol_step_kernel<<<dim3(nblk - c - 1, 1, batch), 256, T64NT8::SMEM_BYTES, main>>>(A, np, strideA, c, dinv, strideD, dInfo,
                                                                                             flags, nblk, 2, kprev);
            BO_CHECK_LAUNCH(ctx);
        }
        return BO_OK;
    };
    // Trailing updates run per GROUP of G panels (K = 64 G): a 64 x 64 tile of C is read and written once per 64 G panel
    // columns instead of once per 64 -- the update is bound by L2 traffic (C in, C out, two operand tiles per 0.5 MFLOP
    // at K = 64), not by the FP64 tensor pipe.  G = 2 for one matrix (the panel chain is the critical path and its own
    // pre-update grows with G), G = 4 for many matrices at once (the trailing update is).  Group p = panels [Gp, Gp+G):
    //   step(Gp)     applies group p-1 to block column Gp itself          (kprev = 64 G, nothing to wait for)
    //   step(Gp+i)   applies panels Gp..Gp+i-1 to block column Gp+i       (kprev = 64 i) after rest(p-1) is done with it
    //   rest(p)      A_ij -= L_i,grp L_j,grp^T for i >= j >= G(p+1)+1     (side stream, after the group's last step)
    const int G = ctx->chol_group > 0 ? ctx->chol_group : (split ? 4 : 2);
    auto rest = [&](int p) -> int {
        const int j0 = G * (p + 1) + 1, T = nblk - j0;
        if (T <= 0) return BO_OK;
        double *panel = A + (int64_t)j0 * BO_NB * np + (int64_t)(G * p) * BO_NB;
        double *A22 = A + (int64_t)j0 * BO_NB * (np + 1);
        BO_LAUNCH_ON(ctx, "chol_syrk_kernel", side);
        syrk_tri_kernel<T64NT><<<dim3(T * (T + 1) / 2, 1, batch), T64NT::NTHREADS, T64NT::SMEM_BYTES, side>>>(
            A22, panel, np, strideA, G * BO_NB);
        BO_CHECK_LAUNCH(ctx);
        return BO_OK;
    };
    if (nblk <= G) {
        for (int c = 0; c < nblk; ++c) BO_TRY(step(c, c * BO_NB));
        return BO_OK;
    }
    // fork: the side stream must see everything queued on the main stream so far (the input matrix)
    BO_CUDA(ctx, cudaEventRecord(ln.consumed[0], main));
    BO_CUDA(ctx, cudaStreamWaitEvent(side, ln.consumed[0], 0));
    const int ngrp = (nblk + G - 1) / G;
    for (int p = 0; p < ngrp; ++p) {
        const int e = p & 1;
        for (int i = 0; i < G && G * p + i < nblk; ++i) {
            if (i == 1 && p >= 1) BO_CUDA(ctx, cudaStreamWaitEvent(main, ln.consumed[e ^ 1], 0));   // rest(p-1) done
            BO_TRY(step(G * p + i, i == 0 ? (p > 0 ? G * BO_NB : 0) : i * BO_NB));
        }
        BO_CUDA(ctx, cudaEventRecord(ln.sliced[e], main));            // group p complete
        BO_CUDA(ctx, cudaStreamWaitEvent(side, ln.sliced[e], 0));
        BO_TRY(rest(p));
        BO_CUDA(ctx, cudaEventRecord(ln.consumed[e], side));          // rest(p) done
    }
    BO_CUDA(ctx, cudaStreamWaitEvent(main, ln.consumed[(ngrp - 1) & 1], 0));   // join
    return BO_OK;
}

// extra stream pairs + events of the batched factorisation (created once, outside any stream capture)
static int chol_lanes_init(bo_ctx *ctx) {
    if (ctx->chol_lane_fork) return BO_OK;
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    for (int l = 0; l < BO_CHOL_MAX_LANES - 1; ++l) {
        BO_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->chol_lane_main[l], cudaStreamNonBlocking, hi));
        BO_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->chol_lane_side[l], cudaStreamNonBlocking, lo));
        for (int i = 0; i < 5; ++i) BO_CUDA(ctx, cudaEventCreateWithFlags(&ctx->chol_lane_ev[l][i], cudaEventDisableTiming));
    }
    BO_CUDA(ctx, cudaEventCreateWithFlags(&ctx->chol_lane_fork, cudaEventDisableTiming));
    return BO_OK;
}

// Many matrices at once are factored as independent sub-batches on their own stream pairs: a sub-batch that sits in
// its diagonal-block step (one CTA per matrix) leaves the SMs to the others' trailing updates.
static int chol_enqueue(bo_ctx *ctx, int np, int batch, double *A, double *dinv, int *dInfo) {
    const int nblk = np / BO_NB;
    const int64_t strideA = (int64_t)np * np, strideD = (int64_t)nblk * 4096;
    BO_CUDA(ctx, cudaMemsetAsync(dInfo, 0, sizeof(int) * batch, ctx->stream));
    BO_CUDA(ctx, cudaMemsetAsync(ctx->dCholFlags, 0, sizeof(int) * batch * nblk, ctx->stream));
    // many matrices at once: diagonal and panel as two launches (no CTA spins on the flag); see chol_step_kernel
    const bool split = (int64_t)batch * nblk > 2 * (int64_t)ctx->sm_count;
    chol_lane l0{ctx->stream, ctx->stream2, {ctx->ev_sliced[0], ctx->ev_sliced[1]}, {ctx->ev_consumed[0], ctx->ev_consumed[1]}};
    static const int max_lanes = getenv("BO_CHOL_LANES") ? std::min(BO_CHOL_MAX_LANES, std::max(1, atoi(getenv("BO_CHOL_LANES")))) : 2;
    const int lanes = split ? std::min(max_lanes, batch / 2) : 1;
    if (lanes <= 1) return chol_enqueue_lane(ctx, l0, split, np, batch, A, dinv, dInfo, ctx->dCholFlags);
    BO_TRY(chol_lanes_init(ctx));
    BO_CUDA(ctx, cudaEventRecord(ctx->chol_lane_fork, ctx->stream));                         // fork
    int b0 = 0;
    for (int l = 0; l < lanes; ++l) {
        const int bl = (batch - b0) / (lanes - l);
        if (l == lanes - 1) {
            BO_TRY(chol_enqueue_lane(ctx, l0, split, np, bl, A + b0 * strideA, dinv + b0 * strideD, dInfo + b0,
                                     ctx->dCholFlags + (int64_t)b0 * nblk));
        } else {
            cudaEvent_t *ev = ctx->chol_lane_ev[l];
            chol_lane ln{ctx->chol_lane_main[l], ctx->chol_lane_side[l], {ev[0], ev[1]}, {ev[2], ev[3]}};
            BO_CUDA(ctx, cudaStreamWaitEvent(ln.main, ctx->chol_lane_fork, 0));
            BO_TRY(chol_enqueue_lane(ctx, ln, split, np, bl, A + b0 * strideA, dinv + b0 * strideD, dInfo + b0,
                                     ctx->dCholFlags + (int64_t)b0 * nblk));
            BO_CUDA(ctx, cudaEventRecord(ev[4], ln.main));
        }
        b0 += bl;
    }
    for (int l = 0; l < lanes - 1; ++l) BO_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->chol_lane_ev[l][4], 0));   // join
    return BO_OK;
}

// The factorisation is a fixed pattern of ~3 nblk launches and ~4 nblk event edges across two streams; issued one by one
// the host spends longer queueing them than the device needs to run the serial chain.  The whole pattern is therefore
// captured once per (matrix address, size, batch) into a CUDA graph and replayed (a handle refactors the same buffers
// on every bo_fit / bo_loglik_fit); the event profiler needs real launches, so profiling runs bypass the graph.
void bo_linalg_drop_graphs(bo_ctx *ctx) {
    for (auto &g : ctx->chol_graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    ctx->chol_graphs.clear();
}

// one matrix: the persistent dataflow kernel (chol_flow_kernel)
static int chol_flow(bo_ctx *ctx, int np, double *A, double *dinv, int *dInfo, int grid) {
    const int nblk = np / BO_NB;
    const size_t nflags = 2 * (size_t)nblk * nblk + 8;
    BO_TRY(bo_reserve(ctx, &ctx->dCholFlags, &ctx->cholflags_capacity, nflags));
    BO_CUDA(ctx, cudaMemsetAsync(dInfo, 0, sizeof(int), ctx->stream));
    BO_CUDA(ctx, cudaMemsetAsync(ctx->dCholFlags, 0, sizeof(int) * nflags, ctx->stream));
    CholFlowParams prm;
    prm.A = A; prm.ld = np; prm.nblk = nblk; prm.dinv = dinv; prm.info = dInfo;
    prm.solved = ctx->dCholFlags;
    prm.applied = ctx->dCholFlags + (size_t)nblk * nblk;
    prm.next = ctx->dCholFlags + 2 * (size_t)nblk * nblk;
    void *args[] = {&prm};
    BO_LAUNCH(ctx, "chol_flow_kernel");
    BO_CUDA(ctx, cudaLaunchCooperativeKernel((const void *)chol_flow_kernel, dim3(grid), dim3(256), args, CHOL_STEP_SMEM, ctx->stream));
    return BO_OK;
}

int bo_linalg_cholesky(bo_ctx *ctx, int np, int batch, double *A, double *dinv, int *dInfo) {
    const int nblk = np / BO_NB;
    static const bool use_flow = !(getenv("BO_CHOL_FLOW") && atoi(getenv("BO_CHOL_FLOW")) == 0);
    // (measured: 2.07 vs 2.2 ms at n = 4096; at n <= 2048 the chain of 45 k cycles per diagonal block dominates either
    //  way and the launch-per-step form, replayed as a graph, is a few per cent ahead)
    static const int flow_min = getenv("BO_CHOL_FLOW_MIN") ? atoi(getenv("BO_CHOL_FLOW_MIN")) : 48;   // (tests / sanitizer lower it)
    if (use_flow && batch == 1 && nblk >= flow_min && nblk >= 4 && ctx->chol_flow_grid >= nblk + 16)
        return chol_flow(ctx, np, A, dinv, dInfo, ctx->chol_flow_grid);
    BO_TRY(bo_reserve(ctx, &ctx->dCholFlags, &ctx->cholflags_capacity, (size_t)batch * nblk));
    static const bool use_graph = !(getenv("BO_CHOL_GRAPH") && atoi(getenv("BO_CHOL_GRAPH")) == 0);
    // (measured: the replayed graph wins 7 % at n <= 2048, loses 7 % at n = 4096 where the chain is bound by the device)
    if (!use_graph || ctx->prof_on || nblk < 3 || nblk > 32 || batch > 16) return chol_enqueue(ctx, np, batch, A, dinv, dInfo);
    for (auto &g : ctx->chol_graphs)
        if (g.A == A && g.dinv == dinv && g.info == dInfo && g.flags == ctx->dCholFlags && g.np == np && g.batch == batch) {
            g.stamp = ++ctx->chol_graph_clock;
            ctx->launches += g.launches;
            BO_CUDA(ctx, cudaGraphLaunch(g.exec, ctx->stream));
            return BO_OK;
        }
    // capture only a pattern that comes back (one-shot callers hand in a fresh scratch matrix every time)
    {
        bool seen = false;
        for (auto &k : ctx->chol_seen) seen = seen || (k.A == A && k.dinv == dinv && k.info == dInfo && k.np == np && k.batch == batch);
        if (!seen) {
            bo_chol_graph k;
            k.A = A; k.dinv = dinv; k.info = dInfo; k.np = np; k.batch = batch;
            if (ctx->chol_seen.size() >= 8) ctx->chol_seen.erase(ctx->chol_seen.begin());
            ctx->chol_seen.push_back(k);
            return chol_enqueue(ctx, np, batch, A, dinv, dInfo);
        }
    }
    const int64_t l0 = ctx->launches;
    BO_TRY(chol_lanes_init(ctx));
    BO_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = chol_enqueue(ctx, np, batch, A, dinv, dInfo);
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
    if (rc != BO_OK) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
    }
    if (e != cudaSuccess) return bo_set_err(ctx, BO_ERR_CUDA, "Cholesky graph capture failed: %s", cudaGetErrorString(e));
    bo_chol_graph g;
    g.A = A; g.dinv = dinv; g.info = dInfo; g.flags = ctx->dCholFlags; g.np = np; g.batch = batch;
    g.launches = ctx->launches - l0;
    g.stamp = ++ctx->chol_graph_clock;
    const cudaError_t ei = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ei != cudaSuccess) return bo_set_err(ctx, BO_ERR_CUDA, "Cholesky graph instantiation failed: %s", cudaGetErrorString(ei));
    if (ctx->chol_graphs.size() >= 4) {          // keep the four most recently used patterns
        size_t old = 0;
        for (size_t i = 1; i < ctx->chol_graphs.size(); ++i)
            if (ctx->chol_graphs[i].stamp < ctx->chol_graphs[old].stamp) old = i;
        cudaGraphExecDestroy(ctx->chol_graphs[old].exec);
        ctx->chol_graphs.erase(ctx->chol_graphs.begin() + old);
    }
    ctx->chol_graphs.push_back(g);
    BO_CUDA(ctx, cudaGraphLaunch(g.exec, ctx->stream));
    return BO_OK;
}

// ---------------------------------------------------------------------------
// W = L^-1 by recursive halving over 64-blocks:  W21 = -W22 (L21 W11).
// All nodes of one depth run in one launch; a node's block range is recovered
// arithmetically from (depth, index) by replaying the floor-halving.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void node_range(int nblk, int depth, int idx, int &lo, int &mid, int &hi) {
    lo = 0;
    hi = nblk;
    for (int l = depth - 1; l >= 0; --l) {
        int m = (lo + hi) >> 1;
        if ((idx >> l) & 1) lo = m; else hi = m;
        if (hi - lo < 1) break;
    }
    mid = (lo + hi) >> 1;
}

// phase 0: T21 = L21 * W11  (W11 lower: k >= column block)
// phase 1: W21 = -W22 * T21 (W22 lower: k <= row block)
template <int PHASE>
__global__ void __launch_bounds__(T64NN::NTHREADS)
trtri_node_kernel(const double *L, double *W, double *Tm, int np, int nblk, int depth, int max_side) {
    extern __shared__ __align__(16) double smem[];
    typedef T64NN T;
    const int node = blockIdx.y;
    const int64_t boff = (int64_t)blockIdx.z * np * np;
    int lo, mid, hi;
    node_range(nblk, depth, node, lo, mid, hi);
    if (hi - lo < 2) return;
    const int rows = hi - mid, cols = mid - lo;
    const int tm = blockIdx.x / max_side, tn = blockIdx.x % max_side;
    if (tm >= rows || tn >= cols) return;
    const double *A, *B;
    double *C;
    int kbeg, kend;
    if (PHASE == 0) {
        A = L + boff + (int64_t)(mid + tm) * 64 * np + (int64_t)lo * 64;     // L21 row tile
        B = W + boff + (int64_t)lo * 64 * np + (int64_t)(lo + tn) * 64;       // W11 column tile
        C = Tm + boff + (int64_t)(mid + tm) * 64 * np + (int64_t)(lo + tn) * 64;
        kbeg = tn * 64;
        kend = cols * 64;
    } else {
        A = W + boff + (int64_t)(mid + tm) * 64 * np + (int64_t)mid * 64;    // W22 row tile
        B = Tm + boff + (int64_t)mid * 64 * np + (int64_t)(lo + tn) * 64;     // T21 column tile
        C = W + boff + (int64_t)(mid + tm) * 64 * np + (int64_t)(lo + tn) * 64;
        kbeg = 0;
        kend = (tm + 1) * 64;
    }
    double acc[T::MI][T::NI][2];
#pragma unroll
    for (int mi = 0; mi < T::MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < T::NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    T::mainloop(acc, A, np, B, np, kbeg, kend, smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp / T::WARPS_N, wn = warp % T::WARPS_N;
    const int g = lane >> 2, t = lane & 3;
    const double sgn = (PHASE == 0) ? 1.0 : -1.0;
#pragma unroll
    for (int mi = 0; mi < T::MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < T::NI; ++ni) {
            int r = wm * T::WM + mi * 8 + g, c = wn * T::WN + ni * 8 + 2 * t;
            *reinterpret_cast<double2 *>(&C[(int64_t)r * np + c]) =
                make_double2(sgn * acc[mi][ni][0], sgn * acc[mi][ni][1]);
        }
}

__global__ void trtri_seed_kernel(const double *dinv, double *W, int np, int nblk) {
    // W := 0 except the diagonal 64-blocks := inverted diagonal blocks
    const int64_t boff = (int64_t)blockIdx.z * np * np;
    const int bi = blockIdx.y, bj = blockIdx.x;
    double *dst = W + boff + (int64_t)bi * 64 * np + (int64_t)bj * 64;
    const double *src = dinv + ((int64_t)blockIdx.z * nblk + bi) * 4096;
    for (int e = threadIdx.x; e < 4096; e += blockDim.x) {
        int i = e >> 6, j = e & 63;
        dst[(int64_t)i * np + j] = (bi == bj) ? src[e] : 0.0;
    }
}

int bo_linalg_trtri(bo_ctx *ctx, int np, int batch, const double *L, const double *dinv,
                    double *W, double *tmp) {
    const int nblk = np / 64;
    {
        BO_LAUNCH(ctx, "trtri_seed_kernel");
        trtri_seed_kernel<<<dim3(nblk, nblk, batch), 256, 0, ctx->stream>>>(dinv, W, np, nblk);
        BO_CHECK_LAUNCH(ctx);
    }
    int depth_max = 0;
    while ((1 << depth_max) < nblk) ++depth_max;   // leaves live at depth <= depth_max
    for (int depth = depth_max - 1; depth >= 0; --depth) {
        const int nodes = 1 << depth;
        int max_side = (nblk + nodes - 1) / nodes;      // upper bound on hi - lo
        max_side = (max_side + 1) / 2 + 1;               // upper bound on rows / cols
        dim3 grd(max_side * max_side, nodes, batch);
        {
            BO_LAUNCH(ctx, "trtri_lw_kernel");
            trtri_node_kernel<0><<<grd, T64NN::NTHREADS, T64NN::SMEM_BYTES, ctx->stream>>>(
                L, W, tmp, np, nblk, depth, max_side);
            BO_CHECK_LAUNCH(ctx);
        }
        {
            BO_LAUNCH(ctx, "trtri_wt_kernel");
            trtri_node_kernel<1><<<grd, T64NN::NTHREADS, T64NN::SMEM_BYTES, ctx->stream>>>(
                L, W, tmp, np, nblk, depth, max_side);
            BO_CHECK_LAUNCH(ctx);
        }
    }
    return BO_OK;
}

// ---------------------------------------------------------------------------
// transpose, alpha = W r, beta = W^T alpha, logdet
// ---------------------------------------------------------------------------
__global__ void transpose_kernel(const double *A, double *AT, int np) {
    __shared__ double tile[32][33];
    const int64_t boff = (int64_t)blockIdx.z * np * np;
    int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) tile[r][threadIdx.x] = A[boff + (int64_t)(y0 + r) * np + x];
    __syncthreads();
    int xo = blockIdx.y * 32 + threadIdx.x, yo0 = blockIdx.x * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) AT[boff + (int64_t)(yo0 + r) * np + xo] = tile[threadIdx.x][r];
}

int bo_linalg_transpose(bo_ctx *ctx, int np, int batch, const double *A, double *AT) {
    BO_LAUNCH(ctx, "transpose_kernel");
    transpose_kernel<<<dim3(np / 32, np / 32, batch), dim3(32, 8), 0, ctx->stream>>>(A, AT, np);
    BO_CHECK_LAUNCH(ctx);
    return BO_OK;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// out[s][i] = sum_{j in [jlo(i), jhi(i))} Mx[s][i][j] * (vec[s][j] - shift[s] if j < nshift)
// mode 0: lower rows (j <= i) with r = y - bias;  mode 1: upper rows (j >= i) with alpha.
__global__ void trimv_kernel(int mode, const double *Mx, const double *vec, const double *bias,
                             int n, int np, double *out) {
    const int s = blockIdx.y;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= np) return;
    const double *mrow = Mx + (int64_t)s * np * np + (int64_t)row * np;
    double acc = 0.0;
    if (mode == 0) {
        const double b = bias[s];
        for (int j = lane; j <= row && j < n; j += 32) acc = fma(mrow[j], vec[j] - b, acc);
    } else {
        const double *v = vec + (int64_t)s * np;
        for (int j = row + lane - (row & 31); j < np; j += 32)
            if (j >= row) acc = fma(mrow[j], v[j], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) out[(int64_t)s * np + row] = acc;
}

__global__ void logdet_kernel(const double *L, int n, int np, double *out) {
    __shared__ double red[32];
    const int s = blockIdx.x;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += log(L[(int64_t)s * np * np + (int64_t)i * (np + 1)]);
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
        v = warp_sum(v);
        if (threadIdx.x == 0) out[s] = v;
    }
}

int bo_linalg_finish_fit(bo_ctx *ctx) {
    const int np = ctx->np, S = ctx->S;
    {   // alpha_s = W_s (y - bias_s)
        BO_LAUNCH(ctx, "trimv_kernel");
        trimv_kernel<<<dim3(np / 8, S), 256, 0, ctx->stream>>>(0, ctx->dW, ctx->dY, ctx->dBias, ctx->n, np, ctx->dAlpha);
        BO_CHECK_LAUNCH(ctx);
    }
    {   // beta_s = W_s^T alpha_s
        BO_LAUNCH(ctx, "trimv_kernel");
        trimv_kernel<<<dim3(np / 8, S), 256, 0, ctx->stream>>>(1, ctx->dWT, ctx->dAlpha, ctx->dBias, ctx->n, np, ctx->dBeta);
        BO_CHECK_LAUNCH(ctx);
    }
    {
        BO_LAUNCH(ctx, "logdet_kernel");
        logdet_kernel<<<S, 256, 0, ctx->stream>>>(ctx->dL, ctx->n, np, ctx->dLogdet);
        BO_CHECK_LAUNCH(ctx);
    }
    return BO_OK;
}

int bo_linalg_init(bo_ctx *ctx) {
    BO_TRY(set_smem(ctx, dgemm_kernel<T64NT>, T64NT::SMEM_BYTES));
    BO_TRY(set_smem(ctx, chol_step_kernel, CHOL_STEP_SMEM));
    BO_TRY(set_smem(ctx, chol_flow_kernel, CHOL_STEP_SMEM));
    {
        int per_sm = 0, coop = 0;
        BO_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, chol_flow_kernel, 256, CHOL_STEP_SMEM));
        BO_CUDA(ctx, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
        ctx->chol_flow_grid = coop ? per_sm * ctx->sm_count : 0;
    }
    if (const char *e = getenv("BO_CHOL_GROUP")) {      // panels per trailing update (A/B knob; 0 = per-shape default)
        int g = atoi(e);
        ctx->chol_group = (g >= 1 && g <= 8) ? g : 0;
    }
    BO_TRY(set_smem(ctx, syrk_tri_kernel<T64NT>, T64NT::SMEM_BYTES));
    BO_TRY(set_smem(ctx, trtri_node_kernel<0>, T64NN::SMEM_BYTES));
    BO_TRY(set_smem(ctx, trtri_node_kernel<1>, T64NN::SMEM_BYTES));
    return BO_OK;
}
