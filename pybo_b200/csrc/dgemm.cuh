// dgemm.cuh -- FP64 tensor-core (DMMA m8n8k4) tile main loop shared by the
// Cholesky trailing update, the blocked triangular inverse and the scoring
// contraction V = W K*.  cp.async multi-stage pipeline, conflict-free padded
// shared-memory layouts (row stride == 4 mod 16 doubles).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// D(8x8) += A(8x4, row) * B(4x8, col); lane = 4*g + t holds
// a = A[g][t], b = B[t][g], c0/c1 = C[g][2t], C[g][2t+1].
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile(
        "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// B_NK == true : B is stored [n][k] (k contiguous)  -> C = A * B^T  ("NT")
// B_NK == false: B is stored [k][n] (n contiguous)  -> C = A * B    ("NN")
template <int BM_, int BN_, int WM_, int WN_, int STAGES_, bool B_NK_>
struct DTile {
    static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_, STAGES = STAGES_;
    static constexpr bool B_NK = B_NK_;
    static constexpr int BK = 16;
    static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
    static constexpr int NTHREADS = 32 * WARPS_M * WARPS_N;
    static constexpr int MI = WM / 8, NI = WN / 8;
    static constexpr int LDA_S = BK + 4;
    static constexpr int LDB_S = B_NK ? (BK + 4) : (BN + 4);
    static constexpr int A_STAGE = BM * LDA_S;
    static constexpr int B_STAGE = B_NK ? (BN * LDB_S) : (BK * LDB_S);
    static constexpr int SMEM_DOUBLES = STAGES * (A_STAGE + B_STAGE);
    static constexpr int SMEM_BYTES = SMEM_DOUBLES * 8;

    // A points at the tile's first row (row-major, leading dimension lda);
    // B points at the tile's first column (NN) or first row (NT).
    __device__ static __forceinline__ void load_stage(double *As, double *Bs, const double *A,
                                                      int64_t lda, const double *B, int64_t ldb,
                                                      int k0, int tid) {
#pragma unroll
        for (int c = tid; c < BM * 8; c += NTHREADS) {
            int r = c >> 3, cc = (c & 7) * 2;
            cp_async16(&As[r * LDA_S + cc], &A[(int64_t)r * lda + k0 + cc]);
        }
        if (B_NK) {
#pragma unroll
            for (int c = tid; c < BN * 8; c += NTHREADS) {
                int r = c >> 3, cc = (c & 7) * 2;
                cp_async16(&Bs[r * LDB_S + cc], &B[(int64_t)r * ldb + k0 + cc]);
            }
        } else {
#pragma unroll
            for (int c = tid; c < BK * (BN / 2); c += NTHREADS) {
                int r = c / (BN / 2), cc = (c % (BN / 2)) * 2;
                cp_async16(&Bs[r * LDB_S + cc], &B[(int64_t)(k0 + r) * ldb + cc]);
            }
        }
    }

    // acc[mi][ni][e] accumulates rows wm*WM + mi*8 + g, cols wn*WN + ni*8 + 2t + e.
    __device__ static __forceinline__ void mainloop(double (&acc)[MI][NI][2], const double *A,
                                                    int64_t lda, const double *B, int64_t ldb,
                                                    int kbeg, int kend, double *smem) {
        const int tid = threadIdx.x;
        const int warp = tid >> 5, lane = tid & 31;
        const int wm = warp / WARPS_N, wn = warp % WARPS_N;
        const int g = lane >> 2, t = lane & 3;
        double *As = smem;
        double *Bs = smem + STAGES * A_STAGE;
        const int nk = (kend - kbeg) / BK;

#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s) {
            if (s < nk) load_stage(As + s * A_STAGE, Bs + s * B_STAGE, A, lda, B, ldb, kbeg + s * BK, tid);
            cp_async_commit();
        }
        for (int it = 0; it < nk; ++it) {
            cp_async_wait<STAGES - 2>();
            __syncthreads();
            {
                int nx = it + STAGES - 1;
                if (nx < nk) {
                    int st = nx % STAGES;
                    load_stage(As + st * A_STAGE, Bs + st * B_STAGE, A, lda, B, ldb, kbeg + nx * BK, tid);
                }
                cp_async_commit();
            }
            const double *as = As + (it % STAGES) * A_STAGE + (wm * WM + g) * LDA_S + t;
            const double *bs = B_NK ? (Bs + (it % STAGES) * B_STAGE + (wn * WN + g) * LDB_S + t)
                                    : (Bs + (it % STAGES) * B_STAGE + t * LDB_S + wn * WN + g);
#pragma unroll
            for (int kk = 0; kk < BK; kk += 4) {
                double a[MI], b[NI];
#pragma unroll
                for (int mi = 0; mi < MI; ++mi) a[mi] = as[mi * 8 * LDA_S + kk];
#pragma unroll
                for (int ni = 0; ni < NI; ++ni)
                    b[ni] = B_NK ? bs[ni * 8 * LDB_S + kk] : bs[kk * LDB_S + ni * 8];
#pragma unroll
                for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                    for (int ni = 0; ni < NI; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
            }
        }
        cp_async_wait<0>();
        __syncthreads();
    }
};

// K-range rules for triangular operands (all in elements, multiples of BK).
enum { KR_FULL = 0, KR_A_LOWER = 1, KR_B_LOWER = 2, KR_A_UPPER = 3 };

struct DGemmParams {
    const double *A;
    const double *B;
    double *C;
    int64_t lda, ldb, ldc;
    int64_t strideA, strideB, strideC;   // per blockIdx.z
    int64_t strideA2, strideB2, strideC2; // per outer batch (blockIdx.z / inner)
    int inner;                            // blockIdx.z = outer * inner + in
    int tiles_m, tiles_n;
    int K;
    int krule;
    double alpha, beta;
};

// Generic batched C = alpha * A op(B) + beta * C on full tiles (no edge handling:
// every dimension is padded by the caller).
template <class T>
__global__ void __launch_bounds__(T::NTHREADS) dgemm_kernel(DGemmParams p) {
    extern __shared__ __align__(16) double smem[];
    const int tm = blockIdx.x / p.tiles_n, tn = blockIdx.x % p.tiles_n;
    const int outer = blockIdx.z / p.inner, in = blockIdx.z % p.inner;
    const double *A = p.A + outer * p.strideA2 + in * p.strideA + (int64_t)tm * T::BM * p.lda;
    const double *B = p.B + outer * p.strideB2 + in * p.strideB +
                      (T::B_NK ? (int64_t)tn * T::BN * p.ldb : (int64_t)tn * T::BN);
    double *C = p.C + outer * p.strideC2 + in * p.strideC + (int64_t)tm * T::BM * p.ldc + (int64_t)tn * T::BN;
    int kbeg = 0, kend = p.K;
    if (p.krule == KR_A_LOWER) kend = min(p.K, (tm + 1) * T::BM);
    if (p.krule == KR_B_LOWER) kbeg = tn * T::BN;
    if (p.krule == KR_A_UPPER) kbeg = tm * T::BM;

    double acc[T::MI][T::NI][2];
#pragma unroll
    for (int mi = 0; mi < T::MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < T::NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    T::mainloop(acc, A, p.lda, B, p.ldb, kbeg, kend, smem);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp / T::WARPS_N, wn = warp % T::WARPS_N;
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int mi = 0; mi < T::MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < T::NI; ++ni) {
            int r = wm * T::WM + mi * 8 + g, c = wn * T::WN + ni * 8 + 2 * t;
            double2 *dst = reinterpret_cast<double2 *>(&C[(int64_t)r * p.ldc + c]);
            double2 v = make_double2(p.alpha * acc[mi][ni][0], p.alpha * acc[mi][ni][1]);
            if (p.beta != 0.0) {
                double2 old = *dst;
                v.x += p.beta * old.x;
                v.y += p.beta * old.y;
            }
            *dst = v;
        }
}
