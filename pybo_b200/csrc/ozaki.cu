// ozaki.cu -- error-bounded int8-slice emulation of the scoring contraction
// V = W K* on the 5th-generation tensor cores (tcgen05.mma kind::i8, TMEM
// accumulators, TMA-staged operands).
//
// FP64 tensor cores cap the contraction at 37 TFLOP/s; the int8 path of the same
// SM runs 128x faster.  Both operands are cut into S balanced base-256 digits (the full
// int8 range [-128, 127]) on a fixed-point grid (rows of W share an exponent, K*/rho lies in [0,1]):
//     W_ij  = 2^e_i / 127 * sum_s a_s[i][j] 256^-s,   K*_jm = rho / 127 * sum_t b_t[m][j] 256^-t
// every int8 x int8 product and its int32 accumulation over j is exact, products are
// grouped by g = s + t (pairs with g >= S are dropped), and
//     v_im = 2^e_i rho / 127^2 * sum_g 256^-g D_g[m][i]
// is reassembled in FP64 in the epilogue, where |v|^2 is reduced.
// Digits: x in [-1, 1] -> X = rint(x * 127 * 256^(P-1)) (P = 7 bytes, or 5 in the fast slicer);
// byte k of (X + sum_k 128 * 256^k), minus 128, is the balanced digit k -- one 64-bit add and one
// XOR, no carry chain -- and the top S bytes are kept (dropping balanced low digits rounds to
// nearest).  int32 accumulators hold S pairs * n * 2^14 < 2^31 for n * S < 2^17.
// The truncation error is bounded a priori (see bo_ozaki_choose_slices).
//
// Orientation: candidates are the MMA M dimension (one TMEM lane = one candidate, so
// the row reductions are per-thread sums), rows of W are the N dimension.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

static int ctx_padded_dim(int d) {
    int dp = 2;
    while (dp < d) dp *= 2;
    return dp;
}

#define OZ_BM 128        // candidates per CTA tile (MMA M, TMEM lanes)
#define OZ_BN 64         // rows of W per block (MMA N)
#define OZ_BK 64         // k bytes per stage row (one 64B swizzle atom, two K=32 MMAs)
#define OZ_MAX_S 7
#define OZ_DIGIT_BIAS7 0x0080808080808080ll   // 128 in each of the 7 digit bytes
#define OZ_A_SLICE_BYTES (OZ_BM * OZ_BK)   // 8192
#define OZ_B_SLICE_BYTES (OZ_BN * OZ_BK)   // 4096
#define OZ_DEFAULT_CLUSTER 1   // CTAs per cluster of the scoring contraction (BO_OZ_CLUSTER overrides)
#define OZ_THREADS 320   // warp 0: TMA producer, warp 1: MMA issuer, warps 2-9: epilogue (two per TMEM lane quarter)
#define OZ_THREADS_M1 OZ_THREADS
#define OZ_ROW_SMEM 8192      // MODE 1: row scale + bias of up to 512 draws
#define OZ_ARG_LD 137         // MODE 1: leading dimension (doubles) of the 16 x 128 arg-max transpose buffer; candidate p sits at
                              // p + (p >> 4) so that the eight 16-candidate parts of a column start in different banks
#define OZ_ARG_SMEM (2 * 16 * OZ_ARG_LD * 8)

// ---------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra.uni DONE_%=;\n\t"
        "bra.uni WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const CUtensorMap *tmap, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// 1-D bulk copy global -> shared, completion on an mbarrier (UBLKCP)
__device__ __forceinline__ void bulk_load(uint32_t smem_dst, const void *gsrc, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
}

// K*^T slice planes live in HBM as tile-contiguous blocks: block (slice s, candidate tile ct,
// k block kb) is the 8 KB shared-memory image of a 128 x 64-byte K-major SWIZZLE_64B operand
// tile (16-byte chunk c of row r sits at chunk c ^ ((r >> 1) & 3)).  The slicer writes whole
// blocks with coalesced stores; the contraction fetches one with a single bulk copy.
__host__ __device__ __forceinline__ size_t oz_kss_block(int s, int ct, int kb, int ntiles, int nkb) {
    return (((size_t)s * ntiles + ct) * nkb + kb) * (size_t)OZ_A_SLICE_BYTES;
}
__host__ __device__ __forceinline__ size_t oz_kss_offset(int s, int m, int j, int ntiles, int nkb) {
    const int r = m & 127, c = (j & 63) >> 4;
    return oz_kss_block(s, m >> 7, j >> 6, ntiles, nkb) + (size_t)r * 64 + (size_t)((c ^ ((r >> 1) & 3)) << 4) + (j & 15);
}

// 1-D bulk copy multicast to the CTAs of the cluster in `mask`: the bytes land at the same CTA-relative offset in each
// destination CTA and complete_tx is signalled on the mbarrier at the same offset in each of them
__device__ __forceinline__ void bulk_load_multicast(uint32_t smem_dst, const void *gsrc, uint32_t bytes, uint32_t bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, int8 x int8 -> int32
__device__ __forceinline__ void umma_i8(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// the same arrival delivered to the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_multicast(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int32_t (&r)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile, rows of 64 bytes, SWIZZLE_64B: 8-row groups are 512 B apart.
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);        // start address
    d |= (uint64_t)1 << 16;                             // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(512 >> 4) << 32;                    // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
    d |= (uint64_t)4 << 61;                             // SWIZZLE_64B
    return d;
}

// kind::i8 instruction descriptor: D = s32, A = B = signed int8, both K-major
__host__ __device__ constexpr uint32_t make_idesc_i8(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// high word of the SWIZZLE_64B K-major descriptor is constant; the low word is (addr >> 4) | LBO
#define OZ_DESC_HI ((uint32_t)(512 >> 4) | (1u << 14) | (4u << 29))
__device__ __forceinline__ uint64_t oz_desc(uint32_t smem_addr) {
    return ((uint64_t)OZ_DESC_HI << 32) | (uint64_t)(((smem_addr >> 4) & 0x3FFFu) | (1u << 16));
}

// All MMAs of one pipeline stage.  For a fixed A slice s the B slices t = 0..S-1-s are
// contiguous in shared memory and their accumulators (groups s..S-1) are contiguous in TMEM,
// so they are issued as ONE instruction of N = 64 (S - s) columns (split at 256): the A tile is
// read from shared memory once per s instead of once per (s, t) pair.
// With EXTRA the pairs of group g = S (s + t = S, s, t >= 1) are accumulated too, into one more
// 64-column accumulator group: the dominant dropped-pair error term disappears at +27 % MMA work
// (S = 5: 19 pairs instead of 15) without re-slicing either operand.
template <int S, int EXTRA>
__device__ __forceinline__ void oz_issue_stage(uint32_t sA, uint32_t sB, uint32_t tacc, bool first) {
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
        int s0 = 0;
        if (EXTRA && kk == 0 && first) {
            // Accumulator initialisation when the extra group exists: group S is only written by
            // slices s >= 1, so the first MMA is slice 1 over groups 1..S with accumulate = 0; slice 0
            // then initialises group 0 on its own and accumulates into groups 1..S-1.
            const uint64_t a1 = oz_desc(sA + 1 * OZ_A_SLICE_BYTES);
            const int n1tot = OZ_BN * S;
            umma_i8(tacc + OZ_BN, a1, oz_desc(sB), make_idesc_i8(OZ_BM, n1tot > 256 ? 256 : n1tot), 0u);
            if (n1tot > 256)
                umma_i8(tacc + OZ_BN + 256, a1, oz_desc(sB + 256 * OZ_BK), make_idesc_i8(OZ_BM, n1tot - 256), 0u);
            const uint64_t a0 = oz_desc(sA);
            umma_i8(tacc, a0, oz_desc(sB), make_idesc_i8(OZ_BM, OZ_BN), 0u);
            if (S > 1) {
                const int n0 = OZ_BN * (S - 1);
                umma_i8(tacc + OZ_BN, a0, oz_desc(sB + OZ_B_SLICE_BYTES), make_idesc_i8(OZ_BM, n0 > 256 ? 256 : n0), 1u);
                if (n0 > 256)
                    umma_i8(tacc + OZ_BN + 256, a0, oz_desc(sB + OZ_B_SLICE_BYTES + 256 * OZ_BK),
                            make_idesc_i8(OZ_BM, n0 - 256), 1u);
            }
            s0 = 2;
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
            if (s < s0) continue;
            const int ntot = OZ_BN * (S - s + ((EXTRA && s >= 1) ? 1 : 0));
            const int n1 = ntot > 256 ? 256 : ntot;
            const uint64_t adesc = oz_desc(sA + s * OZ_A_SLICE_BYTES + kk * 32);
            const uint32_t accumulate = (first && kk == 0 && s == 0) ? 0u : 1u;
            umma_i8(tacc + (uint32_t)(s * OZ_BN), adesc, oz_desc(sB + kk * 32), make_idesc_i8(OZ_BM, n1), accumulate);
            if (ntot > 256)
                umma_i8(tacc + (uint32_t)(s * OZ_BN + 256), adesc, oz_desc(sB + 256 * OZ_BK + kk * 32),
                        make_idesc_i8(OZ_BM, ntot - 256), accumulate);
        }
    }
}

// ---------------------------------------------------------------------------
// slicing kernels
// ---------------------------------------------------------------------------
// Row exponents of W and the epilogue scale 2^e_i * rho / 127^2.
__global__ void oz_row_exponent_kernel(const double *__restrict__ W, int np, double rho, int *__restrict__ rowexp,
                                       double *__restrict__ rowscale, int *__restrict__ emax, int row0) {
    const int row = row0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= np) return;
    const double *w = W + (int64_t)row * np;
    double m = 0.0;
    for (int j = lane; j <= row; j += 32) m = fmax(m, fabs(w[j]));
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) {
        int e = 0;
        if (m > 0.0) frexp(m, &e);          // m = f * 2^e, f in [0.5, 1)  =>  |w| * 2^-e < 1
        rowexp[row] = e;
        rowscale[row] = ldexp(rho, e) / 16129.0;
        atomicMax(emax, e);
    }
}

// balanced base-256 digits of x * 127 * 2^48, |x| <= 1: byte k of the result is digit k as an int8
__device__ __forceinline__ unsigned long long oz_digits7(double x) {
    const long long X = __double2ll_rn(x * 35747322042253312.0);      // 127 * 2^48
    return (unsigned long long)(X + OZ_DIGIT_BIAS7) ^ (unsigned long long)OZ_DIGIT_BIAS7;
}

// W (lower triangle) -> S int8 slice planes [s][row][k]; slice s = digit byte 6 - s
__global__ void oz_slice_w_kernel(const double *__restrict__ W, int np, int S, const int *__restrict__ rowexp,
                                  int8_t *__restrict__ Ws, int row0) {
    const int row = row0 + blockIdx.y;
    const int k0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (k0 >= np) return;
    const double sc = ldexp(1.0, -rowexp[row]);
    const double *w = W + (int64_t)row * np + k0;
    unsigned long long dg[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) dg[i] = oz_digits7((k0 + i <= row) ? w[i] * sc : 0.0);
    for (int s = 0; s < S; ++s) {
        alignas(16) int8_t q[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) q[i] = (int8_t)(uint8_t)(dg[i] >> (8 * (6 - s)));
        *reinterpret_cast<int4 *>(Ws + ((int64_t)s * np + row) * np + k0) = *reinterpret_cast<const int4 *>(q);
    }
}

// K*^T slices [s][cand][k]: kappa = k(x_j, xc_m) / rho in [0,1] on the same grid.
// block: 64 observations x 128 candidates; thread = (candidate, 32-wide k half).
template <int DP>
__global__ void __launch_bounds__(256)
oz_kstar_slices_kernel(int kernel, int n, int np, int d, int S, const double *__restrict__ Xs,
                       const double *__restrict__ invell, const double *__restrict__ Xc, int64_t c0, int mc,
                       int mcp, int8_t *__restrict__ Ks, const double *__restrict__ beta,
                       double *__restrict__ mupart) {
    __shared__ double xs[64][DP];
    const int tid = threadIdx.x;
    const int j0 = blockIdx.y * 64;
    for (int e = tid; e < 64 * DP; e += 256) xs[e / DP][e % DP] = Xs[(int64_t)j0 * DP + e];
    const int m = blockIdx.x * 128 + (tid & 127);
    const int half = tid >> 7;
    const bool live = m < mc;
    double xc[DP];
#pragma unroll
    for (int k = 0; k < DP; ++k) xc[k] = (live && k < d) ? Xc[(c0 + m) * d + k] * invell[k] : 0.0;
    __syncthreads();
    double kb = 0.0;                              // sum_j kappa_j beta_j over this thread's 32 observations
    for (int sub = 0; sub < 2; ++sub) {
        const int jj0 = half * 32 + sub * 16;
        unsigned long long dg[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            double D = 0.0;
#pragma unroll
            for (int k = 0; k < DP; ++k) {
                const double t = xc[k] - xs[jj0 + i][k];
                D = fma(t, t, D);
            }
            double v;
            if (kernel == BO_KERNEL_SE) {
                v = exp(-0.5 * D);
            } else {
                const double rr = sqrt(5.0 * D);
                v = (1.0 + rr + rr * rr * (1.0 / 3.0)) * exp(-rr);
            }
            dg[i] = oz_digits7((live && (j0 + jj0 + i) < n) ? v : 0.0);
            if (live && (j0 + jj0 + i) < n) kb = fma(v, beta[j0 + jj0 + i], kb);
        }
        for (int s = 0; s < S; ++s) {
            alignas(16) int8_t q[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) q[i] = (int8_t)(uint8_t)(dg[i] >> (8 * (6 - s)));
            *reinterpret_cast<int4 *>(Ks + oz_kss_offset(s, m, j0 + jj0, mcp >> 7, np >> 6)) = *reinterpret_cast<const int4 *>(q);
        }
    }
    // partial of kappa . beta for (64-observation block, half): fixed-order sum in oz_moments_kernel
    mupart[((int64_t)blockIdx.y * 2 + half) * mcp + m] = kb;
}

// Fast variant (SE and Matern-5/2 kernels) for S <= 5 slices.  The FP64 pipe limits the slicer, so
// (i) the scaled squared distance comes from one dot product, D/2 = |xc|^2/2 + |xs|^2/2 - xc.xs
// (rounding error ~1e-14 relative in kappa, far below the 2^-41 quantum of 6 slices), (ii) exp is
// an inline exp2 (degree-12 Taylor in ln2*f, |f| <= 1/2, error < 2e-16) whose exponent add also
// applies the fixed-point scale, and (iii) the balanced base-256 digits are the bytes of one biased
// 40-bit integer (integer pipe) instead of S rint/subtract rounds on the FP64 pipe.
// Taylor coefficients ln2^i / i!, i = 12 .. 1, kept in constant memory so that each DFMA
// takes its coefficient as a constant-bank operand (no register or uniform-register traffic)
__constant__ double OZ_EXP2_C[12] = {
    2.5678435993488196e-11, 4.44553827187081e-10, 7.054911620801121e-09, 1.0178086009239696e-07,
    1.3215486790144305e-06, 1.5252733804059838e-05, 0.00015403530393381606, 0.0013333558146428441,
    0.009618129107628477,   0.055504108664821576,  0.2402265069591007,     0.6931471805599453};

// 2^(z + shift) for z <= 0 as a double (0 when it would round to 0 on the integer grid).
// DEG 10 has relative error 2.2e-13 (enough below 2^-35, i.e. up to 5 slices), DEG 12 < 2e-16.
// No FP64<->integer conversion instructions (F2I / FRND are very slow here): rint(z) is read from
// the low mantissa word of z + 1.5*2^52.
template <int DEG>
__device__ __forceinline__ double oz_exp2_scaled(double z, int shift) {
    const double zz = z + 6755399441055744.0;          // 1.5 * 2^52
    const int ki = __double2loint(zz) + shift;         // rint(z) + shift
    const double f = z - (zz - 6755399441055744.0);    // [-0.5, 0.5]
    double p = OZ_EXP2_C[12 - DEG];
#pragma unroll
    for (int i = 12 - DEG + 1; i < 12; ++i) p = fma(p, f, OZ_EXP2_C[i]);
    p = fma(p, f, 1.0);
    // (no cut-off at the integer grid: the caller also accumulates the FP64 mean from this value;
    //  values below 1/2 round to a zero digit string by themselves)
    const double v = __hiloint2double(__double2hiint(p) + (ki << 20), __double2loint(p));
    return ki < -1000 ? 0.0 : v;
}

#define OZ_KS_TILES 4      // observation tiles (of 64) per block
#define OZ_KS_THREADS 256  // two threads per candidate: each covers 32 of a tile's 64 observations

// 2^(i / 16), i = 0 .. 15, correctly rounded.  Sixteen doubles fill exactly one 128-byte row of shared memory: two
// lanes reading different entries hit different banks, two lanes reading the same entry are a broadcast, so the
// look-up is free of bank conflicts whatever the indices are (a 128-entry table cost 15.6 M conflicts per launch).
__constant__ double OZ_EXP2_TAB[16] = {
    1.0, 1.0442737824274138, 1.0905077326652577, 1.1387886347566916, 1.189207115002721, 1.241857812073484,
    1.2968395546510096, 1.3542555469368927, 1.4142135623730951, 1.4768261459394993, 1.5422108254079407,
    1.6104903319492543, 1.681792830507429, 1.7562521603732995, 1.8340080864093424, 1.9152065613971474};
// ln2^j / j!, j = 7 .. 1: (2^r - 1) / r for |r| <= 2^-5 to 1.2e-18 (relative to 2^r)
#define OZ_P7 1.5252733804059841e-05
#define OZ_P6 0.0001540353039338161
#define OZ_P5 0.0013333558146428443
#define OZ_P4 0.009618129107628477
#define OZ_P3 0.05550410866482158
#define OZ_P2 0.24022650695910072
#define OZ_P1 0.6931471805599453

__device__ __forceinline__ void oz_cp_async16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src));
}

// block: one candidate tile (128 candidates) x up to OZ_KS_TILES k blocks of 64 observations; thread = (candidate,
// half of the k block), 256 threads so that three blocks (24 warps) fit an SM -- with one thread per candidate
// (round 1) only 15 warps were resident and the kernel sat at 52 % of the FP64 pipe waiting on its own latency.
// Observation tiles (scaled coordinates + |xs|^2/2) are prefetched with cp.async into a double buffer; each finished
// 128 x 64 tile is staged in shared memory as the swizzled operand image and copied out with fully coalesced
// 16-byte-per-lane stores (8 KB per slice).
// exp2: z = k + i/16 + r with |r| <= 2^-5 (rint(16 z) is read from the low mantissa word of z + 1.5 * 2^48),
// 2^z = 2^k T[i] (1 + r P6(r)): a conflict-free shared-memory look-up and 8 FMAs instead of the 12-term Horner chain.
template <int DP, int S, bool MATERN, bool TAB>
__global__ void __launch_bounds__(OZ_KS_THREADS, 3)
oz_kstar_slices_fast_kernel(int n, int np, int d, const double *__restrict__ Xs, const double *__restrict__ XsHalfSq,
                            const double *__restrict__ invell, const double *__restrict__ Xc, int64_t c0, int mc,
                            int mcp, int8_t *__restrict__ Ks, const double *__restrict__ beta,
                            double *__restrict__ mupart) {
    __shared__ __align__(16) double xs[2][64][DP];
    __shared__ __align__(16) double hb[2][64];
    __shared__ __align__(16) double bt[2][64];                      // beta of the tile
    __shared__ __align__(128) double tab[16];
    extern __shared__ __align__(16) uint8_t oz_stage[];          // [S][128 rows][64 B]
    const int tid = threadIdx.x;
    const int row = tid & 127, half = tid >> 7;
    const int nkb = np / 64, ntiles = mcp / 128;
    const int t0 = blockIdx.y * OZ_KS_TILES;
    const int t1 = (t0 + OZ_KS_TILES < nkb) ? t0 + OZ_KS_TILES : nkb;
    auto prefetch = [&](int tile, int buf) {
        const double *src = Xs + (int64_t)tile * 64 * DP;
        for (int e = tid; e < 64 * DP / 2; e += OZ_KS_THREADS) oz_cp_async16(&xs[buf][0][0] + 2 * e, src + 2 * e);
        if (tid < 32) oz_cp_async16(&hb[buf][0] + 2 * tid, XsHalfSq + (int64_t)tile * 64 + 2 * tid);
        else if (tid < 64) oz_cp_async16(&bt[buf][0] + 2 * (tid - 32), beta + (int64_t)tile * 64 + 2 * (tid - 32));
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    prefetch(t0, 0);
    if (TAB && tid < 16) tab[tid] = OZ_EXP2_TAB[tid];
    const int m = blockIdx.x * 128 + row;
    const bool live = m < mc;
    double xc[DP], ha = 0.0;
#pragma unroll
    for (int k = 0; k < DP; ++k) {
        xc[k] = (live && k < d) ? Xc[(c0 + m) * d + k] * invell[k] : 0.0;
        ha = fma(xc[k], xc[k], ha);
    }
    ha *= 0.5;
    constexpr double LOG2E = 1.4426950408889634;
    constexpr double LOG2_127 = 6.988684686772166;      // kappa * 127 * 2^32 = 2^(log2 kappa + LOG2_127 + 32)
    constexpr double MAGIC48 = 422212465065984.0;       // 1.5 * 2^48: ulp = 2^-4
    static_assert(S >= 2 && S <= 5, "fast slicer handles 2..5 slices");
    const int swz = (row >> 1) & 3;
    const int jlimit = live ? n : 0;      // observations j < jlimit contribute for this thread's candidate
    double kb = 0.0;                 // sum_j 127 * 2^32 kappa_j beta_j over this thread's observations (FP64)
    for (int tile = t0; tile < t1; ++tile) {
        const int buf = (tile - t0) & 1;
        if (tile + 1 < t1) {
            prefetch(tile + 1, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();        // tile data landed; previous copy-out has finished reading the stage
        const int j0 = tile * 64;
#pragma unroll 1
        for (int sb = 0; sb < 2; ++sb) {
            const int sub = 2 * half + sb;
            const int jj0 = sub * 16;
            // X = rint(kappa * 127 * 2^32) < 2^39 as a 5-byte integer; the bytes of X + 0x8080808080,
            // each XOR 0x80, are its balanced base-256 digits (slice s = byte 4 - s): the low word is
            // transposed to slice-major words with byte permutes, the top digit sits in the high word.
            uint32_t wlow[16], wtop[4];
            // four elements in lockstep: the distance dot products and the exp2 chains of the
            // four are independent, so the FP64 pipe always has work in flight
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                const int jq = jj0 + 4 * q4;
                double dot[4], r[4], pl[4];
                int ki[4], ti[4];
                // |xs|^2/2 and beta of the four observations as two 16-byte broadcasts each
                const double2 hb01 = *reinterpret_cast<const double2 *>(&hb[buf][jq]);
                const double2 hb23 = *reinterpret_cast<const double2 *>(&hb[buf][jq + 2]);
                const double2 bt01 = *reinterpret_cast<const double2 *>(&bt[buf][jq]);
                const double2 bt23 = *reinterpret_cast<const double2 *>(&bt[buf][jq + 2]);
                const double hbv[4] = {hb01.x, hb01.y, hb23.x, hb23.y};
                const double btv[4] = {bt01.x, bt01.y, bt23.x, bt23.y};
#pragma unroll
                for (int e = 0; e < 4; ++e) dot[e] = -ha - hbv[e];
#pragma unroll
                for (int k = 0; k < DP; k += 2) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const double2 x2 = *reinterpret_cast<const double2 *>(&xs[buf][jq + e][k]);
                        dot[e] = fma(xc[k], x2.x, dot[e]);
                        dot[e] = fma(xc[k + 1], x2.y, dot[e]);
                    }
                }
                // v = 2^(z + 32), z = log2(127 kappa) (dot <= 0 up to rounding; a last-ulp excess rounds away)
                // Matern-5/2: kappa = (1 + r + r^2 / 3) e^-r, r = sqrt(5 D) = sqrt(-10 dot): the exponential part
                // goes through the same exp2 and the polynomial factor q multiplies its mantissa below
                double q[4];
                if (MATERN) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const double sa = fabs(dot[e] * -10.0) + 1e-300;          // |5 D| (rounding may leave -1e-15)
                        double y;
                        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(sa));
                        const double hs = 0.5 * sa;
                        y = y * fma(-hs, y * y, 1.5);
                        y = y * fma(-hs, y * y, 1.5);
                        const double rr = sa * y;
                        q[e] = fma(fma(rr, 1.0 / 3.0, 1.0), rr, 1.0);
                        dot[e] = -rr;                                          // exponent of e
                    }
                }
                if (TAB) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const double z = fma(dot[e], LOG2E, LOG2_127);
                        const double zz = z + MAGIC48;                      // low word = rint(16 z)
                        const int n16 = __double2loint(zz);
                        ki[e] = (n16 >> 4) + 32;
                        ti[e] = n16 & 15;
                        r[e] = z - (zz - MAGIC48);                          // [-2^-5, 2^-5]
                        pl[e] = OZ_P7;
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) pl[e] = fma(pl[e], r[e], OZ_P6);
#pragma unroll
                    for (int e = 0; e < 4; ++e) pl[e] = fma(pl[e], r[e], OZ_P5);
#pragma unroll
                    for (int e = 0; e < 4; ++e) pl[e] = fma(pl[e], r[e], OZ_P4);
#pragma unroll
                    for (int e = 0; e < 4; ++e) pl[e] = fma(pl[e], r[e], OZ_P3);
#pragma unroll
                    for (int e = 0; e < 4; ++e) pl[e] = fma(pl[e], r[e], OZ_P2);
#pragma unroll
                    for (int e = 0; e < 4; ++e) pl[e] = fma(pl[e], r[e], OZ_P1);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const double t = tab[ti[e]];
                        pl[e] = fma(pl[e] * r[e], t, t);                    // T[i] (1 + r P(r))
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const double z = fma(dot[e], LOG2E, LOG2_127);
                        const double zz = z + 6755399441055744.0;          // 1.5 * 2^52: low word = rint(z)
                        ki[e] = __double2loint(zz) + 32;
                        r[e] = z - (zz - 6755399441055744.0);              // [-0.5, 0.5]
                        pl[e] = OZ_EXP2_C[0];
                    }
#pragma unroll
                    for (int c = 1; c < 12; ++c) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) pl[e] = fma(pl[e], r[e], OZ_EXP2_C[c]);
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) pl[e] = fma(pl[e], r[e], 1.0);
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int i = 4 * q4 + e;
                    if (MATERN) pl[e] *= q[e];
                    // one predicate (live candidate, real observation, no exponent underflow) -> one select
                    const bool on = (((j0 + jq + e) - jlimit) & (-961 - ki[e])) < 0;
                    double v = __hiloint2double(__double2hiint(pl[e]) + (ki[e] << 20), __double2loint(pl[e]));
                    v = on ? v : 0.0;                                           // in [0, 127 * 2^32]
                    kb = fma(v, btv[e], kb);
                    // + 2^52 puts rint(v) into the mantissa; the digit bias 0x8080808080 rides in the same addition
                    // (exact: the sum stays below 2^53), so the carry between the words costs no integer instruction
                    const double vv = v + (4503599627370496.0 + 551911719040.0);
                    wlow[i] = (uint32_t)__double2loint(vv) ^ 0x80808080u;
                    const uint32_t top = ((uint32_t)__double2hiint(vv) & 0xFFu) ^ 0x80u;
                    if (e == 0) wtop[q4] = top;
                    else wtop[q4] |= top << (8 * e);
                }
            }
            uint8_t *o0 = oz_stage + row * 64 + ((sub ^ swz) << 4);
            *reinterpret_cast<uint4 *>(o0) = make_uint4(wtop[0], wtop[1], wtop[2], wtop[3]);
#pragma unroll
            for (int s = 1; s < S; ++s) {
                const uint32_t b = (uint32_t)(4 - s);                   // byte lane holding slice s
                const uint32_t sel2 = b | ((4u + b) << 4);             // {x.b, y.b}
                uint32_t w[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const uint32_t lo2 = __byte_perm(wlow[4 * g + 0], wlow[4 * g + 1], sel2);
                    const uint32_t hi2 = __byte_perm(wlow[4 * g + 2], wlow[4 * g + 3], sel2);
                    w[g] = __byte_perm(lo2, hi2, 0x5410);
                }
                *reinterpret_cast<uint4 *>(o0 + s * OZ_A_SLICE_BYTES) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
        __syncthreads();        // stage complete
#pragma unroll
        for (int s = 0; s < S; ++s) {
            uint4 *dst = reinterpret_cast<uint4 *>(Ks + oz_kss_block(s, blockIdx.x, tile, ntiles, nkb));
            const uint4 *src = reinterpret_cast<const uint4 *>(oz_stage + s * OZ_A_SLICE_BYTES);
#pragma unroll
            for (int e = 0; e < OZ_A_SLICE_BYTES / 16 / OZ_KS_THREADS; ++e) dst[e * OZ_KS_THREADS + tid] = src[e * OZ_KS_THREADS + tid];
        }
    }
    // the posterior mean needs no contraction with W: mu = bias + rho * kappa . beta (beta = K^-1 r);
    // per-(block, half) partials, summed in a fixed order by oz_moments_kernel
    mupart[((int64_t)blockIdx.y * 2 + half) * mcp + m] = kb * (1.0 / 545460846592.0);      // 127 * 2^32
}

// |xs_j|^2 / 2 per observation (once per fit)
__global__ void oz_halfsq_kernel(const double *__restrict__ Xs, int rows, int dp, double *__restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= rows) return;
    double b = 0.0;
    for (int k = 0; k < dp; ++k) b = fma(Xs[(int64_t)j * dp + k], Xs[(int64_t)j * dp + k], b);
    out[j] = 0.5 * b;
}

// ---------------------------------------------------------------------------
// the contraction kernel
// ---------------------------------------------------------------------------
struct OzParams {
    int nrb, nkb, full_k;     // 64-row blocks of the B operand, 64-wide k blocks, 1: every unit walks all k blocks
    int S, nstages, ntiles, nacc, tiles_per_group;
    int mcp;
    const double *rowscale;   // per B row: scale of the reassembled integer
    double *qpart;            // MODE 0: [nrb][mcp] partial |v|^2 per row block
    int32_t *dbg;             // optional: [rb][g][128][64] accumulators of tile 0
    const int8_t *kss;        // A operand slice blocks (oz_kss_block layout)
    // MODE 1 (Thompson draws: B rows = draws, A = cosine features): f = rowbias + rowscale * contraction
    const double *rowbias;
    double *out;              // optional [row][out_ld] values
    int64_t out_ld, c0;       // leading dimension of out (total candidates), first candidate of this chunk
    int mc, nrows_live;
    double *blkval;           // optional per-(row, 32-candidate block) maximum ...
    int64_t *blkidx;          // ... and its first arg max (global candidate index)
    int64_t blk_ld, blk0;
};

// Work unit = (candidate tile, 64-row block of W).  Units are ordered group by group
// (a group = `tiles_per_group` candidate tiles whose K* slices fit in L2 next to the W
// slices), heaviest row block first, and dealt round-robin to the persistent CTAs, so
// the CTAs running concurrently share a small set of candidate tiles (L2 hits on the A
// stream) and all stream the same few row blocks of W.
struct OzUnit { int tile, rb; };
__device__ __forceinline__ OzUnit oz_decode(int u, int nb, int T, int ntiles) {
    const int per = T * nb;
    const int grp = u / per, r = u - grp * per;
    const int rem = ntiles - grp * T;
    const int Tg = rem < T ? rem : T;
    OzUnit o;
    o.rb = nb - 1 - r / Tg;
    o.tile = grp * T + r % Tg;
    return o;
}

// MODE 0: scoring (B = slices of the triangular W, k range cut at the diagonal, epilogue reduces |v|^2);
// MODE 1: Thompson draws (B = slices of Theta, full k range, epilogue writes values / per-draw arg max).
// CL > 1 (MODE 0): thread-block clusters of CL CTAs work on the SAME candidate tile and CL adjacent row blocks of W.
// The K* slice tile of a k block -- two thirds of a stage's bytes -- is fetched ONCE per cluster: CTA r loads the slices
// s = r (mod CL) and multicasts them into the shared memory of all CL CTAs; every CTA loads its own W slices.  A stage may
// be refilled only when every CTA of the cluster has consumed it, so each MMA thread commits its stage release to the
// `empty` barrier of ALL CTAs (count CL).  The CL row blocks have k ranges that differ by up to CL - 1 blocks: the
// CTAs with the shorter range step through those stages without loads or MMAs, so the rings stay in lockstep; the k
// blocks past the shortest range are loaded privately by the CTAs that need them.
template <int S, int EXTRA, int MODE, int CL>
__global__ void __launch_bounds__(OZ_THREADS, 1)
oz_score_kernel(const __grid_constant__ CUtensorMap tmapB, OzParams p) {
    extern __shared__ uint8_t oz_smem_raw[];
    // 1024-byte aligned operand ring
    const uint32_t raw = smem_u32(oz_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *gen_base = oz_smem_raw + (base - raw);
    constexpr int G = S - 1 + EXTRA;          // highest accumulator group
    constexpr int NG = G + 1;
    const int nst = p.nstages, nacc = p.nacc;
    const uint32_t stage_bytes = (uint32_t)S * (OZ_A_SLICE_BYTES + OZ_B_SLICE_BYTES);
    uint64_t *bars = reinterpret_cast<uint64_t *>(gen_base + (size_t)nst * stage_bytes);
    // bars[0..nst): full, [nst..2nst): empty, then tmem_full[2], tmem_empty[2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * nst + 4);
    double *row_sm = reinterpret_cast<double *>(bars + 2 * nst + 6);      // MODE 1: [2][nrb * 64] row scale, row bias
    double *argT = row_sm + OZ_ROW_SMEM / 8;                               // MODE 1: [16][OZ_ARG_LD] arg-max transpose
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (nst + s); };
    auto tmem_full = [&](int a) { return bar0 + 8u * (2 * nst + a); };
    auto tmem_empty = [&](int a) { return bar0 + 8u * (2 * nst + 2 + a); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    static_assert(CL == 1 || MODE == 0, "clusters are wired for the scoring mode only");
    const int crank = (CL > 1) ? (int)cluster_ctarank() : 0;
    constexpr uint16_t cmask = (uint16_t)((1u << CL) - 1u);
    const int nb = p.nrb / CL;                   // row-block groups (CL adjacent row blocks each)
    const int nunits = p.ntiles * nb;
    const int ucta = (int)blockIdx.x / CL, ustride = (int)gridDim.x / CL;   // unit walk of this cluster

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmapB);
        for (int s = 0; s < nst; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), CL);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full(a), 1);
            mbar_init(tmem_empty(a), 8);                       // one arrival per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
    if (MODE == 1) {
        for (int i = threadIdx.x; i < p.nrb * OZ_BN; i += (int)blockDim.x) {
            row_sm[i] = p.rowscale[i];
            row_sm[p.nrb * OZ_BN + i] = p.rowbias[i];
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (CL > 1) cluster_sync_all();              // every CTA's barriers exist before anyone multicasts into them
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int u = ucta; u < nunits; u += ustride) {
                const OzUnit un = oz_decode(u, nb, p.tiles_per_group, p.ntiles);
                const int rb = un.rb * CL + crank;                          // this CTA's row block
                const int kown = (MODE == 1 || p.full_k) ? p.nkb - 1 : rb;  // last k block this CTA needs
                const int kshared = (MODE == 1 || p.full_k) ? p.nkb - 1 : un.rb * CL;   // ... every CTA of the cluster needs
                const int klast = (MODE == 1 || p.full_k) ? p.nkb - 1 : un.rb * CL + CL - 1;
                for (int kb = 0; kb <= klast; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    if (kb > kown) {                                        // not mine: keep the ring in step
                        mbar_arrive(full_bar(stage));
                        if (++stage == nst) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    mbar_expect_tx(full_bar(stage), stage_bytes);
                    const uint32_t sA = base + stage * stage_bytes;
                    const uint32_t sB = sA + S * OZ_A_SLICE_BYTES;
                    const bool shared = CL > 1 && kb <= kshared;
                    for (int s = 0; s < S; ++s) {
                        const int8_t *src = p.kss + oz_kss_block(s, un.tile, kb, p.ntiles, p.nkb);
                        if (!shared) bulk_load(sA + s * OZ_A_SLICE_BYTES, src, OZ_A_SLICE_BYTES, full_bar(stage));
                        else if (s % CL == crank)
                            bulk_load_multicast(sA + s * OZ_A_SLICE_BYTES, src, OZ_A_SLICE_BYTES, full_bar(stage), cmask);
                        tma_load_3d(sB + s * OZ_B_SLICE_BYTES, &tmapB, kb * OZ_BK, rb * OZ_BN, s, full_bar(stage));
                    }
                    if (++stage == nst) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int u = ucta; u < nunits; u += ustride) {
                const OzUnit un = oz_decode(u, nb, p.tiles_per_group, p.ntiles);
                mbar_wait(tmem_empty(acc), acc_phase ^ 1);    // epilogue has drained this accumulator set
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(acc * NG * OZ_BN);
                const int kown = (MODE == 1 || p.full_k) ? p.nkb - 1 : un.rb * CL + crank;
                const int klast = (MODE == 1 || p.full_k) ? p.nkb - 1 : un.rb * CL + CL - 1;
                for (int kb = 0; kb <= klast; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sA = base + stage * stage_bytes;
                    if (kb <= kown) oz_issue_stage<S, EXTRA>(sA, sA + S * OZ_A_SLICE_BYTES, tacc, kb == 0);
                    // smem slot free once these MMAs retire -- in every CTA of the cluster that may write into it
                    if (CL > 1) umma_commit_multicast(empty_bar(stage), cmask);
                    else umma_commit(empty_bar(stage));
                    if (++stage == nst) { stage = 0; phase ^= 1; }
                }
                umma_commit(tmem_full(acc));                  // accumulators of this unit are complete
                if (++acc == nacc) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue: one candidate per thread =====================
        const int quarter = warp & 3;                          // TMEM lane quarter this warp may access
        // Eight epilogue warps, two per lane quarter, 32 of the unit's 64 columns each: with one warp per scheduler the
        // epilogue was latency-bound (ncu: 8.7 cycles per issued instruction, 0.17 IPC); in MODE 1 it paced the kernel at
        // 45 % of the tensor pipe, in MODE 0 at 5 slices (a single accumulator set) it held the set for a fifth of the time.
        const int ehalf = (warp - 2) >> 2;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int u = ucta; u < nunits; u += ustride) {
            OzUnit un = oz_decode(u, nb, p.tiles_per_group, p.ntiles);
            un.rb = un.rb * CL + crank;                        // this CTA's row block
            const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * NG * OZ_BN);
            double q = 0.0;
            mbar_wait(tmem_full(acc), acc_phase);
            tc_fence_after();
            // 16 columns (rows of B) of this thread's candidate: TMEM -> sum_g D_g 256^(G - g) / 256^G.
            // The first five groups fit int64 (|D_g| < 2^30): shift-adds on the integer pipe and a single
            // conversion keep the FP64 pipe for the reductions.
            auto reassemble = [&](int c0, double (&v16)[16]) {
                int32_t r[NG][16];
#pragma unroll
                for (int g = 0; g < NG; ++g) tmem_ld16(lane_base + (uint32_t)(g * OZ_BN + c0), r[g]);
                tmem_ld_wait();
                if (p.dbg && un.tile == 0) {
#pragma unroll
                    for (int g = 0; g < NG; ++g) {
                        int32_t *o = p.dbg + (((int64_t)un.rb * NG + g) * 128 + quarter * 32 + lane) * 64 + c0;
#pragma unroll
                        for (int i = 0; i < 16; ++i) o[i] = r[g][i];
                    }
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    constexpr int GI = G < 4 ? G : 4;
                    long long a = (long long)r[0][i];
#pragma unroll
                    for (int g = 1; g <= GI; ++g) a = a * 256 + (long long)r[g][i];
                    double v = (double)a;
#pragma unroll
                    for (int g = GI + 1; g <= G; ++g) v = fma(v, 256.0, (double)r[g][i]);
                    v16[i] = v * (1.0 / (double)(1ull << (8 * G)));
                }
            };
            if (MODE == 0) {
                const int cbase = ehalf * 32;                  // this warp's 32 of the unit's 64 rows of W
                if (NG <= 5) {
                    // Drain first, compute later: the warp's NG x 32 accumulators go to registers, the TMEM set goes back
                    // to the MMA warp at once, and the reassembly / reduction runs while the next unit's MMAs are already
                    // issuing.  (At 5 slices only ONE accumulator set fits TMEM, so every cycle the set is held after
                    // the last MMA of a unit is a cycle the tensor pipe idles: with the in-place epilogue of round 1 --
                    // one warp per scheduler, 8.7 cycles per dependent instruction -- that was a fifth of the kernel.)
                    int32_t r[NG][2][16];
#pragma unroll
                    for (int g = 0; g < NG; ++g) {
                        tmem_ld16(lane_base + (uint32_t)(g * OZ_BN + cbase), r[g][0]);
                        tmem_ld16(lane_base + (uint32_t)(g * OZ_BN + cbase + 16), r[g][1]);
                    }
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tmem_empty(acc));
                    if (p.dbg && un.tile == 0) {
#pragma unroll
                        for (int g = 0; g < NG; ++g) {
                            int32_t *o = p.dbg + (((int64_t)un.rb * NG + g) * 128 + quarter * 32 + lane) * 64 + cbase;
#pragma unroll
                            for (int i = 0; i < 32; ++i) o[i] = r[g][i >> 4][i & 15];
                        }
                    }
                    const int row0 = un.rb * OZ_BN + cbase;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        constexpr int GI = G < 4 ? G : 4;
                        long long a = (long long)r[0][i >> 4][i & 15];
#pragma unroll
                        for (int g = 1; g <= GI; ++g) a = a * 256 + (long long)r[g][i >> 4][i & 15];
                        double v = (double)a;
#pragma unroll
                        for (int g = GI + 1; g <= G; ++g) v = fma(v, 256.0, (double)r[g][i >> 4][i & 15]);
                        const double vv = v * (1.0 / (double)(1ull << (8 * G))) * __ldg(p.rowscale + row0 + i);
                        q = fma(vv, vv, q);
                    }
                } else {
#pragma unroll 1
                    for (int c0 = cbase; c0 < cbase + 32; c0 += 16) {
                        double v16[16];
                        reassemble(c0, v16);
                        const int row0 = un.rb * OZ_BN + c0;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const double vv = v16[i] * __ldg(p.rowscale + row0 + i);
                            q = fma(vv, vv, q);
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tmem_empty(acc));
                }
                if (++acc == nacc) { acc = 0; acc_phase ^= 1; }
                // two partial sums per (row block, candidate): one per half of the epilogue warps
                const int64_t o = ((int64_t)un.rb * 2 + ehalf) * p.mcp + (int64_t)un.tile * OZ_BM + quarter * 32 + lane;
                p.qpart[o] = q;
            } else {
                // drain this warp's 32 columns into registers first so the accumulators go back to the MMA warp
                // before the stores and the per-draw arg max
                double vcol[32];
#pragma unroll
                for (int cb = 0; cb < 2; ++cb) {
                    double v16[16];
                    reassemble((ehalf * 2 + cb) * 16, v16);
#pragma unroll
                    for (int i = 0; i < 16; ++i) vcol[cb * 16 + i] = v16[i];
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tmem_empty(acc));
                if (++acc == nacc) { acc = 0; acc_phase ^= 1; }
                const int cand = un.tile * OZ_BM + quarter * 32 + lane;
                const bool cand_live = cand < p.mc;
                const int rowb = un.rb * OZ_BN + ehalf * 32;
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    vcol[c] = fma(vcol[c], row_sm[rowb + c], row_sm[p.nrb * OZ_BN + rowb + c]);
                    if (p.out && cand_live && rowb + c < p.nrows_live) p.out[(int64_t)(rowb + c) * p.out_ld + p.c0 + cand] = vcol[c];
                }
                if (p.blkval) {
                    // Per-draw first arg max over the tile's 128 candidates.  Round 1 ran a register butterfly per warp
                    // (62 exchanges of a double and an int per lane, 250 registers): ncu showed the epilogue, not the
                    // tensor pipe, pacing the kernel (39 % of the int8 roof).  Now each half of the epilogue warps
                    // transposes its 32 draws through shared memory, 16 draws per round: thread (candidate) writes its
                    // values, then thread (draw, 16-candidate part) scans in index order and three shuffles join the
                    // eight parts.  NaN never wins; ties go to the lower candidate (strict > in scan order).
                    const int pos = quarter * 32 + lane;
                    const int cc = pos >> 3, part = pos & 7;            // scan role: draw cc of the round, candidates 16 part ..
                    double *argH = argT + ehalf * 16 * OZ_ARG_LD;       // each half of the epilogue warps has its own buffer ...
                    const int barid = 1 + ehalf;                        // ... and its own named barrier (4 warps)
#pragma unroll
                    for (int rd = 0; rd < 2; ++rd) {
#pragma unroll
                        for (int c = 0; c < 16; ++c) {
                            const double v = vcol[rd * 16 + c];
                            argH[c * OZ_ARG_LD + pos + (pos >> 4)] = (cand_live && v == v) ? v : -INFINITY;
                        }
                        asm volatile("bar.sync %0, 128;" ::"r"(barid) : "memory");
                        const double *col = argH + cc * OZ_ARG_LD + part * 17;
                        double bv = -INFINITY;
                        int bi = 0x7fffffff;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const double v = col[i];
                            if (v > bv) { bv = v; bi = part * 16 + i; }
                        }
#pragma unroll
                        for (int o = 1; o < 8; o <<= 1) {               // join the eight parts of a draw (adjacent lanes)
                            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                        }
                        const int row = rowb + rd * 16 + cc;
                        if (part == 0 && row < p.nrows_live) {
                            const int64_t o = (int64_t)row * p.blk_ld + p.blk0 + un.tile;
                            p.blkval[o] = bv;
                            p.blkidx[o] = bi == 0x7fffffff ? INT64_MAX : p.c0 + (int64_t)un.tile * OZ_BM + bi;
                        }
                        asm volatile("bar.sync %0, 128;" ::"r"(barid) : "memory");  // the transpose buffer is free again
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();              // nobody leaves while a peer may still multicast into its shared memory
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// s2 = rho - sum_rb qpart (int8 contraction); mu = bias + rho * sum_blocks mupart (FP64 dot product
// kappa . beta accumulated by the slicer); both in a fixed summation order
// (the column sums keep their fixed order; the loads of 16 partials are issued together so that each thread has 16
//  requests in flight instead of one dependent load per addition)
__device__ __forceinline__ double oz_sum_ordered(const double *__restrict__ p, int n, int64_t stride) {
    double acc = 0.0;
    int i = 0;
    for (; i + 16 <= n; i += 16) {
        double v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = p[(int64_t)(i + k) * stride];
#pragma unroll
        for (int k = 0; k < 16; ++k) acc += v[k];
    }
    for (; i < n; ++i) acc += p[(int64_t)i * stride];
    return acc;
}

__global__ void __launch_bounds__(128)
oz_moments_kernel(int nb, int nmu, int mcp, const double *__restrict__ qpart, const double *__restrict__ mupart, double rho,
                  double bias, double *__restrict__ mu, double *__restrict__ s2) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= mcp) return;
    const double q = oz_sum_ordered(qpart + m, nb, mcp);
    const double pm = oz_sum_ordered(mupart + m, nmu, mcp);
    mu[m] = fma(rho, pm, bias);
    s2[m] = rho - q;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// 3-D int8 tensor (k, row, slice) with a (64, box_rows, 1) box and 64-byte swizzle
static int make_tmap(bo_ctx *ctx, CUtensorMap *tm, const void *base, uint64_t kdim, uint64_t rows, uint64_t slices,
                     uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return bo_set_err(ctx, BO_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[3] = {kdim, rows, slices};
    cuuint64_t strides[2] = {kdim, kdim * rows};
    cuuint32_t box[3] = {OZ_BK, box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return bo_set_err(ctx, BO_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return BO_OK;
}

static int oz_stage_count(int S) {
    const int stage = S * (OZ_A_SLICE_BYTES + OZ_B_SLICE_BYTES);
    int n = (200 * 1024) / stage;
    return n > 4 ? 4 : (n < 2 ? 2 : n);
}
static size_t oz_smem_bytes(int S) {
    return (size_t)oz_stage_count(S) * S * (OZ_A_SLICE_BYTES + OZ_B_SLICE_BYTES) + 1024 + 256;
}
// MODE 1 (Thompson) also holds the row scale / bias and the arg-max transpose buffer: fewer ring stages
static int oz_stage_count_m1(int S) {
    const int stage = S * (OZ_A_SLICE_BYTES + OZ_B_SLICE_BYTES);
    int n = (227 * 1024 - 1280 - OZ_ROW_SMEM - OZ_ARG_SMEM) / stage;
    return n > 4 ? 4 : (n < 2 ? 2 : n);
}
static size_t oz_smem_bytes_m1(int S) {
    return (size_t)oz_stage_count_m1(S) * S * (OZ_A_SLICE_BYTES + OZ_B_SLICE_BYTES) + 1024 + 256 + OZ_ROW_SMEM + OZ_ARG_SMEM;
}

int bo_ozaki_init(bo_ctx *ctx) {
#define OZ_ATTR1(SS, EE, MM) BO_CUDA(ctx, cudaFuncSetAttribute(oz_score_kernel<SS, EE, MM, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MM ? oz_smem_bytes_m1(SS) : oz_smem_bytes(SS))))
#define OZ_ATTR_CL(SS, EE) BO_CUDA(ctx, cudaFuncSetAttribute(oz_score_kernel<SS, EE, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)oz_smem_bytes(SS))); \
                           BO_CUDA(ctx, cudaFuncSetAttribute(oz_score_kernel<SS, EE, 0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)oz_smem_bytes(SS)))
#define OZ_ATTR(SS) OZ_ATTR1(SS, 0, 0); OZ_ATTR1(SS, 1, 0)
#define OZ_ATTR_M1(SS) OZ_ATTR1(SS, 0, 1); OZ_ATTR1(SS, 1, 1)
    OZ_ATTR(2); OZ_ATTR(3); OZ_ATTR(4); OZ_ATTR(5); OZ_ATTR(6); OZ_ATTR(7);
    OZ_ATTR_M1(3); OZ_ATTR_M1(4); OZ_ATTR_M1(5);          // the Thompson path runs 3..5 slices
#undef OZ_ATTR_M1
    OZ_ATTR_CL(4, 0); OZ_ATTR_CL(4, 1); OZ_ATTR_CL(5, 0); OZ_ATTR_CL(5, 1);     // cluster variants of the common levels
#undef OZ_ATTR_CL
#undef OZ_ATTR
#undef OZ_ATTR1
#define OZ_KATTR(DP, SS) BO_CUDA(ctx, cudaFuncSetAttribute(oz_kstar_slices_fast_kernel<DP, SS, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SS * OZ_A_SLICE_BYTES)); \
                         BO_CUDA(ctx, cudaFuncSetAttribute(oz_kstar_slices_fast_kernel<DP, SS, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SS * OZ_A_SLICE_BYTES)); \
                         BO_CUDA(ctx, cudaFuncSetAttribute(oz_kstar_slices_fast_kernel<DP, SS, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SS * OZ_A_SLICE_BYTES)); \
                         BO_CUDA(ctx, cudaFuncSetAttribute(oz_kstar_slices_fast_kernel<DP, SS, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SS * OZ_A_SLICE_BYTES))
#define OZ_KATTR_ALL(DP) OZ_KATTR(DP, 2); OZ_KATTR(DP, 3); OZ_KATTR(DP, 4); OZ_KATTR(DP, 5)
    OZ_KATTR_ALL(2); OZ_KATTR_ALL(4); OZ_KATTR_ALL(8); OZ_KATTR_ALL(16);
#undef OZ_KATTR_ALL
#undef OZ_KATTR
    return BO_OK;
}

// Build the slice planes of W for hyper-sample s (S slices) into ctx->dWs.
int bo_ozaki_prepare(bo_ctx *ctx, int S) {
    const int np = ctx->np, ns = ctx->S;
    if (ctx->oz_ready && ctx->oz_slices >= S) return BO_OK;       // digit planes are a prefix code: more is fine
    BO_TRY(bo_reserve(ctx, &ctx->dWs, &ctx->ws_capacity, (size_t)ns * S * np * np));
    BO_TRY(bo_reserve(ctx, &ctx->dRowScale, &ctx->rowscale_capacity, (size_t)ns * np));
    BO_TRY(bo_reserve(ctx, &ctx->dRowExp, &ctx->rowexp_capacity, (size_t)ns * np + ns));
    int *emax_dev = ctx->dRowExp + (size_t)ns * np;
    std::vector<int> init(ns, -100000);
    BO_CUDA(ctx, cudaMemcpyAsync(emax_dev, init.data(), sizeof(int) * ns, cudaMemcpyHostToDevice, ctx->stream));
    BO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int s = 0; s < ns; ++s) {
        {
            BO_LAUNCH(ctx, "oz_row_exponent_kernel");
            oz_row_exponent_kernel<<<np / 8, 256, 0, ctx->stream>>>(ctx->dW + (size_t)s * np * np, np, ctx->h_rho[s],
                                                                   ctx->dRowExp + (size_t)s * np,
                                                                   ctx->dRowScale + (size_t)s * np, emax_dev + s, 0);
            BO_CHECK_LAUNCH(ctx);
        }
        {
            BO_LAUNCH(ctx, "oz_slice_w_kernel");
            oz_slice_w_kernel<<<dim3((np / 16 + 127) / 128, np), 128, 0, ctx->stream>>>(
                ctx->dW + (size_t)s * np * np, np, S, ctx->dRowExp + (size_t)s * np,
                ctx->dWs + (size_t)s * S * np * np, 0);
            BO_CHECK_LAUNCH(ctx);
        }
    }
    BO_TRY(bo_reserve(ctx, &ctx->dXsHalfSq, &ctx->halfsq_capacity, (size_t)ns * np));
    {
        BO_LAUNCH(ctx, "oz_halfsq_kernel");
        oz_halfsq_kernel<<<(ns * np + 255) / 256, 256, 0, ctx->stream>>>(ctx->dXs, ns * np, ctx->dp, ctx->dXsHalfSq);
        BO_CHECK_LAUNCH(ctx);
    }
    ctx->h_emax.assign(ns, 0);
    BO_CUDA(ctx, cudaMemcpyAsync(ctx->h_emax.data(), emax_dev, sizeof(int) * ns, cudaMemcpyDeviceToHost, ctx->stream));
    BO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->oz_ready = true;
    ctx->oz_slices = S;
    return BO_OK;
}

// bo_append wrote row `row` of W (and of Xs): refresh that row's exponent, scale, slices and |xs|^2/2.
int bo_ozaki_append_row(bo_ctx *ctx, int row) {
    const int np = ctx->np, ns = ctx->S, S = ctx->oz_slices;
    int *emax_dev = ctx->dRowExp + (size_t)ns * np;
    for (int s = 0; s < ns; ++s) {
        {
            BO_LAUNCH(ctx, "oz_row_exponent_kernel");
            oz_row_exponent_kernel<<<1, 32, 0, ctx->stream>>>(ctx->dW + (size_t)s * np * np, np, ctx->h_rho[s],
                                                              ctx->dRowExp + (size_t)s * np, ctx->dRowScale + (size_t)s * np,
                                                              emax_dev + s, row);
            BO_CHECK_LAUNCH(ctx);
        }
        {
            BO_LAUNCH(ctx, "oz_slice_w_kernel");
            oz_slice_w_kernel<<<dim3((np / 16 + 127) / 128, 1), 128, 0, ctx->stream>>>(
                ctx->dW + (size_t)s * np * np, np, S, ctx->dRowExp + (size_t)s * np, ctx->dWs + (size_t)s * S * np * np, row);
            BO_CHECK_LAUNCH(ctx);
        }
    }
    {
        BO_LAUNCH(ctx, "oz_halfsq_kernel");
        oz_halfsq_kernel<<<(ns * np + 255) / 256, 256, 0, ctx->stream>>>(ctx->dXs, ns * np, ctx->dp, ctx->dXsHalfSq);
        BO_CHECK_LAUNCH(ctx);
    }
    BO_CUDA(ctx, cudaMemcpyAsync(ctx->h_emax.data(), emax_dev, sizeof(int) * ns, cudaMemcpyDeviceToHost, ctx->stream));
    BO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BO_OK;
}

// Error model (calibrated on the headline shape, tools/oz_err.py): with S base-256 slices per operand
// the truncation term of v is ~ sqrt(n) 2^e sqrt(rho) 2^(-8S) and the dropped digit pairs of group g = S
// are ~30x larger; accumulating that group too (`extra`) leaves only the truncation term.
// Levels in order of cost: (S, extra) = (3,0) (3,1) (4,0) (4,1) (5,0) (5,1) (6,0) ...; pick the first
// whose estimate is below tol * sqrt(rho).  tol >= 2 pins S = floor(tol), extra = (tol - S >= 0.5).
int bo_ozaki_choose_slices(bo_ctx *ctx, double tol) {
    if (tol >= 2.0) {
        int S = (int)tol;
        ctx->oz_extra = (tol - S) >= 0.5;
        return S > OZ_MAX_S ? OZ_MAX_S : (S < 2 ? 2 : S);
    }
    // exponents need W: use 4 slices provisionally just to obtain e_max
    if (!ctx->oz_ready) {
        if (bo_ozaki_prepare(ctx, 4) != BO_OK) return -1;
    }
    double worst = 0.0;
    for (int s = 0; s < ctx->S; ++s) {
        double est = 8.0 * sqrt((double)ctx->np) * ldexp(1.0, ctx->h_emax[s]) * sqrt(ctx->h_rho[s]);
        worst = est > worst ? est : worst;
    }
    for (int S = 3; S <= OZ_MAX_S; ++S) {
        const double base = worst * ldexp(1.0, -8 * S);
        if (base <= tol) { ctx->oz_extra = false; return S; }
        if (base / 32.0 <= tol) { ctx->oz_extra = true; return S; }
    }
    ctx->oz_extra = true;
    return OZ_MAX_S;
}

// A-priori bound on the error of s2 = rho - |v|^2 on this path, used by the rescue pass (score.cu run_oz):
//     |ds2| <= errK * sqrt(q rho),  q = rho - s2 = |v|^2,  errK = OZ_ERR_SAFETY * est(level)
// with est(level) the model above (8 sqrt(np) 2^emax sqrt(rho) 256^-S, / 32 with the extra pair group).  ds2 = 2 v . dv
// with independent zero-mean entry errors dv_i, hence the sqrt(q) factor.  Calibration (tools/oz_calib.py, nine shapes
// x six levels x 73 k candidates incl. 8 k placed on top of observations): max |ds2| / (est sqrt(q rho)) = 0.85,
// median 0.03 -- the factor 2 leaves > 2.3x on the worst case seen (profiles/r2_oz_calib.txt).
#define OZ_ERR_SAFETY 2.0
int bo_ozaki_error_scale(bo_ctx *ctx, int S, bool extra, int slot) {
    const int ns = ctx->S;
    ctx->h_errk.assign(ns, 0.0);
    for (int s = 0; s < ns; ++s) {
        const double est = 8.0 * sqrt((double)ctx->np) * ldexp(1.0, ctx->h_emax[s]) * sqrt(ctx->h_rho[s]) * ldexp(1.0, -8 * S) /
                           (extra ? 32.0 : 1.0);
        ctx->h_errk[s] = OZ_ERR_SAFETY * est;
    }
    // BO_OZ_ERRK_SLOTS scales per context: one per level a pass can mix (score.cu run_oz)
    BO_TRY(bo_reserve(ctx, &ctx->dErrK, &ctx->errk_capacity, (size_t)ns * BO_OZ_ERRK_SLOTS));
    // (pageable source: the copy is staged before the call returns, h_errk may be rewritten right away)
    BO_CUDA(ctx, cudaMemcpyAsync(ctx->dErrK + (size_t)slot * ns, ctx->h_errk.data(), sizeof(double) * ns, cudaMemcpyHostToDevice,
                                 ctx->stream));
    return BO_OK;
}

template <int DP, int S>
static void launch_oz_kstar_fast(bo_ctx *ctx, int s, const double *dXc, int64_t c0, int mc, int mcp, int8_t *Kss,
                                 cudaStream_t st) {
    const int ntile = ctx->np / 64;
    const dim3 grid(mcp / 128, (ntile + OZ_KS_TILES - 1) / OZ_KS_TILES);
    const double *xs = ctx->dXs + (int64_t)s * ctx->np * ctx->dp, *hsq = ctx->dXsHalfSq + (int64_t)s * ctx->np;
    const double *ie = ctx->dInvEll + (int64_t)s * ctx->dp, *beta = ctx->dBeta + (int64_t)s * ctx->np;
    double *mup = ctx->dOzMu + (size_t)ctx->oz_mu_slot * ctx->ozmu_stride;
    // exp2 through the 16-entry table or the 12-term Horner chain (BO_OZ_EXP2_TABLE=0)
    static const bool tab = !(getenv("BO_OZ_EXP2_TABLE") && atoi(getenv("BO_OZ_EXP2_TABLE")) == 0);
#define OZ_KS_RUN(MAT, TB) oz_kstar_slices_fast_kernel<DP, S, MAT, TB><<<grid, OZ_KS_THREADS, S * OZ_A_SLICE_BYTES, st>>>( \
        ctx->n, ctx->np, ctx->d, xs, hsq, ie, dXc, c0, mc, mcp, Kss, beta, mup)
    if (ctx->kernel == BO_KERNEL_MATERN52) { if (tab) OZ_KS_RUN(true, true); else OZ_KS_RUN(true, false); }
    else { if (tab) OZ_KS_RUN(false, true); else OZ_KS_RUN(false, false); }
#undef OZ_KS_RUN
    ctx->oz_mu_rows[ctx->oz_mu_slot] = 2 * ((ntile + OZ_KS_TILES - 1) / OZ_KS_TILES);
}

template <int DP>
static int launch_oz_kstar(bo_ctx *ctx, int s, int S, const double *dXc, int64_t c0, int mc, int mcp, int8_t *Kss,
                           cudaStream_t st) {
    BO_LAUNCH_ON(ctx, "oz_kstar_slices_kernel", st);
    if (S >= 2 && S <= 5 && DP <= 16) {
        switch (S) {
            case 2: launch_oz_kstar_fast<DP, 2>(ctx, s, dXc, c0, mc, mcp, Kss, st); break;
            case 3: launch_oz_kstar_fast<DP, 3>(ctx, s, dXc, c0, mc, mcp, Kss, st); break;
            case 4: launch_oz_kstar_fast<DP, 4>(ctx, s, dXc, c0, mc, mcp, Kss, st); break;
            default: launch_oz_kstar_fast<DP, 5>(ctx, s, dXc, c0, mc, mcp, Kss, st); break;
        }
        BO_CHECK_LAUNCH(ctx);
        return BO_OK;
    }
    oz_kstar_slices_kernel<DP><<<dim3(mcp / 128, ctx->np / 64), 256, 0, st>>>(
        ctx->kernel, ctx->n, ctx->np, ctx->d, S, ctx->dXs + (int64_t)s * ctx->np * ctx->dp,
        ctx->dInvEll + (int64_t)s * ctx->dp, dXc, c0, mc, mcp, Kss, ctx->dBeta + (int64_t)s * ctx->np,
        ctx->dOzMu + (size_t)ctx->oz_mu_slot * ctx->ozmu_stride);
    ctx->oz_mu_rows[ctx->oz_mu_slot] = 2 * (ctx->np / 64);
    BO_CHECK_LAUNCH(ctx);
    return BO_OK;
}

// Scratch for `nbuf` slice buffers of up to mcp_max candidates each.
int bo_ozaki_reserve(bo_ctx *ctx, int S, int mcp_max, int nbuf) {
    ctx->kss_stride = (size_t)S * mcp_max * ctx->np;
    BO_TRY(bo_reserve(ctx, &ctx->dKss, &ctx->kss_capacity, ctx->kss_stride * nbuf));
    ctx->ozmu_stride = (size_t)2 * (ctx->np / 64) * mcp_max;       // per-buffer mean partials
    BO_TRY(bo_reserve(ctx, &ctx->dOzMu, &ctx->ozmu_capacity, ctx->ozmu_stride * nbuf));
    return BO_OK;
}

// Slice planes of K*^T for candidates [c0, c0 + mc) and hyper-sample s into buffer `buf`.
int bo_ozaki_slice(bo_ctx *ctx, int s, int S, const double *dXc, int64_t c0, int mc, int mcp, int buf,
                   cudaStream_t st) {
    int8_t *Kss = ctx->dKss + (size_t)buf * ctx->kss_stride;
    ctx->oz_mu_slot = buf;
    switch (ctx->dp) {
        case 2: BO_TRY(launch_oz_kstar<2>(ctx, s, S, dXc, c0, mc, mcp, Kss, st)); break;
        case 4: BO_TRY(launch_oz_kstar<4>(ctx, s, S, dXc, c0, mc, mcp, Kss, st)); break;
        case 8: BO_TRY(launch_oz_kstar<8>(ctx, s, S, dXc, c0, mc, mcp, Kss, st)); break;
        case 16: BO_TRY(launch_oz_kstar<16>(ctx, s, S, dXc, c0, mc, mcp, Kss, st)); break;
        case 32: BO_TRY(launch_oz_kstar<32>(ctx, s, S, dXc, c0, mc, mcp, Kss, st)); break;
        default: return bo_set_err(ctx, BO_ERR_ARG, "unsupported padded dimension %d", ctx->dp);
    }
    return BO_OK;
}

// mu_s, s2_s of the candidates sliced into buffer `buf`: the tcgen05 contraction + reductions.
int bo_ozaki_contract(bo_ctx *ctx, int s, int S, bool extra, int mcp, int buf, double *mu, double *s2, int32_t *dbg) {
    const int np = ctx->np;
    const int8_t *Kss = ctx->dKss + (size_t)buf * ctx->kss_stride;
    CUtensorMap tmB;
    // (the planes were built for ctx->oz_slices >= S digits; a lower level reads the leading S of them)
    BO_TRY(make_tmap(ctx, &tmB, ctx->dWs + (size_t)s * ctx->oz_slices * np * np, np, np, S, OZ_BN));
    const int nb = np / OZ_BN;
    {
        size_t need = (size_t)2 * nb * mcp;              // two partial sums per row block (one per half of the epilogue warps)
        BO_TRY(bo_reserve(ctx, &ctx->dOzQ, &ctx->ozpart_capacity, need));
    }
    OzParams p = {};
    p.nrb = np / OZ_BN; p.nkb = np / OZ_BK; p.full_k = 0;
    p.S = S; p.nstages = oz_stage_count(S); p.ntiles = mcp / OZ_BM; p.mcp = mcp;
    // exact int32 accumulation: up to S digit pairs per group, |digit| <= 128, k range <= np
    if ((int64_t)np * S >= (1 << 17))
        return bo_set_err(ctx, BO_ERR_ARG, "int8 path: n = %d with %d slices would overflow the int32 accumulators; use the FP64 path", np, S);
    p.nacc = (2 * (S + (extra ? 1 : 0)) * OZ_BN <= 512) ? 2 : 1;
    // candidate tiles whose K* slices (S * 128 * np bytes each) share L2 with the W slices
    {
        const double tile_bytes = (double)S * OZ_BM * np;
        const double budget = 0.35 * (double)ctx->prop.l2CacheSize;
        int T = (int)(budget / tile_bytes);
        p.tiles_per_group = T < 2 ? 2 : (T > 64 ? 64 : T);
    }
    p.rowscale = ctx->dRowScale + (size_t)s * np;
    p.qpart = ctx->dOzQ; p.dbg = dbg; p.kss = Kss;
    const int nunits = p.ntiles * nb;
    // Clusters of CL CTAs share the K* tile by TMA multicast (see oz_score_kernel): bo_set_option("oz_cluster", 1 / 2 / 4)
    // or BO_OZ_CLUSTER.  Measured at the headline shape: CL = 2 cuts the L2 -> SM operand traffic by a third and changes
    // neither the time (2.73 vs 2.72 ms per launch) nor the clock under the power cap (1687 MHz): the cap is set by the
    // tensor pipe itself working on random digits, not by operand movement.  CL = 4 is slower (4.87 ms: four rings in
    // lockstep, 37 clusters).  Default 1.
    static const int env_cl = getenv("BO_OZ_CLUSTER") ? atoi(getenv("BO_OZ_CLUSTER")) : 0;
    const int want_cl = ctx->oz_cluster > 0 ? ctx->oz_cluster : (env_cl > 0 ? env_cl : OZ_DEFAULT_CLUSTER);
    int cl = (S == 4 || S == 5) ? want_cl : 1;
    while (cl > 1 && (nb % cl != 0 || ctx->sm_count % cl != 0)) cl >>= 1;
    if (cl != 1 && cl != 2 && cl != 4) cl = 1;
    {
        BO_LAUNCH(ctx, "oz_score_kernel");
        if (cl == 1) {
            const int grid = nunits < ctx->sm_count ? nunits : ctx->sm_count;
            switch (S) {
#define OZ_RUN(SS) case SS: if (extra) oz_score_kernel<SS, 1, 0, 1><<<grid, OZ_THREADS, oz_smem_bytes(SS), ctx->stream>>>(tmB, p); \
                         else oz_score_kernel<SS, 0, 0, 1><<<grid, OZ_THREADS, oz_smem_bytes(SS), ctx->stream>>>(tmB, p); break
                OZ_RUN(2); OZ_RUN(3); OZ_RUN(4); OZ_RUN(5); OZ_RUN(6); OZ_RUN(7);
#undef OZ_RUN
                default: return bo_set_err(ctx, BO_ERR_ARG, "int8 path needs 2..7 slices, got %d", S);
            }
        } else {
            const int ncl_units = nunits / cl;                 // units of a cluster: (tile, group of cl row blocks)
            int nclusters = ctx->sm_count / cl;
            if (ncl_units < nclusters) nclusters = ncl_units;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(nclusters * cl));
            cfg.blockDim = dim3(OZ_THREADS);
            cfg.dynamicSmemBytes = oz_smem_bytes(S);
            cfg.stream = ctx->stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = (unsigned)cl;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
#define OZ_RUN_CL(SS, EE, CC) BO_CUDA(ctx, cudaLaunchKernelEx(&cfg, oz_score_kernel<SS, EE, 0, CC>, tmB, p))
            if (S == 5 && !extra && cl == 2) OZ_RUN_CL(5, 0, 2);
            else if (S == 5 && !extra && cl == 4) OZ_RUN_CL(5, 0, 4);
            else if (S == 5 && extra && cl == 2) OZ_RUN_CL(5, 1, 2);
            else if (S == 5 && extra && cl == 4) OZ_RUN_CL(5, 1, 4);
            else if (S == 4 && !extra && cl == 2) OZ_RUN_CL(4, 0, 2);
            else if (S == 4 && !extra && cl == 4) OZ_RUN_CL(4, 0, 4);
            else if (S == 4 && extra && cl == 2) OZ_RUN_CL(4, 1, 2);
            else OZ_RUN_CL(4, 1, 4);
#undef OZ_RUN_CL
        }
        BO_CHECK_LAUNCH(ctx);
    }
    {
        BO_LAUNCH(ctx, "oz_moments_kernel");
        oz_moments_kernel<<<(mcp + 127) / 128, 128, 0, ctx->stream>>>(
            2 * nb, ctx->oz_mu_rows[buf], mcp, ctx->dOzQ, ctx->dOzMu + (size_t)buf * ctx->ozmu_stride, ctx->h_rho[s],
            ctx->h_bias[s], mu, s2);
        BO_CHECK_LAUNCH(ctx);
    }
    return BO_OK;
}

// ---------------------------------------------------------------------------
// Thompson draws on the same int8 machinery (shared spectral basis):
//   F[r][i] = bias_r + scale_r * sum_j cos(w_j . x_i + b_j) theta_r[j]
// A operand = balanced base-256 digits of the cosine features (values in [-1, 1], generated and
// sliced on the fly per candidate chunk), B operand = digits of Theta (rows = draws, one exponent per
// draw, sliced once on the host when the draws are set), K = features.  MODE 1 of the contraction
// kernel writes the values and / or the per-draw first arg max.
// ---------------------------------------------------------------------------
// cos(a) for |a| < ~1e5: Cody-Waite reduction by pi/2 in two pieces, Taylor polynomials on |r| <= pi/4
// (cos to r^16, sin to r^15: truncation < 2e-17); no FP64 <-> integer conversion instructions.
__constant__ double OZ_COS_C[8] = {4.779477332387385e-14, -1.1470745597729725e-11, 2.08767569878681e-09, -2.755731922398589e-07,
                                   2.48015873015873e-05, -0.001388888888888889, 0.041666666666666664, -0.5};
__constant__ double OZ_SIN_C[7] = {-7.647163731819816e-13, 1.6059043836821613e-10, -2.505210838544172e-08, 2.7557319223985893e-06,
                                   -0.0001984126984126984, 0.008333333333333333, -0.16666666666666666};

// SKIP leading Taylor terms are dropped when the digits carry 32 bits or fewer (<= 4 slices): cos to r^12
// (truncation 4e-13), sin to r^11 (7e-12) against a digit quantum of 2^-32 / 127.
template <int SKIP>
__device__ __forceinline__ void oz_cos4(const double (&a)[4], double (&out)[4]) {
    double r[4], r2[4], pc[4], ps[4];
    int q[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const double zz = fma(a[e], 0.6366197723675814, 6755399441055744.0);       // a * 2/pi + 1.5 * 2^52
        q[e] = __double2loint(zz);
        const double qd = zz - 6755399441055744.0;
        r[e] = fma(-qd, 1.5707963267948966, a[e]);
        r[e] = fma(-qd, 6.123233995736766e-17, r[e]);
        r2[e] = r[e] * r[e];
        pc[e] = OZ_COS_C[SKIP];
        ps[e] = OZ_SIN_C[SKIP];
    }
#pragma unroll
    for (int c = SKIP + 1; c < 8; ++c) {
#pragma unroll
        for (int e = 0; e < 4; ++e) pc[e] = fma(pc[e], r2[e], OZ_COS_C[c]);
    }
#pragma unroll
    for (int c = SKIP + 1; c < 7; ++c) {
#pragma unroll
        for (int e = 0; e < 4; ++e) ps[e] = fma(ps[e], r2[e], OZ_SIN_C[c]);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const double cs = fma(pc[e], r2[e], 1.0);                 // cos r
        const double sn = fma(ps[e] * r2[e], r[e], r[e]);        // sin r
        const double v = (q[e] & 1) ? sn : cs;                    // cos(r + q pi/2): cs, -sn, -cs, sn
        out[e] = ((q[e] + 1) & 2) ? -v : v;
    }
}

// block: one candidate tile (128 candidates, one per thread) x up to OZ_KS_TILES blocks of 64 features;
// same staging and layout as oz_kstar_slices_fast_kernel.
template <int DP, int S>
__global__ void __launch_bounds__(128)
oz_cosine_slices_kernel(int m, int mp, int d, const double *__restrict__ Wp, const double *__restrict__ bp,
                        const double *__restrict__ Xc, int64_t c0, int mc, int mcp, int8_t *__restrict__ Ks) {
    __shared__ __align__(16) double ws[2][64][DP];
    __shared__ __align__(16) double bs[2][64];
    extern __shared__ __align__(16) uint8_t oz_stage[];          // [S][128 rows][64 B]
    const int tid = threadIdx.x;
    const int nkb = mp / 64, ntiles = mcp / 128;
    const int t0 = blockIdx.y * OZ_KS_TILES;
    const int t1 = (t0 + OZ_KS_TILES < nkb) ? t0 + OZ_KS_TILES : nkb;
    auto prefetch = [&](int tile, int buf) {
        const double *src = Wp + (int64_t)tile * 64 * DP;
        for (int e = tid; e < 64 * DP / 2; e += 128) oz_cp_async16(&ws[buf][0][0] + 2 * e, src + 2 * e);
        if (tid < 32) oz_cp_async16(&bs[buf][0] + 2 * tid, bp + (int64_t)tile * 64 + 2 * tid);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    prefetch(t0, 0);
    const int mm = blockIdx.x * 128 + tid;
    const bool live = mm < mc;
    double x[DP];
#pragma unroll
    for (int k = 0; k < DP; ++k) x[k] = (live && k < d) ? Xc[(c0 + mm) * d + k] : 0.0;
    const int swz = (tid >> 1) & 3;
    for (int tile = t0; tile < t1; ++tile) {
        const int buf = (tile - t0) & 1;
        if (tile + 1 < t1) {
            prefetch(tile + 1, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const int j0 = tile * 64;
#pragma unroll 1
        for (int sub = 0; sub < 4; ++sub) {
            const int jj0 = sub * 16;
            uint32_t wlow[16], wtop[4];
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                const int jq = jj0 + 4 * q4;
                double a[4], cv[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) a[e] = bs[buf][jq + e];
#pragma unroll
                for (int k = 0; k < DP; k += 2) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const double2 w2 = *reinterpret_cast<const double2 *>(&ws[buf][jq + e][k]);
                        a[e] = fma(x[k], w2.x, a[e]);
                        a[e] = fma(x[k + 1], w2.y, a[e]);
                    }
                }
                oz_cos4<(S <= 4 ? 2 : 0)>(a, cv);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int i = 4 * q4 + e;
                    const bool on = live && (j0 + jq + e) < m;
                    // X = rint(cos * 127 * 2^32) as a 40-bit two's complement integer (low mantissa bits of
                    // v + 1.5 * 2^52); bytes of X + 0x8080808080, each XOR 0x80, are its balanced digits
                    const double vv = fma(on ? cv[e] : 0.0, 545460846592.0, 6755399441055744.0);
                    const uint32_t lo = (uint32_t)__double2loint(vv) + 0x80808080u;
                    const uint32_t hi = ((uint32_t)__double2hiint(vv) + 0x80u + (lo < 0x80808080u ? 1u : 0u)) & 0xFFu;
                    wlow[i] = lo ^ 0x80808080u;
                    const uint32_t top = hi ^ 0x80u;
                    if (e == 0) wtop[q4] = top;
                    else wtop[q4] |= top << (8 * e);
                }
            }
            uint8_t *o0 = oz_stage + tid * 64 + ((sub ^ swz) << 4);
            *reinterpret_cast<uint4 *>(o0) = make_uint4(wtop[0], wtop[1], wtop[2], wtop[3]);
#pragma unroll
            for (int s = 1; s < S; ++s) {
                const uint32_t b = (uint32_t)(4 - s);                   // byte lane holding slice s
                const uint32_t sel2 = b | ((4u + b) << 4);
                uint32_t w[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const uint32_t lo2 = __byte_perm(wlow[4 * g + 0], wlow[4 * g + 1], sel2);
                    const uint32_t hi2 = __byte_perm(wlow[4 * g + 2], wlow[4 * g + 3], sel2);
                    w[g] = __byte_perm(lo2, hi2, 0x5410);
                }
                *reinterpret_cast<uint4 *>(o0 + s * OZ_A_SLICE_BYTES) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
        __syncthreads();
#pragma unroll
        for (int s = 0; s < S; ++s) {
            uint4 *dst = reinterpret_cast<uint4 *>(Ks + oz_kss_block(s, blockIdx.x, tile, ntiles, nkb));
            const uint4 *src = reinterpret_cast<const uint4 *>(oz_stage + s * OZ_A_SLICE_BYTES);
#pragma unroll
            for (int e = 0; e < OZ_A_SLICE_BYTES / 16 / 128; ++e) dst[e * 128 + tid] = src[e * 128 + tid];
        }
    }
}

template <int DP>
static int launch_oz_cosine(bo_ctx *ctx, int S, const double *dXc, int64_t c0, int mc, int mcp) {
    bo_thompson_state &th = ctx->th;
    const int nkb = th.oz_mp / 64;
    const dim3 grid(mcp / 128, (nkb + OZ_KS_TILES - 1) / OZ_KS_TILES);
    BO_LAUNCH(ctx, "oz_cosine_slices_kernel");
#define OZ_COS_RUN(SS) case SS: BO_CUDA(ctx, cudaFuncSetAttribute(oz_cosine_slices_kernel<DP, SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SS * OZ_A_SLICE_BYTES)); \
                       oz_cosine_slices_kernel<DP, SS><<<grid, 128, SS * OZ_A_SLICE_BYTES, ctx->stream>>>(th.m, th.oz_mp, th.d, th.ozWp, th.ozBp, dXc, c0, mc, mcp, th.ozPhi); break
    switch (S) {
        OZ_COS_RUN(3); OZ_COS_RUN(4); OZ_COS_RUN(5);
        default: return bo_set_err(ctx, BO_ERR_ARG, "Thompson int8 path handles 3..5 slices, got %d", S);
    }
#undef OZ_COS_RUN
    BO_CHECK_LAUNCH(ctx);
    return BO_OK;
}

// Level (slices, extra group) for the draws: the truncation error of F_r is ~ 8 sqrt(m) scale_r 2^e_r 256^-S
// (/32 with the extra group); `tol` is relative to the prior standard deviation of a draw,
// scale_r sqrt(m / 2) = sqrt(rho) -- the same reference as the scoring path.  (Posterior weights of an
// ill-conditioned feature system are large and cancel; their exponent, not their sum, sets the error.)
static void th_oz_choose(bo_ctx *ctx, int *S_out, int *extra_out) {
    bo_thompson_state &th = ctx->th;
    const double tol = ctx->prec_tol;
    if (tol >= 2.0) {
        int S = (int)tol;
        *extra_out = (tol - S) >= 0.5;
        *S_out = S > 5 ? 5 : (S < 3 ? 3 : S);
        return;
    }
    double mx = 0.0;
    for (size_t i = 0; i < th.h_theta.size(); ++i) mx = fabs(th.h_theta[i]) > mx ? fabs(th.h_theta[i]) : mx;
    int e = 0;
    if (mx > 0.0) frexp(mx, &e);
    const double worst = 8.0 * sqrt((double)th.oz_mp_hint) * ldexp(1.0, e) * sqrt(2.0 / (double)th.m);
    for (int S = 3; S <= 5; ++S) {
        const double base = worst * ldexp(1.0, -8 * S);
        if (base <= tol) { *S_out = S; *extra_out = 0; return; }
        if (base / 32.0 <= tol) { *S_out = S; *extra_out = 1; return; }
    }
    *S_out = 5;
    *extra_out = 1;
}

// Slice Theta on the host (ndraw x m values: microseconds) and upload the planes, scales and the padded basis.
static int th_oz_prepare(bo_ctx *ctx, int S) {
    bo_thompson_state &th = ctx->th;
    if (th.oz_ready && th.oz_S == S) return BO_OK;
    const int mp = bo_round_up(th.m, 64), ndp = bo_round_up(th.ndraw, 64), dpad = ctx_padded_dim(th.d);
    std::vector<int8_t> planes((size_t)S * ndp * mp, 0);
    std::vector<double> rs(ndp, 0.0), rb(ndp, 0.0);
    for (int r = 0; r < th.ndraw; ++r) {
        const double *t = th.h_theta.data() + (size_t)r * th.m;
        double mx = 0.0;
        for (int j = 0; j < th.m; ++j) mx = fabs(t[j]) > mx ? fabs(t[j]) : mx;
        int e = 0;
        if (mx > 0.0) frexp(mx, &e);
        const double sc = ldexp(1.0, -e);
        for (int j = 0; j < th.m; ++j) {
            const long long X = llrint(t[j] * sc * 35747322042253312.0);              // 127 * 2^48
            const unsigned long long Y = (unsigned long long)(X + OZ_DIGIT_BIAS7) ^ (unsigned long long)OZ_DIGIT_BIAS7;
            for (int s = 0; s < S; ++s) planes[((size_t)s * ndp + r) * mp + j] = (int8_t)(uint8_t)(Y >> (8 * (6 - s)));
        }
        rs[r] = th.h_scale[r] * ldexp(1.0, e) / 16129.0;
        rb[r] = th.h_bias[r];
    }
    std::vector<double> wp((size_t)mp * dpad, 0.0), bpv(mp, 0.0);
    for (int j = 0; j < th.m; ++j) {
        for (int k = 0; k < th.d; ++k) wp[(size_t)j * dpad + k] = th.h_W[(size_t)j * th.d + k];
        bpv[j] = th.h_b[j];
    }
    BO_TRY(bo_reserve(ctx, &th.ozTheta, &th.ozTheta_capacity, planes.size()));
    BO_TRY(bo_reserve(ctx, &th.ozRowScale, &th.ozRow_capacity, (size_t)2 * ndp));
    BO_TRY(bo_reserve(ctx, &th.ozWp, &th.ozWp_capacity, wp.size() + bpv.size()));
    th.ozRowBias = th.ozRowScale + ndp;
    th.ozBp = th.ozWp + wp.size();
    BO_CUDA(ctx, cudaMemcpyAsync(th.ozTheta, planes.data(), planes.size(), cudaMemcpyHostToDevice, ctx->stream));
    BO_CUDA(ctx, cudaMemcpyAsync(th.ozRowScale, rs.data(), sizeof(double) * ndp, cudaMemcpyHostToDevice, ctx->stream));
    BO_CUDA(ctx, cudaMemcpyAsync(th.ozRowBias, rb.data(), sizeof(double) * ndp, cudaMemcpyHostToDevice, ctx->stream));
    BO_CUDA(ctx, cudaMemcpyAsync(th.ozWp, wp.data(), sizeof(double) * wp.size(), cudaMemcpyHostToDevice, ctx->stream));
    BO_CUDA(ctx, cudaMemcpyAsync(th.ozBp, bpv.data(), sizeof(double) * bpv.size(), cudaMemcpyHostToDevice, ctx->stream));
    BO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    th.oz_mp = mp; th.oz_ndp = ndp; th.oz_dpad = dpad; th.oz_S = S; th.oz_ready = true;
    return BO_OK;
}

// 128-candidate tiles the int8 path reports per draw for M candidates
int64_t bo_thompson_ozaki_blocks(int64_t M) {
    const int64_t chunk = 256 * 128;
    const int64_t full = M / chunk, rem = M - full * chunk;
    return full * (chunk / OZ_BM) + bo_round_up64(rem, OZ_BM) / OZ_BM;
}

bool bo_thompson_ozaki_usable(bo_ctx *ctx, int64_t M) {
    const bo_thompson_state &th = ctx->th;
    return ctx->prec == BO_PREC_OZAKI && th.nW == 1 && M >= 1024 && !th.h_theta.empty() && th.ndraw <= 512 &&
           (int64_t)bo_round_up(th.m, 64) * 5 < (1 << 17);
}

int bo_thompson_ozaki_run(bo_ctx *ctx, int64_t M, const double *dXc, double *dOut, double *blkval, int64_t *blkidx,
                          int64_t blk_ld) {
    bo_thompson_state &th = ctx->th;
    int S = 5, extra = 0;
    th.oz_mp_hint = bo_round_up(th.m, 64);
    th_oz_choose(ctx, &S, &extra);
    BO_TRY(th_oz_prepare(ctx, S));
    th.oz_extra = extra;
    const int mp = th.oz_mp, ndp = th.oz_ndp;
    const int64_t chunk = 256 * 128;
    const int64_t cap = bo_round_up64(M < chunk ? M : chunk, 128);
    BO_TRY(bo_reserve(ctx, &th.ozPhi, &th.ozPhi_capacity, (size_t)S * cap * mp));
    CUtensorMap tmB;
    BO_TRY(make_tmap(ctx, &tmB, th.ozTheta, mp, ndp, S, OZ_BN));
    int64_t blk0 = 0;
    for (int64_t c0 = 0; c0 < M; c0 += chunk) {
        const int mc = (int)((M - c0) < chunk ? (M - c0) : chunk);
        const int mcp = bo_round_up(mc, 128);
        switch (th.oz_dpad) {
            case 2: BO_TRY(launch_oz_cosine<2>(ctx, S, dXc, c0, mc, mcp)); break;
            case 4: BO_TRY(launch_oz_cosine<4>(ctx, S, dXc, c0, mc, mcp)); break;
            case 8: BO_TRY(launch_oz_cosine<8>(ctx, S, dXc, c0, mc, mcp)); break;
            case 16: BO_TRY(launch_oz_cosine<16>(ctx, S, dXc, c0, mc, mcp)); break;
            default: BO_TRY(launch_oz_cosine<32>(ctx, S, dXc, c0, mc, mcp)); break;
        }
        OzParams p = {};
        p.nrb = ndp / OZ_BN; p.nkb = mp / OZ_BK; p.full_k = 1;
        p.S = S; p.nstages = oz_stage_count_m1(S); p.ntiles = mcp / OZ_BM; p.mcp = mcp;
        p.nacc = (2 * (S + extra) * OZ_BN <= 512) ? 2 : 1;
        p.tiles_per_group = 64;
        p.rowscale = th.ozRowScale; p.rowbias = th.ozRowBias; p.kss = th.ozPhi;
        p.out = dOut; p.out_ld = M; p.c0 = c0; p.mc = mc; p.nrows_live = th.ndraw;
        p.blkval = blkval; p.blkidx = blkidx; p.blk_ld = blk_ld; p.blk0 = blk0;
        const int nunits = p.ntiles * p.nrb;
        const int grid = nunits < ctx->sm_count ? nunits : ctx->sm_count;
        {
            BO_LAUNCH(ctx, "oz_thompson_kernel");
            switch (S) {
#define OZ_TRUN(SS) case SS: if (extra) oz_score_kernel<SS, 1, 1, 1><<<grid, OZ_THREADS_M1, oz_smem_bytes_m1(SS), ctx->stream>>>(tmB, p); \
                          else oz_score_kernel<SS, 0, 1, 1><<<grid, OZ_THREADS_M1, oz_smem_bytes_m1(SS), ctx->stream>>>(tmB, p); break
                OZ_TRUN(3); OZ_TRUN(4); OZ_TRUN(5);
#undef OZ_TRUN
                default: return bo_set_err(ctx, BO_ERR_ARG, "Thompson int8 path handles 3..5 slices, got %d", S);
            }
            BO_CHECK_LAUNCH(ctx);
        }
        blk0 += mcp / OZ_BM;
    }
    return BO_OK;
}

// Debug / self-test entry point: runs the int8-slice path for the first `mc` candidates
// of Xc (host) on hyper-sample 0 and returns mu, s2 plus the raw int32 group accumulators
// of candidate tile 0 ([np/64][S][128][64]), the W slices and the K* slices.
extern "C" int bo_ozaki_debug(bo_ctx *ctx, int S, int extra, int mc, const double *Xc, double *mu, double *s2,
                              int32_t *acc, int8_t *wslices, int8_t *kslices, double *rowscale) {
    if (!ctx) return BO_ERR_ARG;
    BO_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->fitted) return bo_set_err(ctx, BO_ERR_STATE, "bo_ozaki_debug before bo_fit");
    if (S < 2 || S > OZ_MAX_S || mc < 1) return bo_set_err(ctx, BO_ERR_ARG, "bad S / mc");
    const int np = ctx->np, mcp = bo_round_up(mc, OZ_BM), nb = np / OZ_BN;
    ctx->oz_ready = false;
    ctx->oz_extra = extra != 0;
    const int NG = S + (ctx->oz_extra ? 1 : 0);
    BO_TRY(bo_ozaki_prepare(ctx, S));
    double *dXc = nullptr, *dmu = nullptr;
    int32_t *dacc = nullptr;
    BO_CUDA(ctx, cudaMalloc(&dXc, sizeof(double) * mc * ctx->d));
    BO_CUDA(ctx, cudaMalloc(&dmu, sizeof(double) * 2 * mcp));
    BO_CUDA(ctx, cudaMalloc(&dacc, sizeof(int32_t) * (size_t)nb * NG * 128 * 64));
    BO_CUDA(ctx, cudaMemsetAsync(dacc, 0xff, sizeof(int32_t) * (size_t)nb * NG * 128 * 64, ctx->stream));
    BO_CUDA(ctx, cudaMemcpyAsync(dXc, Xc, sizeof(double) * mc * ctx->d, cudaMemcpyHostToDevice, ctx->stream));
    int rc = bo_ozaki_reserve(ctx, S, mcp, 1);
    if (rc == BO_OK) rc = bo_ozaki_slice(ctx, 0, S, dXc, 0, mc, mcp, 0, ctx->stream);
    if (rc == BO_OK) rc = bo_ozaki_contract(ctx, 0, S, ctx->oz_extra, mcp, 0, dmu, dmu + mcp, acc ? dacc : nullptr);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (rc == BO_OK && e == cudaSuccess) {
        if (mu) cudaMemcpy(mu, dmu, sizeof(double) * mc, cudaMemcpyDeviceToHost);
        if (s2) cudaMemcpy(s2, dmu + mcp, sizeof(double) * mc, cudaMemcpyDeviceToHost);
        if (acc) cudaMemcpy(acc, dacc, sizeof(int32_t) * (size_t)nb * NG * 128 * 64, cudaMemcpyDeviceToHost);
        if (wslices) cudaMemcpy(wslices, ctx->dWs, (size_t)S * np * np, cudaMemcpyDeviceToHost);
        if (kslices) {       // present the blocked / swizzled planes as plain [s][candidate][k]
            std::vector<int8_t> raw((size_t)S * mcp * np);
            cudaMemcpy(raw.data(), ctx->dKss, raw.size(), cudaMemcpyDeviceToHost);
            for (int sl = 0; sl < S; ++sl)
                for (int m = 0; m < mcp; ++m)
                    for (int j = 0; j < np; ++j)
                        kslices[((size_t)sl * mcp + m) * np + j] = raw[oz_kss_offset(sl, m, j, mcp >> 7, np >> 6)];
        }
        if (rowscale) cudaMemcpy(rowscale, ctx->dRowScale, sizeof(double) * np, cudaMemcpyDeviceToHost);
    }
    cudaFree(dXc); cudaFree(dmu); cudaFree(dacc);
    if (rc != BO_OK) return rc;
    if (e != cudaSuccess) return bo_set_err(ctx, BO_ERR_CUDA, "bo_ozaki_debug: %s", cudaGetErrorString(e));
    return BO_OK;
}
