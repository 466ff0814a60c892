// Float64 dense layers on the 5th-generation tensor cores (tcgen05, kind::i8) - sm_100a only.
//
// The PPO update (agents/agent_pg.py:19-26, agents/agent_ppo.py:44-51) is dominated by the GEMMs of the two MLPs
// (models/mlp.py:22-25, core/policy_gaussian.py:19-24, core/critic.py:15-18) over the whole trajbatch in float64
// (ego_pose/ego_mimic.py:31-32).  tcgen05 has no FP64 kind, so the product is evaluated with the Ozaki scheme on the
// int8 tensor cores (B200 has them, 2x the bf16 rate).  With the default radix (ozaki.cuh: RB = 7):
//
//   row i of A:   a_ik = 2^ea_i * sum_{t=1..S} qa_t[i][k] 2^(1-7t),  qa_t in [-64, 64]  (int8 "slices": the signed
//   row j of B:   b_jk = 2^eb_j * sum_{u=1..S} qb_u[j][k] 2^(1-7u)    base-128 digits of round(a 2^(7S-1-e)))
//   C_ij = sum_k a_ik b_jk = 2^(ea_i + eb_j - 12) * sum_{d=0..S-1} 2^(-7d) * [ sum_{t+u-2=d} sum_k qa_t qb_u ]
//
// Every bracket is an int8 x int8 -> int32 GEMM accumulated EXACTLY in Tensor Memory (one accumulator per d, all
// slice pairs of equal weight share it); pairs with t+u > S+1 are below the slicing residual and dropped.  The
// result differs from the float64 product by <= (S+2) 2^(-7S) |a_i|_max |b_j|_max K  (S = 6: 2e-12, S = 7: 1.6e-14,
// the rounding level of a DGEMM itself) and is independent of tile shapes and launch geometry (integer
// accumulation), so the kernel is tested bit-exactly against an integer matmul of the same slices.
//
// Kernels:
//   oz_slice_rows_kernel   f64 [M][K] -> int8 [S][M][Kp] + per-row exponent (scale constant along K = columns)
//   oz_slice_colsT_kernel  f64 [N][F] -> int8 [S][F][Np] + per-column exponent, transposed (for the weight-gradient
//                          GEMMs whose contraction runs over the samples); optional extra row of ones (bias gradient)
//   oz_gemm_kernel         persistent warp-specialised tcgen05 GEMM, one CTA per SM looping over (split, m, n) tiles:
//                          warp 0 TMA producer (3-D boxes {32 B, rows, S slices}, 32-byte swizzle) -> mbarrier ring that
//                          runs ahead across tiles -> warp 1 (uniform loop, one elected lane) tcgen05.mma.kind::i8 into S TMEM
//                          accumulators, A slice kept in the operand collector across the B slices it pairs with
//                          -> 8 epilogue warps: tcgen05.ld, exact int64 Horner over d, one rounding to float64, scale,
//                          bias, relu / relu-backward mask, store; TMEM is handed back to the MMA warp as soon as the
//                          last accumulator chunk is in registers
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "ozaki.cuh"

namespace egp {
namespace oz {

constexpr int UMMA_K = 32;                      // K per tcgen05.mma for 8-bit operands
#ifndef OZ_EPI_WARPS
#define OZ_EPI_WARPS 16
#endif
#ifndef OZ_COLC_SHFL
#define OZ_COLC_SHFL 0
#endif
constexpr int EPI_WARPS = OZ_EPI_WARPS;             // 8: two warps per lane quadrant (half a tile's columns each); 16: four (a quarter each)
constexpr int EPI_PARTS = EPI_WARPS / 4;            // column parts of a tile
constexpr int CW = EPI_PARTS == 2 ? 8 : 4;          // columns per drain / post chunk (registers: 576 threads leave 96 per thread)
constexpr int OUT_BUF_BYTES = 32 * CW * 8;          // one output staging buffer: 32 rows x CW doubles
static_assert(EPI_WARPS == 8 || EPI_WARPS == 16, "epilogue warps");
#ifndef OZ_OUT_BUFS
#define OZ_OUT_BUFS 2
#endif
constexpr int OUT_BUFS = OZ_OUT_BUFS;               // 32 x 64 B output staging buffers per epilogue warp (TMA stores in flight)
constexpr int GEMM_THREADS = 64 + 32 * EPI_WARPS;   // warp 0: TMA producer, warp 1: MMA issuer, the others: epilogue
constexpr int SMEM_LIMIT = 227 * 1024 - 2048;   // dynamic shared memory: 227 KB minus the static barriers
constexpr int MAX_STAGES = 12;
#ifndef OZ_WIDE_N
#define OZ_WIDE_N 0      // 1: merge the B slices of consecutive accumulators into one wide instruction (measured: no change)
#endif

// ---------------------------------------------------------------------------------------------------------------
// slicing
// exponent e with |x| / 2^e < 1 for |x| <= amax (amax / 2^e in [0.5, 1)); 0 for a zero / non-finite row; rows below
// 2^-900 are flushed to zero (the scale 2^(7S-1-e) must stay finite)
__device__ __forceinline__ int oz_exponent(double amax) {
    if (!(amax > 0.0) || !isfinite(amax)) return 0;
    int e;
    frexp(amax, &e);
    return e < -900 ? -900 : e;
}

__device__ __forceinline__ double pow2(int k) {            // 2^k, k in [-1022, 1023]
    return __longlong_as_double((long long)(k + 1023) << 52);
}

// Slices q_1..q_S of X = round(x * scale), X = sum_t q_t 2^(RB (S - t)).
// RB = 8: the two's-complement bytes of X (q_1 signed, q_2.. unsigned), |X| clamped to 2^(8S-1) - 1.
// RB = 7: signed digits in [-64, 64], peeled off the low end in 21-bit groups with 32-bit integer arithmetic.
__device__ __forceinline__ int oz_low_digit(int &x) {
    const int d = ((x + 64) & 127) - 64;
    x = (x - d) >> 7;
    return d;
}
template <int S>
__device__ __forceinline__ void oz_digits(double x, double scale, int *q) {
    double r = x * scale;
    if (!(fabs(r) <= 9.2e18)) r = 0.0;         // NaN / inf / out of int64 range
    long long X = __double2ll_rn(r);
    if constexpr (RB == 8) {
        constexpr long long LIM = (1LL << (8 * S - 1)) - 1;
        X = X > LIM ? LIM : (X < -LIM ? -LIM : X);
        const unsigned lo = (unsigned)X, hi = (unsigned)(X >> 32);
#pragma unroll
        for (int t = 0; t < S; t++) {
            const int sh = 8 * (S - 1 - t);                     // byte t (0 = most significant) of the 8S-bit value
            const unsigned b = sh >= 32 ? (hi >> (sh - 32)) : (sh == 0 ? lo : ((lo >> sh) | (sh > 24 ? (hi << (32 - sh)) : 0u)));
            q[t] = (int)(b & 255u);
        }
    } else {
        constexpr int NG = (S - 1) / 3;            // full 3-digit groups taken from the low end (leaves 1..3 digits)
#pragma unroll
        for (int gi = 0; gi < NG; gi++) {
            const int t = S - 1 - 3 * gi;
            int xl = (int)(((unsigned)X + (1u << 20)) & ((1u << 21) - 1u)) - (1 << 20);
            X = (X - xl) >> 21;
            q[t] = oz_low_digit(xl);
            q[t - 1] = oz_low_digit(xl);
            q[t - 2] = xl;
        }
        int xh = (int)X;
#pragma unroll
        for (int t = S - 1 - 3 * NG; t > 0; t--) q[t] = oz_low_digit(xh);
        q[0] = xh;
    }
}

// Radix-128 fast path for four elements at once.  Adding C = sum_i 64 * 128^i to X turns the balanced digits into the plain
// 7-bit fields of Y = X + C (digit_i = field_i - 64, carries resolved by the one 64-bit add), so every slice is a shift +
// mask per element, a byte merge and one SWAR per-byte subtraction - no sequential digit peeling.
template <int S>
__device__ __forceinline__ void oz_y(double x, double scale, unsigned &lo, unsigned &hi) {
    double r = x * scale;
    if (!(fabs(r) <= 9.2e18)) r = 0.0;         // NaN / inf / out of int64 range
    long long C = 0;
#pragma unroll
    for (int i = 0; i < S; i++) C = C * 128 + 64;
    const long long Y = __double2ll_rn(r) + C;
    lo = (unsigned)Y;
    hi = (unsigned)(Y >> 32);
}
// packed int8 digits of slice t (0 = most significant) of four elements
template <int S>
__device__ __forceinline__ uint32_t oz_pack4(const unsigned (&lo)[4], const unsigned (&hi)[4], int t) {
    const int sh = 7 * (S - 1 - t);
    const unsigned mask = t == 0 ? 255u : 127u;               // the top field can reach 128 (digit +64)
    unsigned f[4];
#pragma unroll
    for (int j = 0; j < 4; j++) f[j] = (sh >= 32 ? (hi[j] >> (sh - 32)) : __funnelshift_r(lo[j], hi[j], sh)) & mask;
    const unsigned W = f[0] | (f[1] << 8) | (f[2] << 16) | (f[3] << 24);
    const unsigned H = 0x80808080u, Cb = 0x40404040u;
    return ((W | H) - Cb) ^ (~W & H);                         // per-byte W - 64 (two's complement bytes)
}

// One warp per row (grid-stride).  Lane l owns the 4-element groups l, l + 32, ... of the row: 32 B loads, 4 B stores
// per slice (128 B per warp and slice).  Optionally accumulates the column abs-max of the matrix into colmax (bit
// patterns of |x| >= 0 ordered like unsigned integers).
// exponent from the high word of the row / column abs-max (monotonic for non-negative doubles): e with max / 2^e in
// [0.5, 1); zero / subnormal rows get the floor -900 (their digits are zero), non-finite rows 0
__device__ __forceinline__ int oz_exponent_hi(unsigned h) {
    if (h >= 0x7ff00000u) return 0;
    const int e = (int)(h >> 20) - 1022;
    return e < -900 ? -900 : e;
}

__device__ __forceinline__ void cp_async_n(void *smem_dst, const void *gsrc, int bytes, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? bytes : 0;            // src-size 0: the destination bytes are zero-filled
    if (bytes == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}

// One warp per row (grid-stride), 4 warps per block.  The next row of a warp streams into the other half of its private
// shared-memory double buffer with cp.async while the current one is sliced (every lane reads back exactly the bytes it
// requested, so no barrier is involved).
constexpr int ROWS_WARPS = 4;
template <int S, int P>
__global__ void __launch_bounds__(32 * ROWS_WARPS)
oz_slice_rows_kernel(const double *__restrict__ x, long long M, int K, long long ldx, int8_t *__restrict__ out, int Kp,
                     int32_t *__restrict__ exps, unsigned long long *__restrict__ colmax) {
    extern __shared__ __align__(16) double rows_smem[];             // [ROWS_WARPS][2][P * 128] doubles, then the reduction words
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long warp = (long long)blockIdx.x * ROWS_WARPS + w;
    const long long nwarp = (long long)gridDim.x * ROWS_WARPS;
    const int ngroup = Kp / 4;
    const bool vec = ((ldx & 1) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    double *mybuf = rows_smem + (size_t)w * 2 * (P * 128);
    unsigned cmax[P][4];                        // running column maxima as high words
#pragma unroll
    for (int p = 0; p < P; p++)
#pragma unroll
        for (int j = 0; j < 4; j++) cmax[p][j] = 0u;
    auto issue = [&](long long row, int buf) {
        const double *xr = x + row * ldx;
        double *b = mybuf + buf * (P * 128);
#pragma unroll
        for (int p = 0; p < P; p++) {
            const int k0 = (lane + 32 * p) * 4;
            double *dst = b + k0;
            if (vec && k0 + 3 < K) {
                cp_async_n(dst, xr + k0, 16, true);
                cp_async_n(dst + 2, xr + k0 + 2, 16, true);
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) cp_async_n(dst + j, k0 + j < K ? xr + k0 + j : x, 8, k0 + j < K);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (warp < M) issue(warp, 0);
    int buf = 0;
    for (long long row = warp; row < M; row += nwarp, buf ^= 1) {
        if (row + nwarp < M) {
            issue(row + nwarp, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        const double *b = mybuf + buf * (P * 128);
        double v[P][4];
        unsigned hmax = 0u;
#pragma unroll
        for (int p = 0; p < P; p++) {
            const int k0 = (lane + 32 * p) * 4;
            const double2 a0 = *reinterpret_cast<const double2 *>(b + k0);
            const double2 a1 = *reinterpret_cast<const double2 *>(b + k0 + 2);
            v[p][0] = a0.x; v[p][1] = a0.y; v[p][2] = a1.x; v[p][3] = a1.y;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const unsigned h = (unsigned)__double2hiint(v[p][j]) & 0x7fffffffu;
                hmax = max(hmax, h);
                cmax[p][j] = max(cmax[p][j], h);
            }
        }
        hmax = __reduce_max_sync(0xffffffffu, hmax);
        const int e = oz_exponent_hi(hmax);
        if (lane == 0) exps[row] = e;
        const double scale = pow2(RB * S - 1 - e);
#pragma unroll
        for (int p = 0; p < P; p++) {
            const int gi = lane + 32 * p;
            if (gi >= ngroup) continue;
            uint32_t pk[S];
            if constexpr (RB == 7) {
                unsigned ylo[4], yhi[4];
#pragma unroll
                for (int j = 0; j < 4; j++) oz_y<S>(v[p][j], scale, ylo[j], yhi[j]);
#pragma unroll
                for (int t = 0; t < S; t++) pk[t] = oz_pack4<S>(ylo, yhi, t);
            } else {
#pragma unroll
                for (int t = 0; t < S; t++) pk[t] = 0u;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    int q[S];
                    oz_digits<S>(v[p][j], scale, q);
#pragma unroll
                    for (int t = 0; t < S; t++) pk[t] |= (uint32_t)(q[t] & 255) << (8 * j);
                }
            }
#pragma unroll
            for (int t = 0; t < S; t++)
                *reinterpret_cast<uint32_t *>(out + ((size_t)t * M + row) * Kp + gi * 4) = pk[t];
        }
    }
    if (colmax) {
        // block-level maximum first (all warps own the same columns), then one atomic per column and block
        // (only the exponent of the maximum is consumed: the high words suffice)
        uint32_t *red = reinterpret_cast<uint32_t *>(rows_smem + (size_t)ROWS_WARPS * 2 * (P * 128));
#pragma unroll
        for (int p = 0; p < P; p++)
#pragma unroll
            for (int j = 0; j < 4; j++) red[w * (P * 128) + (lane + 32 * p) * 4 + j] = cmax[p][j];
        __syncthreads();
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            uint32_t m = red[k];
#pragma unroll
            for (int i = 1; i < ROWS_WARPS; i++) m = max(m, red[i * (P * 128) + k]);
            if (m > 0u) atomicMax(colmax + k, (unsigned long long)m << 32);
        }
    }
}

// column abs-max only (when the row slices are not needed): block = 256 threads = 8 row lanes x 32 features
__global__ void __launch_bounds__(256)
oz_colmax_kernel(const double *__restrict__ x, long long N, int F, long long ldx, long long rows_per_block,
                 unsigned long long *__restrict__ colmax) {
    __shared__ double red[8][33];
    const long long r0 = blockIdx.x * rows_per_block;
    const long long r1 = r0 + rows_per_block < N ? r0 + rows_per_block : N;
    const int fl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    for (int f0 = 0; f0 < F; f0 += 32) {
        const int f = f0 + fl;
        double m = 0.0;
        if (f < F)
            for (long long r = r0 + rl; r < r1; r += 8) m = fmax(m, fabs(x[r * ldx + f]));
        red[rl][fl] = m;
        __syncthreads();
        if (rl == 0 && f < F) {
#pragma unroll
            for (int i = 1; i < 8; i++) m = fmax(m, red[i][fl]);
            if (m > 0.0) atomicMax(colmax + f, (unsigned long long)__double_as_longlong(m));
        }
        __syncthreads();
    }
}

__global__ void oz_col_exps_kernel(const unsigned long long *__restrict__ colmax, int F, int ones_row, int32_t *__restrict__ exps) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < F) exps[f] = oz_exponent(__longlong_as_double((long long)colmax[f]));
    else if (f == F && ones_row) exps[f] = 1;
}

// Transposed, column-scaled slices: x [N][F] -> out [S][F (+1)][Np].  Block tile = 128 samples x 32 features: coalesced
// f64 reads along the features, shared-memory transpose, 16 B stores along the samples.  With ones_row the virtual
// feature F is 1.0 for every valid sample (its products are the column sums = bias gradients).
// Each block walks TPB consecutive 128-sample tiles of one 32-feature panel; the next tile streams into the other half of
// a double buffer with cp.async (8 B per element, zero-filled outside the matrix) while the current one is sliced.
constexpr int COLST_TPB = 4;

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 8 : 0;                // src-size 0: the 8 destination bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}

template <int S>
__global__ void __launch_bounds__(256)
oz_slice_colsT_kernel(const double *__restrict__ x, long long N, int F, long long ldx, const int32_t *__restrict__ exps,
                      int8_t *__restrict__ out, long long Np, int ones_row) {
    extern __shared__ __align__(16) double colst_tile[];           // [2][128][32], columns rotated by 2 per 16-row group
    const int f0 = blockIdx.y * 32;
    const int FT = F + (ones_row ? 1 : 0);
    const long long ntile = (Np + 127) / 128;
    const long long t0 = (long long)blockIdx.x * COLST_TPB;
    const long long t1 = t0 + COLST_TPB < ntile ? t0 + COLST_TPB : ntile;
    const int lfl = threadIdx.x & 31, lrl = threadIdx.x >> 5;       // loader: feature lane, 8 rows per pass
    auto issue = [&](long long tile_idx, int buf) {
        double *tb = colst_tile + buf * (128 * 32);
        const long long n0 = tile_idx * 128;
        const int f = f0 + lfl;
#pragma unroll 4
        for (int r = lrl; r < 128; r += 8) {
            const long long n = n0 + r;
            double *dst = tb + r * 32 + ((lfl + 2 * (r >> 4)) & 31);
            const bool ok = n < N && f < F;
            cp_async8(dst, ok ? x + n * ldx + f : x, ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // thread -> (feature fl, group of 16 samples g): 32 x 8 = 256 threads
    const int g = threadIdx.x & 7, fl = threadIdx.x >> 3;
    const int f = f0 + fl;
    const int fc = (fl + 2 * g) & 31;
    const double scale = f < FT ? pow2(RB * S - 1 - exps[f]) : 0.0;
    const bool ones = ones_row && f == F;
    if (t0 < t1) issue(t0, 0);
    for (long long ti = t0; ti < t1; ti++) {
        const int buf = (int)((ti - t0) & 1);
        if (ti + 1 < t1) {
            issue(ti + 1, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const double *tb = colst_tile + buf * (128 * 32);
        const long long nb = ti * 128 + g * 16;
        if (f < FT && nb < Np) {
            uint32_t pk[S][4];
#pragma unroll
            for (int jq = 0; jq < 4; jq++) {
                double v[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    v[j] = tb[(g * 16 + jq * 4 + j) * 32 + fc];
                    if (ones) v[j] = (nb + jq * 4 + j < N) ? 1.0 : 0.0;
                }
                if constexpr (RB == 7) {
                    unsigned ylo[4], yhi[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) oz_y<S>(v[j], scale, ylo[j], yhi[j]);
#pragma unroll
                    for (int t = 0; t < S; t++) pk[t][jq] = oz_pack4<S>(ylo, yhi, t);
                } else {
#pragma unroll
                    for (int t = 0; t < S; t++) pk[t][jq] = 0u;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        int q[S];
                        oz_digits<S>(v[j], scale, q);
#pragma unroll
                        for (int t = 0; t < S; t++) pk[t][jq] |= (uint32_t)(q[t] & 255) << (8 * j);
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < S; t++)
                *reinterpret_cast<uint4 *>(out + ((size_t)t * FT + f) * Np + nb) = make_uint4(pk[t][0], pk[t][1], pk[t][2], pk[t][3]);
        }
        __syncthreads();                                            // the buffer is refilled two iterations later
    }
}


// packed int8 digits of four elements for all S slices
template <int S>
__device__ __forceinline__ void oz_slice4(const double (&v)[4], double scale, uint32_t (&pk)[S]) {
    if constexpr (RB == 7) {
        unsigned ylo[4], yhi[4];
#pragma unroll
        for (int j = 0; j < 4; j++) oz_y<S>(v[j], scale, ylo[j], yhi[j]);
#pragma unroll
        for (int t = 0; t < S; t++) pk[t] = oz_pack4<S>(ylo, yhi, t);
    } else {
#pragma unroll
        for (int t = 0; t < S; t++) pk[t] = 0u;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int q[S];
            oz_digits<S>(v[j], scale, q);
#pragma unroll
            for (int t = 0; t < S; t++) pk[t] |= (uint32_t)(q[t] & 255) << (8 * j);
        }
    }
}

// BOTH slice orientations of x [N][F] from ONE read: the transposed column-scaled slices of oz_slice_colsT_kernel and the
// row-scaled slices of oz_slice_rows_kernel, when the row and column abs-maxima are already known (the producing GEMM's
// epilogue or the value head's backward kernel recorded them).  Same tiles, loader and column rotation as the transposed
// slicer; after the transposed pass over a 128 x 32 tile, warp w re-reads rows 16 w .. 16 w + 15 of the tile (lane ->
// row l / 8, features 4 (l % 8) ..) and writes 4 bytes per slice: 8 lanes fill one 32-byte sector of a slice row (the row
// pitch Kp is a multiple of 32).  Columns F .. Kp - 1 come out as zero digits (zero-filled loads).  The results are
// bit-identical to the two separate kernels (same exponents, same digit arithmetic).
template <int S>
__global__ void __launch_bounds__(256)
oz_slice_both_kernel(const double *__restrict__ x, long long N, int F, long long ldx, const int32_t *__restrict__ exps,
                     int8_t *__restrict__ out, long long Np, int ones_row, const uint32_t *__restrict__ rowmax,
                     int8_t *__restrict__ outR, int Kp, int32_t *__restrict__ expsR) {
    extern __shared__ __align__(16) double colst_tile[];           // [2][128][32], columns rotated by 2 per 16-row group
    const int f0 = blockIdx.x * 32;         // panels vary fastest: the blocks that fill the sectors of one slice row run together
    const int FT = F + (ones_row ? 1 : 0);
    const long long ntile = (Np + 127) / 128;
    const long long t0 = (long long)blockIdx.y * COLST_TPB;
    const long long t1 = t0 + COLST_TPB < ntile ? t0 + COLST_TPB : ntile;
    const int lfl = threadIdx.x & 31, lrl = threadIdx.x >> 5;       // loader: feature lane, 8 rows per pass
    auto issue = [&](long long tile_idx, int buf) {
        double *tb = colst_tile + buf * (128 * 32);
        const long long n0 = tile_idx * 128;
        const int f = f0 + lfl;
#pragma unroll 4
        for (int r = lrl; r < 128; r += 8) {
            const long long n = n0 + r;
            double *dst = tb + r * 32 + ((lfl + 2 * (r >> 4)) & 31);
            const bool ok = n < N && f < F;
            cp_async8(dst, ok ? x + n * ldx + f : x, ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int g = threadIdx.x & 7, fl = threadIdx.x >> 3;
    const int f = f0 + fl;
    const int fc = (fl + 2 * g) & 31;
    const double scale = f < FT ? pow2(RB * S - 1 - exps[f]) : 0.0;
    const bool ones = ones_row && f == F;
    // row pass: warp lrl owns tile rows 16 lrl .., lane -> (row lfl / 8 of a 4-row step, feature group lfl % 8)
    const int rfg = lfl & 7, rc0 = f0 + 4 * rfg;
    const int rp0 = (4 * rfg + 2 * lrl) & 31, rp1 = (4 * rfg + 2 + 2 * lrl) & 31;
    if (t0 < t1) issue(t0, 0);
    for (long long ti = t0; ti < t1; ti++) {
        const int buf = (int)((ti - t0) & 1);
        if (ti + 1 < t1) {
            issue(ti + 1, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const double *tb = colst_tile + buf * (128 * 32);
        const long long nb = ti * 128 + g * 16;
        if (f < FT && nb < Np) {
            uint32_t pk[S][4];
#pragma unroll
            for (int jq = 0; jq < 4; jq++) {
                double v[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    v[j] = tb[(g * 16 + jq * 4 + j) * 32 + fc];
                    if (ones) v[j] = (nb + jq * 4 + j < N) ? 1.0 : 0.0;
                }
                uint32_t p4[S];
                oz_slice4<S>(v, scale, p4);
#pragma unroll
                for (int t = 0; t < S; t++) pk[t][jq] = p4[t];
            }
#pragma unroll
            for (int t = 0; t < S; t++)
                *reinterpret_cast<uint4 *>(out + ((size_t)t * FT + f) * Np + nb) = make_uint4(pk[t][0], pk[t][1], pk[t][2], pk[t][3]);
        }
        if (rc0 < Kp) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int r = lrl * 16 + i * 4 + (lfl >> 3);
                const long long n = ti * 128 + r;
                if (n < N) {
                    const int e = oz_exponent_hi(rowmax[n]);
                    if (blockIdx.x == 0 && rfg == 0) expsR[n] = e;
                    const double rscale = pow2(RB * S - 1 - e);
                    const double2 a0 = *reinterpret_cast<const double2 *>(tb + r * 32 + rp0);
                    const double2 a1 = *reinterpret_cast<const double2 *>(tb + r * 32 + rp1);
                    const double v[4] = {a0.x, a0.y, a1.x, a1.y};
                    uint32_t p4[S];
                    oz_slice4<S>(v, rscale, p4);
#pragma unroll
                    for (int t = 0; t < S; t++)
                        *reinterpret_cast<uint32_t *>(outR + ((size_t)t * N + n) * Kp + rc0) = p4[t];
                }
            }
        }
        __syncthreads();                                            // the buffer is refilled two iterations later
    }
}

// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers (sm_100a)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"((uint64_t)tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], int8 x int8 -> int32.  Shared-memory descriptors are passed as their low words (start
// address >> 4, leading byte offset 1) plus the constant high word, so the issuing loop only does 32-bit uniform adds.
// COLL: A-operand collector usage - 0 discard (default), 1 fill, 2 use, 3 lastuse: consecutive MMAs of one A slice with
// different B slices keep A in the collector buffer instead of re-reading it from shared memory.
template <int COLL>
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate) {
#define OZ_MMA_ASM(Q)                                                                                     \
    asm volatile(                                                                                         \
        "{\n\t"                                                                                           \
        ".reg .pred p;\n\t"                                                                               \
        ".reg .b64 da, db;\n\t"                                                                           \
        "setp.ne.b32 p, %5, 0;\n\t"                                                                       \
        "mov.b64 da, {%1, %3};\n\t"                                                                       \
        "mov.b64 db, {%2, %3};\n\t"                                                                       \
        "tcgen05.mma.cta_group::1.kind::i8" Q " [%0], da, db, %4, p;\n\t"                                 \
        "}" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate) : "memory")
#ifdef OZ_NO_COLLECTOR
    OZ_MMA_ASM("");
#else
    if constexpr (COLL == 1) OZ_MMA_ASM(".collector::a::fill");
    else if constexpr (COLL == 2) OZ_MMA_ASM(".collector::a::use");
    else if constexpr (COLL == 3) OZ_MMA_ASM(".collector::a::lastuse");
    else OZ_MMA_ASM("");
#endif
#undef OZ_MMA_ASM
}
// K-major operand tile in the canonical swizzled layout that TMA writes: rows of BK bytes, 8-row groups of 8 BK bytes
// (stride byte offset), descriptor version 1 (Blackwell), layout type 6 = SWIZZLE_32B / 4 = SWIZZLE_64B
constexpr uint32_t DESC_HI = (uint32_t)((8 * BK) >> 4) | (1u << 14) | ((BK == 64 ? 4u : 6u) << 29);
constexpr uint32_t DESC_LO_FLAGS = 1u << 16;
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(pred));
    return pred != 0;
}
// instruction descriptor: dense, S32 accumulate, signed int8 A and B, both K-major, M = 128
__host__ __device__ constexpr uint32_t idesc_i8(int bn, bool a_signed = true, bool b_signed = true) {
    return (2u << 4) | ((a_signed ? 1u : 0u) << 7) | ((b_signed ? 1u : 0u) << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, int *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}

// sum_d acc_d 2^(-RB d) for column j: two exact integer Horner groups (d < G | d >= G) joined by one rounding in the fma.
// Radix 128: G = 4; SMALLK (contraction <= 448): neighbours are first merged in int32 (|acc_d| <= (d + 1) K 2^12).
// Radix 256: G = 3 (|acc_d| < 2^31 -> every group stays below 2^53).
template <int S, bool SMALLK>
__device__ __forceinline__ double oz_horner(const int (&acc)[S][CW], int j, double rs_hi, double rs_lo) {
    constexpr int GMAX = RB == 8 ? 3 : 4;
    constexpr int G = S < GMAX ? S : GMAX, L = S - G;
    long long hi, lo = 0;
    if constexpr (RB == 7 && SMALLK) {
        const int c01 = acc[0][j] * 128 + acc[1][j];
        if constexpr (G == 3) hi = (long long)c01 * 128 + acc[2][j];
        else hi = (long long)c01 * 16384 + (acc[2][j] * 128 + acc[3][j]);
        if constexpr (L == 1) lo = acc[G][j];
        if constexpr (L >= 2) {
            const int c45 = acc[G][j] * 128 + acc[G + 1][j];
            if constexpr (L == 2) lo = c45;
            if constexpr (L == 3) lo = (long long)c45 * 128 + acc[G + 2][j];
            if constexpr (L == 4) lo = (long long)c45 * 16384 + (acc[G + 2][j] * 128 + acc[G + 3][j]);
        }
    } else {
        hi = acc[0][j];
#pragma unroll
        for (int d = 1; d < G; d++) hi = hi * (1 << RB) + acc[d][j];
        if constexpr (L > 0) {
            lo = acc[G][j];
#pragma unroll
            for (int d = G + 1; d < S; d++) lo = lo * (1 << RB) + acc[d][j];
        }
    }
    double h = (double)hi * rs_hi;              // rs_hi = 2^(row exponent - RB (G - 1)), rs_lo = 2^(row exponent - RB (S - 1))
    if constexpr (L > 0) h = fma((double)lo, rs_lo, h);
    return h;
}

struct GemmArgs {
    long long M;                // rows of A (output rows)
    int N;                      // rows of B (output columns)
    int nkb;                    // number of BK-wide k blocks over the (padded) contraction
    int kb_per_split;           // split-K: k blocks per split (== nkb when splits == 1)
    int mt, nt, splits;         // tile counts
    int stages, tmem_cols;
    const int32_t *ea, *eb;     // row exponents of A / B
    const double *bias;         // [N] or null
    int relu;
    const double *mask;         // [M][ldm] or null
    long long ldm;
    double *C;                  // [M][ldc] final output, or split-K partials [splits][M][ldc] (raw, unscaled)
    long long ldc;
    int partial;                // 1: write raw partial sums (scales applied by the caller's reduce kernel)
    int smallk;                 // contraction per split <= 448: int32 pre-merge of neighbouring accumulators is exact
    int mask_vec;               // mask rows can be read with 16-byte loads
    int tma_store;              // output through shared memory + TMA (needs even ldc and a 16-byte aligned base)
    unsigned long long *relu_bits;      // optional out [M][2 nt]: bit c of word 2 n_blk + half = (C[m][n_blk BN + half BN/2 + c] > 0)
    const unsigned long long *mask_bits;  // optional in, same layout: replaces the float64 mask (one 8-byte load per thread and tile)
    uint32_t *rowmax;           // optional [M]: high word of max_n |C[m][n]| (atomicMax, zeroed by the caller)
    unsigned long long *colmax; // optional [N]: bit pattern (high word << 32) of max_m |C[m][n]| (atomicMax, zeroed by the caller)
};

template <int S, int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
oz_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const GemmArgs g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr uint32_t A_SLICE = BM * BK, B_SLICE = BN * BK;
    constexpr uint32_t STAGE = S * (A_SLICE + B_SLICE);
    constexpr uint32_t IDESC = idesc_i8(BN);
    __shared__ __align__(8) uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], tfull_bar, tempty_bar;
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stages = g.stages;
    const long long ntiles = (long long)g.mt * g.nt * g.splits;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
        for (int s = 0; s < stages; s++) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
        mbar_init(smem_u32(&tfull_bar), 1);
        mbar_init(smem_u32(&tempty_bar), EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"((uint32_t)g.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    if (warp == 0) {
        // ===== TMA producer: one 3-D box per operand and stage (all S slices at once); the ring runs ahead across tiles
        if (lane == 0) {
            uint32_t it = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int n_blk = (int)(tile % g.nt);
                const long long rest = tile / g.nt;
                const int m_blk = (int)(rest % g.mt);
                const int z = (int)(rest / g.mt);
                const int kb0 = z * g.kb_per_split;
                const int kb1 = kb0 + g.kb_per_split < g.nkb ? kb0 + g.kb_per_split : g.nkb;
                for (int kb = kb0; kb < kb1; kb++, it++) {
                    const uint32_t s = it % stages;
                    if (it >= (uint32_t)stages) mbar_wait(smem_u32(&empty_bar[s]), ((it / stages) - 1) & 1);
                    const uint32_t fb = smem_u32(&full_bar[s]);
                    mbar_expect_tx(fb, STAGE);
                    const uint32_t sa = smem_u32(smem + (size_t)s * STAGE);
                    tma_load_3d(sa, &tmA, fb, kb * BK, m_blk * BM, 0);
                    tma_load_3d(sa + S * A_SLICE, &tmB, fb, kb * BK, n_blk * BN, 0);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: slice pair (t, u) accumulates into TMEM accumulator d = t + u (columns d * BN ...).  The whole
        // warp walks the (uniform) tile / stage loop so that descriptors live in uniform registers; one elected lane
        // issues the tcgen05 instructions.
        uint32_t it = 0, tc = 0;
        const uint32_t smem_base4 = __shfl_sync(0xffffffffu, smem_u32(smem) >> 4, 0);
#ifdef OZ_PROFILE
        long long t_te = 0, t_full = 0, t_tot = clock64(), t0;
#define OZ_T0 t0 = clock64()
#define OZ_ACC(v) v += clock64() - t0
#else
#define OZ_T0
#define OZ_ACC(v)
#endif
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, tc++) {
            const int z = (int)(tile / ((long long)g.nt * g.mt));
            const int kb0 = z * g.kb_per_split;
            const int kb1 = kb0 + g.kb_per_split < g.nkb ? kb0 + g.kb_per_split : g.nkb;
            if (tc > 0) {                                         // epilogue has drained the previous tile
                OZ_T0;
                mbar_wait(smem_u32(&tempty_bar), (tc - 1) & 1);
                OZ_ACC(t_te);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            for (int kb = kb0; kb < kb1; kb++, it++) {
                const uint32_t s = it % stages;
                OZ_T0;
                mbar_wait(smem_u32(&full_bar[s]), (it / stages) & 1);
                OZ_ACC(t_full);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa4 = (smem_base4 + s * (STAGE >> 4)) | DESC_LO_FLAGS, sb4 = sa4 + ((S * A_SLICE) >> 4);
                const uint32_t first = kb > kb0 ? 1u : 0u;
                if (elect_one()) {
#pragma unroll
                    for (int k2 = 0; k2 < BK / UMMA_K; k2++) {
#pragma unroll
                        for (int t = 0; t < S; t++) {
#if OZ_WIDE_N
                            // A slice t pairs with B slices u = 0 .. S-t-1 into accumulators t+u = CONSECUTIVE column blocks
                            // of Tensor Memory, and the B slices of a stage are consecutive 8-row groups of shared memory:
                            // up to 256 / BN of them go into ONE instruction of N = g BN columns (same products, same exact
                            // integer sums).  21 -> 9 instructions per K step at S = 6, BN = 80.  Bit-exact, but NOT faster on
                            // B200 (1.998 vs 1.94 ms for the [1.2M, 243] x [300, 243]^T product): the time goes with the
                            // accumulator columns written, not with the instruction count.
                            constexpr int G = RB == 8 ? 1 : 256 / BN;
#pragma unroll
                            for (int u = 0; u < S - t; u += G) {
                                const int g = S - t - u < G ? S - t - u : G;
                                const uint32_t td = tmem + (uint32_t)((t + u) * BN), al = sa4 + ((t * A_SLICE + k2 * UMMA_K) >> 4),
                                               bl = sb4 + ((u * B_SLICE + k2 * UMMA_K) >> 4), acc_flag = (t > 0 || k2 > 0) ? 1u : first;
                                const uint32_t id = idesc_i8(g * BN, RB != 8 || t == 0, RB != 8 || u == 0);
                                const bool first_u = u == 0, last_u = u + G >= S - t;
                                if (first_u && last_u) umma_i8<0>(td, al, bl, DESC_HI, id, acc_flag);
                                else if (first_u) umma_i8<1>(td, al, bl, DESC_HI, id, acc_flag);
                                else if (last_u) umma_i8<3>(td, al, bl, DESC_HI, id, acc_flag);
                                else umma_i8<2>(td, al, bl, DESC_HI, id, acc_flag);
                            }
#else
#pragma unroll
                            for (int u = 0; u < S - t; u++) {
                                const uint32_t td = tmem + (uint32_t)((t + u) * BN), al = sa4 + ((t * A_SLICE + k2 * UMMA_K) >> 4),
                                               bl = sb4 + ((u * B_SLICE + k2 * UMMA_K) >> 4), acc_flag = (t > 0 || k2 > 0) ? 1u : first;
                                const uint32_t id = idesc_i8(BN, RB != 8 || t == 0, RB != 8 || u == 0);     // radix 256: only the top slices are signed
                                const int n_u = S - t;
                                if (n_u == 1) umma_i8<0>(td, al, bl, DESC_HI, id, acc_flag);
                                else if (u == 0) umma_i8<1>(td, al, bl, DESC_HI, id, acc_flag);
                                else if (u == n_u - 1) umma_i8<3>(td, al, bl, DESC_HI, id, acc_flag);
                                else umma_i8<2>(td, al, bl, DESC_HI, id, acc_flag);
                            }
#endif
                        }
                    }
                    umma_commit(smem_u32(&empty_bar[s]));            // frees the smem stage when these MMAs retire
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(smem_u32(&tfull_bar));       // accumulators of this tile complete
            __syncwarp();
        }
#ifdef OZ_PROFILE
        if (blockIdx.x == 0 && lane == 0) printf("mma warp: total %lld  wait tempty %lld  wait full %lld  tiles %u\n", clock64() - t_tot, t_te, t_full, tc);
#endif
    } else {
        // ===== epilogue: warp w reads TMEM lanes 32 (w % 4) ..; one output row per thread, half of the tile's columns.
        // Drain: accumulators -> exact integer Horner -> one double per element in registers, then TMEM goes back to the
        // MMA warp.  Post: scale / bias / relu / mask, 32 x 8 blocks through a 64-byte-swizzled shared buffer -> TMA store
        // (full sectors, clipped at the matrix edge by the tensor map); overlaps the next tile's MMAs.
        const int q = warp & 3, half = (warp - 2) >> 2;
        constexpr int HC = BN / EPI_PARTS;                                     // columns per warp
        uint8_t *obuf = smem + (size_t)stages * STAGE + (size_t)(warp - 2) * (OUT_BUFS * OUT_BUF_BYTES);
        double *colc = reinterpret_cast<double *>(smem + (size_t)stages * STAGE + EPI_WARPS * (OUT_BUFS * OUT_BUF_BYTES)) + (warp - 2) * 2 * HC;
        uint32_t tc = 0, nstore = 0;
#ifdef OZ_PROFILE
        long long e_wait = 0, e_drain = 0, e_post = 0, e_tot = clock64(), e0, e1, e_wg = 0, e_st = 0, e_fence = 0, e_issue = 0, e_math = 0;
#define OZ_E1 e1 = clock64()
#define OZ_EACC(v) v += clock64() - e1
#else
#define OZ_E1
#define OZ_EACC(v)
#endif
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, tc++) {
            const int n_blk = (int)(tile % g.nt);
            const long long rest = tile / g.nt;
            const int m_blk = (int)(rest % g.mt);
            const int z = (int)(rest / g.mt);
            const long long row = (long long)m_blk * BM + q * 32 + lane;
            const bool row_ok = row < g.M;
            const int col0 = n_blk * BN + half * HC;
            // per-row scale folded into the Horner constants; per-column scale 2^eb and bias staged once per tile
            int ea = (!g.partial && row_ok) ? g.ea[row] + EOFF : 0;
            ea = ea < -960 ? -960 : (ea > 960 ? 960 : ea);
            constexpr int GH = RB == 8 ? (S < 3 ? S : 3) : (S < 4 ? S : 4);
            const double rs_hi = pow2(ea - RB * (GH - 1)), rs_lo = pow2(ea - RB * (S - 1));
            double *crow = g.C + ((size_t)z * (g.partial ? g.M : 0) + (row_ok ? row : 0)) * g.ldc;
            const double *mrow = (g.mask && row_ok) ? g.mask + row * g.ldm : nullptr;
#if OZ_COLC_SHFL
            // column constants of this warp's columns kept in lane c's registers (2^eb as its exponent, the bias) and broadcast
            // by shuffles in the post phase instead of staged in shared memory
            static_assert(!OZ_COLC_SHFL || BN / EPI_PARTS <= 32, "one column per lane");
            int c_eb = 0;
            double c_bias = 0.0;
            if (!g.partial && lane < HC) {
                const int col = col0 + lane;
                int eb = col < g.N ? g.eb[col] : 0;
                c_eb = eb < -1022 ? -1022 : (eb > 1023 ? 1023 : eb);
                if (g.bias && col < g.N) c_bias = g.bias[col];
            }
#else
            if (!g.partial) {
                __syncwarp();
                for (int c = lane; c < HC; c += 32) {
                    const int col = col0 + c;
                    int eb = col < g.N ? g.eb[col] : 0;
                    eb = eb < -1022 ? -1022 : (eb > 1023 ? 1023 : eb);
                    colc[2 * c] = pow2(eb);
                    colc[2 * c + 1] = (g.bias && col < g.N) ? g.bias[col] : 0.0;
                }
                __syncwarp();
            }
#endif
            uint32_t mb_lo = ~0u, mb_hi = ~0u, rb_lo = 0u, rb_hi = 0u;       // columns 0-31 | 32.. of this thread's half tile
            if (g.mask_bits && row_ok) {
                const unsigned long long mb = __ldcs(g.mask_bits + row * (EPI_PARTS * g.nt) + EPI_PARTS * n_blk + half);
                mb_lo = (uint32_t)mb; mb_hi = (uint32_t)(mb >> 32);
            }
#ifdef OZ_PROFILE
            e0 = clock64();
#endif
            mbar_wait(smem_u32(&tfull_bar), tc & 1);
#ifdef OZ_PROFILE
            e_wait += clock64() - e0; e0 = clock64();
#endif
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            double hv[HC];
            uint32_t hrow = 0u;                                        // running |.| maximum of this thread's row (high words)
            const uint32_t hmask = row_ok ? 0x7fffffffu : 0u;         // rows outside the matrix hold the bias only
#pragma unroll
            for (int c0 = 0; c0 < HC; c0 += CW) {
                int acc[S][CW];
#pragma unroll
                for (int d = 0; d < S; d++) {
                    const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(d * BN + half * HC + c0);
                    if constexpr (CW == 8) tmem_ld8(ta, acc[d]); else tmem_ld4(ta, acc[d]);
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c0 + CW >= HC) {                                    // last chunk is in registers: release TMEM
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&tempty_bar));
                }
                if (g.smallk) {
#pragma unroll
                    for (int jj = 0; jj < CW; jj++) hv[c0 + jj] = oz_horner<S, true>(acc, jj, rs_hi, rs_lo);
                } else {
#pragma unroll
                    for (int jj = 0; jj < CW; jj++) hv[c0 + jj] = oz_horner<S, false>(acc, jj, rs_hi, rs_lo);
                }
            }
#ifdef OZ_PROFILE
            e_drain += clock64() - e0; e0 = clock64();
#endif
#pragma unroll
            for (int c0 = 0; c0 < HC; c0 += CW) {
                const int colb = col0 + c0;
                double v[CW];
                OZ_E1;
#pragma unroll
                for (int jj = 0; jj < CW; jj++) v[jj] = hv[c0 + jj];
                if (!g.partial) {
#if OZ_COLC_SHFL
#pragma unroll
                    for (int jj = 0; jj < CW; jj++) {
                        const double sc = pow2(__shfl_sync(0xffffffffu, c_eb, c0 + jj));
                        const double bi = g.bias ? __shfl_sync(0xffffffffu, c_bias, c0 + jj) : 0.0;
                        v[jj] = fma(v[jj], sc, bi);
                    }
#else
#pragma unroll
                    for (int jj = 0; jj < CW; jj += 2) {
                        const double4 cb = *reinterpret_cast<const double4 *>(colc + 2 * (c0 + jj));    // broadcast LDS.128 x2
                        v[jj] = fma(v[jj], cb.x, cb.y);
                        v[jj + 1] = fma(v[jj + 1], cb.z, cb.w);
                    }
#endif
                    if (g.relu) {
#pragma unroll
                        for (int jj = 0; jj < CW; jj++) {               // max(v, 0) on the integer pipe
                            long long b = __double_as_longlong(v[jj]);
                            b &= ~(b >> 63);
                            v[jj] = __longlong_as_double(b);
                        }
                    }
                    if (g.relu_bits) {                                 // after the relu v >= +0: (v > 0) <=> any bit set
#pragma unroll
                        for (int jj = 0; jj < CW; jj++) {
                            const uint32_t nz = ((uint32_t)__double2hiint(v[jj]) | (uint32_t)__double2loint(v[jj])) != 0u ? 1u : 0u;
                            if (c0 + jj < 32) rb_lo |= nz << ((c0 + jj) & 31);
                            else rb_hi |= nz << ((c0 + jj) & 31);
                        }
                    }
                    if (g.mask_bits) {
#pragma unroll
                        for (int jj = 0; jj < CW; jj++)
                            if (!(((c0 + jj < 32 ? mb_lo : mb_hi) >> ((c0 + jj) & 31)) & 1u)) v[jj] = 0.0;
                    } else if (mrow) {
                        if (g.mask_vec && colb + CW <= g.N) {
#pragma unroll
                            for (int jj = 0; jj < CW; jj += 2) {
                                const longlong2 mk = *reinterpret_cast<const longlong2 *>(mrow + colb + jj);
                                if (!(mk.x > 0)) v[jj] = 0.0;           // bit pattern > 0  <=>  value > +0
                                if (!(mk.y > 0)) v[jj + 1] = 0.0;
                            }
                        } else {
#pragma unroll
                            for (int jj = 0; jj < CW; jj++)
                                if (colb + jj < g.N && !(mrow[colb + jj] > 0.0)) v[jj] = 0.0;
                        }
                    }
                }
                if (g.rowmax || g.colmax) {
                    // abs-max of the FINAL values for the slicers of this output (high words order like the doubles): the
                    // consumer reads C once instead of once per orientation plus a reduction pass
                    uint32_t hcol = 0u;
#pragma unroll
                    for (int jj = 0; jj < CW; jj++) {
                        const uint32_t h = (uint32_t)__double2hiint(v[jj]) & hmask;
                        hrow = max(hrow, h);
                        if (g.colmax) {
                            const uint32_t cm = __reduce_max_sync(0xffffffffu, h);
                            if (lane == jj) hcol = cm;
                        }
                    }
                    // (combining these in shared memory first was tried: slower, 255 -> 303 us - shared-memory atomics queue behind the
                    // tensor cores' operand reads; the global reductions are fire-and-forget)
                    if (g.colmax && lane < CW && colb + lane < g.N && hcol > 0u)
                        atomicMax(g.colmax + colb + lane, (unsigned long long)hcol << 32);
                }
                OZ_EACC(e_math);
                if (g.tma_store) {
                    uint8_t *buf = obuf + (nstore % OUT_BUFS) * OUT_BUF_BYTES;
                    OZ_E1;
                    if (nstore >= OUT_BUFS) {                           // the store that last read this buffer has drained it
                        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(OUT_BUFS - 1) : "memory");
                        __syncwarp();
                    }
                    OZ_EACC(e_wg); OZ_E1;
                    // row = lane, CW * 8 bytes per row, 16-byte chunks swizzled as the output tensor map expects (64 B / 32 B swizzle)
                    const uint32_t rb = smem_u32(buf) + lane * (CW * 8), sw = CW == 8 ? ((lane >> 1) & 3) : ((lane >> 2) & 1);
#pragma unroll
                    for (int c = 0; c < CW / 2; c++)
                        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(rb + ((c ^ sw) << 4)), "d"(v[2 * c]), "d"(v[2 * c + 1]) : "memory");
                    OZ_EACC(e_st); OZ_E1;
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    OZ_EACC(e_fence); OZ_E1;
                    if (lane == 0) {
                        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                                     ::"l"((uint64_t)&tmC), "r"(smem_u32(buf)), "r"(colb), "r"(m_blk * BM + q * 32), "r"(z) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    OZ_EACC(e_issue);
                    nstore++;
                } else if (row_ok) {
#pragma unroll
                    for (int jj = 0; jj < CW; jj++)
                        if (colb + jj < g.N) crow[colb + jj] = v[jj];
                }
            }
            if (g.rowmax && row_ok && hrow > 0u) atomicMax(g.rowmax + row, hrow);
            if (g.relu_bits && row_ok) g.relu_bits[row * (EPI_PARTS * g.nt) + EPI_PARTS * n_blk + half] = ((unsigned long long)rb_hi << 32) | rb_lo;
#ifdef OZ_PROFILE
            e_post += clock64() - e0;
#endif
        }
#ifdef OZ_PROFILE
        if (blockIdx.x == 0 && warp == 2 && lane == 0) printf("epi warp: total %lld  wait tfull %lld  drain %lld  post %lld (math %lld  wait_group %lld  st.shared %lld  fence %lld  tma issue %lld)\n", clock64() - e_tot, e_wait, e_drain, e_post, e_math, e_wg, e_st, e_fence, e_issue);
#endif
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)g.tmem_cols) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// int8 slices [S][rows][Kp] -> 3-D tensor map, box {32 B, box_rows, S}, 32-byte swizzle, zero fill out of bounds
static int make_map(CUtensorMap *tm, const int8_t *base, long long rows, long long Kp, int S, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return EGP_ECUDA; }
    cuuint64_t dims[3] = {(cuuint64_t)Kp, (cuuint64_t)rows, (cuuint64_t)S};
    cuuint64_t strides[2] = {(cuuint64_t)Kp, (cuuint64_t)Kp * (cuuint64_t)rows};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, (cuuint32_t)S};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    BK == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with %d (rows %lld Kp %lld S %d)", (int)r, rows, Kp, S); return EGP_ECUDA; }
    return EGP_OK;
}

// float64 output [splits][rows][ld] -> 3-D tensor map, box {8 columns, 32 rows, 1}, 64-byte swizzle (the epilogue's staging layout)
static int make_map_out(CUtensorMap *tm, const double *base, long long rows, int cols, long long ld, int splits) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return EGP_ECUDA; }
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)splits};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 8, (cuuint64_t)ld * 8 * (cuuint64_t)rows};
    cuuint32_t box[3] = {(cuuint32_t)CW, 32, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CW == 8 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (output) failed with %d (rows %lld cols %d ld %lld)", (int)r, rows, cols, ld); return EGP_ECUDA; }
    return EGP_OK;
}

constexpr int OUT_STAGE_BYTES = EPI_WARPS * OUT_BUFS * OUT_BUF_BYTES + EPI_WARPS * 2 * (80 / EPI_PARTS) * 8;   // per epilogue warp: two 32 x 64 B staging buffers + (2^eb, bias) per column

template <int S, int BN>
static int launch_gemm(const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmC, GemmArgs g, cudaStream_t st) {
    static_assert(S * BN <= 512, "accumulators exceed Tensor Memory");
    const size_t stage = (size_t)S * (BM * BK + BN * BK);
    int stages = (int)((SMEM_LIMIT - 1024 - OUT_STAGE_BYTES) / stage);
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 2) { set_error("egp_oz_gemm_f64: S = %d does not fit shared memory", S); return EGP_ESIZE; }
    g.stages = stages;
    int cols = S * BN, p2 = 32;
    while (p2 < cols) p2 <<= 1;
    g.tmem_cols = p2;
    const size_t smem = stage * stages + OUT_STAGE_BYTES + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        EGP_CUDA(cudaFuncSetAttribute(oz_gemm_kernel<S, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        attr_set = true;
    }
    const long long ntiles = (long long)g.mt * g.nt * g.splits;
    const int grid = (int)(ntiles < num_sms() ? ntiles : num_sms());
    oz_gemm_kernel<S, BN><<<grid, GEMM_THREADS, smem, st>>>(tmA, tmB, tmC, g);
    EGP_CHECK_LAUNCH("oz_gemm_kernel");
    return EGP_OK;
}

static inline int pick_bn(int n, int S) {
    // N tile: 80 covers 300/301-wide layers in 4 tiles; 64 otherwise and whenever S * 80 exceeds the 512 TMEM columns
    if (S > 6) return 64;
    static const char *force = getenv("EGP_OZ_BN");
    if (force && atoi(force) == 64) return 64;
    if (force && atoi(force) == 80) return 80;
    const int t64 = (n + 63) / 64, t80 = (n + 79) / 80;
    return (t80 * 80 <= t64 * 64 || t80 < t64) ? 80 : 64;
}

int gemm_ntiles(int n, int S) {
    const int bn = pick_bn(n, S);
    return (n + bn - 1) / bn;
}
int gemm_bits_words(int n, int S) { return EPI_PARTS * gemm_ntiles(n, S); }

long long gemm_tiles(long long m, int n, int S) {
    const int bn = pick_bn(n, S);
    return ((m + BM - 1) / BM) * ((n + bn - 1) / bn);
}

int choose_splits(long long m, int n, long long kp, int S) {
    const int bn = pick_bn(n, S);
    const long long tiles = ((m + BM - 1) / BM) * ((n + bn - 1) / bn);
    const long long nkb = (kp + BK - 1) / BK;
    const long long max_kb = max_kblocks(S);
    long long splits = (nkb + max_kb - 1) / max_kb;
    if (tiles < num_sms() && nkb >= 64) {                     // few output tiles, long contraction: fill the SMs
        long long want = num_sms() / tiles;
        if (want > nkb / 8) want = nkb / 8;
        if (want > splits) splits = want;
    }
    if (splits < 1) splits = 1;
    return (int)splits;
}

long long gemm_work_bytes(long long m, int n, long long kp, int force_splits) {
    return (long long)force_splits * m * (long long)((n + 1) & ~1) * 8;
}

int gemm(const int8_t *a, const int32_t *ea, long long m, const int8_t *b, const int32_t *eb, int n, long long kp, int S,
         GemmOut &o, cudaStream_t st) {
    if (!a || !b || m < 1 || n < 1 || kp < 16 || (kp & 15) || S < 3 || S > MAX_S) {
        set_error("egp_oz_gemm_f64: bad argument (m %lld n %d kp %lld S %d)", m, n, kp, S);
        return EGP_EINVAL;
    }
    const int bn = pick_bn(n, S);
    CUtensorMap tmA, tmB;
    int rc = make_map(&tmA, a, m, kp, S, BM);
    if (rc) return rc;
    rc = make_map(&tmB, b, n, kp, S, bn);
    if (rc) return rc;
    GemmArgs g;
    memset(&g, 0, sizeof g);
    g.M = m; g.N = n; g.nkb = (int)((kp + BK - 1) / BK);
    g.ea = ea; g.eb = eb; g.bias = o.bias; g.relu = o.relu; g.mask = o.mask; g.ldm = o.ldm;
    g.mt = (int)((m + BM - 1) / BM);
    g.nt = (n + bn - 1) / bn;
    int splits = o.force_splits > 0 ? o.force_splits : 1;
    const long long max_kb = max_kblocks(S);
    if (o.force_splits <= 0 && g.nkb > max_kb) {
        set_error("egp_oz_gemm_f64: contraction of %lld needs split-K (int32 accumulators)", kp);
        return EGP_ESIZE;
    }
    g.kb_per_split = (g.nkb + splits - 1) / splits;
    splits = (g.nkb + g.kb_per_split - 1) / g.kb_per_split;            // no empty split
    if (g.kb_per_split > max_kb) { set_error("egp_oz_gemm_f64: %d k blocks per split overflow int32 accumulators", g.kb_per_split); return EGP_ESIZE; }
    g.splits = splits;
    if (o.force_splits > 0) {
        const long long ldp = (n + 1) & ~1;
        if (!o.work || o.work_bytes < splits * m * ldp * 8) { set_error("egp_oz_gemm_f64: split-K workspace too small"); return EGP_EINVAL; }
        if (o.bias || o.relu || o.mask) { set_error("egp_oz_gemm_f64: bias / relu / mask are not supported on the split-K path"); return EGP_EINVAL; }
        g.partial = 1; g.C = o.work; g.ldc = ldp;
        o.ldp = ldp;
    } else {
        if (!o.C || !ea || !eb || o.ldc < n) { set_error("egp_oz_gemm_f64: bad output argument"); return EGP_EINVAL; }
        g.partial = 0; g.C = o.C; g.ldc = o.ldc;
        g.rowmax = o.rowmax; g.colmax = o.colmax;
        if (o.relu_bits && !o.relu) { set_error("egp_oz_gemm_f64: relu_bits needs relu"); return EGP_EINVAL; }
        g.relu_bits = o.relu_bits; g.mask_bits = o.mask_bits;
        if (g.mask_bits) g.mask = nullptr;
    }
    o.splits_used = splits;
    g.smallk = RB == 7 && (long long)g.kb_per_split * BK <= 448;
    g.mask_vec = g.mask && ((g.ldm & 1) == 0) && ((reinterpret_cast<uintptr_t>(g.mask) & 15) == 0);
    g.tma_store = ((g.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0);
    CUtensorMap tmC;
    memset(&tmC, 0, sizeof tmC);
    if (g.tma_store) {
        rc = make_map_out(&tmC, g.C, m, n, g.ldc, g.partial ? splits : 1);
        if (rc) return rc;
    }
#define OZ_LAUNCH(S_, BN_) rc = launch_gemm<S_, BN_>(tmA, tmB, tmC, g, st)
    if (bn == 80) {
        switch (S) {
            case 3: OZ_LAUNCH(3, 80); break;
            case 4: OZ_LAUNCH(4, 80); break;
            case 5: OZ_LAUNCH(5, 80); break;
            default: OZ_LAUNCH(6, 80); break;
        }
    } else {
        switch (S) {
            case 3: OZ_LAUNCH(3, 64); break;
            case 4: OZ_LAUNCH(4, 64); break;
            case 5: OZ_LAUNCH(5, 64); break;
#if OZ_RADIX_BITS == 8
            default: OZ_LAUNCH(6, 64); break;
#else
            case 6: OZ_LAUNCH(6, 64); break;
            case 7: OZ_LAUNCH(7, 64); break;
            default: OZ_LAUNCH(8, 64); break;
#endif
        }
    }
#undef OZ_LAUNCH
    return rc;
}

// C[m][n] = 2^(ea_m + eb_n - 12) * sum_z partial[z][m][n]
__global__ void __launch_bounds__(256)
oz_splitk_reduce_kernel(const double *__restrict__ part, int splits, long long M, int N, long long ldp, const int32_t *__restrict__ ea,
                        const int32_t *__restrict__ eb, double *__restrict__ C, long long ldc) {
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= M * N) return;
    const long long m = idx / N;
    const int n = (int)(idx % N);
    double s = 0.0;
    for (int z = 0; z < splits; z++) s += part[((size_t)z * M + m) * ldp + n];
    C[m * ldc + n] = s * ldexp(1.0, ea[m] + EOFF) * ldexp(1.0, eb[n]);
}

template <int S, int P>
static int launch_slice_rows_p(const double *x, long long m, int k, long long ldx, int8_t *out, int kp, int32_t *exps,
                               unsigned long long *colmax, cudaStream_t st) {
    const int smem = ROWS_WARPS * 2 * (P * 128) * 8 + ROWS_WARPS * (P * 128) * 4;
    static int per_sm = 0;
    if (!per_sm) {
        EGP_CUDA(cudaFuncSetAttribute(oz_slice_rows_kernel<S, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        EGP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, oz_slice_rows_kernel<S, P>, 32 * ROWS_WARPS, smem));
        if (per_sm < 1) per_sm = 1;
    }
    long long blocks = (m + ROWS_WARPS - 1) / ROWS_WARPS;
    const long long cap = (long long)num_sms() * per_sm;              // one resident wave, rows grid-strided over it
    if (blocks > cap) blocks = cap;
    oz_slice_rows_kernel<S, P><<<(unsigned)blocks, 32 * ROWS_WARPS, smem, st>>>(x, m, k, ldx, out, kp, exps, colmax);
    EGP_CHECK_LAUNCH("oz_slice_rows_kernel");
    return EGP_OK;
}

template <int S>
static int launch_slice_rows(const double *x, long long m, int k, long long ldx, int8_t *out, int kp, int32_t *exps,
                             unsigned long long *colmax, cudaStream_t st) {
    const int P = (kp + 127) / 128;
    if (P <= 1) return launch_slice_rows_p<S, 1>(x, m, k, ldx, out, kp, exps, colmax, st);
    if (P == 2) return launch_slice_rows_p<S, 2>(x, m, k, ldx, out, kp, exps, colmax, st);
    if (P == 3) return launch_slice_rows_p<S, 3>(x, m, k, ldx, out, kp, exps, colmax, st);
    return launch_slice_rows_p<S, 6>(x, m, k, ldx, out, kp, exps, colmax, st);
}

#if OZ_RADIX_BITS == 8
#define OZ_CASES_78(CALL)
#else
#define OZ_CASES_78(CALL) case 7: { constexpr int S_ = 7; CALL; } break; case 8: { constexpr int S_ = 8; CALL; } break;
#endif
#define OZ_DISPATCH_S(S, CALL)                       \
    switch (S) {                                     \
        case 3: { constexpr int S_ = 3; CALL; } break; \
        case 4: { constexpr int S_ = 4; CALL; } break; \
        case 5: { constexpr int S_ = 5; CALL; } break; \
        case 6: { constexpr int S_ = 6; CALL; } break; \
        OZ_CASES_78(CALL)                                \
        default: set_error("Ozaki slice count %d outside [3, %d]", S, MAX_S); return EGP_EINVAL; \
    }

int slice_rows(const double *x, long long m, int k, long long ldx, int S, int8_t *out, int kp, int32_t *exps,
               unsigned long long *colmax, cudaStream_t st) {
    if (!x || !out || !exps || m < 1 || k < 1 || ldx < k || kp < k || (kp & 15) || kp > 768) {
        set_error("egp_oz_slice_rows_f64: bad argument (k %d kp %d must satisfy k <= kp <= 768, kp %% 16 == 0)", k, kp);
        return EGP_EINVAL;
    }
    int rc = EGP_OK;
    OZ_DISPATCH_S(S, rc = launch_slice_rows<S_>(x, m, k, ldx, out, kp, exps, colmax, st));
    return rc;
}

int col_absmax(const double *x, long long n, int f, long long ldx, unsigned long long *colmax, cudaStream_t st) {
    if (!x || !colmax || n < 1 || f < 1 || ldx < f) { set_error("egp_oz_colmax_f64: bad argument"); return EGP_EINVAL; }
    int blocks = num_sms() * 4;
    long long rpb = (n + blocks - 1) / blocks;
    if (rpb < 32) rpb = 32;
    blocks = (int)((n + rpb - 1) / rpb);
    oz_colmax_kernel<<<blocks, 256, 0, st>>>(x, n, f, ldx, rpb, colmax);
    EGP_CHECK_LAUNCH("oz_colmax_kernel");
    return EGP_OK;
}

template <int S>
static int launch_colsT(dim3 grid, int smem, const double *x, long long n, int f, long long ldx, const int32_t *exps, int8_t *out,
                        long long np, int ones_row, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        EGP_CUDA(cudaFuncSetAttribute(oz_slice_colsT_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    oz_slice_colsT_kernel<S><<<grid, 256, smem, st>>>(x, n, f, ldx, exps, out, np, ones_row);
    return EGP_OK;
}

int slice_colsT(const double *x, long long n, int f, long long ldx, int S, const unsigned long long *colmax, int8_t *out,
                long long np, int32_t *exps, int ones_row, cudaStream_t st) {
    if (!x || !out || !exps || !colmax || n < 1 || f < 1 || ldx < f || np < n || (np & 15)) {
        set_error("egp_oz_slice_cols_t_f64: bad argument (np %lld must be >= n and a multiple of 16)", np);
        return EGP_EINVAL;
    }
    const int ft = f + (ones_row ? 1 : 0);
    oz_col_exps_kernel<<<(ft + 127) / 128, 128, 0, st>>>(colmax, f, ones_row, exps);
    const long long ntile = (np + 127) / 128;
    dim3 grid((unsigned)((ntile + COLST_TPB - 1) / COLST_TPB), (unsigned)((ft + 31) / 32));
    const int smem = 2 * 128 * 32 * 8;
    int rc = EGP_OK;
    OZ_DISPATCH_S(S, rc = launch_colsT<S_>(grid, smem, x, n, f, ldx, exps, out, np, ones_row, st));
    if (rc) return rc;
    EGP_CHECK_LAUNCH("oz_slice_colsT_kernel");
    return EGP_OK;
}


template <int S>
static int launch_both(dim3 grid, int smem, const double *x, long long n, int f, long long ldx, const int32_t *exps, int8_t *out,
                       long long np, int ones_row, const uint32_t *rowmax, int8_t *outR, int kp, int32_t *expsR, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        EGP_CUDA(cudaFuncSetAttribute(oz_slice_both_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    oz_slice_both_kernel<S><<<grid, 256, smem, st>>>(x, n, f, ldx, exps, out, np, ones_row, rowmax, outR, kp, expsR);
    return EGP_OK;
}

int slice_both(const double *x, long long n, int f, long long ldx, int S, const uint32_t *rowmax, const unsigned long long *colmax,
               int8_t *outR, int kp, int32_t *expsR, int8_t *outT, long long np, int32_t *expsT, int ones_row, cudaStream_t st) {
    if (!x || !outR || !expsR || !outT || !expsT || !rowmax || !colmax || n < 1 || f < 1 || ldx < f || np < n || (np & 15) || kp < f ||
        (kp & 31) || kp > 768) {
        set_error("egp_oz_slice_both_f64: bad argument (np %lld >= n multiple of 16, kp %d >= f multiple of 32)", np, kp);
        return EGP_EINVAL;
    }
    const int ft = f + (ones_row ? 1 : 0);
    oz_col_exps_kernel<<<(ft + 127) / 128, 128, 0, st>>>(colmax, f, ones_row, expsT);
    const long long ntile = (np + 127) / 128;
    dim3 grid((unsigned)((ft + 31) / 32), (unsigned)((ntile + COLST_TPB - 1) / COLST_TPB));
    if (grid.y > 65535u) { set_error("egp_oz_slice_both_f64: more than %d rows", 65535 * COLST_TPB * 128); return EGP_ESIZE; }
    const int smem = 2 * 128 * 32 * 8;
    int rc = EGP_OK;
    OZ_DISPATCH_S(S, rc = launch_both<S_>(grid, smem, x, n, f, ldx, expsT, outT, np, ones_row, rowmax, outR, kp, expsR, st));
    if (rc) return rc;
    EGP_CHECK_LAUNCH("oz_slice_both_kernel");
    return EGP_OK;
}

}  // namespace oz
}  // namespace egp

using namespace egp;
using namespace egp::oz;

extern "C" {

int egp_oz_radix_bits(void) { return RB; }

int egp_oz_slice_rows_f64(const double *d_x, int64_t m, int k, int64_t ldx, int n_slices, int8_t *d_out, int kp,
                          int32_t *d_exps, double *d_colmax, void *stream) {
    return slice_rows(d_x, m, k, ldx, n_slices, d_out, kp, d_exps, (unsigned long long *)d_colmax, (cudaStream_t)stream);
}

int egp_oz_colmax_f64(const double *d_x, int64_t n, int f, int64_t ldx, double *d_colmax, void *stream) {
    return col_absmax(d_x, n, f, ldx, (unsigned long long *)d_colmax, (cudaStream_t)stream);
}

int egp_oz_slice_cols_t_f64(const double *d_x, int64_t n, int f, int64_t ldx, int n_slices, const double *d_colmax,
                           int8_t *d_out, int64_t np, int32_t *d_exps, int ones_row, void *stream) {
    return slice_colsT(d_x, n, f, ldx, n_slices, (const unsigned long long *)d_colmax, d_out, np, d_exps, ones_row, (cudaStream_t)stream);
}

int64_t egp_oz_gemm_work_bytes(int64_t m, int n, int64_t kp, int n_slices) {
    const int splits = choose_splits(m, n, kp, n_slices);
    return splits > 1 ? gemm_work_bytes(m, n, kp, splits) : 0;
}

int egp_oz_gemm_f64(const int8_t *d_a, const int32_t *d_ea, int64_t m, const int8_t *d_b, const int32_t *d_eb, int n, int64_t kp,
                    int n_slices, const double *d_bias, int relu, const double *d_mask, int64_t ldm, double *d_c, int64_t ldc,
                    void *d_work, int64_t work_bytes, void *stream) {
    return egp_oz_gemm_max_f64(d_a, d_ea, m, d_b, d_eb, n, kp, n_slices, d_bias, relu, d_mask, ldm, d_c, ldc, nullptr, nullptr, d_work,
                               work_bytes, stream);
}

int egp_oz_slice_both_f64(const double *d_x, int64_t n, int f, int64_t ldx, int n_slices, const uint32_t *d_rowmax,
                          const double *d_colmax, int8_t *d_out_rows, int kp, int32_t *d_exps_rows, int8_t *d_out_t, int64_t np,
                          int32_t *d_exps_t, int ones_row, void *stream) {
    return slice_both(d_x, n, f, ldx, n_slices, d_rowmax, (const unsigned long long *)d_colmax, d_out_rows, kp, d_exps_rows, d_out_t,
                      np, d_exps_t, ones_row, (cudaStream_t)stream);
}

int egp_oz_gemm_max_f64(const int8_t *d_a, const int32_t *d_ea, int64_t m, const int8_t *d_b, const int32_t *d_eb, int n, int64_t kp,
                        int n_slices, const double *d_bias, int relu, const double *d_mask, int64_t ldm, double *d_c, int64_t ldc,
                        uint32_t *d_rowmax, double *d_colmax, void *d_work, int64_t work_bytes, void *stream) {
    if (!d_a || !d_b || !d_ea || !d_eb || !d_c || m < 1 || n < 1 || ldc < n) {
        set_error("egp_oz_gemm_f64: bad argument (m %lld n %d kp %lld ldc %lld)", (long long)m, n, (long long)kp, (long long)ldc);
        return EGP_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    GemmOut o;
    o.C = d_c; o.ldc = ldc; o.bias = d_bias; o.relu = relu; o.mask = d_mask; o.ldm = ldm;
    const int splits = choose_splits(m, n, kp, n_slices);
    if (splits > 1) {
        if (d_rowmax || d_colmax) { set_error("egp_oz_gemm_max_f64: abs-maxima are not recorded on the split-K path"); return EGP_EINVAL; }
        o.force_splits = splits; o.work = (double *)d_work; o.work_bytes = work_bytes;
    }
    o.rowmax = d_rowmax; o.colmax = (unsigned long long *)d_colmax;
    int rc = gemm(d_a, d_ea, m, d_b, d_eb, n, kp, n_slices, o, st);
    if (rc) return rc;
    if (splits > 1) {
        long long tot = m * (long long)n;
        oz_splitk_reduce_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(o.work, o.splits_used, m, n, o.ldp, d_ea, d_eb, d_c, ldc);
        EGP_CHECK_LAUNCH("oz_splitk_reduce_kernel");
    }
    return EGP_OK;
}

}  // extern "C"
