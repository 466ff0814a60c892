// Float64 dense layers on the 5th-generation tensor cores (tcgen05, kind::i8) - sm_100a only.
//
// The PPO update (agents/agent_pg.py:19-26, agents/agent_ppo.py:44-51) is dominated by the GEMMs of the two MLPs
// (models/mlp.py:22-25, core/policy_gaussian.py:19-24, core/critic.py:15-18) over the whole trajbatch in float64
// (ego_pose/ego_mimic.py:31-32).  tcgen05 has no FP64 kind, so the product is evaluated with the Ozaki scheme on the
// int8 tensor cores (B200 has them, 2x the bf16 rate):
//
//   row i of A:   a_ik = 2^ea_i * sum_{t=1..S} qa_t[i][k] 2^(1-7t),  qa_t in [-64, 64]  (int8 "slices", exact
//   row j of B:   b_jk = 2^eb_j * sum_{u=1..S} qb_u[j][k] 2^(1-7u)    residual < 2^(-7S) relative to the row maximum)
//   C_ij = sum_k a_ik b_jk = 2^(ea_i + eb_j - 12) * sum_{d=0..S-1} 2^(-7d) * [ sum_{t+u-2=d} sum_k qa_t qb_u ]
//
// Every bracket is an int8 x int8 -> int32 GEMM accumulated EXACTLY in Tensor Memory (one accumulator per d, all
// slice pairs of equal weight share it); pairs with t+u > S+1 are below the slicing residual and dropped.  The
// result differs from the float64 product by <= (S+2) 2^(-7S) |a_i|_max |b_j|_max K  (S = 6: 2e-12, S = 5: 2e-10) and
// is independent of tile shapes and launch geometry (integer accumulation), so the kernel is tested bit-exactly
// against an integer matmul of the same slices.
//
// Kernels:
//   oz_slice_rows_kernel   f64 [M][K] -> int8 [S][M][Kp] + per-row exponent (scale constant along K = columns)
//   oz_slice_colsT_kernel  f64 [N][F] -> int8 [S][F][Np] + per-column exponent, transposed (for the weight-gradient
//                          GEMMs whose contraction runs over the samples)
//   oz_gemm_kernel         warp-specialised tcgen05 GEMM: TMA (3-D boxes {64 B, rows, S slices}, 64-byte swizzle)
//                          -> mbarrier ring -> single-thread tcgen05.mma.kind::i8 into S TMEM accumulators ->
//                          4 epilogue warps tcgen05.ld, Horner over d in float64, scale, bias, relu, store
//   oz_splitk_reduce_kernel  sums the split-K partials of the weight-gradient GEMMs and applies the scales
#include <cuda.h>
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace egp {
namespace oz {

constexpr int BM = 128, BN = 64, BK = 64;       // CTA tile; BK in int8 elements = bytes (one 64 B swizzle row)
constexpr int UMMA_K = 32;                      // K per tcgen05.mma for 8-bit operands
constexpr int MAX_S = 8;
constexpr int GEMM_THREADS = 192;               // warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: epilogue
constexpr int SMEM_LIMIT = 227 * 1024;

// ---------------------------------------------------------------------------------------------------------------
// slicing
// exponent e with |x| / 2^e < 1 for |x| <= amax (amax / 2^e in [0.5, 1)); 0 for a zero row
__device__ __forceinline__ int oz_exponent(double amax) {
    if (!(amax > 0.0) || !isfinite(amax)) return 0;
    int e;
    frexp(amax, &e);
    return e;
}

// q_1..q_S of x / 2^e, packed per slice by the caller
template <int S>
__device__ __forceinline__ void oz_slices(double x, int e, int8_t *q) {
    double r = ldexp(x, 6 - e);                 // |r| <= 64
    if (!isfinite(r)) r = 0.0;
#pragma unroll
    for (int t = 0; t < S; t++) {
        double qi = rint(r);
        q[t] = (int8_t)(int)qi;
        r = (r - qi) * 128.0;
    }
}

// One warp per row (grid-stride).  Lane l owns the 8-element chunks l, l + 32, ... of the row: 64 B loads, 8 B stores
// per slice.  Optionally accumulates the column abs-max of the matrix into colmax (bit pattern max of |x| >= 0).
template <int S>
__global__ void __launch_bounds__(256)
oz_slice_rows_kernel(const double *__restrict__ x, long long M, int K, long long ldx, int8_t *__restrict__ out, int Kp,
                     int32_t *__restrict__ exps, unsigned long long *__restrict__ colmax) {
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
    const int nchunk = Kp / 8;
    constexpr int MAXC = 3;                     // K <= 768
    double cmax[MAXC][8];
#pragma unroll
    for (int c = 0; c < MAXC; c++)
#pragma unroll
        for (int j = 0; j < 8; j++) cmax[c][j] = 0.0;
    for (long long row = warp; row < M; row += nwarp) {
        double v[MAXC][8];
        double amax = 0.0;
#pragma unroll
        for (int c = 0; c < MAXC; c++) {
            const int ch = lane + 32 * c;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int k = ch * 8 + j;
                v[c][j] = (ch < nchunk && k < K) ? x[row * ldx + k] : 0.0;
                amax = fmax(amax, fabs(v[c][j]));
                cmax[c][j] = fmax(cmax[c][j], fabs(v[c][j]));
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        const int e = oz_exponent(amax);
        if (lane == 0) exps[row] = e;
#pragma unroll
        for (int c = 0; c < MAXC; c++) {
            const int ch = lane + 32 * c;
            if (ch >= nchunk) continue;
            int8_t q[8][S];
#pragma unroll
            for (int j = 0; j < 8; j++) oz_slices<S>(v[c][j], e, q[j]);
#pragma unroll
            for (int t = 0; t < S; t++) {
                unsigned long long pk = 0;
#pragma unroll
                for (int j = 0; j < 8; j++) pk |= (unsigned long long)(uint8_t)q[j][t] << (8 * j);
                *reinterpret_cast<unsigned long long *>(out + ((size_t)t * M + row) * Kp + ch * 8) = pk;
            }
        }
    }
    if (colmax) {
#pragma unroll
        for (int c = 0; c < MAXC; c++) {
            const int ch = lane + 32 * c;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int k = ch * 8 + j;
                if (ch < nchunk && k < K && cmax[c][j] > 0.0) atomicMax(colmax + k, (unsigned long long)__double_as_longlong(cmax[c][j]));
            }
        }
    }
}

// column abs-max only (when the row slices are not needed): block = 256 threads, rows strided over blocks
__global__ void __launch_bounds__(256)
oz_colmax_kernel(const double *__restrict__ x, long long N, int F, long long ldx, long long rows_per_block,
                 unsigned long long *__restrict__ colmax) {
    const long long r0 = blockIdx.x * rows_per_block;
    const long long r1 = r0 + rows_per_block < N ? r0 + rows_per_block : N;
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        double m = 0.0;
        for (long long r = r0; r < r1; r++) m = fmax(m, fabs(x[r * ldx + f]));
        if (m > 0.0) atomicMax(colmax + f, (unsigned long long)__double_as_longlong(m));
    }
}

__global__ void oz_col_exps_kernel(const unsigned long long *__restrict__ colmax, int F, int32_t *__restrict__ exps) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < F) exps[f] = oz_exponent(__longlong_as_double((long long)colmax[f]));
}

// Transposed, column-scaled slices: x [N][F] -> out [S][F][Np].  Block tile = 128 samples x 32 features: coalesced
// f64 reads along the features, shared-memory transpose, 16 B stores along the samples.
template <int S>
__global__ void __launch_bounds__(256)
oz_slice_colsT_kernel(const double *__restrict__ x, long long N, int F, long long ldx, const int32_t *__restrict__ exps,
                      int8_t *__restrict__ out, long long Np) {
    __shared__ double tile[128][33];
    const long long n0 = (long long)blockIdx.x * 128;
    const int f0 = blockIdx.y * 32;
    {
        const int fl = threadIdx.x & 31, rl = threadIdx.x >> 5;       // 8 rows per pass
#pragma unroll 4
        for (int r = rl; r < 128; r += 8) {
            const long long n = n0 + r;
            const int f = f0 + fl;
            tile[r][fl] = (n < N && f < F) ? x[n * ldx + f] : 0.0;
        }
    }
    __syncthreads();
    // thread -> (feature fl, group of 16 samples g): 32 x 8 = 256 threads
    const int g = threadIdx.x & 7, fl = threadIdx.x >> 3;
    const int f = f0 + fl;
    if (f >= F) return;
    const long long nb = n0 + g * 16;
    if (nb >= Np) return;
    const int e = exps[f];
    uint32_t pk[S][4];
#pragma unroll
    for (int t = 0; t < S; t++) pk[t][0] = pk[t][1] = pk[t][2] = pk[t][3] = 0u;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        int8_t q[S];
        oz_slices<S>(tile[g * 16 + j][fl], e, q);
#pragma unroll
        for (int t = 0; t < S; t++) pk[t][j >> 2] |= (uint32_t)(uint8_t)q[t] << (8 * (j & 3));
    }
#pragma unroll
    for (int t = 0; t < S; t++)
        *reinterpret_cast<uint4 *>(out + ((size_t)t * F + f) * Np + nb) = make_uint4(pk[t][0], pk[t][1], pk[t][2], pk[t][3]);
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
// D[tmem] (+)= A[smem] * B[smem], int8 x int8 -> int32
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
// K-major operand tile in the canonical 64-byte-swizzle layout (what TMA SWIZZLE_64B writes): rows of 64 B, 8-row
// groups of 512 B (stride byte offset), descriptor version 1 (Blackwell), layout type 4 = SWIZZLE_64B
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
// instruction descriptor: dense, S32 accumulate, signed int8 A and B, both K-major, N = 64, M = 128
constexpr uint32_t IDESC_I8 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
}

struct GemmArgs {
    long long M;                // rows of A (output rows)
    int N;                      // rows of B (output columns)
    int nkb;                    // number of 64-wide k blocks over the (padded) contraction
    int kb_per_split;           // split-K: k blocks per grid.z slice (== nkb when gridDim.z == 1)
    int S, stages, tmem_cols;
    const int32_t *ea, *eb;     // row exponents of A / B
    const double *bias;         // [N] or null
    int relu;
    double *C;                  // [M][ldc] final output, or split-K partials [gridDim.z][M][ldc] (raw, unscaled)
    long long ldc;
    int partial;                // 1: write raw partial sums (scales applied by the reduce kernel)
};

template <int S>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
oz_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr uint32_t A_SLICE = BM * BK, B_SLICE = BN * BK;            // 8192, 4096 bytes
    constexpr uint32_t STAGE = S * (A_SLICE + B_SLICE);
    __shared__ __align__(8) uint64_t full_bar[8], empty_bar[8], acc_bar;
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_blk = blockIdx.x, m_blk = blockIdx.y;
    const int kb0 = blockIdx.z * g.kb_per_split;
    const int kb1 = kb0 + g.kb_per_split < g.nkb ? kb0 + g.kb_per_split : g.nkb;
    const int stages = g.stages;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
        for (int s = 0; s < stages; s++) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
        mbar_init(smem_u32(&acc_bar), 1);
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
        // ===== TMA producer: one 3-D box per operand and stage, all S slices at once
        if (lane == 0) {
            for (int kb = kb0, it = 0; kb < kb1; kb++, it++) {
                const int s = it % stages;
                if (it >= stages) mbar_wait(smem_u32(&empty_bar[s]), ((it / stages) - 1) & 1);
                const uint32_t fb = smem_u32(&full_bar[s]);
                mbar_expect_tx(fb, STAGE);
                const uint32_t sa = smem_u32(smem + (size_t)s * STAGE);
                tma_load_3d(sa, &tmA, fb, kb * BK, m_blk * BM, 0);
                tma_load_3d(sa + S * A_SLICE, &tmB, fb, kb * BK, n_blk * BN, 0);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: slice pair (t, u) accumulates into TMEM accumulator d = t + u (columns d * BN ...)
        if (lane == 0) {
            for (int kb = kb0, it = 0; kb < kb1; kb++, it++) {
                const int s = it % stages;
                mbar_wait(smem_u32(&full_bar[s]), (it / stages) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem_u32(smem + (size_t)s * STAGE), sb = sa + S * A_SLICE;
#pragma unroll
                for (int t = 0; t < S; t++) {
#pragma unroll
                    for (int u = 0; u < S - t; u++) {
#pragma unroll
                        for (int k2 = 0; k2 < BK / UMMA_K; k2++) {
                            const uint64_t ad = umma_desc_sw64(sa + t * A_SLICE + k2 * UMMA_K);
                            const uint64_t bd = umma_desc_sw64(sb + u * B_SLICE + k2 * UMMA_K);
                            umma_i8(tmem + (uint32_t)((t + u) * BN), ad, bd, IDESC_I8, (it > 0 || t > 0 || k2 > 0) ? 1u : 0u);
                        }
                    }
                }
                umma_commit(smem_u32(&empty_bar[s]));            // frees the smem stage when these MMAs retire
            }
            umma_commit(smem_u32(&acc_bar));                      // accumulators complete
        }
    } else {
        // ===== epilogue: warp w reads TMEM lanes 32 (w % 4) ..; one output row per thread
        const int q = warp & 3;
        const long long row = (long long)m_blk * BM + q * 32 + lane;
        mbar_wait(smem_u32(&acc_bar), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const bool row_ok = row < g.M;
        double sa_row = 1.0;
        if (!g.partial && row_ok) sa_row = ldexp(1.0, g.ea[row] - 12);
        double *crow = g.C + ((size_t)blockIdx.z * (g.partial ? g.M : 0) + (row_ok ? row : 0)) * g.ldc;
        const bool empty = kb1 <= kb0;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
            int acc[S][16];
#pragma unroll
            for (int d = 0; d < S; d++) tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(d * BN + c0), acc[d]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                double v[2];
#pragma unroll
                for (int jj = 0; jj < 2; jj++) {
                    double h = (double)acc[S - 1][j + jj];
#pragma unroll
                    for (int d = S - 2; d >= 0; d--) h = h * 0.0078125 + (double)acc[d][j + jj];
                    if (empty) h = 0.0;
                    const int col = n_blk * BN + c0 + j + jj;
                    if (!g.partial && col < g.N) {
                        h = h * sa_row * ldexp(1.0, g.eb[col]);
                        if (g.bias) h += g.bias[col];
                        if (g.relu) h = fmax(h, 0.0);
                    }
                    v[jj] = h;
                }
                const int col = n_blk * BN + c0 + j;
                if (row_ok) {
                    if (col + 1 < g.N && ((g.ldc & 1) == 0)) *reinterpret_cast<double2 *>(crow + col) = make_double2(v[0], v[1]);
                    else {
                        if (col < g.N) crow[col] = v[0];
                        if (col + 1 < g.N) crow[col + 1] = v[1];
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    __syncwarp();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)g.tmem_cols) : "memory");
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
    C[m * ldc + n] = s * ldexp(1.0, ea[m] - 12) * ldexp(1.0, eb[n]);
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

// int8 slices [S][rows][Kp] -> 3-D tensor map, box {64 B, box_rows, S}, 64-byte swizzle, zero fill out of bounds
static int make_map(CUtensorMap *tm, const int8_t *base, long long rows, long long Kp, int S, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return EGP_ECUDA; }
    cuuint64_t dims[3] = {(cuuint64_t)Kp, (cuuint64_t)rows, (cuuint64_t)S};
    cuuint64_t strides[2] = {(cuuint64_t)Kp, (cuuint64_t)Kp * (cuuint64_t)rows};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, (cuuint32_t)S};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with %d (rows %lld Kp %lld S %d)", (int)r, rows, Kp, S); return EGP_ECUDA; }
    return EGP_OK;
}

template <int S>
static int launch_gemm(const CUtensorMap &tmA, const CUtensorMap &tmB, GemmArgs g, dim3 grid, cudaStream_t st) {
    const size_t stage = (size_t)S * (BM * BK + BN * BK);
    int stages = (int)((SMEM_LIMIT - 2048) / stage);
    if (stages > 8) stages = 8;
    if (stages > g.kb_per_split) stages = g.kb_per_split > 0 ? g.kb_per_split : 1;
    if (stages < 1) { set_error("egp_oz_gemm_f64: S = %d does not fit shared memory", S); return EGP_ESIZE; }
    g.stages = stages;
    g.S = S;
    int cols = S * BN, p2 = 32;
    while (p2 < cols) p2 <<= 1;
    g.tmem_cols = p2;
    const size_t smem = stage * stages + 1024;
    EGP_CUDA(cudaFuncSetAttribute(oz_gemm_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    oz_gemm_kernel<S><<<grid, GEMM_THREADS, smem, st>>>(tmA, tmB, g);
    EGP_CHECK_LAUNCH("oz_gemm_kernel");
    return EGP_OK;
}

}  // namespace oz
}  // namespace egp

using namespace egp;
using namespace egp::oz;

#define OZ_DISPATCH_S(S, CALL)                       \
    switch (S) {                                     \
        case 3: { constexpr int S_ = 3; CALL; } break; \
        case 4: { constexpr int S_ = 4; CALL; } break; \
        case 5: { constexpr int S_ = 5; CALL; } break; \
        case 6: { constexpr int S_ = 6; CALL; } break; \
        case 7: { constexpr int S_ = 7; CALL; } break; \
        case 8: { constexpr int S_ = 8; CALL; } break; \
        default: set_error("Ozaki slice count %d outside [3, 8]", S); return EGP_EINVAL; \
    }

extern "C" {

int egp_oz_slice_rows_f64(const double *d_x, int64_t m, int k, int64_t ldx, int n_slices, int8_t *d_out, int kp,
                          int32_t *d_exps, double *d_colmax, void *stream) {
    if (!d_x || !d_out || !d_exps || m < 1 || k < 1 || ldx < k || kp < k || (kp & 15) || kp > 768) {
        set_error("egp_oz_slice_rows_f64: bad argument (k %d kp %d must satisfy k <= kp <= 768, kp %% 16 == 0)", k, kp);
        return EGP_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    long long warps = m < (long long)num_sms() * 64 ? m : (long long)num_sms() * 64;
    int blocks = (int)((warps * 32 + 255) / 256);
    OZ_DISPATCH_S(n_slices, (oz_slice_rows_kernel<S_><<<blocks, 256, 0, st>>>(d_x, m, k, ldx, d_out, kp, d_exps, (unsigned long long *)d_colmax)));
    EGP_CHECK_LAUNCH("oz_slice_rows_kernel");
    return EGP_OK;
}

int egp_oz_colmax_f64(const double *d_x, int64_t n, int f, int64_t ldx, double *d_colmax, void *stream) {
    if (!d_x || !d_colmax || n < 1 || f < 1 || ldx < f) { set_error("egp_oz_colmax_f64: bad argument"); return EGP_EINVAL; }
    int blocks = num_sms() * 8;
    long long rpb = (n + blocks - 1) / blocks;
    if (rpb < 16) rpb = 16;
    blocks = (int)((n + rpb - 1) / rpb);
    oz_colmax_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_x, n, f, ldx, rpb, (unsigned long long *)d_colmax);
    EGP_CHECK_LAUNCH("oz_colmax_kernel");
    return EGP_OK;
}

int egp_oz_slice_cols_t_f64(const double *d_x, int64_t n, int f, int64_t ldx, int n_slices, const double *d_colmax,
                           int8_t *d_out, int64_t np, int32_t *d_exps, void *stream) {
    if (!d_x || !d_out || !d_exps || !d_colmax || n < 1 || f < 1 || ldx < f || np < n || (np & 15)) {
        set_error("egp_oz_slice_cols_t_f64: bad argument (np %lld must be >= n and a multiple of 16)", (long long)np);
        return EGP_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    oz_col_exps_kernel<<<(f + 127) / 128, 128, 0, st>>>((const unsigned long long *)d_colmax, f, d_exps);
    dim3 grid((unsigned)((np + 127) / 128), (unsigned)((f + 31) / 32));
    OZ_DISPATCH_S(n_slices, (oz_slice_colsT_kernel<S_><<<grid, 256, 0, st>>>(d_x, n, f, ldx, d_exps, d_out, np)));
    EGP_CHECK_LAUNCH("oz_slice_colsT_kernel");
    return EGP_OK;
}

int64_t egp_oz_gemm_work_bytes(int64_t m, int n, int64_t kp) {
    // split-K partials are only used when the output has few tiles and the contraction is long
    long long tiles = ((m + BM - 1) / BM) * ((n + BN - 1) / BN);
    long long nkb = (kp + BK - 1) / BK;
    if (tiles >= 2LL * 148 || nkb < 64) return 0;
    long long splits = (nkb + 511) / 512;                 // <= 32768 samples per split: int32 accumulation cannot overflow
    long long want = (4LL * 148 + tiles - 1) / tiles;
    if (want > splits) splits = want;
    if (splits > nkb) splits = nkb;
    return (int64_t)(splits * m * (long long)((n + 1) & ~1) * 8);
}

int egp_oz_gemm_f64(const int8_t *d_a, const int32_t *d_ea, int64_t m, const int8_t *d_b, const int32_t *d_eb, int n, int64_t kp,
                    int n_slices, const double *d_bias, int relu, double *d_c, int64_t ldc, void *d_work, int64_t work_bytes,
                    void *stream) {
    if (!d_a || !d_b || !d_ea || !d_eb || !d_c || m < 1 || n < 1 || kp < 16 || (kp & 15) || ldc < n) {
        set_error("egp_oz_gemm_f64: bad argument (m %lld n %d kp %lld ldc %lld)", (long long)m, n, (long long)kp, (long long)ldc);
        return EGP_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    CUtensorMap tmA, tmB;
    int rc = make_map(&tmA, d_a, m, kp, n_slices, BM);
    if (rc) return rc;
    rc = make_map(&tmB, d_b, n, kp, n_slices, BN);
    if (rc) return rc;
    GemmArgs g;
    memset(&g, 0, sizeof g);
    g.M = m; g.N = n; g.nkb = (int)((kp + BK - 1) / BK);
    g.ea = d_ea; g.eb = d_eb; g.bias = d_bias; g.relu = relu;
    const long long mt = (m + BM - 1) / BM, nt = (n + BN - 1) / BN;
    const int64_t need = egp_oz_gemm_work_bytes(m, n, kp);
    int splits = 1;
    long long ldp = (n + 1) & ~1;
    if (need > 0) {
        if (!d_work || work_bytes < need) { set_error("egp_oz_gemm_f64: split-K workspace too small (%lld < %lld bytes)", (long long)work_bytes, (long long)need); return EGP_EINVAL; }
        splits = (int)(need / (m * ldp * 8));
    } else if ((long long)g.nkb * BK * 4096LL * n_slices >= (1LL << 31)) {
        set_error("egp_oz_gemm_f64: contraction of %lld needs split-K (int32 accumulators)", (long long)kp);
        return EGP_ESIZE;
    }
    g.kb_per_split = (g.nkb + splits - 1) / splits;
    splits = (g.nkb + g.kb_per_split - 1) / g.kb_per_split;
    if (mt > 65535 || splits > 65535) {
        // grid.y limit: rows beyond 65535 * 128 = 8.4 M per call are not needed by the configurations in scope
        set_error("egp_oz_gemm_f64: too many row tiles (%lld) for one launch", mt);
        return EGP_ESIZE;
    }
    if (splits > 1) { g.partial = 1; g.C = (double *)d_work; g.ldc = ldp; g.bias = nullptr; g.relu = 0; }
    else { g.partial = 0; g.C = d_c; g.ldc = ldc; }
    dim3 grid((unsigned)nt, (unsigned)mt, (unsigned)splits);
    OZ_DISPATCH_S(n_slices, { rc = launch_gemm<S_>(tmA, tmB, g, grid, st); if (rc) return rc; });
    if (splits > 1) {
        if (d_bias || relu) { set_error("egp_oz_gemm_f64: bias / relu are not supported on the split-K path"); return EGP_EINVAL; }
        long long tot = m * (long long)n;
        oz_splitk_reduce_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>((const double *)d_work, splits, m, n, ldp, d_ea, d_eb, d_c, ldc);
        EGP_CHECK_LAUNCH("oz_splitk_reduce_kernel");
    }
    return EGP_OK;
}

}  // extern "C"
