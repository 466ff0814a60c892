// Chunked forward / loss / backward of the two-hidden-layer relu MLPs of the PPO update on the int8 tensor cores.
//
// Replaces, for one optimisation step of one net, the autograd graph behind agents/agent_pg.py:19-26 (value) and
// agents/agent_ppo.py:44-51,58-65 (policy): models/mlp.py:22-25 trunk + core/policy_gaussian.py:19-24 /
// core/critic.py:15-18 head, loss, backward to the six parameter gradients.
//
// The batch is processed in chunks of 148 x 128 rows (one 128-row tile per SM and n-block wave).  For one chunk every
// intermediate (activations in float64, their int8 slices in both orientations, split-K partials) is a few tens of MB,
// so producer -> consumer traffic between the kernels of a chunk stays in the 126 MB L2 instead of HBM; only the
// cached input slices stream from HBM.  Per chunk:
//   x slices (cached per update) -> F1 (+bias, relu) -> slice h1 -> F2 -> slice h2 -> F3 -> loss kernel -> dy ->
//   slice dy -> gW3 | D2 (x relu mask) -> slice dh2 -> gW2 | D1 -> slice dh1 -> gW1
// Weight gradients contract over the samples: transposed column-scaled slices with a row of ones appended to the
// activation side, so the bias gradient is one more output row of the same GEMM; every chunk's split-K partials are
// reduced in a fixed order and accumulated into the gradient in chunk order (deterministic).
#include <string.h>

#include "ozaki.cuh"

namespace egp {
namespace oz {

static inline long long al(long long v) { return (v + 1023) & ~1023LL; }
static inline int pad16(long long k) { return (int)((k + 15) / 16 * 16); }
// Row pitch of slice tensors whose contraction runs along the FEATURES: a multiple of the 32-byte k block, so that every
// 32-byte row segment a TMA box fetches is one aligned sector (pitch 304 for a 300-wide layer put every other row across
// two sectors: the 300-wide forward / data-gradient GEMMs ran 25 % slower than the 243 -> 256 wide first layer per k block).
static inline int padk(long long k) { return (int)((k + 31) / 32 * 32); }

// gW[o][i] (+)= 2^(ea_i + eb_o + EOFF) * sum_z part[z][i][o] for i < in; gb[o] (+)= the same for i == in (ones row)
__global__ void __launch_bounds__(256)
oz_wgrad_reduce_kernel(const double *__restrict__ part, int splits, int in, int out, long long ldp, const int32_t *__restrict__ ea,
                       const int32_t *__restrict__ eb, double *__restrict__ gW, double *__restrict__ gb, int accumulate) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;      // o fastest: coalesced reads of the partials
    if (idx >= (in + 1) * out) return;
    const int i = idx / out, o = idx % out;
    double s = 0.0;
    for (int z = 0; z < splits; z++) s += part[((size_t)z * (in + 1) + i) * ldp + o];
    s *= ldexp(1.0, ea[i] + EOFF) * ldexp(1.0, eb[o]);
    double *dst = i < in ? gW + (size_t)o * in + i : gb + o;
    *dst = accumulate ? *dst + s : s;
}

// ---- scalar head (critic.py:15-18, value_head Linear(h2, 1)) in plain float64, HBM-bound -------------------------------
// As a tensor-core GEMM the 1-wide head pads to a 64-column tile and needs the row AND the transposed slices of the
// last hidden layer; as three streaming kernels it reads that layer twice and writes its gradient once.
constexpr int HEAD_MAXJ = 16;       // hidden width <= 512 (wider heads take the GEMM path); 8 x 513 doubles of shared memory
constexpr int HEAD_WARPS = 8;

// y[n] = b + a2[n] . w            (one warp per row, fixed shuffle order: deterministic)
__global__ void __launch_bounds__(32 * HEAD_WARPS)
head1_fwd_kernel(const double *__restrict__ a2, long long m, int h2, const double *__restrict__ w, const double *__restrict__ b,
                 double *__restrict__ y) {
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * HEAD_WARPS + (threadIdx.x >> 5), nwarp = (long long)gridDim.x * HEAD_WARPS;
    double wl[HEAD_MAXJ];
#pragma unroll
    for (int j = 0; j < HEAD_MAXJ; j++) wl[j] = lane + 32 * j < h2 ? w[lane + 32 * j] : 0.0;
    const double bias = b[0];
    for (long long r = warp; r < m; r += nwarp) {
        const double *row = a2 + r * h2;
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < HEAD_MAXJ; j++)
            if (lane + 32 * j < h2) s += row[lane + 32 * j] * wl[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) y[r] = s + bias;
    }
}

// d2[n][k] = a2[n][k] > 0 ? dy[n] w[k] : 0 ; part[block][k] = sum over the block's rows of dy[n] a2[n][k], part[block][h2] = sum dy[n]
// J = ceil(h2 / 32) column groups per lane (compile-time: weights and partial sums stay in registers, 3 blocks per SM).
// Optionally records what the slicers of d2 need: rowmax[n] = high word of max_k |d2[n][k]|, colmax[k] = bit pattern of
// max_n |d2[n][k]| (atomicMax, buffer zeroed by the caller) - so d2 is read ONCE afterwards (oz_slice_both).
template <int J>
__global__ void __launch_bounds__(32 * HEAD_WARPS, J <= 10 ? 3 : 2)
head1_bwd_kernel(const double *__restrict__ a2, const double *__restrict__ dy, long long m, int h2, const double *__restrict__ w,
                 double *__restrict__ d2, double *__restrict__ part, uint32_t *__restrict__ rowmax,
                 unsigned long long *__restrict__ colmax) {
    __shared__ double red[HEAD_WARPS][32 * J + 1];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const long long per = (m + gridDim.x - 1) / gridDim.x, r0 = (long long)blockIdx.x * per, r1 = r0 + per < m ? r0 + per : m;
    double wl[J], acc[J], sdy = 0.0;
    uint32_t cm[J];
#pragma unroll
    for (int j = 0; j < J; j++) { wl[j] = lane + 32 * j < h2 ? w[lane + 32 * j] : 0.0; acc[j] = 0.0; cm[j] = 0u; }
    for (long long r = r0 + wp; r < r1; r += HEAD_WARPS) {
        const double g = dy[r];
        const double *row = a2 + r * h2;
        double *out = d2 + r * h2;
        double a[J];
#pragma unroll
        for (int j = 0; j < J; j++) a[j] = lane + 32 * j < h2 ? __ldcs(row + lane + 32 * j) : 0.0;
        sdy += g;
        uint32_t hm = 0u;
#pragma unroll
        for (int j = 0; j < J; j++)
            if (lane + 32 * j < h2) {
                const double v = a[j] > 0.0 ? g * wl[j] : 0.0;
                out[lane + 32 * j] = v;
                acc[j] += g * a[j];
                const uint32_t h = (uint32_t)__double2hiint(v) & 0x7fffffffu;
                hm = max(hm, h);
                cm[j] = max(cm[j], h);
            }
        if (rowmax) {
            hm = __reduce_max_sync(0xffffffffu, hm);
            if (lane == 0) rowmax[r] = hm;
        }
    }
#pragma unroll
    for (int j = 0; j < J; j++) red[wp][lane + 32 * j] = acc[j];
    if (lane == 0) red[wp][32 * J] = sdy;
    __syncthreads();
    for (int k = threadIdx.x; k <= h2; k += blockDim.x) {
        const int src = k < h2 ? k : 32 * J;
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < HEAD_WARPS; q++) t += red[q][src];
        part[(size_t)blockIdx.x * (h2 + 1) + k] = t;
    }
    if (colmax) {
        __syncthreads();
        uint32_t *redu = reinterpret_cast<uint32_t *>(&red[0][0]);
#pragma unroll
        for (int j = 0; j < J; j++) redu[wp * (32 * J) + lane + 32 * j] = cm[j];
        __syncthreads();
        for (int k = threadIdx.x; k < h2; k += blockDim.x) {
            uint32_t t = redu[k];
#pragma unroll
            for (int q = 1; q < HEAD_WARPS; q++) t = max(t, redu[q * (32 * J) + k]);
            if (t > 0u) atomicMax(colmax + k, (unsigned long long)t << 32);
        }
    }
}

// gW[k] (+)= sum_blocks part[block][k] in a fixed order ; gb (+)= part[.][h2].  One block per k: thread t sums blocks
// t, t + 128, ... then a fixed shared-memory tree (deterministic).
__global__ void __launch_bounds__(128)
head1_reduce_kernel(const double *__restrict__ part, int nblocks, int h2, double *__restrict__ gW, double *__restrict__ gb,
                    int accumulate) {
    __shared__ double red[128];
    const int k = blockIdx.x;
    double t = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += 128) t += part[(size_t)b * (h2 + 1) + k];
    red[threadIdx.x] = t;
    __syncthreads();
#pragma unroll
    for (int o = 64; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double *dst = k < h2 ? gW + k : gb;
        *dst = accumulate ? *dst + red[0] : red[0];
    }
}

struct Buf {
    char *base;
    long long off = 0, cap;
    template <typename T> T *take(long long n) {
        T *p = reinterpret_cast<T *>(base + off);
        off += al(n * (long long)sizeof(T));
        return p;
    }
};

struct Sl {                 // one sliced operand
    int8_t *q = nullptr;
    int32_t *e = nullptr;
};

}  // namespace oz
}  // namespace egp

using namespace egp;
using namespace egp::oz;

extern "C" {

int64_t egp_oz_mlp_chunk_rows(void) { return (int64_t)num_sms() * BM; }

/* 1 (default; EGP_OZ_FUSED_SLICE=0 in the environment starts with 0): producers record row / column abs-maxima and every
 * intermediate is sliced in both orientations from one read; 0: separate row / column-maximum / transposed passes.
 * Negative: query only.  Returns the setting in force (before the call's change when querying). */
int egp_oz_mlp_set_fused_slicing(int on) {
    static int state = -1;
    if (state < 0) { const char *e = getenv("EGP_OZ_FUSED_SLICE"); state = (e && atoi(e) == 0) ? 0 : 1; }
    if (on >= 0) state = on ? 1 : 0;
    return state;
}

/* bytes of the input-slice cache for n rows of width in_dim (row slices + transposed slices with the ones row, per chunk) */
static long long xcache_chunk_bytes(int in_dim, long long chunk, int S) {
    const int kp = padk(in_dim);
    const long long cp = pad16(chunk);
    return al((long long)S * chunk * kp) + al(chunk * 4) + al((long long)S * (in_dim + 1) * cp) + al((in_dim + 1) * 4LL);
}

int64_t egp_oz_mlp_xcache_bytes(int in_dim, int64_t n, int64_t chunk_rows, int n_slices) {
    const long long nchunks = (n + chunk_rows - 1) / chunk_rows;
    return nchunks * xcache_chunk_bytes(in_dim, chunk_rows, n_slices);
}

struct MlpPlan {
    Sl W1s, W2s, W3s, W3T, W2T, a1s, a2s, d2s, dys, a1T, a2T, d2T, d1T, dyT, d1s, W1Tc;
    unsigned long long *cmax[8];
    double *a1, *a2, *d1, *d2, *ybuf, *dy, *part;
    uint32_t *rmax;
    unsigned long long *a1b, *a2b;      // relu sign bits of a1 / a2 (one word per row and half n-tile)
    long long part_bytes;
    char *xlocal;
    long long total;
};

static MlpPlan mlp_plan(char *base, int in, int h1, int h2, int od, long long M, int S, int dx_cols = 0) {
    MlpPlan P;
    const long long MP = pad16(M);
    const int kin = padk(in), kh1 = padk(h1), kh2 = padk(h2), kod = padk(od);
    const int hmax = h1 > h2 ? h1 : h2;
    Buf B{base, 0, 0};
    P.W1s = Sl{B.take<int8_t>((long long)S * h1 * kin), B.take<int32_t>(h1 + 16)};
    P.W2s = Sl{B.take<int8_t>((long long)S * h2 * kh1), B.take<int32_t>(h2 + 16)};
    P.W3s = Sl{B.take<int8_t>((long long)S * od * kh2), B.take<int32_t>(od + 16)};
    P.W3T = Sl{B.take<int8_t>((long long)S * h2 * kod), B.take<int32_t>(h2 + 16)};     // rows = h2 index, contraction over out
    P.W2T = Sl{B.take<int8_t>((long long)S * h1 * kh2), B.take<int32_t>(h1 + 16)};     // rows = h1 index, contraction over h2
    for (int i = 0; i < 8; i++) P.cmax[i] = B.take<unsigned long long>(hmax + in + od + 16);
    P.a1 = B.take<double>(M * h1); P.a2 = B.take<double>(M * h2); P.d1 = B.take<double>(M * h1); P.d2 = B.take<double>(M * h2);
    P.ybuf = B.take<double>(M * od); P.dy = B.take<double>(M * od);
    P.rmax = B.take<uint32_t>(M);
    P.a1b = B.take<unsigned long long>(M * gemm_bits_words(h1, S)); P.a2b = B.take<unsigned long long>(M * gemm_bits_words(h2, S));
    P.a1s = Sl{B.take<int8_t>(S * M * kh1), B.take<int32_t>(M)}; P.a2s = Sl{B.take<int8_t>(S * M * kh2), B.take<int32_t>(M)};
    P.d2s = Sl{B.take<int8_t>(S * M * kh2), B.take<int32_t>(M)}; P.dys = Sl{B.take<int8_t>(S * M * kod), B.take<int32_t>(M)};
    P.a1T = Sl{B.take<int8_t>(S * (h1 + 1) * MP), B.take<int32_t>(h1 + 16)}; P.a2T = Sl{B.take<int8_t>(S * (h2 + 1) * MP), B.take<int32_t>(h2 + 16)};
    P.d2T = Sl{B.take<int8_t>(S * h2 * MP), B.take<int32_t>(h2 + 16)}; P.d1T = Sl{B.take<int8_t>(S * h1 * MP), B.take<int32_t>(h1 + 16)};
    P.dyT = Sl{B.take<int8_t>(S * od * MP), B.take<int32_t>(od + 16)};
    if (dx_cols > 0) {              // data gradient of the first layer: row slices of dh1, W1[:, :dx_cols]^T slices
        P.d1s = Sl{B.take<int8_t>(S * M * kh1), B.take<int32_t>(M)};
        P.W1Tc = Sl{B.take<int8_t>((long long)S * dx_cols * kh1), B.take<int32_t>(dx_cols + 16)};
    }
    P.part_bytes = al((long long)num_sms() * BM * 80 * 8 * 2);     // splits * rows * ldp <= SMs * one 128 x 80 tile
    {
        const long long nkb = (MP + BK - 1) / BK, need = (nkb + max_kblocks(S) - 1) / max_kblocks(S) + 1;
        const long long rows = (in > hmax ? in : hmax) + 1, ldp = (hmax + 1) & ~1;
        if (P.part_bytes < al(need * rows * ldp * 8)) P.part_bytes = al(need * rows * ldp * 8);
    }
    P.part = B.take<double>(P.part_bytes / 8);
    P.xlocal = B.take<char>(xcache_chunk_bytes(in, M, S));         // used when the caller passes no cache
    P.total = B.off;
    return P;
}

int64_t egp_oz_mlp_work_bytes(int in_dim, int h1, int h2, int out_dim, int64_t chunk_rows, int n_slices) {
    return mlp_plan(nullptr, in_dim, h1, h2, out_dim, chunk_rows, n_slices, in_dim).total;     // room for any dx_cols <= in_dim
}

/* One forward (+ loss + backward) pass of a two-hidden-layer relu MLP over x [n][in_dim] (leading dimension ldx).
 * loss->kind 0: forward only, y [n][out_dim] is written to d_y.  kind 1 / 2: PPO clipped surrogate / value MSE on the
 * chunk outputs, gradients of the six parameter tensors are written to net->d_g* (overwritten, not accumulated);
 * d_y (optional) still receives the forward output.  d_xcache / xcache_state: 0 no cache, 1 fill the cache while
 * running, 2 use the cache (x unchanged since it was filled with the same n, chunk_rows and n_slices). */
int egp_oz_mlp_step_f64(const EgpMlpNet *net, const double *d_x, int64_t ldx, int64_t n, const EgpMlpLoss *loss, double *d_y,
                        int n_slices, int64_t chunk_rows, void *d_xcache, int xcache_state, void *d_work, int64_t work_bytes,
                        void *stream) {
    if (!net || !d_x || !loss || n < 1 || chunk_rows < 1 || !d_work || n_slices < 3 || n_slices > MAX_S) {
        set_error("egp_oz_mlp_step_f64: bad argument");
        return EGP_EINVAL;
    }
    const int in = net->in_dim, h1 = net->h1, h2 = net->h2, od = net->out_dim, S = n_slices;
    const bool bwd = loss->kind != 0;
    if (in < 1 || h1 < 1 || h2 < 1 || od < 1 || in > 768 || h1 > 768 || h2 > 768 || od > 768) { set_error("egp_oz_mlp_step_f64: layer widths must be in [1, 768]"); return EGP_ESIZE; }
    const int dxc = (bwd && net->d_dx) ? net->dx_cols : 0;
    if (dxc < 0 || dxc > in) { set_error("egp_oz_mlp_step_f64: dx_cols outside [0, in_dim]"); return EGP_EINVAL; }
    if (work_bytes < mlp_plan(nullptr, in, h1, h2, od, chunk_rows, S, dxc).total) { set_error("egp_oz_mlp_step_f64: workspace too small"); return EGP_EINVAL; }
    if (bwd && (!net->d_gW1 || !net->d_gb1 || !net->d_gW2 || !net->d_gb2 || !net->d_gW3 || !net->d_gb3)) { set_error("egp_oz_mlp_step_f64: gradient pointers missing"); return EGP_EINVAL; }
    if (!bwd && !d_y) { set_error("egp_oz_mlp_step_f64: forward-only pass needs d_y"); return EGP_EINVAL; }
    if (xcache_state && !d_xcache) { set_error("egp_oz_mlp_step_f64: cache state without cache"); return EGP_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    const long long M = chunk_rows, MP = pad16(M);
    const int kin = padk(in), kh1 = padk(h1), kh2 = padk(h2), kod = padk(od);
    const int hmax = h1 > h2 ? h1 : h2;
    const MlpPlan P = mlp_plan((char *)d_work, in, h1, h2, od, M, S, dxc);
    const Sl &W1s = P.W1s, &W2s = P.W2s, &W3s = P.W3s, &W3T = P.W3T, &W2T = P.W2T, &a1s = P.a1s, &a2s = P.a2s, &d2s = P.d2s, &dys = P.dys,
             &a1T = P.a1T, &a2T = P.a2T, &d2T = P.d2T, &d1T = P.d1T, &dyT = P.dyT, &d1s = P.d1s, &W1Tc = P.W1Tc;
    unsigned long long *const *cmax = P.cmax;
    double *a1 = P.a1, *a2 = P.a2, *d1 = P.d1, *d2 = P.d2, *ybuf = P.ybuf, *dy = P.dy, *part = P.part;
    uint32_t *rmax = P.rmax;
    unsigned long long *a1b = P.a1b, *a2b = P.a2b;
    // Row / column abs-maxima of every GEMM output come from the producing kernel's epilogue, so each intermediate is read
    // ONCE by oz_slice_both instead of by a row slicer, (a column-maximum pass,) and a transposed slicer.
    // EGP_OZ_FUSED_SLICE=0 keeps the separate passes (bit-identical results; tests compare the two).
    const bool fused = egp_oz_mlp_set_fused_slicing(-1) != 0;
    const long long part_bytes = P.part_bytes;
    char *xlocal = P.xlocal;
    const long long cmax_bytes = (hmax + in + od + 16) * 8LL;

    int rc;
#define OZ_TRY(call) do { rc = (call); if (rc) return rc; } while (0)
    OZ_TRY(slice_rows(net->d_W1, h1, in, in, S, W1s.q, kin, W1s.e, nullptr, st));
    OZ_TRY(slice_rows(net->d_W2, h2, h1, h1, S, W2s.q, kh1, W2s.e, nullptr, st));
    // scalar head (the critic): streamed in plain float64 instead of padding a 1-wide GEMM to a tensor-core tile
    const bool head1 = od == 1 && h2 <= 32 * HEAD_MAXJ;
    const int head_blocks = num_sms() * 3;
    if (!head1) OZ_TRY(slice_rows(net->d_W3, od, h2, h2, S, W3s.q, kh2, W3s.e, nullptr, st));
    if (bwd) {
        EGP_CUDA(cudaMemsetAsync(cmax[0], 0, cmax_bytes, st));
        EGP_CUDA(cudaMemsetAsync(cmax[1], 0, cmax_bytes, st));
        if (!head1) {
            OZ_TRY(col_absmax(net->d_W3, od, h2, h2, cmax[0], st));
            OZ_TRY(slice_colsT(net->d_W3, od, h2, h2, S, cmax[0], W3T.q, kod, W3T.e, 0, st));
        }
        OZ_TRY(col_absmax(net->d_W2, h2, h1, h1, cmax[1], st));
        OZ_TRY(slice_colsT(net->d_W2, h2, h1, h1, S, cmax[1], W2T.q, kh2, W2T.e, 0, st));
        if (dxc > 0) {          // W1[:, :dxc]^T: rows = input feature, contraction over h1
            EGP_CUDA(cudaMemsetAsync(cmax[2], 0, cmax_bytes, st));
            OZ_TRY(col_absmax(net->d_W1, h1, dxc, in, cmax[2], st));
            OZ_TRY(slice_colsT(net->d_W1, h1, dxc, in, S, cmax[2], W1Tc.q, kh1, W1Tc.e, 0, st));
        }
    }
    const long long xc_stride = xcache_chunk_bytes(in, M, S);
    long long chunk_idx = 0;
    for (long long r0 = 0; r0 < n; r0 += M, chunk_idx++) {
        const long long m = n - r0 < M ? n - r0 : M, mp = pad16(m);
        // ---- input slices (cache laid out for full chunks; a short tail chunk uses the same strides)
        char *xb = xcache_state ? (char *)d_xcache + chunk_idx * xc_stride : xlocal;
        Sl xs{(int8_t *)xb, (int32_t *)(xb + al((long long)S * M * kin))};
        Sl xT{(int8_t *)(xb + al((long long)S * M * kin) + al(M * 4)), (int32_t *)(xb + al((long long)S * M * kin) + al(M * 4) + al((long long)S * (in + 1) * MP))};
        const double *xc = d_x + r0 * ldx;
        if (xcache_state != 2) {
            if (bwd || xcache_state == 1) EGP_CUDA(cudaMemsetAsync(cmax[2], 0, cmax_bytes, st));
            OZ_TRY(slice_rows(xc, m, in, ldx, S, xs.q, kin, xs.e, (bwd || xcache_state == 1) ? cmax[2] : nullptr, st));
            if (bwd || xcache_state == 1) OZ_TRY(slice_colsT(xc, m, in, ldx, S, cmax[2], xT.q, mp, xT.e, 1, st));
        }
        GemmOut o;
        // ---- forward
        o = GemmOut(); o.C = a1; o.ldc = h1; o.bias = net->d_b1; o.relu = 1;
        if (bwd) EGP_CUDA(cudaMemsetAsync(cmax[3], 0, cmax_bytes, st));
        if (bwd && fused) { EGP_CUDA(cudaMemsetAsync(rmax, 0, m * 4, st)); o.rowmax = rmax; o.colmax = cmax[3]; o.relu_bits = a1b; }
        OZ_TRY(gemm(xs.q, xs.e, m, W1s.q, W1s.e, h1, kin, S, o, st));
        if (bwd && fused) {
            OZ_TRY(slice_both(a1, m, h1, h1, S, rmax, cmax[3], a1s.q, kh1, a1s.e, a1T.q, mp, a1T.e, 1, st));
        } else {
            OZ_TRY(slice_rows(a1, m, h1, h1, S, a1s.q, kh1, a1s.e, bwd ? cmax[3] : nullptr, st));
            if (bwd) OZ_TRY(slice_colsT(a1, m, h1, h1, S, cmax[3], a1T.q, mp, a1T.e, 1, st));
        }
        o = GemmOut(); o.C = a2; o.ldc = h2; o.bias = net->d_b2; o.relu = 1;
        const bool a2_fused = bwd && fused && !head1;
        if (a2_fused) {
            EGP_CUDA(cudaMemsetAsync(cmax[4], 0, cmax_bytes, st));
            EGP_CUDA(cudaMemsetAsync(rmax, 0, m * 4, st));
            o.rowmax = rmax; o.colmax = cmax[4]; o.relu_bits = a2b;
        }
        OZ_TRY(gemm(a1s.q, a1s.e, m, W2s.q, W2s.e, h2, kh1, S, o, st));
        double *yc = d_y ? d_y + r0 * od : ybuf;
        if (head1) {
            head1_fwd_kernel<<<head_blocks, 32 * HEAD_WARPS, 0, st>>>(a2, m, h2, net->d_W3, net->d_b3, yc);
            EGP_CHECK_LAUNCH("head1_fwd_kernel");
        } else {
            if (a2_fused) {
                OZ_TRY(slice_both(a2, m, h2, h2, S, rmax, cmax[4], a2s.q, kh2, a2s.e, a2T.q, mp, a2T.e, 1, st));
            } else {
                if (bwd) EGP_CUDA(cudaMemsetAsync(cmax[4], 0, cmax_bytes, st));
                OZ_TRY(slice_rows(a2, m, h2, h2, S, a2s.q, kh2, a2s.e, bwd ? cmax[4] : nullptr, st));
                if (bwd) OZ_TRY(slice_colsT(a2, m, h2, h2, S, cmax[4], a2T.q, mp, a2T.e, 1, st));
            }
            o = GemmOut(); o.C = yc; o.ldc = od; o.bias = net->d_b3;
            OZ_TRY(gemm(a2s.q, a2s.e, m, W3s.q, W3s.e, od, kh2, S, o, st));
        }
        if (!bwd) continue;
        // ---- loss on the chunk: dL/dy
        if (loss->kind == 1) {
            OZ_TRY(ppo_loss_grad_launch(yc, loss->d_actions + r0 * od, loss->d_log_std, loss->d_adv + r0, loss->d_stats,
                                        const_cast<double *>(loss->d_logp0) + r0, loss->init_logp0, loss->d_exps + r0, loss->clip_eps,
                                        loss->inv_count, m, od, dy, loss->d_dlogstd, loss->d_loss, st));
        } else if (loss->kind == 2) {
            if (od != 1) { set_error("egp_oz_mlp_step_f64: value loss needs out_dim 1"); return EGP_EINVAL; }
            OZ_TRY(egp_value_loss_grad_f64(yc, loss->d_returns + r0, loss->inv_n, m, dy, loss->d_loss, st));
        } else { set_error("egp_oz_mlp_step_f64: unknown loss kind %d", loss->kind); return EGP_EINVAL; }
        // ---- backward
        const int acc = chunk_idx > 0;
        auto wgrad = [&](const Sl &actT, int fin, const Sl &gT, int fout, double *gW, double *gb) -> int {
            GemmOut w;
            const long long tiles = gemm_tiles(fin + 1, fout, S);
            long long nkb = (mp + BK - 1) / BK;
            long long splits = num_sms() / tiles;
            if (splits > nkb / 4) splits = nkb / 4;
            const long long need = (nkb + max_kblocks(S) - 1) / max_kblocks(S);      // int32 accumulators must not overflow
            if (splits < need) splits = need;
            if (splits < 1) splits = 1;
            w.force_splits = (int)splits; w.work = part; w.work_bytes = part_bytes;
            int r = gemm(actT.q, actT.e, fin + 1, gT.q, gT.e, fout, mp, S, w, st);
            if (r) return r;
            const int tot = (fin + 1) * fout;
            oz_wgrad_reduce_kernel<<<(tot + 255) / 256, 256, 0, st>>>(part, w.splits_used, fin, fout, w.ldp, actT.e, gT.e, gW, gb, acc);
            EGP_CHECK_LAUNCH("oz_wgrad_reduce_kernel");
            return EGP_OK;
        };
        if (head1) {
            if ((long long)head_blocks * (h2 + 1) * 8 > part_bytes) { set_error("egp_oz_mlp_step_f64: head workspace"); return EGP_EINVAL; }
            EGP_CUDA(cudaMemsetAsync(cmax[6], 0, cmax_bytes, st));
            uint32_t *rmx = fused ? rmax : nullptr;
            unsigned long long *cmx = fused ? cmax[6] : nullptr;
            const int jn = (h2 + 31) / 32;
            if (jn <= 4) head1_bwd_kernel<4><<<head_blocks, 32 * HEAD_WARPS, 0, st>>>(a2, dy, m, h2, net->d_W3, d2, part, rmx, cmx);
            else if (jn <= 8) head1_bwd_kernel<8><<<head_blocks, 32 * HEAD_WARPS, 0, st>>>(a2, dy, m, h2, net->d_W3, d2, part, rmx, cmx);
            else if (jn <= 10) head1_bwd_kernel<10><<<head_blocks, 32 * HEAD_WARPS, 0, st>>>(a2, dy, m, h2, net->d_W3, d2, part, rmx, cmx);
            else head1_bwd_kernel<HEAD_MAXJ><<<head_blocks, 32 * HEAD_WARPS, 0, st>>>(a2, dy, m, h2, net->d_W3, d2, part, rmx, cmx);
            EGP_CHECK_LAUNCH("head1_bwd_kernel");
            head1_reduce_kernel<<<h2 + 1, 128, 0, st>>>(part, head_blocks, h2, net->d_gW3, net->d_gb3, acc);
            EGP_CHECK_LAUNCH("head1_reduce_kernel");
        } else {
            EGP_CUDA(cudaMemsetAsync(cmax[5], 0, cmax_bytes, st));
            OZ_TRY(slice_rows(dy, m, od, od, S, dys.q, kod, dys.e, cmax[5], st));
            OZ_TRY(slice_colsT(dy, m, od, od, S, cmax[5], dyT.q, mp, dyT.e, 0, st));
            OZ_TRY(wgrad(a2T, h2, dyT, od, net->d_gW3, net->d_gb3));
            EGP_CUDA(cudaMemsetAsync(cmax[6], 0, cmax_bytes, st));
            o = GemmOut(); o.C = d2; o.ldc = h2; o.mask = a2; o.ldm = h2;
            if (fused) { EGP_CUDA(cudaMemsetAsync(rmax, 0, m * 4, st)); o.rowmax = rmax; o.colmax = cmax[6]; o.mask_bits = a2b; }
            OZ_TRY(gemm(dys.q, dys.e, m, W3T.q, W3T.e, h2, kod, S, o, st));
        }
        if (fused) {
            OZ_TRY(slice_both(d2, m, h2, h2, S, rmax, cmax[6], d2s.q, kh2, d2s.e, d2T.q, mp, d2T.e, 0, st));
        } else {
            OZ_TRY(slice_rows(d2, m, h2, h2, S, d2s.q, kh2, d2s.e, cmax[6], st));
            OZ_TRY(slice_colsT(d2, m, h2, h2, S, cmax[6], d2T.q, mp, d2T.e, 0, st));
        }
        OZ_TRY(wgrad(a1T, h1, d2T, h2, net->d_gW2, net->d_gb2));
        EGP_CUDA(cudaMemsetAsync(cmax[7], 0, cmax_bytes, st));
        o = GemmOut(); o.C = d1; o.ldc = h1; o.mask = a1; o.ldm = h1;
        if (fused) {
            o.colmax = cmax[7]; o.mask_bits = a1b;
            if (dxc > 0) { EGP_CUDA(cudaMemsetAsync(rmax, 0, m * 4, st)); o.rowmax = rmax; }
        }
        OZ_TRY(gemm(d2s.q, d2s.e, m, W2T.q, W2T.e, h1, kh2, S, o, st));
        if (dxc > 0) {
            if (fused) {
                OZ_TRY(slice_both(d1, m, h1, h1, S, rmax, cmax[7], d1s.q, kh1, d1s.e, d1T.q, mp, d1T.e, 0, st));
            } else {
                OZ_TRY(slice_rows(d1, m, h1, h1, S, d1s.q, kh1, d1s.e, cmax[7], st));
                OZ_TRY(slice_colsT(d1, m, h1, h1, S, cmax[7], d1T.q, mp, d1T.e, 0, st));
            }
            o = GemmOut(); o.C = net->d_dx + r0 * dxc; o.ldc = dxc;
            OZ_TRY(gemm(d1s.q, d1s.e, m, W1Tc.q, W1Tc.e, dxc, kh1, S, o, st));
        } else {
            if (!fused) OZ_TRY(col_absmax(d1, m, h1, h1, cmax[7], st));
            OZ_TRY(slice_colsT(d1, m, h1, h1, S, cmax[7], d1T.q, mp, d1T.e, 0, st));
        }
        OZ_TRY(wgrad(xT, in, d1T, h1, net->d_gW1, net->d_gb1));
    }
#undef OZ_TRY
    return EGP_OK;
}

}  // extern "C"
