// PPO update kernels (K4 GAE scan, K5 loss forward+backward, K6 clip+Adam) - sm_100a, float64.
// All of these are HBM-bound streaming kernels: coalesced 128-bit accesses, grid sized in multiples
// of the SM count, one pass over the data (SURVEY.md 8d gives the algorithmic bytes per unit).
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace egp {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return EGP_ECUDA;
}
int num_sms(int device) {
    static int cached[64] = {0};
    if (device < 0) cudaGetDevice(&device);
    if (device >= 0 && device < 64 && cached[device]) return cached[device];
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
    if (device >= 0 && device < 64) cached[device] = n;
    return n;
}

// ------------------------------------------------------------------------------------------ K4 GAE
// core/common.py:5-25.  A_i = delta_i + c_i A_{i+1},  delta_i = r_i + gamma V_{i+1} m_i - V_i,
// c_i = gamma tau m_i, flat over the batch with A_N = V_N = 0: an affine recurrence, scanned in reverse.
//   pass 1 (gae_agg_kernel)   : every 2048-sample tile reduces its samples to one affine map (C, D) with
//                               A_first = D + C A_in  (thread-local compose -> warp-shuffle scan -> 8 warps)
//   pass 2 (gae_apply_kernel) : a tile folds the maps of the tiles after it until C hits 0 (the first episode
//                               boundary - normally the very next tile), rescans with that carry-in, writes
//                               advantages and returns = V + A, and a Chan (n, mean, M2) partial of A
//   pass 3 (gae_moments_kernel): merges the partials in tile order (deterministic) into stats[3]
// No atomics, fences or spinning.  The second read of (r, m, V) comes out of L2 (inputs << 126 MB), so DRAM
// traffic stays at the algorithmic 5 doubles / sample (3 in, 2 out).
constexpr int GAE_THREADS = 256;
#ifndef GAE_ITEMS_N
#define GAE_ITEMS_N 8
#endif
constexpr int GAE_ITEMS = GAE_ITEMS_N;
constexpr int GAE_TILE = GAE_THREADS * GAE_ITEMS;
#ifndef GAE_EXP
#define GAE_EXP 0
#endif
#ifndef GAE_ONEPASS_CTAS
#define GAE_ONEPASS_CTAS 4
#endif

struct Moments { double n, mean, m2; };

__device__ __forceinline__ Moments merge(Moments a, Moments b) {
    if (b.n == 0.0) return a;
    if (a.n == 0.0) return b;
    Moments r;
    r.n = a.n + b.n;
    double d = b.mean - a.mean;
    if (a.n == b.n) {                               // equal counts (every merge inside a full tile): the same arithmetic as
        r.mean = a.mean + d * 0.5;                  // below with b.n / r.n = 1/2 and a.n b.n / r.n = a.n / 2, no division
        r.m2 = a.m2 + b.m2 + d * d * (0.5 * a.n);
        return r;
    }
    r.mean = a.mean + d * (b.n / r.n);
    r.m2 = a.m2 + b.m2 + d * d * (a.n * b.n / r.n);
    return r;
}
__device__ __forceinline__ Moments shfl_down(Moments a, int o) {
    Moments r;
    r.n = __shfl_down_sync(0xffffffffu, a.n, o);
    r.mean = __shfl_down_sync(0xffffffffu, a.mean, o);
    r.m2 = __shfl_down_sync(0xffffffffu, a.m2, o);
    return r;
}

// loads this thread's GAE_ITEMS samples (thread 0 owns the HIGHEST indices of the tile) and composes them:
// dl[k] = delta, c[k] = gamma tau m; (C, D) = map of the thread's segment
__device__ __forceinline__ void gae_load_compose(const double *__restrict__ rew, const double *__restrict__ msk,
                                                 const double *__restrict__ val, double gamma, double tau, long long n,
                                                 long long lo, bool vec, double *dl, double *c, double *v, double &C, double &D) {
    if (vec && lo + GAE_ITEMS < n) {    // interior, 16-byte aligned arrays: 128-bit loads
#pragma unroll
        for (int k = 0; k < GAE_ITEMS; k += 2) {
            double2 r2 = __ldcs(reinterpret_cast<const double2 *>(rew + lo + k));       // read once: streaming
            double2 m2 = __ldcs(reinterpret_cast<const double2 *>(msk + lo + k));
            double2 v2 = __ldcs(reinterpret_cast<const double2 *>(val + lo + k));
            dl[k] = r2.x; dl[k + 1] = r2.y; c[k] = m2.x; c[k + 1] = m2.y; v[k] = v2.x; v[k + 1] = v2.y;
        }
        v[GAE_ITEMS] = val[lo + GAE_ITEMS];
    } else {
#pragma unroll
        for (int k = 0; k <= GAE_ITEMS; k++) {
            const long long i = lo + k;
            const bool ok = i < n;
            if (k < GAE_ITEMS) { dl[k] = ok ? rew[i] : 0.0; c[k] = ok ? msk[i] : 0.0; }
            v[k] = ok ? val[i] : 0.0;
        }
    }
    C = 1.0; D = 0.0;
#pragma unroll
    for (int k = GAE_ITEMS - 1; k >= 0; k--) {      // highest index first
        const double m = c[k];
        dl[k] = dl[k] + gamma * v[k + 1] * m - v[k];
        c[k] = gamma * tau * m;
        D = dl[k] + c[k] * D;
        C = c[k] * C;
    }
}

// (n, mean, M2) of this thread's GAE_ITEMS advantages: two short sums instead of a Chan merge (with its divisions) per sample
__device__ __forceinline__ Moments gae_thread_moments(const double *out_a, long long lo, long long n) {
    Moments mom = {0.0, 0.0, 0.0};
    if (lo + GAE_ITEMS <= n) {
        double sum = 0.0;
#pragma unroll
        for (int k = 0; k < GAE_ITEMS; k++) sum += out_a[k];
        mom.n = GAE_ITEMS; mom.mean = sum * (1.0 / GAE_ITEMS);
#pragma unroll
        for (int k = 0; k < GAE_ITEMS; k++) { const double d = out_a[k] - mom.mean; mom.m2 = fma(d, d, mom.m2); }
    } else if (lo < n) {
        double sum = 0.0;
        for (int k = 0; k < GAE_ITEMS; k++) if (lo + k < n) { sum += out_a[k]; mom.n += 1.0; }
        mom.mean = sum / mom.n;
        for (int k = 0; k < GAE_ITEMS; k++) if (lo + k < n) { const double d = out_a[k] - mom.mean; mom.m2 = fma(d, d, mom.m2); }
    }
    return mom;
}

// (n, mean, M2) of a tile from the per-thread moments: shuffle tree per warp, then an ordered tree over the 8 warps
// (equal counts at every level of a full tile: no divisions); valid in thread 0.  Contains one __syncthreads.
__device__ __forceinline__ Moments gae_tile_moments(Moments mom, Moments *s_mom) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mom = merge(mom, shfl_down(mom, o));
    if (lane == 0) s_mom[warp] = mom;
    __syncthreads();
    Moments t = {0.0, 0.0, 0.0};
    if (warp == 0) {
        if (lane < GAE_THREADS / 32) t = s_mom[lane];
#pragma unroll
        for (int o = 1; o < GAE_THREADS / 32; o <<= 1) {
            Moments nb = shfl_down(t, o);
            if ((lane & (2 * o - 1)) == 0) t = merge(t, nb);
        }
    }
    return t;
}

// inclusive scan of the per-thread maps over the block in thread order (descending sample index);
// returns the exclusive map of this thread in (eC, eD) and the tile map in (tC, tD) (valid in every thread)
__device__ __forceinline__ void gae_block_scan(double C, double D, double &eC, double &eD, double &tC, double &tD) {
    __shared__ double s_c[GAE_THREADS / 32], s_d[GAE_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double iC = C, iD = D;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double pc = __shfl_up_sync(0xffffffffu, iC, o), pd = __shfl_up_sync(0xffffffffu, iD, o);
        if (lane >= o) { iD = iD + iC * pd; iC = iC * pc; }
    }
    if (lane == 31) { s_c[warp] = iC; s_d[warp] = iD; }
    __syncthreads();
    double wC = 1.0, wD = 0.0;                      // composition of all earlier warps
    for (int w = 0; w < warp; w++) { wD = s_d[w] + s_c[w] * wD; wC = s_c[w] * wC; }
    tC = 1.0; tD = 0.0;
    for (int w = 0; w < GAE_THREADS / 32; w++) { tD = s_d[w] + s_c[w] * tD; tC = s_c[w] * tC; }
    iD = iD + iC * wD;
    iC = iC * wC;
    eC = __shfl_up_sync(0xffffffffu, iC, 1);
    eD = __shfl_up_sync(0xffffffffu, iD, 1);
    if (lane == 0) { eC = wC; eD = wD; }
}

__global__ void __launch_bounds__(GAE_THREADS)
gae_agg_kernel(const double *__restrict__ rew, const double *__restrict__ msk, const double *__restrict__ val, double gamma,
               double tau, long long n, bool vec, double *__restrict__ agg /* [ntiles][2] */) {
    const long long lo = (long long)blockIdx.x * GAE_TILE + (long long)(GAE_THREADS - 1 - threadIdx.x) * GAE_ITEMS;
    double dl[GAE_ITEMS], c[GAE_ITEMS], v[GAE_ITEMS + 1], C, D, eC, eD, tC, tD;
    gae_load_compose(rew, msk, val, gamma, tau, n, lo, vec, dl, c, v, C, D);
    gae_block_scan(C, D, eC, eD, tC, tD);
    if (threadIdx.x == 0) { agg[2 * (size_t)blockIdx.x] = tC; agg[2 * (size_t)blockIdx.x + 1] = tD; }
}

__global__ void __launch_bounds__(GAE_THREADS)
gae_apply_kernel(const double *__restrict__ rew, const double *__restrict__ msk, const double *__restrict__ val, double gamma,
                 double tau, long long n, bool vec, unsigned int ntiles, const double *__restrict__ agg, double *__restrict__ adv,
                 double *__restrict__ ret, double *__restrict__ partial /* [ntiles][3] */) {
    __shared__ Moments s_mom[GAE_THREADS / 32];
    const int tid = threadIdx.x;
    const long long lo = (long long)blockIdx.x * GAE_TILE + (long long)(GAE_THREADS - 1 - tid) * GAE_ITEMS;
    double dl[GAE_ITEMS], c[GAE_ITEMS], v[GAE_ITEMS + 1], C, D, eC, eD, tC, tD;
    gae_load_compose(rew, msk, val, gamma, tau, n, lo, vec, dl, c, v, C, D);
    // carry-in: advantage of the first sample of the next tile, folded through the following tiles' maps
    double aC = 1.0, ain = 0.0;
    for (unsigned int q = blockIdx.x + 1; q < ntiles && aC != 0.0; q++) {
        ain += aC * agg[2 * (size_t)q + 1];
        aC *= agg[2 * (size_t)q];
    }
    gae_block_scan(C, D, eC, eD, tC, tD);
    double a = eD + eC * ain;                       // advantage just above this thread's highest sample
    double out_a[GAE_ITEMS];
#pragma unroll
    for (int k = GAE_ITEMS - 1; k >= 0; k--) {
        a = dl[k] + c[k] * a;
        out_a[k] = a;
    }
    Moments mom = gae_thread_moments(out_a, lo, n);
    if (vec && lo + GAE_ITEMS <= n) {
#pragma unroll
        for (int k = 0; k < GAE_ITEMS; k += 2) {
            *reinterpret_cast<double2 *>(adv + lo + k) = make_double2(out_a[k], out_a[k + 1]);
            *reinterpret_cast<double2 *>(ret + lo + k) = make_double2(v[k] + out_a[k], v[k + 1] + out_a[k + 1]);
        }
    } else {
#pragma unroll
        for (int k = 0; k < GAE_ITEMS; k++)
            if (lo + k < n) { adv[lo + k] = out_a[k]; ret[lo + k] = v[k] + out_a[k]; }
    }
    const Moments t = gae_tile_moments(mom, s_mom);
    if (tid == 0) {
        partial[3 * (size_t)blockIdx.x + 0] = t.n; partial[3 * (size_t)blockIdx.x + 1] = t.mean; partial[3 * (size_t)blockIdx.x + 2] = t.m2;
    }
}

// ordered merge of the per-tile partials by one 256-thread block (deterministic: contiguous slices per thread in tile
// order, ordered shuffle tree, warps in order)
__device__ __forceinline__ void gae_merge_partials(const double *partial, unsigned int ntiles, double *__restrict__ stats) {
    __shared__ Moments s_mom[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int per = (ntiles + 255) / 256, k0 = threadIdx.x * per;
    // this thread's contiguous slice, in tile order.  Partials with the count of the slice's first one (all full tiles)
    // combine without a division each: mean of the means, M2 = sum of M2 + count * sum of squared mean deviations
    // (second read of the slice: L2); the others (the ragged last tile) go through the general merge.
    Moments t = {0.0, 0.0, 0.0}, odd = {0.0, 0.0, 0.0};
    if (k0 < ntiles) {
        const unsigned int k1 = min(k0 + per, ntiles);
        const double n0 = __ldcg(partial + 3 * (size_t)k0);
        double cnt = 0.0, sum = 0.0, m2 = 0.0;
        for (unsigned int k = k0; k < k1; k++) {
            Moments o = {__ldcg(partial + 3 * (size_t)k), __ldcg(partial + 3 * (size_t)k + 1), __ldcg(partial + 3 * (size_t)k + 2)};
            if (o.n == n0) { cnt += 1.0; sum += o.mean; m2 += o.m2; }
            else odd = merge(odd, o);
        }
        const double mean = sum / cnt;
        double dev = 0.0;
        for (unsigned int k = k0; k < k1; k++)
            if (__ldcg(partial + 3 * (size_t)k) == n0) { const double d = __ldcg(partial + 3 * (size_t)k + 1) - mean; dev = fma(d, d, dev); }
        t.n = cnt * n0; t.mean = mean; t.m2 = m2 + n0 * dev;
        if (n0 == 0.0) t = Moments{0.0, 0.0, 0.0};
        t = merge(t, odd);
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {              // ordered tree: lane l absorbs lane l + o
        Moments nb = shfl_down(t, o);
        if ((lane & (2 * o - 1)) == 0) t = merge(t, nb);
    }
    if (lane == 0) s_mom[warp] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        Moments r = s_mom[0];
        for (int w = 1; w < 8; w++) r = merge(r, s_mom[w]);
        stats[0] = r.n; stats[1] = r.mean; stats[2] = r.m2;
    }
}

__global__ void __launch_bounds__(256)
gae_moments_kernel(const double *__restrict__ partial, unsigned int ntiles, double *__restrict__ stats) {
    gae_merge_partials(partial, ntiles, stats);
}

// Large batches (inputs beyond what L2 keeps between two passes): ONE pass by a persistent grid (every CTA resident, so
// waiting on another CTA cannot deadlock).  CTA b takes the tiles ntiles-1-b, ntiles-1-b-G, ... (from the END of the
// batch: the direction of the recurrence).  A tile publishes its affine map (C, D) right after its block scan as ONE
// 16-byte store into a slot the caller filled with an all-ones bit pattern, then folds the maps of the tiles after it
// until C hits 0 (the first episode boundary: normally inside the very next tile, which the neighbouring CTA is
// processing at the same moment), re-reading (one 16-byte load) only slots that still hold the pattern: no flags and no
// fences.  The last CTA to finish merges the moment partials in tile order.  5 doubles of traffic per sample (3 in,
// 2 out), the algorithmic minimum; results are bit-identical to the two-pass kernels (same per-tile arithmetic, same
// merge order).  `done` starts at 2^32 - 1 (the caller's one memset fills slots and counter with ones).
constexpr unsigned long long GAE_EMPTY = 0xffffffffffffffffull;

__device__ __forceinline__ void gae_publish(double *slot, double C, double D) {
    unsigned long long c = (unsigned long long)__double_as_longlong(C), d = (unsigned long long)__double_as_longlong(D);
    if (c == GAE_EMPTY) c = 0x7ff8000000000000ull;          // a NaN carrying the reserved payload stays a NaN
    if (d == GAE_EMPTY) d = 0x7ff8000000000000ull;
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"(c), "l"(d) : "memory");
}
__device__ __forceinline__ void gae_fetch(const double *slot, double &C, double &D) {
    unsigned long long c, d;
    for (;;) {
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(c), "=l"(d) : "l"(slot) : "memory");
        if (c != GAE_EMPTY && d != GAE_EMPTY) break;
        __nanosleep(20);
    }
    C = __longlong_as_double((long long)c); D = __longlong_as_double((long long)d);
}

__global__ void __launch_bounds__(GAE_THREADS, GAE_ONEPASS_CTAS)
gae_onepass_kernel(const double *__restrict__ rew, const double *__restrict__ msk, const double *__restrict__ val, double gamma,
                   double tau, long long n, bool vec, unsigned int ntiles, double *agg, unsigned int *done,
                   double *__restrict__ adv, double *__restrict__ ret, double *partial, double *__restrict__ stats) {
    __shared__ Moments s_mom[GAE_THREADS / 32];
    __shared__ unsigned int s_last;
    __shared__ double s_ain;
    const int tid = threadIdx.x;
    for (long long tl = (long long)ntiles - 1 - blockIdx.x; tl >= 0; tl -= gridDim.x) {
        const unsigned int tile = (unsigned int)tl;
        const long long lo = (long long)tile * GAE_TILE + (long long)(GAE_THREADS - 1 - tid) * GAE_ITEMS;
        double dl[GAE_ITEMS], c[GAE_ITEMS], v[GAE_ITEMS + 1], C, D, eC, eD, tC, tD;
        gae_load_compose(rew, msk, val, gamma, tau, n, lo, vec, dl, c, v, C, D);
        gae_block_scan(C, D, eC, eD, tC, tD);
        if (tid == 0) {
            gae_publish(agg + 2 * (size_t)tile, tC, tD);
            double aC = 1.0, ain = 0.0;
            for (unsigned int q = tile + 1; q < ntiles && aC != 0.0; q++) {
                double qC, qD;
                gae_fetch(agg + 2 * (size_t)q, qC, qD);
                ain += aC * qD;
                aC *= qC;
            }
            s_ain = ain;
        }
        __syncthreads();
        double a = eD + eC * s_ain;
        double out_a[GAE_ITEMS];
#pragma unroll
        for (int k = GAE_ITEMS - 1; k >= 0; k--) {
            a = dl[k] + c[k] * a;
            out_a[k] = a;
        }
        Moments mom = gae_thread_moments(out_a, lo, n);
        if (vec && lo + GAE_ITEMS <= n) {
#pragma unroll
            for (int k = 0; k < GAE_ITEMS; k += 2) {
                __stcs(reinterpret_cast<double2 *>(adv + lo + k), make_double2(out_a[k], out_a[k + 1]));
                __stcs(reinterpret_cast<double2 *>(ret + lo + k), make_double2(v[k] + out_a[k], v[k + 1] + out_a[k + 1]));
            }
        } else {
#pragma unroll
            for (int k = 0; k < GAE_ITEMS; k++)
                if (lo + k < n) { adv[lo + k] = out_a[k]; ret[lo + k] = v[k] + out_a[k]; }
        }
        const Moments t = gae_tile_moments(mom, s_mom);
        if (tid == 0) {
            __stcg(partial + 3 * (size_t)tile + 0, t.n); __stcg(partial + 3 * (size_t)tile + 1, t.mean); __stcg(partial + 3 * (size_t)tile + 2, t.m2);
        }
        // the next iteration's block scan has a barrier between these shared-memory reads and its own writes of s_mom / s_ain
    }
    if (tid == 0) {
        __threadfence();
        s_last = atomicAdd(done, 1u) + 1u == gridDim.x - 1u ? 1u : 0u;     // counts up from 2^32 - 1 (one memset fills slots and counter)
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        gae_merge_partials(partial, ntiles, stats);
    }
}

__global__ void standardize_kernel(double *x, long long n, const double *stats) {
    const double mean = stats[1], inv = 1.0 / sqrt(stats[2] / (stats[0] - 1.0));
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        x[i] = (x[i] - mean) * inv;
}

// ------------------------------------------------------------------------- K5 loss fwd + bwd
// 16 lanes per row; lane s of a group handles action dims s, s+16, s+32, ... (coalesced 128 B runs).
constexpr int LPR = 16;
constexpr double HALF_LOG_2PI = 0.91893853320467274178;

__device__ __forceinline__ double group_sum16(double v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(256)
gauss_logp_kernel(const double *__restrict__ mu, const double *__restrict__ act, const double *__restrict__ log_std,
                  long long n, int adim, double *__restrict__ logp) {
    const int sub = threadIdx.x & (LPR - 1);
    const long long g0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / LPR;
    const long long stride = (long long)gridDim.x * blockDim.x / LPR;
    // warp-uniform trip count: both 16-lane halves of a warp run the same number of iterations (a half whose row is
    // past the end computes on row n - 1 and discards), so every full-mask shuffle below is executed by all 32 lanes
    const long long g_lo = (blockIdx.x * (long long)blockDim.x + (threadIdx.x & ~31)) / LPR;
    for (long long row_u = g_lo, row_t = g0; row_u < n; row_u += stride, row_t += stride) {
        const bool valid = row_t < n;
        const long long row = valid ? row_t : n - 1;
        double s = 0.0;
        for (int j = sub; j < adim; j += LPR) {
            double ls = log_std[j], d = act[row * adim + j] - mu[row * adim + j];
            double var = exp(ls) * exp(ls);
            s += -(d * d) / (2.0 * var) - ls - HALF_LOG_2PI;
        }
        s = group_sum16(s);
        if (sub == 0 && valid) logp[row] = s;
    }
}

__global__ void __launch_bounds__(256)
ppo_loss_grad_kernel(const double *__restrict__ mu, const double *__restrict__ act, const double *__restrict__ log_std,
                     const double *__restrict__ adv, const double *__restrict__ stats, double *logp0,
                     const double *__restrict__ exps, double clip_eps, double inv_count, long long n, int adim,
                     double *__restrict__ dmu, double *__restrict__ dlogstd, double *__restrict__ loss, int record) {
    const int sub = threadIdx.x & (LPR - 1);
    const long long g0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / LPR;
    const long long stride = (long long)gridDim.x * blockDim.x / LPR;
    const double a_mean = stats[1], a_inv = 1.0 / sqrt(stats[2] / (stats[0] - 1.0));
    double loss_acc = 0.0;
    double dls_acc[4] = {0.0, 0.0, 0.0, 0.0};       // adim <= 64
    // warp-uniform trip count: both 16-lane halves of a warp run the same number of iterations (a half whose row is
    // past the end computes on row n - 1 and discards), so every full-mask shuffle below is executed by all 32 lanes
    const long long g_lo = (blockIdx.x * (long long)blockDim.x + (threadIdx.x & ~31)) / LPR;
    for (long long row_u = g_lo, row_t = g0; row_u < n; row_u += stride, row_t += stride) {
        const bool valid = row_t < n;
        const long long row = valid ? row_t : n - 1;
        double d[4], iv[4];
        double s = 0.0;
        int cnt = 0;
        for (int j = sub; j < adim; j += LPR, cnt++) {
            double ls = log_std[j];
            double sd = exp(ls);
            double var = sd * sd;
            d[cnt] = act[row * adim + j] - mu[row * adim + j];
            iv[cnt] = 1.0 / var;
            s += -(d[cnt] * d[cnt]) / (2.0 * var) - ls - HALF_LOG_2PI;
        }
        s = group_sum16(s);
        const bool on = valid && exps[row] != 0.0;
        // record: this IS the pass that defines fixed_log_probs (agent_ppo.py:18-20: same parameters, same arithmetic as
        // the first epoch's forward, so the reference's ratio there is exp(0) = 1 as well)
        if (record && valid && sub == 0) logp0[row] = s;
        double coef = 0.0;                           // dL/dlogp
        if (on) {
            double ratio = record ? 1.0 : exp(s - logp0[row]);
            double ah = (adv[row] - a_mean) * a_inv;
            double lo = 1.0 - clip_eps, hi = 1.0 + clip_eps;
            double clamped = fmin(fmax(ratio, lo), hi);
            double s1 = ratio * ah, s2 = clamped * ah;
            double inrange = (ratio >= lo && ratio <= hi) ? 1.0 : 0.0;
            // torch.min backward: ties split evenly; clamp backward passes inside the closed range
            double w = s1 < s2 ? 1.0 : (s1 > s2 ? inrange : 0.5 + 0.5 * inrange);
            coef = -inv_count * ah * w * ratio;
            if (sub == 0) loss_acc += -fmin(s1, s2) * inv_count;
        }
        cnt = 0;
        for (int j = sub; j < adim && valid; j += LPR, cnt++) {
            dmu[row * adim + j] = coef * d[cnt] * iv[cnt];
            dls_acc[cnt] += coef * (d[cnt] * d[cnt] * iv[cnt] - 1.0);
        }
    }
    // block reduction of the scalar loss
    loss_acc = warp_sum(loss_acc);
    if ((threadIdx.x & 31) == 0 && loss_acc != 0.0) atomicAdd(loss, loss_acc);
    if (dlogstd) {
#pragma unroll
        for (int cnt = 0; cnt < 4; cnt++) {         // uniform trip count: the shuffle is executed by all 32 lanes
            const int j = sub + cnt * LPR;
            double v = dls_acc[cnt] + __shfl_xor_sync(0xffffffffu, dls_acc[cnt], 16);
            if (j < adim && (threadIdx.x & 31) < LPR && v != 0.0) atomicAdd(dlogstd + j, v);
        }
    }
}

__global__ void __launch_bounds__(256)
value_loss_grad_kernel(const double *__restrict__ v, const double *__restrict__ ret, double inv_n, long long n,
                       double *__restrict__ dv, double *__restrict__ loss) {
    double acc = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double e = v[i] - ret[i];
        dv[i] = 2.0 * e * inv_n;
        acc += e * e;
    }
    acc = warp_sum(acc);
    __shared__ double s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += s[w];
        atomicAdd(loss, t * inv_n);
    }
}

// ---------------------------------------------------------------- elementwise helpers around GEMMs
__global__ void __launch_bounds__(256)
bias_relu_kernel(double *__restrict__ y, const double *__restrict__ b, long long total, int dim) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        double t = y[i] + b[i % dim];
        y[i] = t > 0.0 ? t : 0.0;
    }
}
__global__ void __launch_bounds__(256)
relu_bwd_kernel(double *__restrict__ dy, const double *__restrict__ y, long long total) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        if (!(y[i] > 0.0)) dy[i] = 0.0;
}
// column sums: block = 32 x 8 threads; each block owns a band of rows, lanes sweep columns (coalesced)
__global__ void __launch_bounds__(256)
colsum_kernel(const double *__restrict__ x, long long n, int dim, long long rows_per_block, double *__restrict__ out) {
    __shared__ double s[8][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    long long r0 = blockIdx.x * rows_per_block, r1 = r0 + rows_per_block;
    if (r1 > n) r1 = n;
    for (int c0 = 0; c0 < dim; c0 += 32) {
        int c = c0 + lane;
        double acc = 0.0;
        if (c < dim)
            for (long long r = r0 + w; r < r1; r += 8) acc += x[r * dim + c];
        s[w][lane] = acc;
        __syncthreads();
        if (w == 0 && c < dim) {
            double t = 0.0;
#pragma unroll
            for (int k = 0; k < 8; k++) t += s[k][lane];
            atomicAdd(out + c, t);
        }
        __syncthreads();
    }
}

// out[i][:] = in[perm[i]][:]  (agents/agent_ppo.py:30-32 'states[perm].clone()' for every batch array)
__global__ void __launch_bounds__(256)
gather_rows_kernel(const double *__restrict__ in, const long long *__restrict__ perm, long long n, int dim,
                   double *__restrict__ out) {
    const long long total = n * dim;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long i = idx / dim;
        const int j = (int)(idx - i * dim);
        out[idx] = in[perm[i] * dim + j];
    }
}

// fused relu backward + bias gradient: dy *= (y > 0) in place and out[c] += sum_rows dy[:, c] (one pass)
__global__ void __launch_bounds__(256)
relu_bwd_colsum_kernel(double *__restrict__ dy, const double *__restrict__ y, long long n, int dim, long long rows_per_block,
                       double *__restrict__ out) {
    __shared__ double s[8][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    long long r0 = blockIdx.x * rows_per_block, r1 = r0 + rows_per_block;
    if (r1 > n) r1 = n;
    for (int c0 = 0; c0 < dim; c0 += 32) {
        int c = c0 + lane;
        double acc = 0.0;
        if (c < dim)
            for (long long r = r0 + w; r < r1; r += 8) {
                const long long idx = r * dim + c;
                double g = y[idx] > 0.0 ? dy[idx] : 0.0;
                dy[idx] = g;
                acc += g;
            }
        s[w][lane] = acc;
        __syncthreads();
        if (w == 0 && c < dim) {
            double t = 0.0;
#pragma unroll
            for (int k = 0; k < 8; k++) t += s[k][lane];
            atomicAdd(out + c, t);
        }
        __syncthreads();
    }
}

// shifted first/second column moments (batched ZFilter update)
__global__ void __launch_bounds__(256)
col_moments_kernel(const double *__restrict__ x, long long n, int dim, long long rows_per_block,
                   const double *__restrict__ shift, double *__restrict__ out) {
    __shared__ double s1[8][33], s2[8][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    long long r0 = blockIdx.x * rows_per_block, r1 = r0 + rows_per_block;
    if (r1 > n) r1 = n;
    for (int c0 = 0; c0 < dim; c0 += 32) {
        int c = c0 + lane;
        double a1 = 0.0, a2 = 0.0;
        if (c < dim) {
            const double sh = shift ? shift[c] : 0.0;
            for (long long r = r0 + w; r < r1; r += 8) { double d = x[r * dim + c] - sh; a1 += d; a2 += d * d; }
        }
        s1[w][lane] = a1; s2[w][lane] = a2;
        __syncthreads();
        if (w == 0 && c < dim) {
            double t1 = 0.0, t2 = 0.0;
#pragma unroll
            for (int k = 0; k < 8; k++) { t1 += s1[k][lane]; t2 += s2[k][lane]; }
            atomicAdd(out + c, t1);
            atomicAdd(out + dim + c, t2);
        }
        __syncthreads();
    }
}

// -------------------------------------------------------------------------------- K6 clip + Adam
// Deterministic sum of squares (fixed grid, last block folds the partials in index order) so that
// data-parallel replicas compute bit-identical clip coefficients after the gradient all-reduce.
constexpr int SUMSQ_BLOCKS = 148;
__device__ double g_sumsq_partial[SUMSQ_BLOCKS];
__device__ unsigned int g_sumsq_ticket = 0;

__global__ void __launch_bounds__(256)
sumsq_kernel(const double *__restrict__ g, long long n, double *__restrict__ norm2) {
    double acc = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        acc += g[i] * g[i];
    acc = warp_sum(acc);
    __shared__ double s[8];
    __shared__ bool last;
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += s[w];
        g_sumsq_partial[blockIdx.x] = t;
        __threadfence();
        last = atomicAdd(&g_sumsq_ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        double t = 0.0;
        for (unsigned int b = 0; b < gridDim.x; b++) t += g_sumsq_partial[b];
        norm2[0] = t;
        g_sumsq_ticket = 0;
    }
}

// torch.optim.Adam (amsgrad False, weight_decay 0) with the clip_grad_norm_ coefficient folded in:
// coef = min(1, max_norm / (||g|| + 1e-6)) when max_norm > 0 (torch.nn.utils.clip_grad_norm_).
__global__ void __launch_bounds__(256)
adam_kernel(double *__restrict__ p, const double *__restrict__ g, double *__restrict__ m, double *__restrict__ v,
            long long n, double lr, double b1, double b2, double eps, double bc1, double bc2_sqrt, double max_norm,
            const double *__restrict__ norm2) {
    double coef = 1.0;
    if (max_norm > 0.0) {
        double c = max_norm / (sqrt(norm2[0]) + 1e-6);
        coef = c < 1.0 ? c : 1.0;
    }
    const double step_size = lr / bc1;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double gi = g[i] * coef;
        double mi = m[i] + (gi - m[i]) * (1.0 - b1);        // exp_avg.lerp_(grad, 1 - beta1)
        double vi = v[i] * b2 + (1.0 - b2) * gi * gi;       // mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        m[i] = mi;
        v[i] = vi;
        double denom = sqrt(vi) / bc2_sqrt + eps;
        p[i] = p[i] - step_size * (mi / denom);
    }
}

static int grid_for(long long n, int threads, int per_sm = 8) {
    long long want = (n + threads - 1) / threads;
    long long cap = (long long)num_sms() * per_sm;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

}  // namespace egp

using namespace egp;

extern "C" {

const char *egp_last_error_string(void) { return egp::g_err; }
int egp_version(void) { return 100; }

int64_t egp_gae_work_bytes(int64_t n) {
    int64_t nt = (n + GAE_TILE - 1) / GAE_TILE;
    if (nt < 1) nt = 1;
    return nt * (int64_t)(5 * sizeof(double)) + 64;      /* agg[nt][2] (16-byte aligned), partial[nt][3], done */
}

static int64_t g_gae_onepass_min = 1ll << 16;   // samples (32 tiles); measured faster from config 2's 1.2 M samples (43 vs 48 us) upwards

int64_t egp_gae_set_onepass_min(int64_t n) {
    const int64_t old = g_gae_onepass_min;
    if (n >= 0) g_gae_onepass_min = n;
    return old;
}

int egp_gae_f64(const double *d_rewards, const double *d_masks, const double *d_values, double gamma, double tau,
                int64_t n, double *d_adv, double *d_ret, double *d_stats, void *d_work, void *stream) {
    if (n <= 0 || !d_rewards || !d_masks || !d_values || !d_adv || !d_ret || !d_stats || !d_work || ((uintptr_t)d_work & 15)) {
        set_error("egp_gae_f64: bad argument (null pointer, n <= 0 or d_work not 16-byte aligned)");
        return EGP_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    unsigned int nt = (unsigned int)((n + GAE_TILE - 1) / GAE_TILE);
    double *agg = (double *)d_work;                                  // [nt][2], then 16 bytes for `done`, then partial [nt][3]
    unsigned int *done = (unsigned int *)(agg + 2 * (size_t)nt);
    double *partial = agg + 2 * (size_t)nt + 2;
    const bool vec = (((uintptr_t)d_rewards | (uintptr_t)d_masks | (uintptr_t)d_values | (uintptr_t)d_adv | (uintptr_t)d_ret) & 15) == 0;
    if (n >= g_gae_onepass_min) {
        if (cudaMemsetAsync(agg, 0xff, (2 * (size_t)nt + 2) * sizeof(double), st) != cudaSuccess) {     // empty slots; done = 2^32 - 1
            set_error("egp_gae_f64: cudaMemsetAsync failed");
            return EGP_ECUDA;
        }
        static int ctas_per_sm = 0;
        if (!ctas_per_sm) {
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, gae_onepass_kernel, GAE_THREADS, 0) != cudaSuccess || ctas_per_sm < 1) {
                ctas_per_sm = 0;
                set_error("egp_gae_f64: occupancy query failed");
                return EGP_ECUDA;
            }
        }
        int dev = 0;
        cudaGetDevice(&dev);
        unsigned int grid = (unsigned int)(num_sms(dev) * ctas_per_sm);        // every CTA resident: spinning on a neighbour is safe
        if (grid > nt) grid = nt;
        gae_onepass_kernel<<<grid, GAE_THREADS, 0, st>>>(d_rewards, d_masks, d_values, gamma, tau, (long long)n, vec, nt, agg, done,
                                                       d_adv, d_ret, partial, d_stats);
        EGP_CHECK_LAUNCH("gae_onepass_kernel");
        return EGP_OK;
    }
    gae_agg_kernel<<<nt, GAE_THREADS, 0, st>>>(d_rewards, d_masks, d_values, gamma, tau, (long long)n, vec, agg);
    gae_apply_kernel<<<nt, GAE_THREADS, 0, st>>>(d_rewards, d_masks, d_values, gamma, tau, (long long)n, vec, nt, agg, d_adv,
                                                 d_ret, partial);
    gae_moments_kernel<<<1, 256, 0, st>>>(partial, nt, d_stats);
    EGP_CHECK_LAUNCH("gae kernels");
    return EGP_OK;
}

int egp_standardize_f64(double *d_x, int64_t n, const double *d_stats, void *stream) {
    if (n <= 0 || !d_x || !d_stats) { set_error("egp_standardize_f64: bad argument"); return EGP_EINVAL; }
    standardize_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(d_x, (long long)n, d_stats);
    EGP_CHECK_LAUNCH("standardize_kernel");
    return EGP_OK;
}

int egp_gauss_logp_f64(const double *d_mu, const double *d_actions, const double *d_log_std, int64_t n, int adim,
                       double *d_logp, void *stream) {
    if (n <= 0 || adim <= 0 || !d_mu || !d_actions || !d_log_std || !d_logp) {
        set_error("egp_gauss_logp_f64: bad argument");
        return EGP_EINVAL;
    }
    gauss_logp_kernel<<<grid_for(n * LPR, 256), 256, 0, (cudaStream_t)stream>>>(d_mu, d_actions, d_log_std, (long long)n,
                                                                              adim, d_logp);
    EGP_CHECK_LAUNCH("gauss_logp_kernel");
    return EGP_OK;
}

int egp_ppo_loss_grad_f64(const double *d_mu, const double *d_actions, const double *d_log_std, const double *d_adv,
                          const double *d_stats, const double *d_logp0, const double *d_exps, double clip_eps,
                          double inv_count, int64_t n, int adim, double *d_dmu, double *d_dlogstd, double *d_loss,
                          void *stream) {
    return egp::ppo_loss_grad_launch(d_mu, d_actions, d_log_std, d_adv, d_stats, const_cast<double *>(d_logp0), 0, d_exps, clip_eps,
                                     inv_count, n, adim, d_dmu, d_dlogstd, d_loss, stream);
}

}  // extern "C" (the launcher below has C++ linkage)

int egp::ppo_loss_grad_launch(const double *d_mu, const double *d_actions, const double *d_log_std, const double *d_adv,
                              const double *d_stats, double *d_logp0, int record, const double *d_exps, double clip_eps,
                              double inv_count, int64_t n, int adim, double *d_dmu, double *d_dlogstd, double *d_loss,
                              void *stream) {
    if (n <= 0 || adim <= 0 || adim > 4 * LPR || !d_mu || !d_actions || !d_log_std || !d_adv || !d_stats || !d_logp0 ||
        !d_exps || !d_dmu || !d_loss) {
        set_error("egp_ppo_loss_grad_f64: bad argument (adim must be <= %d)", 4 * LPR);
        return EGP_EINVAL;
    }
    ppo_loss_grad_kernel<<<grid_for(n * LPR, 256), 256, 0, (cudaStream_t)stream>>>(
        d_mu, d_actions, d_log_std, d_adv, d_stats, d_logp0, d_exps, clip_eps, inv_count, (long long)n, adim, d_dmu,
        d_dlogstd, d_loss, record);
    EGP_CHECK_LAUNCH("ppo_loss_grad_kernel");
    return EGP_OK;
}

extern "C" {

int egp_value_loss_grad_f64(const double *d_v, const double *d_ret, double inv_n, int64_t n, double *d_dv,
                            double *d_loss, void *stream) {
    if (n <= 0 || !d_v || !d_ret || !d_dv || !d_loss) { set_error("egp_value_loss_grad_f64: bad argument"); return EGP_EINVAL; }
    value_loss_grad_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(d_v, d_ret, inv_n, (long long)n, d_dv, d_loss);
    EGP_CHECK_LAUNCH("value_loss_grad_kernel");
    return EGP_OK;
}

int egp_bias_relu_f64(double *d_y, const double *d_b, int64_t n, int dim, void *stream) {
    if (n <= 0 || dim <= 0 || !d_y || !d_b) { set_error("egp_bias_relu_f64: bad argument"); return EGP_EINVAL; }
    bias_relu_kernel<<<grid_for(n * dim, 256), 256, 0, (cudaStream_t)stream>>>(d_y, d_b, (long long)n * dim, dim);
    EGP_CHECK_LAUNCH("bias_relu_kernel");
    return EGP_OK;
}

int egp_relu_bwd_f64(double *d_dy, const double *d_y, int64_t n, int dim, void *stream) {
    if (n <= 0 || dim <= 0 || !d_dy || !d_y) { set_error("egp_relu_bwd_f64: bad argument"); return EGP_EINVAL; }
    relu_bwd_kernel<<<grid_for(n * dim, 256), 256, 0, (cudaStream_t)stream>>>(d_dy, d_y, (long long)n * dim);
    EGP_CHECK_LAUNCH("relu_bwd_kernel");
    return EGP_OK;
}

int egp_colsum_f64(const double *d_x, int64_t n, int dim, double *d_out, void *stream) {
    if (n <= 0 || dim <= 0 || !d_x || !d_out) { set_error("egp_colsum_f64: bad argument"); return EGP_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    EGP_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * dim, st));
    int blocks = num_sms() * 4;
    long long rpb = (n + blocks - 1) / blocks;
    if (rpb < 8) rpb = 8;
    blocks = (int)((n + rpb - 1) / rpb);
    colsum_kernel<<<blocks, 256, 0, st>>>(d_x, (long long)n, dim, rpb, d_out);
    EGP_CHECK_LAUNCH("colsum_kernel");
    return EGP_OK;
}

int egp_gather_rows_f64(const double *d_in, const int64_t *d_perm, int64_t n, int dim, double *d_out, void *stream) {
    if (n <= 0 || dim <= 0 || !d_in || !d_perm || !d_out) { set_error("egp_gather_rows_f64: bad argument"); return EGP_EINVAL; }
    gather_rows_kernel<<<grid_for(n * dim, 256), 256, 0, (cudaStream_t)stream>>>(d_in, (const long long *)d_perm, (long long)n, dim, d_out);
    EGP_CHECK_LAUNCH("gather_rows_kernel");
    return EGP_OK;
}

int egp_relu_bwd_colsum_f64(double *d_dy, const double *d_y, int64_t n, int dim, double *d_out, void *stream) {
    if (n <= 0 || dim <= 0 || !d_dy || !d_y || !d_out) { set_error("egp_relu_bwd_colsum_f64: bad argument"); return EGP_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    EGP_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * dim, st));
    int blocks = num_sms() * 8;
    long long rpb = (n + blocks - 1) / blocks;
    if (rpb < 8) rpb = 8;
    blocks = (int)((n + rpb - 1) / rpb);
    relu_bwd_colsum_kernel<<<blocks, 256, 0, st>>>(d_dy, d_y, (long long)n, dim, rpb, d_out);
    EGP_CHECK_LAUNCH("relu_bwd_colsum_kernel");
    return EGP_OK;
}

int egp_col_moments_f64(const double *d_x, int64_t n, int dim, const double *d_shift, double *d_out, void *stream) {
    if (n <= 0 || dim <= 0 || !d_x || !d_out) { set_error("egp_col_moments_f64: bad argument"); return EGP_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    EGP_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * 2 * dim, st));
    int blocks = num_sms() * 4;
    long long rpb = (n + blocks - 1) / blocks;
    if (rpb < 8) rpb = 8;
    blocks = (int)((n + rpb - 1) / rpb);
    col_moments_kernel<<<blocks, 256, 0, st>>>(d_x, (long long)n, dim, rpb, d_shift, d_out);
    EGP_CHECK_LAUNCH("col_moments_kernel");
    return EGP_OK;
}

int egp_sumsq_f64(const double *d_g, int64_t n, double *d_norm2, void *stream) {
    if (n <= 0 || !d_g || !d_norm2) { set_error("egp_sumsq_f64: bad argument"); return EGP_EINVAL; }
    int blocks = (int)((n + 255) / 256);
    if (blocks > SUMSQ_BLOCKS) blocks = SUMSQ_BLOCKS;
    sumsq_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_g, (long long)n, d_norm2);
    EGP_CHECK_LAUNCH("sumsq_kernel");
    return EGP_OK;
}

int egp_adam_step_f64(double *d_p, const double *d_g, double *d_m, double *d_v, int64_t n, double lr, double beta1,
                      double beta2, double eps, int64_t step, double max_norm, const double *d_norm2, void *stream) {
    if (n <= 0 || step < 1 || !d_p || !d_g || !d_m || !d_v || (max_norm > 0.0 && !d_norm2)) {
        set_error("egp_adam_step_f64: bad argument");
        return EGP_EINVAL;
    }
    double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    adam_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(d_p, d_g, d_m, d_v, (long long)n, lr, beta1, beta2, eps,
                                                                   bc1, sqrt(bc2), max_norm, d_norm2);
    EGP_CHECK_LAUNCH("adam_kernel");
    return EGP_OK;
}

}  // extern "C"
