// Fused LSTM sequence recurrence, forward and backward, float64, sm_100a ("fused BiLSTM fwd/bwd").
//
// Replaces the per-time-step LSTMCell loop of models/rnn.py:45-61 (RNN.batch_forward, both directions of the BiLSTM in
// models/video_state_net.py:36-70) and the padded state-LSTM unroll of models/video_forecast_net.py:95-111 in the PPO update,
// where the reference re-runs them on every forward (21 forward + 20 backward passes per iteration).
//
// Split of the work: everything that is NOT recurrent is a plain GEMM and stays one (cuBLAS through torch): the input projection
// xi = x W_ih^T + b_ih + b_hh of ALL time steps, and in the backward pass dW_ih = dG^T X, db = colsum dG, dW_hh = dG^T H_prev.
// These kernels do the sequential part in ONE launch per sequence sweep:
//   forward   gates_t = xi_t + h_{t-1} W_hh^T ; i, f, o = sigmoid, g = tanh ; c_t = f c_{t-1} + i g ; h_t = o tanh c_t
//   backward  dG_t from (dh_t + dh_rec, dc_rec, saved gates / cell states) ; dh_rec = dG_t W_hh ; dc_rec = dc f
// Layout: "time-major packed" rows (like a PackedSequence): step s owns rows [off[s], off[s+1]) = the batch elements alive at
// that step, off non-decreasing with non-increasing counts (ragged episodes sorted longest first; a dense [L, B] sequence has
// off[s] = s B).  A CTA owns 32 batch elements for the whole sequence: h / dh live feature-major in shared memory as the
// operand tile of the FP64 tensor-core products (csrc/dmma.cuh, W_hh pre-packed in fragment order, 2 x 2 fragment work items
// dealt to 8 warps), cell states / recurrent cell gradients in registers, gate pre-activations pass through a shared tile.
#include "common.cuh"
#include "dmma.cuh"

namespace egp {

constexpr int LS_WARPS = 8, LS_THREADS = LS_WARPS * 32, LS_XS = 36;

struct LstmArgs {
    const double *xi;           // [Np, 4H] input projections (+ biases), gate order i | f | g | o like torch.nn.LSTMCell
    const long long *off;       // [L + 1] packed row offsets per step (iteration order)
    const double *Wf;           // W_hh fragments: forward [4H / 8][H / 4][32] (out = 4H, K = H); backward [H / 8][4H / 4][32]
    double *h, *gates, *c;      // forward outputs [Np, H], [Np, 4H] (post-activation), [Np, H]
    const double *dh;           // backward: upstream gradient of h [Np, H]
    double *dxi;                // backward output [Np, 4H] = gradient of the gate pre-activations
    int L, B;
};

__device__ __forceinline__ double sigmoid(double x) { return 1.0 / (1.0 + exp(-x)); }

template <int H>
__global__ void __launch_bounds__(LS_THREADS, 1) lstm_fwd_kernel(const LstmArgs A) {
    extern __shared__ double sm[];
    double *hT = sm;                                    // [H][XS]
    double *gT = sm + (size_t)H * LS_XS;                // [4H][XS]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long j0 = (long long)blockIdx.x * 32;
    constexpr int PER = H * 32 / LS_THREADS;            // (unit, column) pairs per thread: unit = lane + 32 * (k % (H / 32)), ...
    double c[PER];
#pragma unroll
    for (int k = 0; k < PER; k++) c[k] = 0.0;
    for (int k = threadIdx.x; k < H * LS_XS; k += LS_THREADS) hT[k] = 0.0;
    __syncthreads();
    for (int s = 0; s < A.L; s++) {
        const long long o0 = A.off[s], na = A.off[s + 1] - o0;
        if (j0 >= na) break;                            // counts are non-increasing: this tile is done
        if (s > 0) t4_mlp_layer<false, LS_WARPS>(A.Wf, nullptr, H / 4, 0, 4 * H / MLP_NT, 0, hT, gT, LS_XS, lane, w);
        __syncthreads();
        // elementwise: pair p = threadIdx.x + LS_THREADS * k  ->  unit u = p % H (lanes run over units: coalesced rows), column p / H
#pragma unroll
        for (int k = 0; k < PER; k++) {
            const int p = threadIdx.x + LS_THREADS * k, u = p % H, col = p / H;
            if (j0 + col < na) {
                const size_t r = (size_t)(o0 + j0 + col);
                const double *x4 = A.xi + r * 4 * H;
                double gi = x4[u], gf = x4[H + u], gg = x4[2 * H + u], go = x4[3 * H + u];
                if (s > 0) {
                    gi += gT[(size_t)u * LS_XS + col]; gf += gT[(size_t)(H + u) * LS_XS + col];
                    gg += gT[(size_t)(2 * H + u) * LS_XS + col]; go += gT[(size_t)(3 * H + u) * LS_XS + col];
                }
                const double i = sigmoid(gi), f = sigmoid(gf), g = tanh(gg), o = sigmoid(go);
                const double cn = f * c[k] + i * g, hn = o * tanh(cn);
                c[k] = cn;
                double *g4 = A.gates + r * 4 * H;
                g4[u] = i; g4[H + u] = f; g4[2 * H + u] = g; g4[3 * H + u] = o;
                A.c[r * H + u] = cn;
                A.h[r * H + u] = hn;
                hT[(size_t)u * LS_XS + col] = hn;
            }
        }
        __syncthreads();
    }
}

template <int H>
__global__ void __launch_bounds__(LS_THREADS, 1) lstm_bwd_kernel(const LstmArgs A) {
    extern __shared__ double sm[];
    double *dhT = sm;                                   // [H][XS] recurrent gradient dG_{s+1} W_hh
    double *dgT = sm + (size_t)H * LS_XS;               // [4H][XS]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long j0 = (long long)blockIdx.x * 32;
    constexpr int PER = H * 32 / LS_THREADS;
    double dc[PER];
#pragma unroll
    for (int k = 0; k < PER; k++) dc[k] = 0.0;
    for (int k = threadIdx.x; k < 5 * H * LS_XS; k += LS_THREADS) sm[k] = 0.0;
    __syncthreads();
    // first step (in iteration order) at which this tile is not alive any more
    int s_end = 0;
    while (s_end < A.L && j0 < A.off[s_end + 1] - A.off[s_end]) s_end++;
    for (int s = s_end - 1; s >= 0; s--) {
        const long long o0 = A.off[s], na = A.off[s + 1] - o0;
        const long long op = s > 0 ? A.off[s - 1] : 0;
#pragma unroll
        for (int k = 0; k < PER; k++) {
            const int p = threadIdx.x + LS_THREADS * k, u = p % H, col = p / H;
            double di = 0.0, df = 0.0, dg = 0.0, dob = 0.0;
            if (j0 + col < na) {
                const size_t r = (size_t)(o0 + j0 + col);
                const double *g4 = A.gates + r * 4 * H;
                const double i = g4[u], f = g4[H + u], g = g4[2 * H + u], o = g4[3 * H + u];
                const double ct = A.c[r * H + u], cp = s > 0 ? A.c[(size_t)(op + j0 + col) * H + u] : 0.0;
                const double tc = tanh(ct);
                const double dh = A.dh[r * H + u] + dhT[(size_t)u * LS_XS + col];
                const double dct = dc[k] + dh * o * (1.0 - tc * tc);
                dob = dh * tc * o * (1.0 - o);
                di = dct * g * i * (1.0 - i);
                dg = dct * i * (1.0 - g * g);
                df = dct * cp * f * (1.0 - f);
                dc[k] = dct * f;
                double *d4 = A.dxi + r * 4 * H;
                d4[u] = di; d4[H + u] = df; d4[2 * H + u] = dg; d4[3 * H + u] = dob;
            }
            dgT[(size_t)u * LS_XS + col] = di; dgT[(size_t)(H + u) * LS_XS + col] = df;
            dgT[(size_t)(2 * H + u) * LS_XS + col] = dg; dgT[(size_t)(3 * H + u) * LS_XS + col] = dob;
        }
        __syncthreads();
        if (s > 0) t4_mlp_layer<false, LS_WARPS>(A.Wf, nullptr, 4 * H / 4, 0, H / MLP_NT, 0, dgT, dhT, LS_XS, lane, w);
        __syncthreads();
    }
}

static int lstm_check(int H, int L, long long B) {
    if (H != 64 && H != 128) { set_error("egp_lstm: hidden size %d not supported by the fused kernels (64 or 128)", H); return EGP_ESIZE; }
    if (L < 1 || B < 1) { set_error("egp_lstm: bad sizes"); return EGP_EINVAL; }
    return EGP_OK;
}

}  // namespace egp

using namespace egp;

extern "C" {

int64_t egp_lstm_wfrag_elems(int H) { return (int64_t)4 * H * H; }

int egp_lstm_pack_whh_f64(const double *d_Whh, int H, double *d_Wf_fwd, double *d_Wf_bwd, void *stream) {
    if (!d_Whh || !d_Wf_fwd) { set_error("egp_lstm_pack_whh_f64: null argument"); return EGP_EINVAL; }
    int rc = lstm_check(H, 1, 1);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)4 * H * H;
    // forward: gates = h W_hh^T, out = 4H rows of W_hh [4H][H], K = H
    pack_frag_kernel_t<false><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_Whh, nullptr, 4 * H, H, 4 * H, H / 4, d_Wf_fwd, nullptr);
    // backward: dh = dG W_hh, out = H, K = 4H, weight(j, k) = W_hh[k][j]
    if (d_Wf_bwd)
        pack_frag_kernel_t<true><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_Whh, nullptr, H, 4 * H, H, 4 * H / 4, d_Wf_bwd, nullptr);
    EGP_CHECK_LAUNCH("pack_frag_kernel_t");
    return EGP_OK;
}

int egp_lstm_seq_fwd_f64(const double *d_xi, const int64_t *d_off, int L, int64_t B, int H, const double *d_Wf_fwd, double *d_h,
                         double *d_gates, double *d_c, void *stream) {
    if (!d_xi || !d_off || !d_Wf_fwd || !d_h || !d_gates || !d_c) { set_error("egp_lstm_seq_fwd_f64: null argument"); return EGP_EINVAL; }
    int rc = lstm_check(H, L, B);
    if (rc) return rc;
    LstmArgs A;
    memset(&A, 0, sizeof A);
    A.xi = d_xi; A.off = (const long long *)d_off; A.Wf = d_Wf_fwd; A.h = d_h; A.gates = d_gates; A.c = d_c; A.L = L; A.B = (int)B;
    const unsigned blocks = (unsigned)((B + 31) / 32);
    const size_t smem = sizeof(double) * 5 * H * LS_XS;
    cudaStream_t st = (cudaStream_t)stream;
    if (H == 64) {
        EGP_CUDA(cudaFuncSetAttribute(lstm_fwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lstm_fwd_kernel<64><<<blocks, LS_THREADS, smem, st>>>(A);
    } else {
        EGP_CUDA(cudaFuncSetAttribute(lstm_fwd_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lstm_fwd_kernel<128><<<blocks, LS_THREADS, smem, st>>>(A);
    }
    EGP_CHECK_LAUNCH("lstm_fwd_kernel");
    return EGP_OK;
}

int egp_lstm_seq_bwd_f64(const double *d_dh, const double *d_gates, const double *d_c, const int64_t *d_off, int L, int64_t B, int H,
                         const double *d_Wf_bwd, double *d_dxi, void *stream) {
    if (!d_dh || !d_gates || !d_c || !d_off || !d_Wf_bwd || !d_dxi) { set_error("egp_lstm_seq_bwd_f64: null argument"); return EGP_EINVAL; }
    int rc = lstm_check(H, L, B);
    if (rc) return rc;
    LstmArgs A;
    memset(&A, 0, sizeof A);
    A.off = (const long long *)d_off; A.Wf = d_Wf_bwd; A.gates = const_cast<double *>(d_gates); A.c = const_cast<double *>(d_c);
    A.dh = d_dh; A.dxi = d_dxi; A.L = L; A.B = (int)B;
    const unsigned blocks = (unsigned)((B + 31) / 32);
    const size_t smem = sizeof(double) * 5 * H * LS_XS;
    cudaStream_t st = (cudaStream_t)stream;
    if (H == 64) {
        EGP_CUDA(cudaFuncSetAttribute(lstm_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lstm_bwd_kernel<64><<<blocks, LS_THREADS, smem, st>>>(A);
    } else {
        EGP_CUDA(cudaFuncSetAttribute(lstm_bwd_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lstm_bwd_kernel<128><<<blocks, LS_THREADS, smem, st>>>(A);
    }
    EGP_CHECK_LAUNCH("lstm_bwd_kernel");
    return EGP_OK;
}

}  // extern "C"
