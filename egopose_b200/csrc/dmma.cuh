// FP64 tensor-core (mma.sync m8n8k4 f64, "DMMA") dense-layer tiles shared by the rollout kernel's policy MLP / state LSTM
// (csrc/rollout.cu) and the fused LSTM sequence kernels (csrc/lstm.cu).
//
// One dense layer for a CTA's 32 batch columns = [32 x K] . [K x N]: a DMMA issues 256 FMAs per instruction at the DFMA
// pipe's FLOP rate (measured: 16 cycles per DMMA and sub-partition, tools/micro/fp64_probe.cu), i.e. 8x fewer issue slots
// than a SIMT loop, and its operands are distributed over the lanes: an activation fragment is one conflict-free 256-byte
// shared load, a weight fragment one coalesced 256-byte global load of weights pre-packed in fragment order.
// Activations live feature-major [k][batch column] with a row stride of XS doubles: with XS = 36 the four k rows of a
// fragment fall on disjoint bank groups (stride 32 is a 4-way conflict).  Work item = 16 batch columns x 16 neurons
// (2 x 2 fragments); 2 * N/16 items are dealt round-robin to the warps of the CTA.
#pragma once
#include <stddef.h>

namespace egp {

constexpr int XS_WIDE = 36;                  // default row stride; 32 (4-way conflicts on the fragment loads) when 36 does not fit
constexpr int MLP_NT = 16;                   // neurons per work item; weights / biases are padded to multiples of 16

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// accumulates acc[mf][nf] += x[16 envs of half eh][k in 4 kc0 .. 4 kc1) . W[neuron tile nt][k]; Wf = [N/8][K4][32] fragments,
// K4 (a multiple of 4) = padded K / 4; xs row r holds k = row_k0 + r
__device__ __forceinline__ void t4_dmma_acc(double (&acc)[2][2][2], const double *__restrict__ Wf, int K4, int nt, int kc0, int kc1,
                                            int row_k0, const double *xs, int XS, int eh, int lane) {
    const double *w0 = Wf + ((size_t)(2 * nt) * K4) * 32 + lane, *w1 = w0 + (size_t)K4 * 32;
    const double *xa = xs + (ptrdiff_t)((lane & 3) - row_k0) * XS + eh * 16 + (lane >> 2);
    double bn[4][2];
#pragma unroll
    for (int u = 0; u < 4; u++) { bn[u][0] = __ldg(w0 + (size_t)(kc0 + u) * 32); bn[u][1] = __ldg(w1 + (size_t)(kc0 + u) * 32); }
    for (int g = kc0; g < kc1; g += 4) {
        double bc[4][2];
#pragma unroll
        for (int u = 0; u < 4; u++) { bc[u][0] = bn[u][0]; bc[u][1] = bn[u][1]; }
        if (g + 4 < kc1) {
#pragma unroll
            for (int u = 0; u < 4; u++) { bn[u][0] = __ldg(w0 + (size_t)(g + 4 + u) * 32); bn[u][1] = __ldg(w1 + (size_t)(g + 4 + u) * 32); }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const double a0 = xa[(size_t)(4 * (g + u)) * XS], a1 = xa[(size_t)(4 * (g + u)) * XS + 8];
            dmma884(acc[0][0], a0, bc[u][0]);
            dmma884(acc[0][1], a0, bc[u][1]);
            dmma884(acc[1][0], a1, bc[u][0]);
            dmma884(acc[1][1], a1, bc[u][1]);
        }
    }
}

__device__ __forceinline__ void t4_dmma_init(double (&acc)[2][2][2], const double *__restrict__ bias, int nt, int lane) {
#pragma unroll
    for (int nf = 0; nf < 2; nf++)
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const double bv = bias ? bias[nt * MLP_NT + nf * 8 + 2 * (lane & 3) + c] : 0.0;
            acc[0][nf][c] = bv; acc[1][nf][c] = bv;
        }
}

template <bool RELU>
__device__ __forceinline__ void t4_dmma_store(const double (&acc)[2][2][2], int row0, double *ys, int XS, int eh, int lane) {
#pragma unroll
    for (int mf = 0; mf < 2; mf++)
#pragma unroll
        for (int nf = 0; nf < 2; nf++)
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const double v = acc[mf][nf][c];
                ys[(size_t)(row0 + nf * 8 + 2 * (lane & 3) + c) * XS + eh * 16 + mf * 8 + (lane >> 2)] = RELU ? fmax(v, 0.0) : v;
            }
}

// neuron tiles [nt0, nt1) of one dense layer: ys[(16 (nt - nt0) + j) + out_row0][env] = act(bias + W x)
template <bool RELU, int NW = 8>
__device__ __forceinline__ void t4_mlp_layer(const double *__restrict__ Wf, const double *__restrict__ bias, int K4, int nt0, int nt1,
                                             int out_row0, const double *xs, double *ys, int XS, int lane, int w) {
    for (int t = w; t < 2 * (nt1 - nt0); t += NW) {
        const int eh = t & 1, nt = nt0 + (t >> 1);
        double acc[2][2][2];
        t4_dmma_init(acc, bias, nt, lane);
        t4_dmma_acc(acc, Wf, K4, nt, 0, K4, 0, xs, XS, eh, lane);
        t4_dmma_store<RELU>(acc, (nt - nt0) * MLP_NT + out_row0, ys, XS, eh, lane);
    }
}

// packs W [out][in] (or its transpose view: element (j, k) = W[k][j] with W [in][out], TRANSPOSED) into DMMA fragments
// Wf[outp / 8][K4][32]: element (nf, kc, lane) = W(8 nf + lane / 4, 4 kc + lane % 4), zero padded both ways; outp a multiple of
// 16, K4 a multiple of 4; biases padded (bp may be null)
template <bool TRANSPOSED>
__global__ void pack_frag_kernel_t(const double *__restrict__ W, const double *__restrict__ b, int out, int in, int outp, int K4,
                                   double *__restrict__ Wf, double *__restrict__ bp) {
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long total = (long long)(outp / 8) * K4 * 32;
    if (idx < total) {
        const int lane = (int)(idx & 31);
        const long long t = idx >> 5;
        const int kc = (int)(t % K4), nf = (int)(t / K4);
        const int j = 8 * nf + (lane >> 2), k = 4 * kc + (lane & 3);
        Wf[idx] = (j < out && k < in) ? (TRANSPOSED ? W[(size_t)k * out + j] : W[(size_t)j * in + k]) : 0.0;
    }
    if (bp && idx < outp) bp[idx] = (b && idx < out) ? b[idx] : 0.0;
}

}  // namespace egp
