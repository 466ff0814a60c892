// Micro-benchmark: FP64 issue / latency on this GPU (one CTA per SM, clock64 around unrolled chains).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_probe fp64_probe.cu && ./fp64_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void dfma_chain(double *out, long long *cyc, int iters) {
    double acc[ILP];
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x * 1e-3 + i;
    double m = 1.0000001, a = 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], m, a);
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < ILP; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int ILP>
__global__ void dmma_chain(double *out, long long *cyc, int iters) {
    double c[ILP][2];
    for (int i = 0; i < ILP; i++) { c[i][0] = threadIdx.x * 1e-3; c[i][1] = i; }
    double a = 1e-3 * threadIdx.x, b = 1e-3;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < ILP; i++) dmma(c[i][0], c[i][1], a, b);
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <typename K>
void run(const char *name, K kern, int threads, int ilp, double flop_per_inst) {
    double *out; long long *cyc, h;
    cudaMalloc(&out, sizeof(double) * 148 * 1024); cudaMalloc(&cyc, 8);
    int iters = 2000;
    kern<<<148, threads>>>(out, cyc, iters);
    kern<<<148, threads>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per_inst = (double)h / (iters * 8.0 * ilp);
    printf("%-28s threads %4d ilp %2d: %.2f cycles per warp-instruction per warp (%s)\n", name, threads, ilp, per_inst, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run("dfma dependent", dfma_chain<1>, 32, 1, 64);
    run("dfma ilp2", dfma_chain<2>, 32, 2, 64);
    run("dfma ilp4", dfma_chain<4>, 32, 4, 64);
    run("dfma ilp8", dfma_chain<8>, 32, 8, 64);
    run("dfma ilp8 4 warps", dfma_chain<8>, 128, 8, 64);
    run("dfma ilp8 8 warps", dfma_chain<8>, 256, 8, 64);
    run("dfma ilp1 16 warps", dfma_chain<1>, 512, 1, 64);
    run("dmma dependent", dmma_chain<1>, 32, 1, 512);
    run("dmma ilp2", dmma_chain<2>, 32, 2, 512);
    run("dmma ilp4", dmma_chain<4>, 32, 4, 512);
    run("dmma ilp8", dmma_chain<8>, 32, 8, 512);
    run("dmma ilp4 4 warps", dmma_chain<4>, 128, 4, 512);
    run("dmma ilp4 8 warps", dmma_chain<4>, 256, 4, 512);
    run("dmma ilp4 16 warps", dmma_chain<4>, 512, 4, 512);
    return 0;
}
