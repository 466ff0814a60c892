// Micro-benchmark: latency of tcgen05.ld / tcgen05.st (+ their waits) used as per-thread scratch, one CTA, 4 warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_probe.cu && ./tmem_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void probe(long long *out, int iters, int nwarps_active) {
    __shared__ uint32_t s_base;
    const int w = threadIdx.x >> 5;
    if (w == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = s_base + ((uint32_t)(w * 32) << 16);
    int r[8];
    long long t[8] = {0};
    if (w < nwarps_active) {
        // (0) ld x8 + wait, dependent chain
        long long t0 = clock64();
        int acc = 0;
        for (int i = 0; i < iters; i++) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\ttcgen05.wait::ld.sync.aligned;"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(tm + ((acc & 1) << 3)));
            acc += r[0];
        }
        long long t1 = clock64();
        t[0] = t1 - t0;
        // (1) wait::ld alone (nothing outstanding)
        t0 = clock64();
        for (int i = 0; i < iters; i++) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        t1 = clock64();
        t[1] = t1 - t0;
        // (2) st x8 (no wait)
        t0 = clock64();
        for (int i = 0; i < iters; i++)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(tm + 16), "r"(i), "r"(i), "r"(i), "r"(i), "r"(i), "r"(i), "r"(i), "r"(i));
        t1 = clock64();
        t[2] = t1 - t0;
        // (3) st x8 + wait::st
        t0 = clock64();
        for (int i = 0; i < iters; i++) {
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(tm + 16), "r"(i), "r"(i), "r"(i), "r"(i), "r"(i), "r"(i), "r"(i), "r"(i));
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        t1 = clock64();
        t[3] = t1 - t0;
        // (4) ld x32 + wait
        int q[32];
        t0 = clock64();
        for (int i = 0; i < iters; i++) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                         "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\ttcgen05.wait::ld.sync.aligned;"
                         : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]), "=r"(q[9]),
                           "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]), "=r"(q[16]), "=r"(q[17]), "=r"(q[18]),
                           "=r"(q[19]), "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]), "=r"(q[24]), "=r"(q[25]), "=r"(q[26]), "=r"(q[27]),
                           "=r"(q[28]), "=r"(q[29]), "=r"(q[30]), "=r"(q[31]) : "r"(tm + 64 + ((acc & 1) << 3)));
            acc += q[0] + q[31];
        }
        t1 = clock64();
        t[4] = t1 - t0;
        // (5) ld x8 issue, 40 dependent DFMAs, then wait (overlap test)
        double d = 1.0 + acc * 1e-30;
        t0 = clock64();
        for (int i = 0; i < iters; i++) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(tm));
#pragma unroll
            for (int k = 0; k < 40; k++) d = fma(d, 1.0000001, 1e-9);
            asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]));
            acc += r[0];
        }
        t1 = clock64();
        t[5] = t1 - t0;
        // (6) the 40 dependent DFMAs alone
        t0 = clock64();
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int k = 0; k < 40; k++) d = fma(d, 1.0000001, 1e-9);
        }
        t1 = clock64();
        t[6] = t1 - t0;
        if (threadIdx.x == 0) { for (int k = 0; k < 7; k++) out[k] = t[k]; out[7] = acc + (long long)d; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_base), "r"(512));
}

int main() {
    long long *d, h[8];
    cudaMalloc(&d, 64);
    const int iters = 2000;
    for (int nw = 1; nw <= 4; nw *= 2) {
        probe<<<1, 128>>>(d, iters, nw);
        cudaDeviceSynchronize();
        cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
        printf("%d warps active: ld.x8+wait %.1f | wait::ld alone %.1f | st.x8 %.1f | st.x8+wait::st %.1f | ld.x32+wait %.1f | ld.x8 + 40 dep DFMA + wait %.1f | 40 dep DFMA %.1f  cycles (%s)\n",
               nw, h[0] / (double)iters, h[1] / (double)iters, h[2] / (double)iters, h[3] / (double)iters, h[4] / (double)iters, h[5] / (double)iters,
               h[6] / (double)iters, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
