// Gradient all-reduce over NVLink / NVSwitch peer memory (sm_100a), one launch per PPO epoch.
//
// The data-parallel exchange of the update (agents/agent_ppo.py:44-56 differentiates one global mini-batch, so
// rank-local gradients are summed) is ONE flat [value | policy] float64 buffer of ~2.2 MB per epoch: latency, not
// bandwidth.  Every rank owns an exchange block (cudaMalloc + CUDA IPC, mapped by all peers of the node):
//     src [n] doubles   this rank's gradient (the flat gradient buffer of the nets lives here: no staging copy)
//     out [n] doubles   the sum, identical bits on every rank
//     flags [2][8]      64-bit epoch counters written by the peers (phase 0: "my src is complete", 1: "I have read yours")
// egp_allreduce_grads_f64 launches one kernel per rank:
//   1. signal phase 0 into every peer's flags (stream order already guarantees that the kernels which wrote src are done),
//      wait until all peers have signalled;
//   2. out[i] = src_0[i] + src_1[i] + ... in RANK ORDER on every rank (peer loads over NVLink): bit-identical replicas
//      without a broadcast;
//   3. the last CTA to finish signals phase 1 and waits for the peers' phase 1: when the kernel has completed nobody is
//      reading this rank's src any more, so the next backward pass may overwrite it.
// No NCCL call, no host synchronisation; the spin loops give up after ~4 s (error word) instead of hanging the GPU if a
// rank never arrives.
#include "common.cuh"

#include <string.h>

namespace egp {

constexpr int P2P_MAX_RANKS = 8;
constexpr int P2P_FLAG_WORDS = 2 * P2P_MAX_RANKS + 8;      // flags [2][8] | done counter | error word | pad

struct P2pBlock {                    // layout of one rank's exchange allocation
    unsigned long long flags[2][P2P_MAX_RANKS];
    unsigned int done, error, pad[30];
};
static_assert(sizeof(P2pBlock) == 256, "exchange block header: 256 bytes keeps src / out 256-byte aligned");

struct P2pPeers {
    double *src[P2P_MAX_RANKS];
    P2pBlock *blk[P2P_MAX_RANKS];
};

__device__ __forceinline__ void st_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ long long c_p2p_timeout = 8000000000ll;              // cycles a barrier waits for a peer: ~4 s at 2 GHz

// all ranks have written `epoch` into my flags[phase]; false on timeout
__device__ bool p2p_wait(const P2pBlock *mine, int phase, int world, unsigned long long epoch) {
    const long long t0 = clock64();
    for (int p = 0; p < world; p++) {
        while (ld_sys(&mine->flags[phase][p]) < epoch) {
            if (clock64() - t0 > c_p2p_timeout) return false;
            __nanosleep(100);
        }
    }
    return true;
}

__global__ void __launch_bounds__(256)
p2p_allreduce_kernel(P2pPeers peers, int rank, int world, long long n, unsigned long long epoch, double *__restrict__ out) {
    P2pBlock *mine = peers.blk[rank];
    __shared__ int s_ok;
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) {
            __threadfence_system();
            for (int p = 0; p < world; p++) st_sys(&peers.blk[p]->flags[0][rank], epoch);
        }
        s_ok = p2p_wait(mine, 0, world, epoch) ? 1 : 0;
    }
    __syncthreads();
    if (s_ok) {
        const long long n2 = n >> 1;
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
            double2 acc = __ldcg(reinterpret_cast<const double2 *>(peers.src[0]) + i);
            for (int p = 1; p < world; p++) {
                const double2 v = __ldcg(reinterpret_cast<const double2 *>(peers.src[p]) + i);
                acc.x += v.x; acc.y += v.y;
            }
            reinterpret_cast<double2 *>(out)[i] = acc;
        }
        if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
            double acc = __ldcg(peers.src[0] + n - 1);
            for (int p = 1; p < world; p++) acc += __ldcg(peers.src[p] + n - 1);
            out[n - 1] = acc;
        }
    } else if (threadIdx.x == 0) atomicOr(&mine->error, 1u);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&mine->done, 1u) == gridDim.x - 1) {          // last CTA of this rank
            mine->done = 0u;
            __threadfence_system();
            for (int p = 0; p < world; p++) st_sys(&peers.blk[p]->flags[1][rank], epoch);
            if (!p2p_wait(mine, 1, world, epoch)) atomicOr(&mine->error, 2u);
        }
    }
}

}  // namespace egp

using namespace egp;

struct EgpComm {
    int rank, world, device;
    long long n;
    unsigned long long epoch;
    void *base;                      // this rank's allocation: P2pBlock | src [n] | out [n]
    void *peer_base[P2P_MAX_RANKS];  // mapped peer allocations (own entry = base)
    P2pPeers peers;
    double *out;
};

extern "C" {

int64_t egp_comm_handle_bytes(void) { return (int64_t)sizeof(cudaIpcMemHandle_t); }

int egp_comm_create(int rank, int world, int device, int64_t n, EgpComm **out, void *handle_out) {
    if (!out || !handle_out || world < 1 || world > P2P_MAX_RANKS || rank < 0 || rank >= world || n < 1) {
        set_error("egp_comm_create: bad argument (1 <= world <= %d, 0 <= rank < world, n >= 1)", P2P_MAX_RANKS);
        return EGP_EINVAL;
    }
    EGP_CUDA(cudaSetDevice(device));
    EgpComm *c = new EgpComm();
    memset(c, 0, sizeof(*c));
    c->rank = rank; c->world = world; c->device = device; c->n = n;
    const size_t n_pad = ((size_t)n + 31) & ~(size_t)31;          // out starts 256-byte aligned like src
    const size_t bytes = sizeof(P2pBlock) + 2 * n_pad * sizeof(double);
    cudaError_t e = cudaMalloc(&c->base, bytes);
    if (e != cudaSuccess) { delete c; return cuda_fail(e, "cudaMalloc (exchange block)"); }
    e = cudaMemset(c->base, 0, bytes);
    if (e == cudaSuccess) e = cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle_out, c->base);
    if (e != cudaSuccess) { cudaFree(c->base); delete c; return cuda_fail(e, "cudaIpcGetMemHandle"); }
    c->out = (double *)((char *)c->base + sizeof(P2pBlock)) + n_pad;
    *out = c;
    return EGP_OK;
}

int egp_comm_connect(EgpComm *c, const void *all_handles) {
    if (!c || !all_handles) { set_error("egp_comm_connect: null argument"); return EGP_EINVAL; }
    EGP_CUDA(cudaSetDevice(c->device));
    const cudaIpcMemHandle_t *h = (const cudaIpcMemHandle_t *)all_handles;
    for (int p = 0; p < c->world; p++) {
        if (p == c->rank) c->peer_base[p] = c->base;
        else {
            cudaError_t e = cudaIpcOpenMemHandle(&c->peer_base[p], h[p], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                set_error("egp_comm_connect: cudaIpcOpenMemHandle of rank %d failed: %s (peer access between the GPUs of one "
                          "node is required)", p, cudaGetErrorString(e));
                return EGP_ECUDA;
            }
        }
        c->peers.blk[p] = (P2pBlock *)c->peer_base[p];
        c->peers.src[p] = (double *)((char *)c->peer_base[p] + sizeof(P2pBlock));
    }
    return EGP_OK;
}

/* Same wiring for several communicators that live in ONE process (one process driving several GPUs, or the in-process
 * tests): CUDA IPC handles cannot be opened by the process that exported them, the raw pointers are used instead. */
int egp_comm_connect_local(EgpComm *c, EgpComm *const *all) {
    if (!c || !all) { set_error("egp_comm_connect_local: null argument"); return EGP_EINVAL; }
    for (int p = 0; p < c->world; p++) {
        if (!all[p] || all[p]->world != c->world || all[p]->rank != p || all[p]->n != c->n) {
            set_error("egp_comm_connect_local: entry %d is not rank %d of the same exchange", p, p);
            return EGP_EINVAL;
        }
        if (all[p]->device != c->device) {
            EGP_CUDA(cudaSetDevice(c->device));
            cudaError_t e = cudaDeviceEnablePeerAccess(all[p]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "cudaDeviceEnablePeerAccess");
            cudaGetLastError();
        }
        c->peer_base[p] = p == c->rank ? c->base : nullptr;         // nothing to close for local peers
        c->peers.blk[p] = (P2pBlock *)all[p]->base;
        c->peers.src[p] = (double *)((char *)all[p]->base + sizeof(P2pBlock));
    }
    return EGP_OK;
}

double *egp_comm_src(EgpComm *c) { return c ? (double *)((char *)c->base + sizeof(P2pBlock)) : nullptr; }
double *egp_comm_out(EgpComm *c) { return c ? c->out : nullptr; }

int egp_allreduce_grads_f64(EgpComm *c, int64_t n, void *stream) {
    if (!c || n < 1 || n > c->n) { set_error("egp_allreduce_grads_f64: bad argument"); return EGP_EINVAL; }
    if (!c->peers.blk[c->world - 1] || !c->peers.blk[0]) { set_error("egp_allreduce_grads_f64: egp_comm_connect has not been called"); return EGP_EINVAL; }
    c->epoch++;
    int grid = (int)((n / 2 + 255) / 256);
    const int cap = 2 * num_sms(c->device);
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    p2p_allreduce_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(c->peers, c->rank, c->world, (long long)n, c->epoch, c->out);
    EGP_CHECK_LAUNCH("p2p_allreduce_kernel");
    return EGP_OK;
}

/* barrier timeout in SM cycles (default 8e9, about 4 s); returns the previous value; cycles <= 0 only queries */
int64_t egp_comm_set_timeout_cycles(int64_t cycles) {
    long long old = 0;
    if (cudaMemcpyFromSymbol(&old, c_p2p_timeout, sizeof old) != cudaSuccess) return -1;
    if (cycles > 0) {
        long long v = cycles;
        if (cudaMemcpyToSymbol(c_p2p_timeout, &v, sizeof v) != cudaSuccess) return -1;
    }
    return old;
}

/* 0 = no rank timed out so far; bit 0: a peer's gradient never became ready, bit 1: a peer never finished reading
 * (reads the error word: synchronises the device) */
int egp_comm_error(EgpComm *c) {
    if (!c) return -1;
    P2pBlock b;
    if (cudaMemcpy(&b, c->base, sizeof b, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (int)b.error;
}

void egp_comm_destroy(EgpComm *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (int p = 0; p < c->world; p++)
        if (p != c->rank && c->peer_base[p]) cudaIpcCloseMemHandle(c->peer_base[p]);
    cudaFree(c->base);
    delete c;
}

}  // extern "C"
