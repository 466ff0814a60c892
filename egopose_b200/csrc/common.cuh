// Shared host/device helpers for the egopose_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/egopose_b200.h"

namespace egp {

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define EGP_CUDA(call)                                             \
    do {                                                           \
        cudaError_t _e = (call);                                   \
        if (_e != cudaSuccess) return egp::cuda_fail(_e, #call);   \
    } while (0)

#define EGP_CHECK_LAUNCH(name)                                     \
    do {                                                           \
        cudaError_t _e = cudaGetLastError();                       \
        if (_e != cudaSuccess) return egp::cuda_fail(_e, name);    \
    } while (0)

int num_sms(int device = -1);

// egp_ppo_loss_grad_f64 with an option: record = 1 writes logp of this pass to d_logp0 (the fixed_log_probs of
// agent_ppo.py:18-20) and uses ratio = 1, i.e. the first epoch's pass doubles as the fixed-log-prob pass
int ppo_loss_grad_launch(const double *d_mu, const double *d_actions, const double *d_log_std, const double *d_adv,
                         const double *d_stats, double *d_logp0, int record, const double *d_exps, double clip_eps,
                         double inv_count, int64_t n, int adim, double *d_dmu, double *d_dlogstd, double *d_loss,
                         void *stream);

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace egp
