// TEST INFRASTRUCTURE ONLY: host build of the rollout kernel's tree sweeps (egopose_b200/csrc/tree.cuh) for ONE
// environment, with plain arrays standing in for the shared-memory rows and the Tensor Memory scratch and the four
// chain warps of the kernel run one after the other inside every level.  tests/test_host_sweeps.py compares it with
// the C oracle (dense CRBA + Cholesky) so that the block articulated-body arithmetic is checked on the CPU; the
// product path never loads this library.
#define _GNU_SOURCE 1
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../../egopose_b200/csrc/tree.cuh"

namespace egp {

struct HostCtx {
    static constexpr bool CONS = false;
    double *sm;         // [rows]
    double *tmem;       // this warp's per-thread scratch, 256 doubles
    int w;
    T4Off o;
    double &at(int off, int idx) const { return sm[off + idx]; }
    template <int NR> void tld_issue(int col, int (&r)[NR]) const { memcpy(r, tmem + col / 2, 4 * NR); }
    template <int NR> void tld_wait(int (&)[NR]) const {}
    static double unpack(const int *r, int k) { double v; memcpy(&v, r + 2 * k, 8); return v; }
    template <int N> void tst(int col, const double *v) const { for (int k = 0; k < N; k++) tmem[col / 2 + k] = v[k]; }
    void twait_st() const {}
};

static double g_sm[4096], g_tm[T4_CW][256 + 8];
static HostCtx g_x[T4_CW];

template <int MODE> static void fwd() {
    for (int L = 0; L < h_m.nlevel; L++)
        for (int w = 0; w < T4_CW; w++) {
            const int c = h_m.lvl_chain[L][w];
            if (c >= 0) t5_fwd_chain<MODE>(g_x[w], c);
        }
}
static void bwd(int mode) {
    for (int L = h_m.nlevel - 1; L >= 0; L--) {
        Bwd W[T4_CW];
        for (int w = 0; w < T4_CW; w++) t5_bwd_gather(g_x[w], h_m.lvl_chain[L][w], W[w]);
        for (int w = 0; w < T4_CW; w++) {
            const int c = h_m.lvl_chain[L][w];
            if (c >= 0) t5_bwd_chain(g_x[w], c, W[w], mode);
        }
    }
}

}  // namespace egp

using namespace egp;

static int dof_warp(int i) {
    for (int b = 0; b < h_m.nbody; b++)
        if (i >= h_m.body_dofadr[b] && i < h_m.body_dofadr[b] + h_m.body_dofnum[b]) return h_m.chain_warp[h_m.body_chain[b]];
    return 0;
}

extern "C" {

// returns t4_ok (1 = the block sweeps support this model), negative on error
int hs_init(const EgpModelDesc *s) {
    const char *why = nullptr;
    int rc = fill_dev_model(s, h_m, &why);
    if (rc != EGP_OK) return rc;
    T4Off O = t4_offsets(h_m);
    if (O.total > 4096) return -100;
    for (int w = 0; w < T4_CW; w++) { g_x[w].sm = g_sm; g_x[w].tmem = g_tm[w]; g_x[w].w = w; g_x[w].o = O; }
    return h_m.t4_ok;
}

// mirrors egp_env_step_debug_f64: sim.forward() at (qpos, qvel), then frame_skip stable-PD sub-steps with
// ctrl = a_ref + action a_scale; qpos / qvel updated in place; torque0 = clipped torque of the first sub-step;
// head_z = head height of the last kinematics refresh (stale by one sub-step); bias = qfrc_bias
// after the reset forward pass
void hs_env_step(double *qpos, double *qvel, const double *action, double *torque0, double *head_z, double *bias0) {
    const T4Off &O = g_x[0].o;
    memset(g_sm, 0, sizeof g_sm);
    memset(g_tm, 0, sizeof g_tm);
    for (int k = 0; k < h_m.nq; k++) g_sm[O.q + k] = qpos[k];
    for (int k = 0; k < h_m.nv; k++) g_sm[O.v + k] = qvel[k];
    fwd<2>();
    for (int i = 6; i < h_m.nv; i++) g_tm[dof_warp(i)][h_m.dof_col_tau[i] / 2] = 0.0;
    bwd(0);
    if (bias0) for (int i = 0; i < h_m.nv; i++) bias0[i] = g_tm[dof_warp(i)][h_m.dof_col_c[i] / 2];
    for (int i = 6; i < h_m.nv; i++)
        g_tm[dof_warp(i)][h_m.dof_col_ctrl[i] / 2] = h_m.a_ref[i] + action[i - 6] * h_m.a_scale[i];
    for (int s = 0; s < h_m.frame_skip; s++) {
        bwd(1);
        fwd<1>();
        if (s == 0 && torque0)
            for (int i = 6; i < h_m.nv; i++) torque0[i - 6] = g_tm[dof_warp(i)][h_m.dof_col_tau[i] / 2];
        bwd(0);
        fwd<0>();
    }
    for (int k = 0; k < h_m.nq; k++) qpos[k] = g_sm[O.q + k];
    for (int k = 0; k < h_m.nv; k++) qvel[k] = g_sm[O.v + k];
    if (head_z) *head_z = g_sm[O.xp + 3 * h_m.head_xp_slot + 2];
}

}  // extern "C"
