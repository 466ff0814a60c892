// Kinematic-tree model tables and the per-body ("block") tree sweeps of the fused rollout kernel.
//
// This header compiles twice: by nvcc into the rollout kernel (csrc/rollout.cu, storage = shared memory rows +
// Tensor Memory scratch) and by g++ into a host test harness (tests/host_sweeps.cpp, storage = plain arrays) that
// runs the same arithmetic for one environment so the CPU test suite can compare it with the CPU restatement
// (tests/test_host_sweeps.py) before any GPU time is spent.  The harness is test infrastructure: nothing in the
// product path calls the host build.
//
// Replaces, per sub-step of HumanoidEnv.do_simulation (ego_pose/envs/humanoid_v1.py:158-177): mj_fullM + the dense
// Cholesky solve of compute_desired_accel (:130-144), compute_torque (:146-156) and mj_step (call site :174).
//
// Formulation (unchanged from round 1): world-aligned spatial algebra about O = root position, bias force by RNE,
// both linear solves as articulated-body sweeps with armature / Kd h in the joint-space pivots.  What is new is
// the GRANULARITY of the sweeps: all hinges of one body share one anchor, so a body's 1-3 dofs are eliminated as
// ONE block (U = I^A S [6 x nd], D = S^T U + diag [nd x nd], one closed-form inverse, W = U D^-1, I^A -= W U^T)
// instead of nd dependent rank-1 steps.  Same result as the scalar recursion (block LDL^T of the same matrix), but
// the serial dependency chain per sub-step shrinks from 28 pivots to 10 blocks on the critical path
// root -> spine -> arm, every block exposes 100+ independent FMAs, and because the joint pattern of a body
// (x-y-z, or a single axis) is a template parameter the frame / inertia / carry arrays are statically indexed
// and live in registers.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/egopose_b200.h"

#if defined(__CUDACC__)
#define EGP_HD __device__ __forceinline__
#define EGP_CONST_M egp::c_m
#else
#define EGP_HD inline
#define EGP_CONST_M egp::h_m
#endif

namespace egp {

constexpr int MAXB = EGP_MAX_BODY;
constexpr int MAXV = EGP_MAX_DOF;
constexpr int MAXC = EGP_MAX_CHAIN;
constexpr int T4_CW = 4;                    // chain warps (own the tree sweeps)

// body kinds of the block sweeps
enum { BK_ROOT = 0, BK_XYZ = 1, BK_X = 2, BK_Y = 3, BK_Z = 4, BK_OTHER = 5 };

struct DevModel {
    int nq, nv, nu, nbody, nchain, frame_skip, head_body, v_ord, decay;
    int ee_body[EGP_NEE];
    double h, grav[3];
    int body_parent[MAXB], body_dofadr[MAXB], body_dofnum[MAXB], body_qposadr[MAXB], body_chain[MAXB], body_kind[MAXB];
    double body_pos[MAXB][3], body_mass[MAXB], body_ipos[MAXB][3], body_inertia[MAXB][6], b_diffw[MAXB];
    int dof_axis_id[MAXV];
    double dof_arm[MAXV], dof_axis[MAXV][3], dof_anchor[MAXV][3];
    double kp[MAXV], kd[MAXV], a_ref[MAXV], a_scale[MAXV], tlim[MAXV];        // indexed by dof (0 on the root)
    int chain_lo[MAXC], chain_hi[MAXC], chain_parent[MAXC];                   // body ranges, inclusive
    // T4 schedule: tree level of a chain, warp that owns it, slot of its forward/accel junction record (chains
    // with children), sibling slot for its backward junction record, number of child chains
    int chain_level[MAXC], chain_warp[MAXC], chain_pslot[MAXC], chain_cslot[MAXC], chain_nchild[MAXC];
    int lvl_chain[MAXC][T4_CW];                                               // chain of (level, warp) or -1
    int nlevel, nparent, max_sib, t4_ok;
    int body_xp_slot[MAXB], ee_xp_slot[EGP_NEE], head_xp_slot;                // rows of the shared body-position record
    // TMEM scratch layout (T4): index of a dof / body among those owned by the same warp, column bases (32-bit units)
    int dof_slot[MAXV], body_slot[MAXB];
    int tm_ctrl, tm_y, tm_tau, tm_c, tm_cin, tm_fb, tm_cols;
    double w_p, w_v, w_e, w_rp, w_rv, k_p, k_v, k_e, k_rh, k_rq, k_rl, k_ra;
};

#if defined(__CUDACC__)
__constant__ DevModel c_m;
#else
static DevModel h_m;
#endif

// shared-memory row offsets (units of one [32 env] row)
struct T4Off { int q, v, ax, anc, U, jf, jb, ja, xp, red, total; };

// ------------------------------------------------------------------------------------------------
// small math
EGP_HD void cross3(const double *a, const double *b, double *o) {
    double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
EGP_HD double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
EGP_HD double dot6(const double *a, const double *b) {
    return (a[0] * b[0] + a[1] * b[1] + a[2] * b[2]) + (a[3] * b[3] + a[4] * b[4] + a[5] * b[5]);
}
// quaternion helpers (w, x, y, z), utils/transformation.py:1379-1421 conventions
EGP_HD void quat_mul(const double *q1, const double *q0, double *o) {
    double w0 = q0[0], x0 = q0[1], y0 = q0[2], z0 = q0[3], w1 = q1[0], x1 = q1[1], y1 = q1[2], z1 = q1[3];
    o[0] = -x1 * x0 - y1 * y0 - z1 * z0 + w1 * w0;
    o[1] = x1 * w0 + y1 * z0 - z1 * y0 + w1 * x0;
    o[2] = -x1 * z0 + y1 * w0 + z1 * x0 + w1 * y0;
    o[3] = x1 * y0 - y1 * x0 + z1 * w0 + w1 * z0;
}
EGP_HD void quat_to_mat(const double *q, double *R) {      // unit quaternion
    double w = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = w * w + x * x - y * y - z * z; R[1] = 2 * (x * y - w * z);           R[2] = 2 * (x * z + w * y);
    R[3] = 2 * (x * y + w * z);           R[4] = w * w - x * x + y * y - z * z; R[5] = 2 * (y * z - w * x);
    R[6] = 2 * (x * z - w * y);           R[7] = 2 * (y * z + w * x);           R[8] = w * w - x * x - y * y + z * z;
}

// symmetric 6x6 in packed upper storage
EGP_HD constexpr int sx(int r, int c) { return r <= c ? r * 6 - r * (r - 1) / 2 + (c - r) : c * 6 - c * (c - 1) / 2 + (r - c); }

// spatial inertia (m, h = m c, I_O) applied to a motion vector [w; v]
EGP_HD void spi_mul(const double *ci /*10*/, const double *x, double *f) {
    const double m = ci[0], *hh = ci + 1, *I = ci + 4;
    double hxl[3], hxw[3];
    cross3(hh, x + 3, hxl);
    cross3(hh, x, hxw);
    f[0] = I[0] * x[0] + I[3] * x[1] + I[4] * x[2] + hxl[0];
    f[1] = I[3] * x[0] + I[1] * x[1] + I[5] * x[2] + hxl[1];
    f[2] = I[4] * x[0] + I[5] * x[1] + I[2] * x[2] + hxl[2];
    f[3] = m * x[3] - hxw[0];
    f[4] = m * x[4] - hxw[1];
    f[5] = m * x[5] - hxw[2];
}

struct Fwd {            // forward carry along a chain: body frame (relative to O), spatial velocity, bias accel
    double p[3], R[9], v[6], a[6];
};
struct Bwd {            // backward carry: articulated inertia, bias force of the pure solve, RNE force
    double IA[21], pA[6], F[6];
};

// ------------------------------------------------------------------------------------------------
// inverse of a small symmetric positive-definite matrix (joint-space pivot block), closed form: one division
template <int ND> struct SymInv;
template <> struct SymInv<1> {
    static EGP_HD void run(const double (&D)[1][1], double (&Di)[1][1]) { Di[0][0] = 1.0 / D[0][0]; }
};
template <> struct SymInv<3> {
    static EGP_HD void run(const double (&D)[3][3], double (&Di)[3][3]) {
        const double c00 = D[1][1] * D[2][2] - D[1][2] * D[1][2];
        const double c01 = D[0][2] * D[1][2] - D[0][1] * D[2][2];
        const double c02 = D[0][1] * D[1][2] - D[0][2] * D[1][1];
        const double c11 = D[0][0] * D[2][2] - D[0][2] * D[0][2];
        const double c12 = D[0][1] * D[0][2] - D[0][0] * D[1][2];
        const double c22 = D[0][0] * D[1][1] - D[0][1] * D[0][1];
        const double det = D[0][0] * c00 + D[0][1] * c01 + D[0][2] * c02;
        const double r = 1.0 / det;
        Di[0][0] = c00 * r; Di[0][1] = c01 * r; Di[0][2] = c02 * r;
        Di[1][0] = Di[0][1]; Di[1][1] = c11 * r; Di[1][2] = c12 * r;
        Di[2][0] = Di[0][2]; Di[2][1] = Di[1][2]; Di[2][2] = c22 * r;
    }
};

// One block of the backward articulated-body sweep.
//   S[j]      motion axes of the block's dofs (world axes about O)
//   diag[j]   armature (+ kd h for the stable-PD solve)
//   rhs[j]    right-hand side of dof j (tau - C for forward dynamics, -C - kp e - kd v for stable PD)
// Updates the carry (I^A -= W U^T, p^A += W u) and returns W = U D^-1 (for the forward sweep) and y = D^-1 u.
// SKIND: 0 general S = [ax; lin], 1 angular only S = [ax; 0] (root rotation), 2 S = [0; e_j] (root translation).
template <int ND, int SKIND, bool LAST>
EGP_HD void blk_backward(Bwd &w, const double (&S)[ND][6], const double (&diag)[ND], const double (&rhs)[ND],
                         double (&W)[ND][6], double (&y)[ND]) {
    double U[ND][6], D[ND][ND], Di[ND][ND], u[ND];
#pragma unroll
    for (int j = 0; j < ND; j++) {
#pragma unroll
        for (int r = 0; r < 6; r++) {
            if (SKIND == 2) U[j][r] = w.IA[sx(r, 3 + j)];
            else {
                double t = (w.IA[sx(r, 0)] * S[j][0] + w.IA[sx(r, 1)] * S[j][1]) + w.IA[sx(r, 2)] * S[j][2];
                if (SKIND == 0) t += (w.IA[sx(r, 3)] * S[j][3] + w.IA[sx(r, 4)] * S[j][4]) + w.IA[sx(r, 5)] * S[j][5];
                U[j][r] = t;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < ND; j++) {
#pragma unroll
        for (int k = j; k < ND; k++) {
            double t;
            if (SKIND == 2) t = U[k][3 + j];
            else if (SKIND == 1) t = dot3(S[j], U[k]);
            else t = dot6(S[j], U[k]);
            if (k == j) t += diag[j];
            D[j][k] = t; D[k][j] = t;
        }
        if (SKIND == 2) u[j] = rhs[j] - w.pA[3 + j];
        else if (SKIND == 1) u[j] = rhs[j] - dot3(S[j], w.pA);
        else u[j] = rhs[j] - dot6(S[j], w.pA);
    }
    SymInv<ND>::run(D, Di);
#pragma unroll
    for (int j = 0; j < ND; j++) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < ND; k++) t += Di[j][k] * u[k];
        y[j] = t;
#pragma unroll
        for (int r = 0; r < 6; r++) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < ND; k++) s += U[k][r] * Di[k][j];
            W[j][r] = s;
        }
    }
    if (LAST) return;           // nothing above the root: the reduced carry is never read
#pragma unroll
    for (int r = 0; r < 6; r++) {
#pragma unroll
        for (int c = r; c < 6; c++) {
            double t = w.IA[sx(r, c)];
#pragma unroll
            for (int j = 0; j < ND; j++) t -= W[j][r] * U[j][c];
            w.IA[sx(r, c)] = t;
        }
        double t = w.pA[r];
#pragma unroll
        for (int j = 0; j < ND; j++) t += W[j][r] * u[j];
        w.pA[r] = t;
    }
}

// add a body's spatial inertia [[I_O, hx], [hx^T, m 1]] (packed ci[10]) to the articulated inertia
EGP_HD void add_body_inertia(Bwd &w, const double *ci) {
    w.IA[sx(0, 0)] += ci[4]; w.IA[sx(1, 1)] += ci[5]; w.IA[sx(2, 2)] += ci[6];
    w.IA[sx(0, 1)] += ci[7]; w.IA[sx(0, 2)] += ci[8]; w.IA[sx(1, 2)] += ci[9];
    w.IA[sx(0, 4)] += -ci[3]; w.IA[sx(0, 5)] += ci[2];
    w.IA[sx(1, 3)] += ci[3];  w.IA[sx(1, 5)] += -ci[1];
    w.IA[sx(2, 3)] += -ci[2]; w.IA[sx(2, 4)] += ci[1];
    w.IA[sx(3, 3)] += ci[0]; w.IA[sx(4, 4)] += ci[0]; w.IA[sx(5, 5)] += ci[0];
}

// ------------------------------------------------------------------------------------------------
// Sweeps over one (level, warp) chain.  X is the storage context:
//   double &at(int row_off, int idx)          shared row [row][env] of this lane's environment
//   tld1/tld2/tld3 ... tst1/tst3 ...          per-thread scratch (Tensor Memory on the device), column = 32-bit unit
//   o                                         T4Off row offsets
// MODE of the backward sweep: 0 forward dynamics (bias C = S.F stored, rhs = tau - C, pivots + armature),
//                             1 stable PD (rhs = -C - kp e - kd v from the current q, v and the stored bias, pivots + kd h).
// MODE of the forward sweep:  1 [PD accel -> clipped torque] on the OLD tree rows, then kinematics refresh;
//                             0 forward-dynamics accel + semi-implicit Euler; 2 kinematics refresh only (sim.forward()).

template <class X>
EGP_HD void t5_bwd_gather(const X &x, int c, Bwd &w) {
#pragma unroll
    for (int k = 0; k < 21; k++) w.IA[k] = 0.0;
#pragma unroll
    for (int k = 0; k < 6; k++) { w.pA[k] = 0.0; w.F[k] = 0.0; }
    if (c >= 0) {
        const int nch = EGP_CONST_M.chain_nchild[c];
        for (int s = 0; s < nch; s++) {
            const int base = x.o.jb + 33 * s;
#pragma unroll
            for (int k = 0; k < 21; k++) w.IA[k] += x.at(base, k);
#pragma unroll
            for (int k = 0; k < 6; k++) { w.pA[k] += x.at(base, 21 + k); w.F[k] += x.at(base, 27 + k); }
        }
    }
}

// hinge body with ND joints through one anchor
template <int ND, class X>
EGP_HD void t5_bwd_body(const X &x, Bwd &w, int b, const int MODE) {
    const DevModel &M = EGP_CONST_M;
    const int da = M.body_dofadr[b], sl = M.dof_slot[da];
    double S[ND][6], diag[ND], rhs[ND], W[ND][6], y[ND];
    const double an[3] = {x.at(x.o.anc, 3 * b), x.at(x.o.anc, 3 * b + 1), x.at(x.o.anc, 3 * b + 2)};
#pragma unroll
    for (int j = 0; j < ND; j++) {
        const int i = da + j;
        S[j][0] = x.at(x.o.ax, 3 * i); S[j][1] = x.at(x.o.ax, 3 * i + 1); S[j][2] = x.at(x.o.ax, 3 * i + 2);
        cross3(an, S[j], S[j] + 3);
        diag[j] = M.dof_arm[i] + (MODE == 1 ? M.kd[i] * M.h : 0.0);
    }
    if (MODE == 0) {
        double tau[ND], C[ND];
        x.template tld<ND>(M.tm_tau + 2 * sl, tau);
#pragma unroll
        for (int j = 0; j < ND; j++) { C[j] = dot6(S[j], w.F); rhs[j] = tau[j] - C[j]; }
        x.template tst<ND>(M.tm_c + 2 * sl, C);
    } else {
        double C[ND], ctrl[ND];
        x.template tld<ND>(M.tm_c + 2 * sl, C);
        x.template tld<ND>(M.tm_ctrl + 2 * sl, ctrl);
#pragma unroll
        for (int j = 0; j < ND; j++) {
            const int i = da + j;
            const double eq = x.at(x.o.q, i + 1) - ctrl[j];
            rhs[j] = -C[j] - M.kp[i] * eq - M.kd[i] * x.at(x.o.v, i);
        }
    }
    blk_backward<ND, 0, false>(w, S, diag, rhs, W, y);
    x.template tst<ND>(M.tm_y + 2 * sl, y);
#pragma unroll
    for (int j = 0; j < ND; j++)
#pragma unroll
        for (int r = 0; r < 6; r++) x.at(x.o.U, 6 * (da + j) + r) = W[j][r];
}

// free-joint root: rotational block (dofs 3..5, S = [R e_k; 0]) then translational block (dofs 0..2, S = [0; e_k]);
// no gains, no armature, no actuation on the root (humanoid_v1.py:137-140)
template <class X>
EGP_HD void t5_bwd_root(const X &x, Bwd &w, int b, const int MODE) {
    const DevModel &M = EGP_CONST_M;
    const int da = M.body_dofadr[b], sl = M.dof_slot[da];
    double W[3][6], y[3], rhs[3], diag[3] = {M.dof_arm[da + 3], M.dof_arm[da + 4], M.dof_arm[da + 5]};
    {
        double S[3][6];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const int i = da + 3 + j;
            S[j][0] = x.at(x.o.ax, 3 * i); S[j][1] = x.at(x.o.ax, 3 * i + 1); S[j][2] = x.at(x.o.ax, 3 * i + 2);
            S[j][3] = S[j][4] = S[j][5] = 0.0;
        }
        double C[3];
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 3; j++) { C[j] = dot3(S[j], w.F); rhs[j] = -C[j]; }
            x.template tst<3>(M.tm_c + 2 * (sl + 3), C);
        } else {
            x.template tld<3>(M.tm_c + 2 * (sl + 3), C);
#pragma unroll
            for (int j = 0; j < 3; j++) rhs[j] = -C[j];
        }
        blk_backward<3, 1, false>(w, S, diag, rhs, W, y);
        x.template tst<3>(M.tm_y + 2 * (sl + 3), y);
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
            for (int r = 0; r < 6; r++) x.at(x.o.U, 6 * (da + 3 + j) + r) = W[j][r];
    }
    {
        double S[3][6] = {};
        double C[3];
        diag[0] = M.dof_arm[da]; diag[1] = M.dof_arm[da + 1]; diag[2] = M.dof_arm[da + 2];
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 3; j++) { C[j] = w.F[3 + j]; rhs[j] = -C[j]; }
            x.template tst<3>(M.tm_c + 2 * sl, C);
        } else {
            x.template tld<3>(M.tm_c + 2 * sl, C);
#pragma unroll
            for (int j = 0; j < 3; j++) rhs[j] = -C[j];
        }
        blk_backward<3, 2, true>(w, S, diag, rhs, W, y);
        x.template tst<3>(M.tm_y + 2 * sl, y);
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
            for (int r = 0; r < 6; r++) x.at(x.o.U, 6 * (da + j) + r) = W[j][r];
    }
}

template <class X>
EGP_HD void t5_bwd_chain(const X &x, int c, Bwd &w, const int MODE) {
    const DevModel &M = EGP_CONST_M;
    for (int b = M.chain_hi[c]; b >= M.chain_lo[c]; b--) {
        double ci[10];
        x.ld_cin(b, ci);
        add_body_inertia(w, ci);
        if (MODE == 0) {
            double fbv[6];
            x.ld_fb(b, fbv);
#pragma unroll
            for (int k = 0; k < 6; k++) w.F[k] += fbv[k];
        }
        const int kind = M.body_kind[b];
        if (kind == BK_XYZ) t5_bwd_body<3>(x, w, b, MODE);
        else if (kind == BK_ROOT) t5_bwd_root(x, w, b, MODE);
        else t5_bwd_body<1>(x, w, b, MODE);
    }
    x.twait_st();
    if (M.chain_parent[c] >= 0) {
        const int base = x.o.jb + 33 * M.chain_cslot[c];
#pragma unroll
        for (int k = 0; k < 21; k++) x.at(base, k) = w.IA[k];
#pragma unroll
        for (int k = 0; k < 6; k++) { x.at(base, 21 + k) = w.pA[k]; x.at(base, 27 + k) = w.F[k]; }
    }
}

// ---- forward sweep pieces ---------------------------------------------------------------------------
// solve part of a hinge body on the rows currently in storage: x = y - W^T a, a += S x, then torque (MODE 1) or
// semi-implicit Euler (MODE 0)
template <int ND, int MODE, class X>
EGP_HD void t5_fwd_solve_body(const X &x, double *a, int b) {
    const DevModel &M = EGP_CONST_M;
    const int da = M.body_dofadr[b], sl = M.dof_slot[da];
    const double h = M.h;
    const double an[3] = {x.at(x.o.anc, 3 * b), x.at(x.o.anc, 3 * b + 1), x.at(x.o.anc, 3 * b + 2)};
    double y[ND], xs[ND], S[ND][6];
    x.template tld<ND>(M.tm_y + 2 * sl, y);
#pragma unroll
    for (int j = 0; j < ND; j++) {
        const int i = da + j;
        double Wj[6];
#pragma unroll
        for (int r = 0; r < 6; r++) Wj[r] = x.at(x.o.U, 6 * i + r);
        xs[j] = y[j] - dot6(Wj, a);
        S[j][0] = x.at(x.o.ax, 3 * i); S[j][1] = x.at(x.o.ax, 3 * i + 1); S[j][2] = x.at(x.o.ax, 3 * i + 2);
        cross3(an, S[j], S[j] + 3);
    }
#pragma unroll
    for (int r = 0; r < 6; r++) {
        double t = a[r];
#pragma unroll
        for (int j = 0; j < ND; j++) t += S[j][r] * xs[j];
        a[r] = t;
    }
    if (MODE == 1) {                // torque = clip(-kp e - kd (v + x h))  (humanoid_v1.py:152-155,172)
        double ctrl[ND], tq[ND];
        x.template tld<ND>(M.tm_ctrl + 2 * sl, ctrl);
#pragma unroll
        for (int j = 0; j < ND; j++) {
            const int i = da + j;
            const double eq = x.at(x.o.q, i + 1) - ctrl[j];
            double t = -M.kp[i] * eq - M.kd[i] * (x.at(x.o.v, i) + xs[j] * h);
            const double lim = M.tlim[i];
            tq[j] = t < -lim ? -lim : (t > lim ? lim : t);
        }
        x.template tst<ND>(M.tm_tau + 2 * sl, tq);
    } else {
#pragma unroll
        for (int j = 0; j < ND; j++) {
            const int i = da + j;
            const double vn = x.at(x.o.v, i) + h * xs[j];
            x.at(x.o.v, i) = vn;
            x.at(x.o.q, i + 1) += h * vn;
        }
    }
}

template <int MODE, class X>
EGP_HD void t5_fwd_solve_root(const X &x, double *a, int b) {
    const DevModel &M = EGP_CONST_M;
    const int da = M.body_dofadr[b], sl = M.dof_slot[da], qa = M.body_qposadr[b];
    const double h = M.h;
    double y[6], xs[6];
    x.template tld<3>(M.tm_y + 2 * sl, y);
    x.template tld<3>(M.tm_y + 2 * (sl + 3), y + 3);
    // translational block first (the root has no parent: a = 0 on entry)
#pragma unroll
    for (int j = 0; j < 3; j++) {
        double Wj[6];
#pragma unroll
        for (int r = 0; r < 6; r++) Wj[r] = x.at(x.o.U, 6 * (da + j) + r);
        xs[j] = y[j] - dot6(Wj, a);
    }
#pragma unroll
    for (int j = 0; j < 3; j++) a[3 + j] += xs[j];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        double Wj[6];
#pragma unroll
        for (int r = 0; r < 6; r++) Wj[r] = x.at(x.o.U, 6 * (da + 3 + j) + r);
        xs[3 + j] = y[3 + j] - dot6(Wj, a);
    }
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const int i = da + 3 + j;
        const double xj = xs[3 + j];
        a[0] += x.at(x.o.ax, 3 * i) * xj; a[1] += x.at(x.o.ax, 3 * i + 1) * xj; a[2] += x.at(x.o.ax, 3 * i + 2) * xj;
    }
    if (MODE == 0) {
        // semi-implicit Euler; root position + quaternion integration with the NEW velocity
        double vn[6];
#pragma unroll
        for (int j = 0; j < 6; j++) { vn[j] = x.at(x.o.v, da + j) + h * xs[j]; x.at(x.o.v, da + j) = vn[j]; }
#pragma unroll
        for (int k = 0; k < 3; k++) x.at(x.o.q, qa + k) += h * vn[k];
        const double *wv = vn + 3;
        double n = sqrt(dot3(wv, wv)), ax[3] = {1.0, 0.0, 0.0};
        if (n > 1e-15) { ax[0] = wv[0] / n; ax[1] = wv[1] / n; ax[2] = wv[2] / n; }
        double sn, cs;
        sincos(0.5 * h * n, &sn, &cs);
        double qr[4] = {cs, ax[0] * sn, ax[1] * sn, ax[2] * sn};
        double q4[4] = {x.at(x.o.q, qa + 3), x.at(x.o.q, qa + 4), x.at(x.o.q, qa + 5), x.at(x.o.q, qa + 6)};
        double qn = sqrt(q4[0] * q4[0] + q4[1] * q4[1] + q4[2] * q4[2] + q4[3] * q4[3]);
#pragma unroll
        for (int k = 0; k < 4; k++) q4[k] /= qn;
        double o4[4];
        quat_mul(q4, qr, o4);
#pragma unroll
        for (int k = 0; k < 4; k++) x.at(x.o.q, qa + 3 + k) = o4[k];
    }
}

// body done: world position record, spatial inertia about O in world axes, RNE body force
template <class X>
EGP_HD void t5_body_finish(const X &x, const Fwd &f, int b) {
    const DevModel &M = EGP_CONST_M;
    const int xs = M.body_xp_slot[b];
    if (xs >= 0)
#pragma unroll
        for (int r = 0; r < 3; r++) x.at(x.o.xp, 3 * xs + r) = f.p[r] + x.at(x.o.q, r);
    double cpos[3];
    const double ip0 = M.body_ipos[b][0], ip1 = M.body_ipos[b][1], ip2 = M.body_ipos[b][2];
#pragma unroll
    for (int r = 0; r < 3; r++) cpos[r] = f.p[r] + f.R[3 * r] * ip0 + f.R[3 * r + 1] * ip1 + f.R[3 * r + 2] * ip2;
    const double *in = M.body_inertia[b];
    const double Ib[9] = {in[0], in[3], in[4], in[3], in[1], in[5], in[4], in[5], in[2]};
    double Tm[9], Iw[6];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int cc = 0; cc < 3; cc++)
            Tm[3 * r + cc] = f.R[3 * r] * Ib[cc] + f.R[3 * r + 1] * Ib[3 + cc] + f.R[3 * r + 2] * Ib[6 + cc];
    Iw[0] = Tm[0] * f.R[0] + Tm[1] * f.R[1] + Tm[2] * f.R[2];
    Iw[1] = Tm[3] * f.R[3] + Tm[4] * f.R[4] + Tm[5] * f.R[5];
    Iw[2] = Tm[6] * f.R[6] + Tm[7] * f.R[7] + Tm[8] * f.R[8];
    Iw[3] = Tm[0] * f.R[3] + Tm[1] * f.R[4] + Tm[2] * f.R[5];
    Iw[4] = Tm[0] * f.R[6] + Tm[1] * f.R[7] + Tm[2] * f.R[8];
    Iw[5] = Tm[3] * f.R[6] + Tm[4] * f.R[7] + Tm[5] * f.R[8];
    const double mass = M.body_mass[b], cc2 = dot3(cpos, cpos);
    double ci[10];
    ci[0] = mass;
    ci[1] = mass * cpos[0]; ci[2] = mass * cpos[1]; ci[3] = mass * cpos[2];
    ci[4] = Iw[0] + mass * (cc2 - cpos[0] * cpos[0]);
    ci[5] = Iw[1] + mass * (cc2 - cpos[1] * cpos[1]);
    ci[6] = Iw[2] + mass * (cc2 - cpos[2] * cpos[2]);
    ci[7] = Iw[3] - mass * cpos[0] * cpos[1];
    ci[8] = Iw[4] - mass * cpos[0] * cpos[2];
    ci[9] = Iw[5] - mass * cpos[1] * cpos[2];
    double Ia[6], Iv[6];
    spi_mul(ci, f.a, Ia);
    spi_mul(ci, f.v, Iv);
    double c0[3], c1[3], c2[3];
    cross3(f.v, Iv, c0);            // v x* f = [w x n + v x f ; w x f]
    cross3(f.v + 3, Iv + 3, c1);
    cross3(f.v, Iv + 3, c2);
    double fbv[6];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        fbv[r] = Ia[r] + c0[r] + c1[r];
        fbv[3 + r] = Ia[3 + r] + c2[r];
    }
    x.st_cin(b, ci);
    x.st_fb(b, fbv);
}

// kinematics refresh of a hinge body: ND joints about coordinate axes (A0 + j) % 3 of the successively rotated
// frame, all through one anchor (MuJoCo kinematics)
template <int ND, int A0, class X>
EGP_HD void t5_fwd_kin_body(const X &x, Fwd &f, int b) {
    const DevModel &M = EGP_CONST_M;
    const int da = M.body_dofadr[b], qa = M.body_qposadr[b];
    const double bp0 = M.body_pos[b][0], bp1 = M.body_pos[b][1], bp2 = M.body_pos[b][2];
    const double da0 = M.dof_anchor[da][0], da1 = M.dof_anchor[da][1], da2 = M.dof_anchor[da][2];
    double anc[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        f.p[r] += f.R[3 * r] * bp0 + f.R[3 * r + 1] * bp1 + f.R[3 * r + 2] * bp2;
        anc[r] = f.p[r] + f.R[3 * r] * da0 + f.R[3 * r + 1] * da1 + f.R[3 * r + 2] * da2;
        x.at(x.o.anc, 3 * b + r) = anc[r];
    }
    double sn[ND], cs[ND], qd[ND];
#pragma unroll
    for (int j = 0; j < ND; j++) { sincos(x.at(x.o.q, qa + j), &sn[j], &cs[j]); qd[j] = x.at(x.o.v, da + j); }
#pragma unroll
    for (int j = 0; j < ND; j++) {
        const int aid = (A0 + j) % 3, c1 = (aid + 1) % 3, c2 = (aid + 2) % 3;
        const int i = da + j;
        double S[6];
        S[0] = f.R[aid]; S[1] = f.R[3 + aid]; S[2] = f.R[6 + aid];
#pragma unroll
        for (int r = 0; r < 3; r++) x.at(x.o.ax, 3 * i + r) = S[r];
        cross3(anc, S, S + 3);
        // cdof_dot = v x S ; a += cdof_dot qd ; v += S qd
        double t0[3], t1[3], t2[3];
        cross3(f.v, S, t0);
        cross3(f.v, S + 3, t1);
        cross3(f.v + 3, S, t2);
#pragma unroll
        for (int r = 0; r < 3; r++) {
            f.a[r] += t0[r] * qd[j];
            f.a[3 + r] += (t1[r] + t2[r]) * qd[j];
        }
#pragma unroll
        for (int r = 0; r < 6; r++) f.v[r] += S[r] * qd[j];
        // rotate the frame about the joint axis through the anchor
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const double a1 = f.R[3 * r + c1], a2 = f.R[3 * r + c2];
            f.R[3 * r + c1] = cs[j] * a1 + sn[j] * a2;
            f.R[3 * r + c2] = -sn[j] * a1 + cs[j] * a2;
        }
    }
#pragma unroll
    for (int r = 0; r < 3; r++) f.p[r] = anc[r] - (f.R[3 * r] * da0 + f.R[3 * r + 1] * da1 + f.R[3 * r + 2] * da2);
}

// free joint: O = root position, so the frame origin is 0 and rotational cdofs have no linear part
template <class X>
EGP_HD void t5_fwd_kin_root(const X &x, Fwd &f, int b) {
    const DevModel &M = EGP_CONST_M;
    const int da = M.body_dofadr[b], qa = M.body_qposadr[b];
    double q4[4] = {x.at(x.o.q, qa + 3), x.at(x.o.q, qa + 4), x.at(x.o.q, qa + 5), x.at(x.o.q, qa + 6)};
    const double n = sqrt(q4[0] * q4[0] + q4[1] * q4[1] + q4[2] * q4[2] + q4[3] * q4[3]);
#pragma unroll
    for (int k = 0; k < 4; k++) q4[k] /= n;
    quat_to_mat(q4, f.R);
    f.p[0] = f.p[1] = f.p[2] = 0.0;
    const double wl[3] = {x.at(x.o.v, da + 3), x.at(x.o.v, da + 4), x.at(x.o.v, da + 5)};
    double ww[3];
#pragma unroll
    for (int r = 0; r < 3; r++) ww[r] = f.R[3 * r] * wl[0] + f.R[3 * r + 1] * wl[1] + f.R[3 * r + 2] * wl[2];
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int r = 0; r < 3; r++) {
            x.at(x.o.ax, 3 * (da + k) + r) = r == k ? 1.0 : 0.0;
            x.at(x.o.ax, 3 * (da + 3 + k) + r) = f.R[3 * r + k];
        }
        x.at(x.o.anc, 3 * b + k) = 0.0;
    }
    const double vl[3] = {x.at(x.o.v, da), x.at(x.o.v, da + 1), x.at(x.o.v, da + 2)};
    double vxw[3];
    cross3(vl, ww, vxw);            // sum_k ([0;v] x [R e_k;0]) w_k = [0; v x (R w)]
#pragma unroll
    for (int k = 0; k < 3; k++) {
        f.v[k] = ww[k]; f.v[3 + k] = vl[k];
        f.a[k] = 0.0; f.a[3 + k] = -M.grav[k] + vxw[k];
    }
}

template <int MODE, class X>
EGP_HD void t5_fwd_chain(const X &x, int c) {
    const DevModel &M = EGP_CONST_M;
    Fwd f;
    double a[6];
    const int pc = M.chain_parent[c];
    if (MODE != 2) {
#pragma unroll
        for (int k = 0; k < 6; k++) a[k] = pc >= 0 ? x.at(x.o.ja + 6 * M.chain_pslot[pc], k) : 0.0;
    }
    if (MODE != 0 && pc >= 0) {
        const int base = x.o.jf + 24 * M.chain_pslot[pc];
#pragma unroll
        for (int k = 0; k < 3; k++) f.p[k] = x.at(base, k);
#pragma unroll
        for (int k = 0; k < 9; k++) f.R[k] = x.at(base, 3 + k);
#pragma unroll
        for (int k = 0; k < 6; k++) { f.v[k] = x.at(base, 12 + k); f.a[k] = x.at(base, 18 + k); }
    }
    for (int b = M.chain_lo[c]; b <= M.chain_hi[c]; b++) {
        const int kind = M.body_kind[b];
        if (MODE != 2) {
            if (kind == BK_XYZ) t5_fwd_solve_body<3, MODE>(x, a, b);
            else if (kind == BK_ROOT) t5_fwd_solve_root<MODE>(x, a, b);
            else t5_fwd_solve_body<1, MODE>(x, a, b);
        }
        if (MODE != 0) {
            if (kind == BK_XYZ) t5_fwd_kin_body<3, 0>(x, f, b);
            else if (kind == BK_ROOT) t5_fwd_kin_root(x, f, b);
            else if (kind == BK_X) t5_fwd_kin_body<1, 0>(x, f, b);
            else if (kind == BK_Y) t5_fwd_kin_body<1, 1>(x, f, b);
            else t5_fwd_kin_body<1, 2>(x, f, b);
            t5_body_finish(x, f, b);
        }
    }
    x.twait_st();
    if (M.chain_pslot[c] >= 0) {
        if (MODE != 2) {
#pragma unroll
            for (int k = 0; k < 6; k++) x.at(x.o.ja + 6 * M.chain_pslot[c], k) = a[k];
        }
        if (MODE != 0) {
            const int base = x.o.jf + 24 * M.chain_pslot[c];
#pragma unroll
            for (int k = 0; k < 3; k++) x.at(base, k) = f.p[k];
#pragma unroll
            for (int k = 0; k < 9; k++) x.at(base, 3 + k) = f.R[k];
#pragma unroll
            for (int k = 0; k < 6; k++) { x.at(base, 12 + k) = f.v[k]; x.at(base, 18 + k) = f.a[k]; }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side: model tables
inline void build_chains(DevModel &d) {
    // a chain is a maximal run of consecutive bodies b, b+1, ... with parent(b+1) == b where b has one child
    int nb = d.nbody, nchild[MAXB] = {0};
    for (int b = 0; b < nb; b++) if (d.body_parent[b] >= 0) nchild[d.body_parent[b]]++;
    int nc = 0;
    for (int b = 0; b < nb; b++) {
        bool cont = b > 0 && d.body_parent[b] == b - 1 && nchild[b - 1] == 1;
        if (!cont) {
            if (nc >= MAXC) { d.nchain = MAXC + 1; d.t4_ok = 0; return; }
            d.chain_lo[nc] = b;
            d.chain_parent[nc] = d.body_parent[b] >= 0 ? d.body_chain[d.body_parent[b]] : -1;
            nc++;
        }
        d.chain_hi[nc - 1] = b;
        d.body_chain[b] = nc - 1;
    }
    d.nchain = nc;
    // ---- T4 schedule
    int cchild[MAXC] = {0}, per_level_parents[MAXC] = {0};
    d.nlevel = 0; d.nparent = 0; d.max_sib = 0; d.t4_ok = 1;
    for (int c = 0; c < nc; c++) {
        d.chain_level[c] = d.chain_parent[c] >= 0 ? d.chain_level[d.chain_parent[c]] + 1 : 0;
        if (d.chain_level[c] + 1 > d.nlevel) d.nlevel = d.chain_level[c] + 1;
        d.chain_cslot[c] = d.chain_parent[c] >= 0 ? cchild[d.chain_parent[c]]++ : -1;
    }
    for (int c = 0; c < nc; c++) {
        d.chain_nchild[c] = cchild[c];
        d.chain_pslot[c] = cchild[c] > 0 ? d.nparent++ : -1;
        if (cchild[c] > d.max_sib) d.max_sib = cchild[c];
        if (cchild[c] > 0) per_level_parents[d.chain_level[c]]++;
    }
    // per level: longest chain to warp 0, next to warp 1, ... (one chain per warp per level)
    for (int L = 0; L < MAXC; L++) for (int w = 0; w < T4_CW; w++) d.lvl_chain[L][w] = -1;
    for (int L = 0; L < d.nlevel; L++) {
        int order[MAXC], n = 0;
        for (int c = 0; c < nc; c++) if (d.chain_level[c] == L) order[n++] = c;
        auto ndof = [&](int c) { return d.body_dofadr[d.chain_hi[c]] + d.body_dofnum[d.chain_hi[c]] - d.body_dofadr[d.chain_lo[c]]; };
        for (int a = 0; a < n; a++) for (int b2 = a + 1; b2 < n; b2++) {
            // parents first (the trunk continues on warp 0), then by length
            bool swap = (cchild[order[b2]] > 0 && cchild[order[a]] == 0) ||
                        ((cchild[order[b2]] > 0) == (cchild[order[a]] > 0) && ndof(order[b2]) > ndof(order[a]));
            if (swap) { int t = order[a]; order[a] = order[b2]; order[b2] = t; }
        }
        for (int a = 0; a < n; a++) {
            d.chain_warp[order[a]] = a % T4_CW;
            if (a < T4_CW) d.lvl_chain[L][a] = order[a];
        }
        if (n > T4_CW || per_level_parents[L] > 1) d.t4_ok = 0;
    }
    // every hinge of a body must share one anchor (the shared rows keep one anchor per body); the block sweeps know
    // the joint patterns x-y-z and single x / y / z
    for (int b = 0; b < nb; b++) {
        const int da = d.body_dofadr[b], nd = d.body_dofnum[b];
        for (int j = 1; j < nd && b > 0; j++)
            for (int k = 0; k < 3; k++)
                if (d.dof_anchor[da + j][k] != d.dof_anchor[da][k]) d.t4_ok = 0;
        int kind = BK_OTHER;
        if (b == 0) kind = BK_ROOT;
        else if (nd == 3 && d.dof_axis_id[da] == 0 && d.dof_axis_id[da + 1] == 1 && d.dof_axis_id[da + 2] == 2) kind = BK_XYZ;
        else if (nd == 1 && d.dof_axis_id[da] >= 0) kind = BK_X + d.dof_axis_id[da];
        d.body_kind[b] = kind;
        if (kind == BK_OTHER) d.t4_ok = 0;
    }
    for (int b = 0; b < nb; b++) d.body_xp_slot[b] = -1;
    int nslot = 0;
    for (int k = 0; k < EGP_NEE; k++) {
        if (d.body_xp_slot[d.ee_body[k]] < 0) d.body_xp_slot[d.ee_body[k]] = nslot++;
        d.ee_xp_slot[k] = d.body_xp_slot[d.ee_body[k]];
    }
    if (d.body_xp_slot[d.head_body] < 0) d.body_xp_slot[d.head_body] = nslot++;
    d.head_xp_slot = d.body_xp_slot[d.head_body];
    // TMEM scratch slots: position of each dof / body among those owned by the same warp.  Per dof: ctrl, y, tau, C
    // (2 columns each; one spare dof slot so that 3-wide block accesses may be issued as one 4-wide access)
    int nd_w[T4_CW] = {0}, nb_w[T4_CW] = {0};
    for (int b = 0; b < nb; b++) {
        int w = d.chain_warp[d.body_chain[b]];
        d.body_slot[b] = nb_w[w]++;
        for (int i = d.body_dofadr[b]; i < d.body_dofadr[b] + d.body_dofnum[b]; i++) d.dof_slot[i] = nd_w[w]++;
    }
    int ND = 0, NB = 0;
    for (int w = 0; w < T4_CW; w++) { if (nd_w[w] > ND) ND = nd_w[w]; if (nb_w[w] > NB) NB = nb_w[w]; }
    d.tm_ctrl = 0; d.tm_y = 2 * ND; d.tm_tau = 4 * ND; d.tm_c = 6 * ND; d.tm_cin = 8 * ND; d.tm_fb = 8 * ND + 20 * NB;
    d.tm_cols = 8 * ND + 32 * NB;
    if (d.tm_cols > 512) d.t4_ok = 0;
}

// EgpModelDesc -> DevModel (egp_model_create); returns 0 or a negative EGP_E* code with *why set
inline int fill_dev_model(const EgpModelDesc *s, DevModel &d, const char **why) {
    *why = nullptr;
    if (s->nbody > MAXB || s->nv > MAXV || s->nbody < 1 || s->nv != s->nq - 1 || s->nu != s->nv - 6) {
        *why = "unsupported sizes";
        return EGP_ESIZE;
    }
    if (s->body_dofnum[0] != 6 || s->body_parent[0] != -1) { *why = "body 0 must be the free-joint root"; return EGP_EINVAL; }
    memset(&d, 0, sizeof d);
    d.nq = s->nq; d.nv = s->nv; d.nu = s->nu; d.nbody = s->nbody;
    d.frame_skip = s->frame_skip; d.head_body = s->head_body; d.v_ord = s->v_ord; d.decay = s->decay;
    for (int k = 0; k < EGP_NEE; k++) d.ee_body[k] = s->ee_body[k];
    d.h = s->timestep;
    for (int k = 0; k < 3; k++) d.grav[k] = s->gravity[k];
    for (int b = 0; b < s->nbody; b++) {
        d.body_parent[b] = s->body_parent[b]; d.body_dofadr[b] = s->body_dofadr[b];
        d.body_dofnum[b] = s->body_dofnum[b]; d.body_qposadr[b] = s->body_qposadr[b];
        if (b > 0 && (s->body_parent[b] < 0 || s->body_parent[b] >= b || s->body_dofnum[b] < 1 || s->body_dofnum[b] > 3)) {
            *why = "only 1-3 hinge joints on non-root bodies, parents before children";
            return EGP_EINVAL;
        }
        d.body_mass[b] = s->body_mass[b];
        for (int k = 0; k < 3; k++) { d.body_pos[b][k] = s->body_pos[3 * b + k]; d.body_ipos[b][k] = s->body_ipos[3 * b + k]; }
        for (int k = 0; k < 6; k++) d.body_inertia[b][k] = s->body_inertia[6 * b + k];
        d.b_diffw[b] = (b < s->nbody - 1 && s->b_diffw) ? s->b_diffw[b] : 1.0;
    }
    for (int i = 0; i < s->nv; i++) {
        d.dof_arm[i] = s->dof_armature[i];
        int aid = -1;
        for (int k = 0; k < 3; k++) {
            d.dof_axis[i][k] = s->dof_axis[3 * i + k];
            d.dof_anchor[i][k] = s->dof_anchor[3 * i + k];
        }
        for (int k = 0; k < 3; k++)
            if (d.dof_axis[i][k] == 1.0 && d.dof_axis[i][(k + 1) % 3] == 0.0 && d.dof_axis[i][(k + 2) % 3] == 0.0) aid = k;
        d.dof_axis_id[i] = aid;
        bool act = i >= 6;
        d.kp[i] = act ? s->jkp[i - 6] : 0.0;
        d.kd[i] = act ? s->jkd[i - 6] : 0.0;
        d.a_ref[i] = act ? s->a_ref[i - 6] : 0.0;
        d.a_scale[i] = act ? s->a_scale[i - 6] : 0.0;
        d.tlim[i] = act ? s->torque_lim[i - 6] : 0.0;
    }
    d.w_p = s->w_p; d.w_v = s->w_v; d.w_e = s->w_e; d.w_rp = s->w_rp; d.w_rv = s->w_rv;
    d.k_p = s->k_p; d.k_v = s->k_v; d.k_e = s->k_e; d.k_rh = s->k_rh; d.k_rq = s->k_rq; d.k_rl = s->k_rl; d.k_ra = s->k_ra;
    build_chains(d);
    if (d.nchain > MAXC) { *why = "too many chains"; return EGP_ESIZE; }
    return EGP_OK;
}

inline T4Off t4_offsets(const DevModel &d) {
    T4Off O;
    O.q = 0; O.v = O.q + d.nq; O.ax = O.v + d.nv; O.anc = O.ax + 3 * d.nv; O.U = O.anc + 3 * d.nbody;
    O.jf = O.U + 6 * d.nv; O.jb = O.jf + 24 * d.nparent; O.ja = O.jb + 33 * d.max_sib; O.xp = O.ja + 6 * d.nparent;
    O.red = O.xp + 3 * (EGP_NEE + 1); O.total = O.red + 16;
    return O;
}

}  // namespace egp
