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

#ifndef EGP_CLK_MARK
#define EGP_CLK_MARK(slot)
#endif

namespace egp {

constexpr int MAXB = EGP_MAX_BODY;
constexpr int MAXV = EGP_MAX_DOF;
constexpr int MAXC = EGP_MAX_CHAIN;
constexpr int T4_CW = 4;                    // chain warps (own the tree sweeps)

constexpr int R_COLS_HOST = 56;              // columns of one body's scratch record (see the sweeps below)

// body kinds of the block sweeps
enum { BK_ROOT = 0, BK_XYZ = 1, BK_X = 2, BK_Y = 3, BK_Z = 4, BK_OTHER = 5 };

// per-body constants of the block sweeps, contiguous so that a body-step fetches them with a few wide constant loads
struct BodyK {
    int da, qa, nd, kind, rec, xp_slot, pad0, pad1;         // dof / qpos address, #dofs, BK_*, scratch-record column, xp row slot
    double pos[3], anchor[3], ipos[3], inertia[6], mass;
    double arm[3], armkd[3], kp[3], kd[3], tlim[3];         // per dof of the body: armature, armature + kd h, gains, torque limit
};

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
    // scratch (Tensor Memory) layout: one 56-column record per body, numbered among the bodies owned by the same warp;
    // column of a dof's ctrl / tau / C entry (-1 where the root has none)
    int body_slot[MAXB], dof_col_ctrl[MAXV], dof_col_tau[MAXV], dof_col_c[MAXV], tm_cols;
    BodyK bk[MAXB];
    double w_p, w_v, w_e, w_rp, w_rv, k_p, k_v, k_e, k_rh, k_rq, k_rl, k_ra;
    // joint limits as MuJoCo soft constraints (egp_model_set_joint_limits; off = smooth dynamics, the north star)
    int limits;
    double lim_lo[MAXV], lim_hi[MAXV], lim_iw[MAXV];       // range per dof (lo >= hi: none), diag(M^-1) at qpos0
    double lim_k, lim_b, lim_d0, lim_dw, lim_width, lim_mid, lim_pow;
    // floor contact (egp_model_set_contacts): one collision geom per body (0 sphere | 1 capsule | 2 box), plane z = 0
    int contacts, geom_type[MAXB];
    double geom_size[MAXB][3], geom_p0[MAXB][3], geom_p1[MAXB][3], body_iw[MAXB];     // body_iw: translational invweight0
    double con_margin, con_mu;
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
// phase A: U = I^A S, D = S^T U + diag, u = rhs - S^T p^A, D^-1
template <int ND, int SKIND>
EGP_HD void blk_backward_a(const Bwd &w, const double (&S)[ND][6], const double (&diag)[ND], const double (&rhs)[ND],
                           double (&U)[ND][6], double (&Di)[ND][ND], double (&u)[ND]) {
    double D[ND][ND];
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
}

// phase B: W = U D^-1, y = D^-1 u, carry update I^A -= W U^T, p^A += W u (skipped at the root: nothing above it).
// W is produced one row (r) at a time - stored through `put(j, r, value)` and folded into row r of the carry - so that
// only ND of its 6 ND entries are live at any moment.
template <int ND, bool LAST, class PUT>
EGP_HD void blk_backward_b(Bwd &w, const double (&U)[ND][6], const double (&Di)[ND][ND], const double (&u)[ND],
                           double (&y)[ND], PUT put) {
#pragma unroll
    for (int j = 0; j < ND; j++) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < ND; k++) t += Di[j][k] * u[k];
        y[j] = t;
    }
#pragma unroll
    for (int r = 0; r < 6; r++) {
        double Wr[ND];
#pragma unroll
        for (int j = 0; j < ND; j++) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < ND; k++) s += U[k][r] * Di[k][j];
            Wr[j] = s;
            put(j, r, s);
        }
        if (!LAST) {
#pragma unroll
            for (int c = r; c < 6; c++) {
                double t = w.IA[sx(r, c)];
#pragma unroll
                for (int j = 0; j < ND; j++) t -= Wr[j] * U[j][c];
                w.IA[sx(r, c)] = t;
            }
            double t = w.pA[r];
#pragma unroll
            for (int j = 0; j < ND; j++) t += Wr[j] * u[j];
            w.pA[r] = t;
        }
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
//   double &at(int row_off, int idx)            shared row [row][env] of this lane's environment
//   tld_issue<NR>(col, int (&r)[NR]) / tld_wait  per-thread scratch (Tensor Memory on the device): asynchronous load of NR
//                                               32-bit columns, complete after the matching tld_wait; unpack(r, k) = double k
//   tst<N>(col, const double *)                 scratch store of N doubles
//   o                                           T4Off row offsets
// MODE of the backward sweep: 0 forward dynamics (bias C = S.F stored, rhs = tau - C, pivots + armature),
//                             1 stable PD (rhs = -C - kp e - kd v from the current q, v and the stored bias, pivots + kd h).
// MODE of the forward sweep:  1 [PD accel -> clipped torque] on the OLD tree rows, then kinematics refresh;
//                             0 forward-dynamics accel + semi-implicit Euler; 2 kinematics refresh only (sim.forward()).
//
// Scheduling: the only values that really recur along a chain are the carries (I^A, p^A, F backward; frame, velocity,
// acceleration forward).  Everything else a body-step reads - its scratch record, its shared rows, its constants - does
// not depend on the carry, so every sweep is software-pipelined BY HAND: the operands of body b+1 are requested before the
// arithmetic of body b starts (one body ahead: what fits in registers next to the carry).  The compiler cannot do this
// itself: the loads of b+1 would have to move above the stores of b into the same shared array.
//
// Scratch record of a body (56 columns = 28 doubles, contiguous so that one access fetches what a sweep needs):
//   hinge body: fb 6 | cin 10 | ctrl 3 | C 3 | tau 3 | y 3        root: fb 6 | cin 10 | C 6 | y 6
// cross-body operand prefetch per sweep (register pressure decides: see DESIGN.md)
#ifndef EGP_PF_BWD
#define EGP_PF_BWD 0
#endif
#ifndef EGP_PF_FWD1
#define EGP_PF_FWD1 1
#endif
#ifndef EGP_PF_FWD0
#define EGP_PF_FWD0 0
#endif
constexpr int R_FB = 0, R_CIN = 12, R_CTRL = 32, R_C = 38, R_TAU = 44, R_Y = 50, R_ROOT_C = 32, R_ROOT_Y = 44, R_COLS = 56;

// ---- constraint rows in the block sweeps (contexts with X::CONS) -----------------------------
// A context with CONS = true adds:  double &cs(int slot)  this environment's constraint scratch (L2-resident global memory,
// [slot][lane] per CTA), and the per-thread state  ccnt (3 bits per body slot: floor contacts of that body), lim_inst / lim_act
// (1 bit per (body slot, joint): range violated / row active), cchg (the active set changed in this pass), cit (active-set
// iteration), cflip (a row of this thread has flipped in this pass), lim_pinst / lim_pact (the same at the end of the
// previous sub-step).  Scratch record of body b at slot CS_PER_BODY * b:  flags (as a double: bit 4 k + e = edge e of contact k active) |
// per contact: point c (relative to O) 3, weight D, aref of the four pyramid edges.
constexpr int CS_PER_BODY = 33;
constexpr int CONS_CAREFUL = 12;     // passes after which only one disagreeing row flips per pass (coupled rows can flip back and forth)

// may a disagreeing row of this thread flip in this pass?  Always in the first CONS_CAREFUL passes; afterwards one row per pass
// and environment: only in the chain warp whose turn it is (round robin), and only the first one it meets
template <class X>
EGP_HD bool cons_may_flip(const X &x) {
    if (x.cit < CONS_CAREFUL) return true;
    if ((x.cit & 3) != x.w || x.cflip) return false;
    x.cflip = 1;
    return true;
}

// one row at signed distance dist (already minus its margin) and row velocity vel -> weight D = 1/R, reference acceleration
EGP_HD void cons_row(double invweight, double dist, double vel, double &Dc, double &aref) {
    const DevModel &M = EGP_CONST_M;
    double imp;
    if (M.lim_d0 == M.lim_dw || M.lim_width <= 1e-15) imp = 0.5 * (M.lim_d0 + M.lim_dw);
    else {
        const double xx = fabs(dist) / M.lim_width, pw = M.lim_pow, mid = M.lim_mid;
        double y;
        if (xx >= 1.0) y = 1.0;
        else if (xx <= 0.0) y = 0.0;
        else if (pw == 1.0) y = xx;
        else if (xx <= mid) y = pow(xx, pw) / pow(mid, pw - 1.0);
        else y = 1.0 - pow(1.0 - xx, pw) / pow(1.0 - mid, pw - 1.0);
        imp = M.lim_d0 + y * (M.lim_dw - M.lim_d0);
    }
    double R = (1.0 - imp) / imp * invweight;
    if (R < 1e-15) R = 1e-15;
    Dc = 1.0 / R;
    aref = -M.lim_b * vel - M.lim_k * imp * dist;
}

// range row of dof i at position q, velocity v: instantiated?, side s, weight, aref
EGP_HD bool cons_limit(int i, double q, double v, double &s, double &Dc, double &aref) {
    const DevModel &M = EGP_CONST_M;
    const double lo = M.lim_lo[i], hi = M.lim_hi[i];
    if (!M.limits || !(lo < hi)) return false;
    double dist;
    if (q - lo < 0.0) { dist = q - lo; s = 1.0; }
    else if (hi - q < 0.0) { dist = hi - q; s = -1.0; }
    else return false;
    cons_row(M.lim_iw[i], dist, s * v, Dc, aref);
    return true;
}

// spatial directions [c x d; d] of the contact normal (0,0,1) and the tangents (0,1,0), (-1,0,0) at point c
EGP_HD void cons_dirs(const double *c, double *pn, double *p1, double *p2) {
    pn[0] = c[1]; pn[1] = -c[0]; pn[2] = 0.0; pn[3] = 0.0; pn[4] = 0.0; pn[5] = 1.0;
    p1[0] = -c[2]; p1[1] = 0.0; p1[2] = c[0]; p1[3] = 0.0; p1[4] = 1.0; p1[5] = 0.0;
    p2[0] = 0.0; p2[1] = -c[2]; p2[2] = c[1]; p2[3] = -1.0; p2[4] = 0.0; p2[5] = 0.0;
}

// one contact of body b at point c (relative to O), distance dist: record k of the body's scratch, all four edges
template <class X>
EGP_HD void cons_emit(const X &x, const Fwd &f, int b, int k, const double *c, double dist) {
    const DevModel &M = EGP_CONST_M;
    const double mu = M.con_mu, tran = M.body_iw[b];
    double pn[6], p1[6], p2[6];
    cons_dirs(c, pn, p1, p2);
    const double vn = dot6(pn, f.v), v1 = dot6(p1, f.v), v2 = dot6(p2, f.v);
    const int r0 = CS_PER_BODY * b + 1 + 8 * k;
    x.cs(r0) = c[0]; x.cs(r0 + 1) = c[1]; x.cs(r0 + 2) = c[2];
    double D0 = 0.0;
#pragma unroll
    for (int ed = 0; ed < 4; ed++) {
        double ar;
        cons_row(tran + mu * mu * tran, dist - M.con_margin, vn + ((ed & 1) ? -mu : mu) * (ed < 2 ? v1 : v2), D0, ar);
        x.cs(r0 + 4 + ed) = ar;
    }
    x.cs(r0 + 3) = D0 / (2.0 * mu * mu);
}

// kinematics pass: body b's geom against the floor plane z = 0 (mjc_PlaneSphere / PlaneCapsule / PlaneBox) -> scratch record,
// all edges active
template <class X>
EGP_HD void cons_collide(const X &x, const Fwd &f, int b) {
    const DevModel &M = EGP_CONST_M;
    const int slot = M.body_slot[b];
    const int old_cnt = (x.ccnt >> (3 * slot)) & 7u;
    int cnt = 0;
    if (M.contacts) {
        const double margin = M.con_margin, zO = x.at(x.o.q, 2);
        const double *sz = M.geom_size[b];
        double c0[3];
#pragma unroll
        for (int r = 0; r < 3; r++)
            c0[r] = f.p[r] + f.R[3 * r] * M.geom_p0[b][0] + f.R[3 * r + 1] * M.geom_p0[b][1] + f.R[3 * r + 2] * M.geom_p0[b][2];
        if (M.geom_type[b] == 2) {
            for (int vtx = 0; vtx < 8 && cnt < 4; vtx++) {
                const double l0 = (vtx & 1) ? sz[0] : -sz[0], l1 = (vtx & 2) ? sz[1] : -sz[1], l2 = (vtx & 4) ? sz[2] : -sz[2];
                double wv[3];
#pragma unroll
                for (int r = 0; r < 3; r++) wv[r] = c0[r] + f.R[3 * r] * l0 + f.R[3 * r + 1] * l1 + f.R[3 * r + 2] * l2;
                const double dist = wv[2] + zO;
                if (dist > margin) continue;
                wv[2] -= 0.5 * dist;
                cons_emit(x, f, b, cnt, wv, dist);
                cnt++;
            }
        } else {
            const int ne = M.geom_type[b] == 1 ? 2 : 1;
            for (int en = 0; en < ne; en++) {
                double cc[3];
                if (ne == 2 && en == 0) {
#pragma unroll
                    for (int r = 0; r < 3; r++)
                        cc[r] = f.p[r] + f.R[3 * r] * M.geom_p1[b][0] + f.R[3 * r + 1] * M.geom_p1[b][1] + f.R[3 * r + 2] * M.geom_p1[b][2];
                } else {
#pragma unroll
                    for (int r = 0; r < 3; r++) cc[r] = c0[r];
                }
                const double dist = cc[2] + zO - sz[0];
                if (dist >= margin) continue;
                cc[2] -= sz[0] + 0.5 * dist;
                cons_emit(x, f, b, cnt, cc, dist);
                cnt++;
            }
        }
        // first guess of the active set: what the previous sub-step ended with if the body still has the same number of contacts
        // (a body at rest keeps its rows; the fixed point of the iteration does not depend on the guess), else every edge
        if (cnt && cnt != old_cnt) x.cs(CS_PER_BODY * b) = (double)((1 << (4 * cnt)) - 1);
    }
    x.ccnt = (x.ccnt & ~(7u << (3 * slot))) | ((unsigned)cnt << (3 * slot));
}

// backward sweep: the active contact rows of body b join its articulated inertia (D p p^T) and its bias force (- D aref p)
template <class X>
EGP_HD void cons_bwd_contacts(const X &x, Bwd &w, int b) {
    const int slot = EGP_CONST_M.body_slot[b], cnt = (x.ccnt >> (3 * slot)) & 7u, base = CS_PER_BODY * b;
    if (!cnt) return;
    const unsigned flags = (unsigned)x.cs(base);
    const double mu = EGP_CONST_M.con_mu;
    for (int k = 0; k < cnt; k++) {
        const unsigned eb = (flags >> (4 * k)) & 15u;
        if (!eb) continue;
        const int r0 = base + 1 + 8 * k;
        const double c[3] = {x.cs(r0), x.cs(r0 + 1), x.cs(r0 + 2)}, D = x.cs(r0 + 3);
        double pn[6], p1[6], p2[6];
        cons_dirs(c, pn, p1, p2);
#pragma unroll
        for (int ed = 0; ed < 4; ed++) {
            if (!((eb >> ed) & 1u)) continue;
            const double sg = (ed & 1) ? -mu : mu, fa = D * x.cs(r0 + 4 + ed);
            double p[6];
#pragma unroll
            for (int r = 0; r < 6; r++) p[r] = pn[r] + sg * (ed < 2 ? p1[r] : p2[r]);
#pragma unroll
            for (int r = 0; r < 6; r++) {
#pragma unroll
                for (int cc = r; cc < 6; cc++) w.IA[sx(r, cc)] += D * p[r] * p[cc];
                w.pA[r] -= fa * p[r];
            }
        }
    }
}

// forward (solve-only) sweep: re-evaluate the contact rows of body b at its acceleration a = J_body qacc
template <class X>
EGP_HD void cons_fwd_contacts(const X &x, const double *a, int b) {
    const int slot = EGP_CONST_M.body_slot[b], cnt = (x.ccnt >> (3 * slot)) & 7u, base = CS_PER_BODY * b;
    if (!cnt) return;
    const unsigned flags = (unsigned)x.cs(base);
    const double mu = EGP_CONST_M.con_mu;
    unsigned nf = 0;
    for (int k = 0; k < cnt; k++) {
        const int r0 = base + 1 + 8 * k;
        const double c[3] = {x.cs(r0), x.cs(r0 + 1), x.cs(r0 + 2)};
        double pn[6], p1[6], p2[6];
        cons_dirs(c, pn, p1, p2);
        const double an = dot6(pn, a), a1 = dot6(p1, a), a2 = dot6(p2, a);
#pragma unroll
        for (int ed = 0; ed < 4; ed++)
            if (an + ((ed & 1) ? -mu : mu) * (ed < 2 ? a1 : a2) - x.cs(r0 + 4 + ed) < 0.0) nf |= 1u << (4 * k + ed);
    }
    if (nf != flags) {
        x.cchg = 1;
        if (x.cit < CONS_CAREFUL) x.cs(base) = (double)nf;
        else if (cons_may_flip(x)) {            // one edge only: the lowest disagreeing bit
            const unsigned df = nf ^ flags;
            x.cs(base) = (double)(flags ^ (df & (0u - df)));
        }
    }
}

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

// operands of one backward body-step (hinge body), requested while the previous body-step finishes
struct BwdIn {
    int da, nd, rec;
    int b;
    double kp[3], g[3];                 // stable PD: kp and g = kp q + kd v (rhs = kp ctrl - g - C)
    double ax[3][3], anc[3];
    int tm[32], tmt[8];                 // MODE 0: fb | cin, tau (+1).  MODE 1: cin | ctrl | C
};

// the scratch record (longest latency) is requested between the two phases of the previous body-step, the shared rows
// (shorter latency, 30 more registers) when its carry update is done
template <class X>
EGP_HD void t5_bwd_load_tm(const X &x, int b, const int MODE, BwdIn &in) {
    const BodyK &K = EGP_CONST_M.bk[b];
    in.b = b; in.da = K.da; in.nd = K.nd; in.rec = K.rec;
    if (MODE == 0) { x.template tld_issue<32>(K.rec + R_FB, in.tm); x.template tld_issue<8>(K.rec + R_TAU, in.tmt); }
    else x.template tld_issue<32>(K.rec + R_CIN, in.tm);
}

template <class X>
EGP_HD void t5_bwd_load_rows(const X &x, int b, const int MODE, BwdIn &in) {
    const BodyK &K = EGP_CONST_M.bk[b];
#pragma unroll
    for (int r = 0; r < 3; r++) in.anc[r] = x.at(x.o.anc, 3 * b + r);
    // branch-free: a 1-dof body re-reads its dof 0 for j = 1, 2 (never used); every field is written on every path, which
    // keeps the operand record in registers (a partially written aggregate would be demoted to local memory)
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const int jj = j < K.nd ? j : 0, i = K.da + jj;
#pragma unroll
        for (int r = 0; r < 3; r++) in.ax[j][r] = x.at(x.o.ax, 3 * i + r);
        if (MODE == 1) { in.kp[j] = K.kp[jj]; in.g[j] = K.kp[jj] * x.at(x.o.q, i + 1) + K.kd[jj] * x.at(x.o.v, i); }
    }
}

// one hinge body; `in` holds this body's operands on entry and (next >= 0) the next body's on exit: they are requested
// between the two phases of the block elimination, when this body's own operands are dead and the carry update
// (the longest stretch of arithmetic that needs no memory) is about to start
template <int ND, class X>
EGP_HD void t5_bwd_body(const X &x, Bwd &w, BwdIn &in, const int MODE, const int next) {
    const int da = in.da, rec = in.rec;
    double ci[10], S[ND][6], diag[ND], rhs[ND], y[ND], U[ND][6], Di[ND][ND], u[ND];
    if (MODE == 0) {
#pragma unroll
        for (int k = 0; k < 6; k++) w.F[k] += X::unpack(in.tm, k);
#pragma unroll
        for (int k = 0; k < 10; k++) ci[k] = X::unpack(in.tm, 6 + k);
    } else {
#pragma unroll
        for (int k = 0; k < 10; k++) ci[k] = X::unpack(in.tm, k);
    }
    add_body_inertia(w, ci);
    if constexpr (X::CONS) { if (MODE == 0) cons_bwd_contacts(x, w, in.b); }
#pragma unroll
    for (int j = 0; j < ND; j++) {
        S[j][0] = in.ax[j][0]; S[j][1] = in.ax[j][1]; S[j][2] = in.ax[j][2];
        cross3(in.anc, S[j], S[j] + 3);
        diag[j] = MODE == 1 ? EGP_CONST_M.bk[in.b].armkd[j] : EGP_CONST_M.bk[in.b].arm[j];
    }
    if (MODE == 0) {
        double C[ND];
#pragma unroll
        for (int j = 0; j < ND; j++) { C[j] = dot6(S[j], w.F); rhs[j] = X::unpack(in.tmt, j) - C[j]; }
        x.template tst<ND>(rec + R_C, C);
        if constexpr (X::CONS) {
            // range rows of this body's joints: a violated range adds its weight to the pivot and D s aref to the right-hand side
            const int slot = EGP_CONST_M.body_slot[in.b];
#pragma unroll
            for (int j = 0; j < ND; j++) {
                const unsigned bit = 1u << (3 * slot + j);
                double sgn, Dc, ar;
                if (!cons_limit(da + j, x.at(x.o.q, da + j + 1), x.at(x.o.v, da + j), sgn, Dc, ar)) continue;
                if (x.cit == 0) {           // first guess: the row's state at the end of the previous sub-step, active if it is new
                    x.lim_inst |= bit;
                    if (!(x.lim_pinst & bit) || (x.lim_pact & bit)) x.lim_act |= bit;
                }
                if (x.lim_act & bit) { diag[j] += Dc; rhs[j] += Dc * sgn * ar; }
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < ND; j++) rhs[j] = in.kp[j] * X::unpack(in.tm, 10 + j) - in.g[j] - X::unpack(in.tm, 13 + j);
    }
    blk_backward_a<ND, 0>(w, S, diag, rhs, U, Di, u);
    if (EGP_PF_BWD && next >= 0) t5_bwd_load_tm(x, next, MODE, in);
    blk_backward_b<ND, false>(w, U, Di, u, y, [&](int j, int r, double v) { x.at(x.o.U, 6 * (da + j) + r) = v; });
    x.template tst<ND>(rec + R_Y, y);
    if (EGP_PF_BWD == 1 && next >= 0) t5_bwd_load_rows(x, next, MODE, in);
}

// free-joint root: rotational block (dofs 3..5, S = [R e_k; 0]) then translational block (dofs 0..2, S = [0; e_k]);
// no gains, no armature, no actuation on the root (humanoid_v1.py:137-140)
template <class X>
EGP_HD void t5_bwd_root(const X &x, Bwd &w, int b, const int MODE) {
    const BodyK &K = EGP_CONST_M.bk[b];
    const int da = K.da, rec = K.rec;
    int tm[32], tmc[16];
    x.template tld_issue<32>(rec + R_FB, tm);
    if (MODE == 1) x.template tld_issue<16>(rec + R_ROOT_C, tmc);
    double S[3][6];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const int i = da + 3 + j;
        S[j][0] = x.at(x.o.ax, 3 * i); S[j][1] = x.at(x.o.ax, 3 * i + 1); S[j][2] = x.at(x.o.ax, 3 * i + 2);
        S[j][3] = S[j][4] = S[j][5] = 0.0;
    }
    x.template tld_wait<32>(tm);
    if (MODE == 1) x.template tld_wait<16>(tmc);
    double ci[10];
#pragma unroll
    for (int k = 0; k < 10; k++) ci[k] = X::unpack(tm, 6 + k);
    add_body_inertia(w, ci);
    if constexpr (X::CONS) { if (MODE == 0) cons_bwd_contacts(x, w, b); }
    if (MODE == 0) {
#pragma unroll
        for (int k = 0; k < 6; k++) w.F[k] += X::unpack(tm, k);
    }
    double y[6], C[6], rhs[3], diag[3] = {K.arm[0], K.arm[1], K.arm[2]};   // free joint: armature of the 6 dofs is one value
    {
#pragma unroll
        for (int j = 0; j < 3; j++) {
            C[3 + j] = MODE == 0 ? dot3(S[j], w.F) : X::unpack(tmc, 3 + j);
            rhs[j] = -C[3 + j];
        }
        double yb[3], U[3][6], Di[3][3], u[3];
        blk_backward_a<3, 1>(w, S, diag, rhs, U, Di, u);
        blk_backward_b<3, false>(w, U, Di, u, yb, [&](int j, int r, double v) { x.at(x.o.U, 6 * (da + 3 + j) + r) = v; });
        y[3] = yb[0]; y[4] = yb[1]; y[5] = yb[2];
    }
    {
        double Sz[3][6] = {};
#pragma unroll
        for (int j = 0; j < 3; j++) {
            C[j] = MODE == 0 ? w.F[3 + j] : X::unpack(tmc, j);
            rhs[j] = -C[j];
        }
        double yb[3], U[3][6], Di[3][3], u[3];
        blk_backward_a<3, 2>(w, Sz, diag, rhs, U, Di, u);
        blk_backward_b<3, true>(w, U, Di, u, yb, [&](int j, int r, double v) { x.at(x.o.U, 6 * (da + j) + r) = v; });
        y[0] = yb[0]; y[1] = yb[1]; y[2] = yb[2];
    }
    if (MODE == 0) x.template tst<6>(rec + R_ROOT_C, C);
    x.template tst<6>(rec + R_ROOT_Y, y);
}

template <class X>
EGP_HD void t5_bwd_chain(const X &x, int c, Bwd &w, const int MODE) {
    const DevModel &M = EGP_CONST_M;
    const int hi = M.chain_hi[c], lo = M.chain_lo[c];
    const bool has_root = M.bk[lo].kind == BK_ROOT;
    const int lo_h = has_root ? lo + 1 : lo;            // hinge bodies [lo_h, hi]
    BwdIn in;
    if (EGP_PF_BWD && hi >= lo_h) { t5_bwd_load_tm(x, hi, MODE, in); if (EGP_PF_BWD == 1) t5_bwd_load_rows(x, hi, MODE, in); }
#pragma unroll 1
    for (int b = hi; b >= lo_h; b--) {
        if (!EGP_PF_BWD) t5_bwd_load_tm(x, b, MODE, in);       // all operands of the body requested at once
        if (EGP_PF_BWD != 1) t5_bwd_load_rows(x, b, MODE, in);
        x.template tld_wait<32>(in.tm);
        if (MODE == 0) x.template tld_wait<8>(in.tmt);
        const int next = b > lo_h ? b - 1 : -1;
        if (in.nd == 3) t5_bwd_body<3>(x, w, in, MODE, next);
        else t5_bwd_body<1>(x, w, in, MODE, next);
    }
    if (has_root) t5_bwd_root(x, w, lo, MODE);
    x.twait_st();
    if (M.chain_parent[c] >= 0) {
        const int base = x.o.jb + 33 * M.chain_cslot[c];
#pragma unroll
        for (int k = 0; k < 21; k++) x.at(base, k) = w.IA[k];
#pragma unroll
        for (int k = 0; k < 6; k++) { x.at(base, 21 + k) = w.pA[k]; x.at(base, 27 + k) = w.F[k]; }
    }
}

// ---- forward sweep ---------------------------------------------------------------------------------
// operands of one forward body-step (hinge body), requested while the previous body-step finishes
template <int MODE>
struct FwdIn {
    int b, da, qa, nd, kind, rec, xp_slot;
    double W[3][6], ax[3][3], anc[3];       // solve part: rows currently in storage (the OLD tree rows for MODE 1)
    double q[3], v[3];
    double sn[3], cs[3];                    // kinematics part: sin / cos of the joint angles
    int tmy[8], tmc[8];                     // y (solve), ctrl (MODE 1)
};

template <int MODE, class X>
EGP_HD void t5_fwd_load(const X &x, int b, FwdIn<MODE> &in) {
    const BodyK &K = EGP_CONST_M.bk[b];
    in.b = b; in.da = K.da; in.qa = K.qa; in.nd = K.nd; in.kind = K.kind; in.rec = K.rec; in.xp_slot = K.xp_slot;
    if (MODE != 2) x.template tld_issue<8>(K.rec + R_Y - 2, in.tmy);         // tau[2] | y      (MODE 3: solve only, like 0 without Euler)
    if (MODE == 1) x.template tld_issue<8>(K.rec + R_CTRL, in.tmc);          // ctrl | C[0]
    if (MODE != 2) {
#pragma unroll
        for (int r = 0; r < 3; r++) in.anc[r] = x.at(x.o.anc, 3 * b + r);
    }
    // branch-free (see t5_bwd_load): a 1-dof body re-reads its dof 0 for j = 1, 2
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const int jj = j < K.nd ? j : 0, i = K.da + jj;
        if (MODE != 2) {
#pragma unroll
            for (int r = 0; r < 6; r++) in.W[j][r] = x.at(x.o.U, 6 * i + r);
#pragma unroll
            for (int r = 0; r < 3; r++) in.ax[j][r] = x.at(x.o.ax, 3 * i + r);
        }
        in.v[j] = x.at(x.o.v, i);
        in.q[j] = x.at(x.o.q, K.qa + jj);
        if (MODE == 1 || MODE == 2) {            // through temporaries: handing sincos() addresses of `in` members forces the whole record into local memory
            double sn_j, cs_j;
            sincos(in.q[j], &sn_j, &cs_j);
            in.sn[j] = sn_j; in.cs[j] = cs_j;
        }
    }
}

// what the kinematics part of a body-step keeps of its operands while the next body's are being requested
struct KinIn {
    int b, da, rec, xp_slot;
    double v[3], sn[3], cs[3];
};

// solve part of a hinge body: x = y - W^T a, a += S x, then torque (MODE 1) or semi-implicit Euler (MODE 0)
template <int ND, int MODE, class X>
EGP_HD void t5_fwd_solve_body(const X &x, double *a, const FwdIn<MODE> &in) {
    const double h = EGP_CONST_M.h;
    double xs[ND], S[ND][6];
#pragma unroll
    for (int j = 0; j < ND; j++) {
        xs[j] = X::unpack(in.tmy, 1 + j) - dot6(in.W[j], a);
        S[j][0] = in.ax[j][0]; S[j][1] = in.ax[j][1]; S[j][2] = in.ax[j][2];
        cross3(in.anc, S[j], S[j] + 3);
    }
#pragma unroll
    for (int r = 0; r < 6; r++) {
        double t = a[r];
#pragma unroll
        for (int j = 0; j < ND; j++) t += S[j][r] * xs[j];
        a[r] = t;
    }
    if (MODE == 1) {                // torque = clip(-kp e - kd (v + x h))  (humanoid_v1.py:152-155,172)
        double tq[ND];
#pragma unroll
        for (int j = 0; j < ND; j++) {
            const double eq = in.q[j] - X::unpack(in.tmc, j);
            const BodyK &K = EGP_CONST_M.bk[in.b];
            const double t = -K.kp[j] * eq - K.kd[j] * (in.v[j] + xs[j] * h);
            const double lim = K.tlim[j];
            tq[j] = t < -lim ? -lim : (t > lim ? lim : t);
        }
        x.template tst<ND>(in.rec + R_TAU, tq);
    } else if (MODE == 3) {
        // constrained solve: keep qacc in the record's y slot (this body has consumed y), re-evaluate the rows of this body
        x.template tst<ND>(in.rec + R_Y, xs);
        if constexpr (X::CONS) {
            const int slot = EGP_CONST_M.body_slot[in.b];
#pragma unroll
            for (int j = 0; j < ND; j++) {
                const unsigned bit = 1u << (3 * slot + j);
                if (!(x.lim_inst & bit)) continue;
                double sgn, Dc, ar;
                cons_limit(in.da + j, in.q[j], in.v[j], sgn, Dc, ar);
                const bool on = sgn * xs[j] - ar < 0.0;
                if (on != ((x.lim_act & bit) != 0)) { x.cchg = 1; if (cons_may_flip(x)) x.lim_act ^= bit; }
            }
            cons_fwd_contacts(x, a, in.b);
        }
    } else {
#pragma unroll
        for (int j = 0; j < ND; j++) {
            const double vn = in.v[j] + h * xs[j];
            x.at(x.o.v, in.da + j) = vn;
            x.at(x.o.q, in.qa + j) = in.q[j] + h * vn;
        }
    }
}

template <int MODE, class X>
EGP_HD void t5_fwd_solve_root(const X &x, double *a, int b) {
    const BodyK &K = EGP_CONST_M.bk[b];
    const int da = K.da, qa = K.qa;
    const double h = EGP_CONST_M.h;
    int tmy[16];
    x.template tld_issue<16>(K.rec + R_ROOT_Y - 4, tmy);       // C[4], C[5] | y 6
    double y[6], xs[6], W[6][6], R3[3][3];
#pragma unroll
    for (int j = 0; j < 6; j++)
#pragma unroll
        for (int r = 0; r < 6; r++) W[j][r] = x.at(x.o.U, 6 * (da + j) + r);
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int r = 0; r < 3; r++) R3[j][r] = x.at(x.o.ax, 3 * (da + 3 + j) + r);
    x.template tld_wait<16>(tmy);
#pragma unroll
    for (int j = 0; j < 6; j++) y[j] = X::unpack(tmy, 2 + j);
    // translational block first (the root has no parent: a = 0 on entry)
#pragma unroll
    for (int j = 0; j < 3; j++) xs[j] = y[j] - dot6(W[j], a);
#pragma unroll
    for (int j = 0; j < 3; j++) a[3 + j] += xs[j];
#pragma unroll
    for (int j = 0; j < 3; j++) xs[3 + j] = y[3 + j] - dot6(W[3 + j], a);
#pragma unroll
    for (int j = 0; j < 3; j++) { a[0] += R3[j][0] * xs[3 + j]; a[1] += R3[j][1] * xs[3 + j]; a[2] += R3[j][2] * xs[3 + j]; }
    if (MODE == 3) {
        x.template tst<6>(K.rec + R_ROOT_Y, xs);
        if constexpr (X::CONS) cons_fwd_contacts(x, a, b);
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

// body done: world position record, spatial inertia about O in world axes, RNE body force -> scratch record
template <class X>
EGP_HD void t5_body_finish(const X &x, const Fwd &f, int b, int rec, int xp_slot) {
    const BodyK &K = EGP_CONST_M.bk[b];
    if (xp_slot >= 0)
#pragma unroll
        for (int r = 0; r < 3; r++) x.at(x.o.xp, 3 * xp_slot + r) = f.p[r] + x.at(x.o.q, r);
    double cpos[3];
    const double ip0 = K.ipos[0], ip1 = K.ipos[1], ip2 = K.ipos[2];
#pragma unroll
    for (int r = 0; r < 3; r++) cpos[r] = f.p[r] + f.R[3 * r] * ip0 + f.R[3 * r + 1] * ip1 + f.R[3 * r + 2] * ip2;
    const double *in = K.inertia;
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
    const double mass = K.mass, cc2 = dot3(cpos, cpos);
    double rc[16];                  // fb 6 | cin 10, the record's leading 32 columns
    double *ci = rc + 6;
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
#pragma unroll
    for (int r = 0; r < 3; r++) {
        rc[r] = Ia[r] + c0[r] + c1[r];
        rc[3 + r] = Ia[3 + r] + c2[r];
    }
    x.template tst<16>(rec + R_FB, rc);
}

// kinematics refresh of a hinge body: ND joints about coordinate axes (A0 + j) % 3 of the successively rotated
// frame, all through one anchor (MuJoCo kinematics)
template <int ND, int A0, class X>
EGP_HD void t5_fwd_kin_body(const X &x, Fwd &f, const KinIn &in) {
    const BodyK &K = EGP_CONST_M.bk[in.b];
    const double bp0 = K.pos[0], bp1 = K.pos[1], bp2 = K.pos[2];
    const double da0 = K.anchor[0], da1 = K.anchor[1], da2 = K.anchor[2];
    double anc[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        f.p[r] += f.R[3 * r] * bp0 + f.R[3 * r + 1] * bp1 + f.R[3 * r + 2] * bp2;
        anc[r] = f.p[r] + f.R[3 * r] * da0 + f.R[3 * r + 1] * da1 + f.R[3 * r + 2] * da2;
        x.at(x.o.anc, 3 * in.b + r) = anc[r];
    }
#pragma unroll
    for (int j = 0; j < ND; j++) {
        const int aid = (A0 + j) % 3, c1 = (aid + 1) % 3, c2 = (aid + 2) % 3;
        const int i = in.da + j;
        const double qd = in.v[j];
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
            f.a[r] += t0[r] * qd;
            f.a[3 + r] += (t1[r] + t2[r]) * qd;
        }
#pragma unroll
        for (int r = 0; r < 6; r++) f.v[r] += S[r] * qd;
        // rotate the frame about the joint axis through the anchor
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const double a1 = f.R[3 * r + c1], a2 = f.R[3 * r + c2];
            f.R[3 * r + c1] = in.cs[j] * a1 + in.sn[j] * a2;
            f.R[3 * r + c2] = -in.sn[j] * a1 + in.cs[j] * a2;
        }
    }
#pragma unroll
    for (int r = 0; r < 3; r++) f.p[r] = anc[r] - (f.R[3 * r] * da0 + f.R[3 * r + 1] * da1 + f.R[3 * r + 2] * da2);
}

// free joint: O = root position, so the frame origin is 0 and rotational cdofs have no linear part
template <class X>
EGP_HD void t5_fwd_kin_root(const X &x, Fwd &f, int b) {
    const DevModel &M = EGP_CONST_M;
    const int da = M.bk[b].da, qa = M.bk[b].qa;
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
    constexpr bool KIN = MODE == 1 || MODE == 2;      // kinematics refresh; 0 / 1 / 3 solve
    if (KIN && pc >= 0) {
        const int base = x.o.jf + 24 * M.chain_pslot[pc];
#pragma unroll
        for (int k = 0; k < 3; k++) f.p[k] = x.at(base, k);
#pragma unroll
        for (int k = 0; k < 9; k++) f.R[k] = x.at(base, 3 + k);
#pragma unroll
        for (int k = 0; k < 6; k++) { f.v[k] = x.at(base, 12 + k); f.a[k] = x.at(base, 18 + k); }
    }
    const int lo = M.chain_lo[c], hi = M.chain_hi[c];
    const bool has_root = M.bk[lo].kind == BK_ROOT;
    const int lo_h = has_root ? lo + 1 : lo;
    if (MODE == 0) { EGP_CLK_MARK(8) }
    if (has_root) {
        if (MODE != 2) t5_fwd_solve_root<MODE>(x, a, lo);
        if (KIN) {
            t5_fwd_kin_root(x, f, lo);
            t5_body_finish(x, f, lo, M.bk[lo].rec, M.bk[lo].xp_slot);
            if constexpr (X::CONS) { if (MODE == 1) cons_collide(x, f, lo); }
        }
    }
    if (MODE == 0) { EGP_CLK_MARK(9) }
    constexpr bool PF = (MODE == 0 && EGP_PF_FWD0) || (MODE == 1 && EGP_PF_FWD1);
    FwdIn<MODE> in;
    if (PF && hi >= lo_h) t5_fwd_load<MODE>(x, lo_h, in);
#pragma unroll 1
    for (int b = lo_h; b <= hi; b++) {
        if (!PF) t5_fwd_load<MODE>(x, b, in);
        if (MODE != 2) x.template tld_wait<8>(in.tmy);
        if (MODE == 1) x.template tld_wait<8>(in.tmc);
        const int kind = in.kind;
        KinIn kin;
        kin.b = in.b; kin.da = in.da; kin.rec = in.rec; kin.xp_slot = in.xp_slot;
#pragma unroll
        for (int j = 0; j < 3; j++) { kin.v[j] = in.v[j]; kin.sn[j] = in.sn[j]; kin.cs[j] = in.cs[j]; }
        if (MODE == 0 && PF) {
            // accel + Euler only: little arithmetic per body, so the next body's operands are requested a full body ahead
            FwdIn<MODE> cur;
            cur.da = in.da; cur.qa = in.qa; cur.nd = in.nd; cur.rec = in.rec;
#pragma unroll
            for (int j = 0; j < 3; j++) {
#pragma unroll
                for (int r = 0; r < 6; r++) cur.W[j][r] = in.W[j][r];
#pragma unroll
                for (int r = 0; r < 3; r++) cur.ax[j][r] = in.ax[j][r];
                cur.anc[j] = in.anc[j]; cur.q[j] = in.q[j]; cur.v[j] = in.v[j];
            }
#pragma unroll
            for (int k = 0; k < 8; k++) cur.tmy[k] = in.tmy[k];
            if (b < hi) t5_fwd_load<MODE>(x, b + 1, in);
            if (cur.nd == 3) t5_fwd_solve_body<3, MODE>(x, a, cur);
            else t5_fwd_solve_body<1, MODE>(x, a, cur);
        } else {
            if (MODE != 2) {
                if (in.nd == 3) t5_fwd_solve_body<3, MODE>(x, a, in);
                else t5_fwd_solve_body<1, MODE>(x, a, in);
            }
            // this body's solve operands are dead: request the next body's while the kinematics / body-force arithmetic runs
            if (PF && b < hi) t5_fwd_load<MODE>(x, b + 1, in);
        }
        if (KIN) {
            if (kind == BK_XYZ) t5_fwd_kin_body<3, 0>(x, f, kin);
            else if (kind == BK_X) t5_fwd_kin_body<1, 0>(x, f, kin);
            else if (kind == BK_Y) t5_fwd_kin_body<1, 1>(x, f, kin);
            else t5_fwd_kin_body<1, 2>(x, f, kin);
            t5_body_finish(x, f, kin.b, kin.rec, kin.xp_slot);
            if constexpr (X::CONS) { if (MODE == 1) cons_collide(x, f, kin.b); }
        }
    }
    if (MODE == 0) { EGP_CLK_MARK(10) }
    x.twait_st();
    if (M.chain_pslot[c] >= 0) {
        if (MODE != 2) {
#pragma unroll
            for (int k = 0; k < 6; k++) x.at(x.o.ja + 6 * M.chain_pslot[c], k) = a[k];
        }
        if (KIN) {
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

// semi-implicit Euler of one chain from the accelerations a solve-only forward sweep (MODE 3) left in the y slots
template <class X>
EGP_HD void t5_integrate_chain(const X &x, int c) {
    const DevModel &M = EGP_CONST_M;
    const double h = M.h;
    for (int b = M.chain_lo[c]; b <= M.chain_hi[c]; b++) {
        const BodyK &K = M.bk[b];
        if (K.kind == BK_ROOT) {
            const int da = K.da, qa = K.qa;
            int tmy[16];
            x.template tld_issue<16>(K.rec + R_ROOT_Y - 4, tmy);
            x.template tld_wait<16>(tmy);
            double vn[6];
#pragma unroll
            for (int j = 0; j < 6; j++) { vn[j] = x.at(x.o.v, da + j) + h * X::unpack(tmy, 2 + j); x.at(x.o.v, da + j) = vn[j]; }
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
        } else {
            int tmy[8];
            x.template tld_issue<8>(K.rec + R_Y - 2, tmy);
            x.template tld_wait<8>(tmy);
            for (int j = 0; j < K.nd; j++) {
                const double vn = x.at(x.o.v, K.da + j) + h * (j == 0 ? X::unpack(tmy, 1) : (j == 1 ? X::unpack(tmy, 2) : X::unpack(tmy, 3)));
                x.at(x.o.v, K.da + j) = vn;
                x.at(x.o.q, K.qa + j) += h * vn;
            }
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
    // scratch records: position of each body among those owned by the same warp
    int nb_w[T4_CW] = {0};
    for (int b = 0; b < nb; b++) d.body_slot[b] = nb_w[d.chain_warp[d.body_chain[b]]]++;
    int NB = 0;
    for (int w = 0; w < T4_CW; w++) if (nb_w[w] > NB) NB = nb_w[w];
    d.tm_cols = R_COLS_HOST * NB;
    if (d.tm_cols > 512) d.t4_ok = 0;
    for (int b = 0; b < nb; b++) {
        BodyK &K = d.bk[b];
        memset(&K, 0, sizeof K);
        const int da = d.body_dofadr[b], nd = d.body_dofnum[b];
        K.da = da; K.qa = d.body_qposadr[b]; K.nd = nd; K.kind = d.body_kind[b]; K.rec = R_COLS_HOST * d.body_slot[b];
        K.xp_slot = d.body_xp_slot[b];
        for (int k = 0; k < 3; k++) { K.pos[k] = d.body_pos[b][k]; K.ipos[k] = d.body_ipos[b][k]; K.anchor[k] = d.dof_anchor[da][k]; }
        for (int k = 0; k < 6; k++) K.inertia[k] = d.body_inertia[b][k];
        K.mass = d.body_mass[b];
        for (int j = 0; j < 3 && j < nd; j++) {
            const int i = da + j;
            K.arm[j] = d.dof_arm[i]; K.armkd[j] = d.dof_arm[i] + d.kd[i] * d.h; K.kp[j] = d.kp[i]; K.kd[j] = d.kd[i]; K.tlim[j] = d.tlim[i];
        }
        for (int j = 0; j < nd; j++) {
            const int i = da + j;
            if (b == 0) { d.dof_col_ctrl[i] = -1; d.dof_col_tau[i] = -1; d.dof_col_c[i] = K.rec + 32 + 2 * j; }
            else { d.dof_col_ctrl[i] = K.rec + 32 + 2 * j; d.dof_col_c[i] = K.rec + 38 + 2 * j; d.dof_col_tau[i] = K.rec + 44 + 2 * j; }
        }
    }
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
