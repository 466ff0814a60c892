// Fused humanoid rollout (K1 physics + K2 policy/sample + K3 reward/done/reset/trajbatch) - sm_100a, f64.
//
// Replaces Agent.sample / sample_worker (agents/agent.py:29-111) + HumanoidEnv.step/reset
// (ego_pose/envs/humanoid_v1.py:130-231) + quat_space_reward_v3 (ego_pose/core/reward_function.py:4-60)
// + the MuJoCo calls behind them (mj_forward / mj_step / mj_fullM) + the dense LAPACK solve of the
// stable-PD controller.
//
// Dynamics formulation (same results as the reference's CRBA + dense Cholesky, O(n) instead of O(n^3)):
//   * every spatial quantity is expressed in world axes about the point O = current root position
//   * bias force C = RNE(q, v, 0) by one forward (velocities / bias accelerations) and one backward
//     (force accumulation) sweep over the kinematic tree
//   * both linear solves  qacc = M^-1 (tau - C)  and  q**  = (M + Kd h)^-1 rhs  are articulated-body
//     sweeps (Featherstone): the diagonal terms (armature, Kd h) enter the joint-space pivots D_i
//   * the stable-PD solve of sub-step k runs BEFORE the kinematics refresh, i.e. on the tree data of
//     sub-step k-1 - exactly the one-sub-step staleness of data.qM / data.qfrc_bias in the reference
//     (humanoid_v1.py:134-136); head / end-effector positions are taken from the
//     last refresh (state before the last sub-step, appendix C.2)
// Mapping (V1): one thread per environment, CTA = 32 environments; tree data in thread-local arrays,
// policy activations staged in shared memory [feature][env] (bank-conflict free), model constants in
// __constant__ memory (warp-uniform indices -> broadcast), policy weights pre-transposed/padded and read
// through the uniform L1/L2 path.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>
#include <vector>

#include "common.cuh"

#ifdef EGP_T4_CLK
namespace egp { __device__ unsigned long long g_t4_clk2[16]; __device__ long long g_t4_last; }
#define EGP_CLK_MARK(slot) { long long _t = clock64(); if (blockIdx.x == 0 && threadIdx.x == 0) { egp::g_t4_clk2[slot] += _t - egp::g_t4_last; egp::g_t4_last = _t; } }
#endif
#include "tree.cuh"
#include "dmma.cuh"

namespace egp {

constexpr int ENVS_PER_CTA = 32;
constexpr int JB = 8;                       // output neurons per register block in the policy MLP

}  // namespace egp

struct EgpModel {
    egp::DevModel host;
    int device;
    int n_takes, ctx_dim;
    long long total_frames;
    int32_t *d_take_off;
    double *d_rows, *d_head_lb, *d_ctx;
    // transposed / padded policy weights (scratch owned by the model)
    double *d_wbuf;
    size_t wbuf_elems;
    double *d_cons;             // constraint scratch of the block-sweep kernel (joint limits / floor contact)
    size_t cons_elems;
};

namespace egp {

static const EgpModel *g_bound_model[64] = {nullptr};

static int bind_model(const EgpModel *m) {
    int dev = m->device;
    if (dev >= 0 && dev < 64 && g_bound_model[dev] == m) return EGP_OK;
    EGP_CUDA(cudaMemcpyToSymbol(c_m, &m->host, sizeof(DevModel)));
    if (dev >= 0 && dev < 64) g_bound_model[dev] = m;
    return EGP_OK;
}

// ------------------------------------------------------------------------------------------------
// small device math
// quaternion helpers (w, x, y, z), utils/transformation.py:1379-1421 conventions
__device__ __forceinline__ void quat_inv(const double *q, double *o) {
    double n = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    o[0] = q[0] / n; o[1] = -q[1] / n; o[2] = -q[2] / n; o[3] = -q[3] / n;
}
// utils/math.py:62-67,80-81: heading quaternion (w,0,0,z)/|.| and de_heading
__device__ __forceinline__ void heading_cs(const double *q, double *hw, double *hz) {
    double n = sqrt(q[0] * q[0] + q[3] * q[3]);
    *hw = q[0] / n; *hz = q[3] / n;
}
__device__ __forceinline__ void de_heading(const double *q, double *o) {
    double hw, hz;
    heading_cs(q, &hw, &hz);
    double hq[4] = {hw, 0.0, 0.0, hz}, ih[4];
    quat_inv(hq, ih);
    quat_mul(ih, q, o);
}
// utils/math.py:47-59 transform_vec(v, q, 'heading'): R(hq)^T v with quaternion_matrix's normalisation
__device__ __forceinline__ void to_heading(const double *v, const double *q, double *o) {
    double hw, hz;
    heading_cs(q, &hw, &hz);
    double n = hw * hw + hz * hz, s = 2.0 / n;
    double czz = hz * hz * s, cwz = hw * hz * s;        // R = [[1-czz, -cwz, 0], [cwz, 1-czz, 0], [0, 0, 1]]
    double x = (1.0 - czz) * v[0] + cwz * v[1];
    double y = -cwz * v[0] + (1.0 - czz) * v[1];
    o[0] = x; o[1] = y; o[2] = v[2];
}
// utils/math.py:47-59 transform_vec(v, q, 'root')
__device__ __forceinline__ void to_root(const double *v, const double *quat, double *o) {
    double n = quat[0] * quat[0] + quat[1] * quat[1] + quat[2] * quat[2] + quat[3] * quat[3];
    double s = sqrt(2.0 / n);
    double q0 = quat[0] * s, q1 = quat[1] * s, q2 = quat[2] * s, q3 = quat[3] * s;
    double R0 = 1.0 - q2 * q2 - q3 * q3, R1 = q1 * q2 - q3 * q0, R2 = q1 * q3 + q2 * q0;
    double R3 = q1 * q2 + q3 * q0, R4 = 1.0 - q1 * q1 - q3 * q3, R5 = q2 * q3 - q1 * q0;
    double R6 = q1 * q3 - q2 * q0, R7 = q2 * q3 + q1 * q0, R8 = 1.0 - q1 * q1 - q2 * q2;
    double x = R0 * v[0] + R3 * v[1] + R6 * v[2], y = R1 * v[0] + R4 * v[1] + R7 * v[2], z = R2 * v[0] + R5 * v[1] + R8 * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
// utils/transformation.py:348-356 rotation_from_quaternion -> axis * angle
__device__ __forceinline__ void rot_from_quat(const double *q, double *o, double *angle_out) {
    if (1.0 - q[0] < 1e-8) { o[0] = o[1] = o[2] = 0.0; *angle_out = 0.0; return; }
    double s = sqrt(1.0 - q[0] * q[0]), ang = 2.0 * acos(q[0]);
    o[0] = q[1] / s; o[1] = q[2] / s; o[2] = q[3] / s;
    *angle_out = ang;
}
// utils/transformation.py:1194-1248 quaternion_from_euler(.., 'sxyz')
__device__ __forceinline__ void quat_from_euler(double ai, double aj, double ak, double *q) {
    double si, ci, sj, cj, sk, ck;
    sincos(0.5 * ai, &si, &ci);
    sincos(0.5 * aj, &sj, &cj);
    sincos(0.5 * ak, &sk, &ck);
    double cc = ci * ck, cs = ci * sk, sc = si * ck, ss = si * sk;
    q[0] = cj * cc + sj * ss; q[1] = cj * sc - sj * cs; q[2] = cj * ss + sj * cc; q[3] = cj * cs - sj * sc;
}

// per-environment scratch (thread-local)

constexpr int MAXCON = 40;             // floor contacts per environment (21 geoms: 15 capsules x 2, 4 spheres, 2 boxes x 4)

struct EnvData {
    double q[MAXV + 1], v[MAXV];
    double S[MAXV][6], U[MAXV][6], Dinv[MAXV], u[MAXV];
    double C[MAXV], tau[MAXV], qacc[MAXV];
    double dadd[MAXV], radd[MAXV];      // joint limits: extra pivot weight / right-hand side of the active rows (0 otherwise)
    // floor contacts of the last kinematics refresh: body, spatial directions [c x d; d] (about O) of the normal and the two
    // tangents, and per pyramid edge (n + mu t1, n - mu t1, n + mu t2, n - mu t2) the weight D, aref and the active flag
    int ncon, con_body[MAXCON];
    double con_p[MAXCON][3][6], row_D[4 * MAXCON], row_aref[4 * MAXCON];
    unsigned char row_act[4 * MAXCON], row_new[4 * MAXCON];
    double cin[MAXB][10], fb[MAXB][6];
    double xp[MAXB][3];                 // body positions (world) from the last kinematics refresh
    Fwd jf[MAXC];
    Bwd jb[MAXC];
    double ja[MAXC][6];
};

__device__ void limit_row(double invweight, double dist, double vel, double &Dc, double &aref);

// one floor contact of body b at point c (relative to O) with penetration distance dist: the three spatial directions
// and the four pyramid-edge rows (condim 3, friction mu; mj_instantiateContact, mj_diagApprox, mj_makeImpedance)
__device__ void add_contact(EnvData &e, int b, const double *c, double dist, const double *vb /* body velocity [w; v_O] */) {
    if (e.ncon >= MAXCON) return;
    const int k = e.ncon++;
    e.con_body[k] = b;
    const double dirs[3][3] = {{0.0, 0.0, 1.0}, {0.0, 1.0, 0.0}, {-1.0, 0.0, 0.0}};     // normal, tangents (mju_makeFrame)
    for (int a = 0; a < 3; a++) {
        cross3(c, dirs[a], e.con_p[k][a]);
        e.con_p[k][a][3] = dirs[a][0]; e.con_p[k][a][4] = dirs[a][1]; e.con_p[k][a][5] = dirs[a][2];
    }
    const double mu = c_m.con_mu, tran = c_m.body_iw[b];
    const double vn = dot6(e.con_p[k][0], vb), v1 = dot6(e.con_p[k][1], vb), v2 = dot6(e.con_p[k][2], vb);
    for (int ed = 0; ed < 4; ed++) {
        const double vel = vn + ((ed & 1) ? -mu : mu) * (ed < 2 ? v1 : v2);
        double D0, aref;
        limit_row(tran + mu * mu * tran, dist - c_m.con_margin, vel, D0, aref);
        e.row_D[4 * k + ed] = D0 / (2.0 * mu * mu);                                    // pyramidal: R = 2 mu^2 R_0
        e.row_aref[4 * k + ed] = aref;
    }
}

// F1: kinematics + velocities + bias accelerations + body inertias/forces at (q, v); refreshes S, cin, fb, xp.
__device__ void pass_kinematics(EnvData &e) {
    const int nchain = c_m.nchain;
    e.ncon = 0;
    for (int c = 0; c < nchain; c++) {
        Fwd f;
        const int pc = c_m.chain_parent[c];
        if (pc >= 0) f = e.jf[pc];
        for (int b = c_m.chain_lo[c]; b <= c_m.chain_hi[c]; b++) {
            const int da = c_m.body_dofadr[b], nd = c_m.body_dofnum[b], qa = c_m.body_qposadr[b];
            if (nd == 6) {
                // free joint; O = root position, so the frame origin is 0 and rotational cdofs have no linear part
                double qn[4], n = sqrt(e.q[qa + 3] * e.q[qa + 3] + e.q[qa + 4] * e.q[qa + 4] + e.q[qa + 5] * e.q[qa + 5] +
                                       e.q[qa + 6] * e.q[qa + 6]);
                for (int k = 0; k < 4; k++) qn[k] = e.q[qa + 3 + k] / n;
                quat_to_mat(qn, f.R);
                f.p[0] = f.p[1] = f.p[2] = 0.0;
                double wl[3] = {e.v[da + 3], e.v[da + 4], e.v[da + 5]}, ww[3];
                for (int r = 0; r < 3; r++) ww[r] = f.R[3 * r] * wl[0] + f.R[3 * r + 1] * wl[1] + f.R[3 * r + 2] * wl[2];
                for (int k = 0; k < 3; k++) {
                    for (int r = 0; r < 6; r++) { e.S[da + k][r] = 0.0; e.S[da + 3 + k][r] = 0.0; }
                    e.S[da + k][3 + k] = 1.0;
                    for (int r = 0; r < 3; r++) e.S[da + 3 + k][r] = f.R[3 * r + k];
                }
                double vl[3] = {e.v[da], e.v[da + 1], e.v[da + 2]}, vxw[3];
                cross3(vl, ww, vxw);            // sum_k ([0;v] x [R e_k;0]) w_k = [0; v x (R w)]
                for (int k = 0; k < 3; k++) {
                    f.v[k] = ww[k]; f.v[3 + k] = vl[k];
                    f.a[k] = 0.0; f.a[3 + k] = -c_m.grav[k] + vxw[k];
                }
            } else {
                // frame before the joints: parent frame shifted by body_pos
                double off[3];
                for (int r = 0; r < 3; r++)
                    off[r] = f.R[3 * r] * c_m.body_pos[b][0] + f.R[3 * r + 1] * c_m.body_pos[b][1] + f.R[3 * r + 2] * c_m.body_pos[b][2];
                for (int r = 0; r < 3; r++) f.p[r] += off[r];
                for (int j = 0; j < nd; j++) {
                    const int i = da + j;
                    double anc[3], ax[3];
                    for (int r = 0; r < 3; r++) {
                        anc[r] = f.p[r] + f.R[3 * r] * c_m.dof_anchor[i][0] + f.R[3 * r + 1] * c_m.dof_anchor[i][1] +
                                 f.R[3 * r + 2] * c_m.dof_anchor[i][2];
                    }
                    const int aid = c_m.dof_axis_id[i];
                    if (aid >= 0) { ax[0] = f.R[aid]; ax[1] = f.R[3 + aid]; ax[2] = f.R[6 + aid]; }
                    else for (int r = 0; r < 3; r++)
                        ax[r] = f.R[3 * r] * c_m.dof_axis[i][0] + f.R[3 * r + 1] * c_m.dof_axis[i][1] + f.R[3 * r + 2] * c_m.dof_axis[i][2];
                    double S[6];
                    S[0] = ax[0]; S[1] = ax[1]; S[2] = ax[2];
                    cross3(anc, ax, S + 3);
                    for (int r = 0; r < 6; r++) e.S[i][r] = S[r];
                    // cdof_dot = v x S ; a += cdof_dot qd ; v += S qd
                    const double qd = e.v[i];
                    double t0[3], t1[3], t2[3];
                    cross3(f.v, S, t0);
                    cross3(f.v, S + 3, t1);
                    cross3(f.v + 3, S, t2);
                    for (int r = 0; r < 3; r++) {
                        f.a[r] += t0[r] * qd;
                        f.a[3 + r] += (t1[r] + t2[r]) * qd;
                    }
                    for (int r = 0; r < 6; r++) f.v[r] += S[r] * qd;
                    // rotate the frame about the joint axis through the anchor
                    double sn, cs;
                    sincos(e.q[qa + j], &sn, &cs);
                    if (aid >= 0) {
                        const int c1 = (aid + 1) % 3, c2 = (aid + 2) % 3;
                        for (int r = 0; r < 3; r++) {
                            double a1 = f.R[3 * r + c1], a2 = f.R[3 * r + c2];
                            f.R[3 * r + c1] = cs * a1 + sn * a2;
                            f.R[3 * r + c2] = -sn * a1 + cs * a2;
                        }
                    } else {
                        const double *a = c_m.dof_axis[i];
                        double K[9] = {0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0}, Rot[9], Rn[9];
                        for (int r = 0; r < 3; r++) for (int cidx = 0; cidx < 3; cidx++)
                            Rot[3 * r + cidx] = (r == cidx ? cs : 0.0) + sn * K[3 * r + cidx] + (1.0 - cs) * a[r] * a[cidx];
                        for (int r = 0; r < 3; r++) for (int cidx = 0; cidx < 3; cidx++)
                            Rn[3 * r + cidx] = f.R[3 * r] * Rot[cidx] + f.R[3 * r + 1] * Rot[3 + cidx] + f.R[3 * r + 2] * Rot[6 + cidx];
                        for (int r = 0; r < 9; r++) f.R[r] = Rn[r];
                    }
                    for (int r = 0; r < 3; r++)
                        f.p[r] = anc[r] - (f.R[3 * r] * c_m.dof_anchor[i][0] + f.R[3 * r + 1] * c_m.dof_anchor[i][1] +
                                           f.R[3 * r + 2] * c_m.dof_anchor[i][2]);
                }
            }
            // body done: world position, spatial inertia about O in world axes, RNE body force
            for (int r = 0; r < 3; r++) e.xp[b][r] = f.p[r] + e.q[r];
            if (c_m.contacts) {
                // geom against the floor plane z = 0 (mjc_PlaneSphere / PlaneCapsule / PlaneBox); heights are world heights,
                // contact points are kept relative to O = root position like every other spatial quantity
                const double margin = c_m.con_margin, zO = e.q[2];
                const double *sz = c_m.geom_size[b];
                double c0[3];
                for (int r = 0; r < 3; r++)
                    c0[r] = f.p[r] + f.R[3 * r] * c_m.geom_p0[b][0] + f.R[3 * r + 1] * c_m.geom_p0[b][1] + f.R[3 * r + 2] * c_m.geom_p0[b][2];
                if (c_m.geom_type[b] == 2) {
                    int cnt = 0;
                    for (int vtx = 0; vtx < 8 && cnt < 4; vtx++) {
                        const double l0 = (vtx & 1) ? sz[0] : -sz[0], l1 = (vtx & 2) ? sz[1] : -sz[1], l2 = (vtx & 4) ? sz[2] : -sz[2];
                        double wv[3];
                        for (int r = 0; r < 3; r++) wv[r] = c0[r] + f.R[3 * r] * l0 + f.R[3 * r + 1] * l1 + f.R[3 * r + 2] * l2;
                        const double dist = wv[2] + zO;
                        if (dist > margin) continue;
                        wv[2] -= 0.5 * dist;
                        add_contact(e, b, wv, dist, f.v);
                        cnt++;
                    }
                } else {
                    const int ne = c_m.geom_type[b] == 1 ? 2 : 1;
                    for (int en = 0; en < ne; en++) {
                        double cc[3];
                        if (ne == 2 && en == 0) {
                            for (int r = 0; r < 3; r++)
                                cc[r] = f.p[r] + f.R[3 * r] * c_m.geom_p1[b][0] + f.R[3 * r + 1] * c_m.geom_p1[b][1] + f.R[3 * r + 2] * c_m.geom_p1[b][2];
                        } else for (int r = 0; r < 3; r++) cc[r] = c0[r];
                        const double dist = cc[2] + zO - sz[0];
                        if (dist >= margin) continue;
                        cc[2] -= sz[0] + 0.5 * dist;
                        add_contact(e, b, cc, dist, f.v);
                    }
                }
            }
            double cpos[3];
            for (int r = 0; r < 3; r++)
                cpos[r] = f.p[r] + f.R[3 * r] * c_m.body_ipos[b][0] + f.R[3 * r + 1] * c_m.body_ipos[b][1] + f.R[3 * r + 2] * c_m.body_ipos[b][2];
            const double *in = c_m.body_inertia[b];
            double Ib[9] = {in[0], in[3], in[4], in[3], in[1], in[5], in[4], in[5], in[2]}, Tm[9], Iw[6];
            for (int r = 0; r < 3; r++) for (int cidx = 0; cidx < 3; cidx++)
                Tm[3 * r + cidx] = f.R[3 * r] * Ib[cidx] + f.R[3 * r + 1] * Ib[3 + cidx] + f.R[3 * r + 2] * Ib[6 + cidx];
            // Iw = Tm R^T, symmetric: xx yy zz xy xz yz
            Iw[0] = Tm[0] * f.R[0] + Tm[1] * f.R[1] + Tm[2] * f.R[2];
            Iw[1] = Tm[3] * f.R[3] + Tm[4] * f.R[4] + Tm[5] * f.R[5];
            Iw[2] = Tm[6] * f.R[6] + Tm[7] * f.R[7] + Tm[8] * f.R[8];
            Iw[3] = Tm[0] * f.R[3] + Tm[1] * f.R[4] + Tm[2] * f.R[5];
            Iw[4] = Tm[0] * f.R[6] + Tm[1] * f.R[7] + Tm[2] * f.R[8];
            Iw[5] = Tm[3] * f.R[6] + Tm[4] * f.R[7] + Tm[5] * f.R[8];
            const double mass = c_m.body_mass[b], cc = dot3(cpos, cpos);
            double *ci = e.cin[b];
            ci[0] = mass;
            ci[1] = mass * cpos[0]; ci[2] = mass * cpos[1]; ci[3] = mass * cpos[2];
            ci[4] = Iw[0] + mass * (cc - cpos[0] * cpos[0]);
            ci[5] = Iw[1] + mass * (cc - cpos[1] * cpos[1]);
            ci[6] = Iw[2] + mass * (cc - cpos[2] * cpos[2]);
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
            for (int r = 0; r < 3; r++) {
                e.fb[b][r] = Ia[r] + c0[r] + c1[r];
                e.fb[b][3 + r] = Ia[3 + r] + c2[r];
            }
        }
        e.jf[c] = f;
    }
}

// Backward articulated-body sweep (MODE 2 / 3: with the active constraint rows, 2 re-uses the bias C, 3 computes it like 0).
// MODE 0: bias C_i = S_i . F and factor/solve with rhs = tau - C,
// pivots S.U + armature (forward dynamics).  MODE 1: rhs = e.tau (pre-filled), pivots + kd h (stable PD).
template <int MODE>
__device__ void pass_backward(EnvData &e) {
    const int nchain = c_m.nchain;
    for (int c = 0; c < nchain; c++) {
        for (int k = 0; k < 21; k++) e.jb[c].IA[k] = 0.0;
        for (int k = 0; k < 6; k++) { e.jb[c].pA[k] = 0.0; e.jb[c].F[k] = 0.0; }
    }
    for (int c = nchain - 1; c >= 0; c--) {
        Bwd w = e.jb[c];
        for (int b = c_m.chain_hi[c]; b >= c_m.chain_lo[c]; b--) {
            const double *ci = e.cin[b];
            // add the body's spatial inertia: [[I_O, hx], [hx^T, m 1]]
            w.IA[sx(0, 0)] += ci[4]; w.IA[sx(1, 1)] += ci[5]; w.IA[sx(2, 2)] += ci[6];
            w.IA[sx(0, 1)] += ci[7]; w.IA[sx(0, 2)] += ci[8]; w.IA[sx(1, 2)] += ci[9];
            w.IA[sx(0, 4)] += -ci[3]; w.IA[sx(0, 5)] += ci[2];
            w.IA[sx(1, 3)] += ci[3];  w.IA[sx(1, 5)] += -ci[1];
            w.IA[sx(2, 3)] += -ci[2]; w.IA[sx(2, 4)] += ci[1];
            w.IA[sx(3, 3)] += ci[0]; w.IA[sx(4, 4)] += ci[0]; w.IA[sx(5, 5)] += ci[0];
            if (MODE == 0 || MODE == 3) for (int k = 0; k < 6; k++) w.F[k] += e.fb[b][k];
            if (MODE >= 2) {
                // active contact rows of this body: D p p^T joins the articulated inertia, D aref p acts as an external force
                for (int k = 0; k < e.ncon; k++) {
                    if (e.con_body[k] != b) continue;
                    for (int ed = 0; ed < 4; ed++) {
                        if (!e.row_act[4 * k + ed]) continue;
                        const double sg = (ed & 1) ? -c_m.con_mu : c_m.con_mu, D = e.row_D[4 * k + ed], fa = D * e.row_aref[4 * k + ed];
                        double p[6];
                        for (int r = 0; r < 6; r++) p[r] = e.con_p[k][0][r] + sg * e.con_p[k][ed < 2 ? 1 : 2][r];
                        for (int r = 0; r < 6; r++) {
                            for (int cc = r; cc < 6; cc++) w.IA[sx(r, cc)] += D * p[r] * p[cc];
                            w.pA[r] -= fa * p[r];
                        }
                    }
                }
            }
            const int da = c_m.body_dofadr[b], nd = c_m.body_dofnum[b];
            for (int i = da + nd - 1; i >= da; i--) {
                double S[6], U[6];
#pragma unroll
                for (int r = 0; r < 6; r++) S[r] = e.S[i][r];
                double rhs;
                if (MODE == 0 || MODE == 3) {
                    double Ci = dot6(S, w.F);
                    e.C[i] = Ci;
                    rhs = e.tau[i] - Ci;
                    if (MODE == 3) rhs += e.radd[i];
                } else if (MODE == 2) rhs = e.tau[i] + e.radd[i] - e.C[i];     // constrained re-solve: radd = the limit rows' D s aref
                else rhs = e.tau[i];
#pragma unroll
                for (int r = 0; r < 6; r++) {
                    double t = 0.0;
#pragma unroll
                    for (int cc = 0; cc < 6; cc++) t += w.IA[sx(r, cc)] * S[cc];
                    U[r] = t;
                }
                double D = dot6(S, U) + c_m.dof_arm[i];
                if (MODE == 1) D += c_m.kd[i] * c_m.h;
                if (MODE >= 2) D += e.dadd[i];
                const double Dinv = 1.0 / D;
                const double ui = rhs - dot6(S, w.pA);
                e.Dinv[i] = Dinv;
                e.u[i] = ui;
#pragma unroll
                for (int r = 0; r < 6; r++) e.U[i][r] = U[r];
#pragma unroll
                for (int r = 0; r < 6; r++) {
                    const double ur = U[r] * Dinv;
#pragma unroll
                    for (int cc = r; cc < 6; cc++) w.IA[sx(r, cc)] -= ur * U[cc];
                    w.pA[r] += ur * ui;
                }
            }
        }
        const int pc = c_m.chain_parent[c];
        if (pc >= 0) {
            for (int k = 0; k < 21; k++) e.jb[pc].IA[k] += w.IA[k];
            for (int k = 0; k < 6; k++) { e.jb[pc].pA[k] += w.pA[k]; e.jb[pc].F[k] += w.F[k]; }
        }
    }
}

// Forward acceleration sweep of the pure solve: out[i] = x_i where (M + diag) x = rhs.
__device__ void pass_accel(EnvData &e, double *out) {
    const int nchain = c_m.nchain;
    for (int c = 0; c < nchain; c++) {
        double a[6];
        const int pc = c_m.chain_parent[c];
        for (int k = 0; k < 6; k++) a[k] = pc >= 0 ? e.ja[pc][k] : 0.0;
        const int lo = c_m.body_dofadr[c_m.chain_lo[c]];
        const int hi = c_m.body_dofadr[c_m.chain_hi[c]] + c_m.body_dofnum[c_m.chain_hi[c]];
        for (int i = lo; i < hi; i++) {
            double x = e.Dinv[i] * (e.u[i] - dot6(e.U[i], a));
            out[i] = x;
#pragma unroll
            for (int r = 0; r < 6; r++) a[r] += e.S[i][r] * x;
        }
        for (int k = 0; k < 6; k++) e.ja[c][k] = a[k];
    }
}

// ---- constraint rows: joint limits and floor contacts (MuJoCo soft constraints; egp_model_set_joint_limits / _contacts) ---
// One row at signed distance dist (already minus its margin) and row velocity vel: impedance d(dist) from solimp, weight
// D = 1/R with R = (1 - d)/d * invweight, reference acceleration aref = -b vel - k d dist.
__device__ void limit_row(double invweight, double dist, double vel, double &Dc, double &aref) {
    double imp;
    if (c_m.lim_d0 == c_m.lim_dw || c_m.lim_width <= 1e-15) imp = 0.5 * (c_m.lim_d0 + c_m.lim_dw);
    else {
        const double xx = fabs(dist) / c_m.lim_width, pw = c_m.lim_pow, mid = c_m.lim_mid;
        double y;
        if (xx >= 1.0) y = 1.0;
        else if (xx <= 0.0) y = 0.0;
        else if (pw == 1.0) y = xx;
        else if (xx <= mid) y = pow(xx, pw) / pow(mid, pw - 1.0);
        else y = 1.0 - pow(1.0 - xx, pw) / pow(1.0 - mid, pw - 1.0);
        imp = c_m.lim_d0 + y * (c_m.lim_dw - c_m.lim_d0);
    }
    double R = (1.0 - imp) / imp * invweight;
    if (R < 1e-15) R = 1e-15;
    Dc = 1.0 / R;
    aref = -c_m.lim_b * vel - c_m.lim_k * imp * dist;
}

// forward acceleration sweep that also evaluates the contact rows: row_new = (p . a_body - aref < 0)
__device__ void pass_accel_rows(EnvData &e, double *out) {
    const int nchain = c_m.nchain;
    for (int c = 0; c < nchain; c++) {
        double a[6];
        const int pc = c_m.chain_parent[c];
        for (int k = 0; k < 6; k++) a[k] = pc >= 0 ? e.ja[pc][k] : 0.0;
        for (int b = c_m.chain_lo[c]; b <= c_m.chain_hi[c]; b++) {
            const int da = c_m.body_dofadr[b], nd = c_m.body_dofnum[b];
            for (int i = da; i < da + nd; i++) {
                double x = e.Dinv[i] * (e.u[i] - dot6(e.U[i], a));
                out[i] = x;
#pragma unroll
                for (int r = 0; r < 6; r++) a[r] += e.S[i][r] * x;
            }
            for (int k = 0; k < e.ncon; k++) {
                if (e.con_body[k] != b) continue;
                const double an = dot6(e.con_p[k][0], a), a1 = dot6(e.con_p[k][1], a), a2 = dot6(e.con_p[k][2], a);
                for (int ed = 0; ed < 4; ed++) {
                    const double r = an + ((ed & 1) ? -c_m.con_mu : c_m.con_mu) * (ed < 2 ? a1 : a2) - e.row_aref[4 * k + ed];
                    e.row_new[4 * k + ed] = r < 0.0;
                }
            }
        }
        for (int k = 0; k < 6; k++) e.ja[c][k] = a[k];
    }
}

// constrained solves that left the active-set loop at its cap without a fixed point (egp_cons_cap_hits)
__device__ unsigned long long g_cons_cap_hits;
__device__ unsigned long long g_cons_passes[2];     // block-sweep kernel, CTA 0 only: sub-steps, active-set passes

// mj_forward's acceleration stage after pass_kinematics, with e.tau = actuation: bias C + qacc.  Without constraint rows
// the plain bias / factor sweep and acceleration sweep.  With rows the solver's optimum is
//   (M + sum_act D_i J_i^T J_i) a = tau - C + sum_act D_i aref_i J_i^T.
// A limit row's Jacobian is a unit vector (adds D to a pivot), a contact row's is p^T J_body (adds D p p^T to that body's
// articulated inertia and D aref p to its external force): the same articulated-body sweeps solve it in O(n).  The
// active set {rows with J a - aref < 0} is iterated to its fixed point, starting from all rows active; that first
// iterate shares its backward sweep with the bias computation (MODE 3), later ones re-use the bias (MODE 2).  The
// smooth bias C and the smooth tree data stay as they are (compute_torque reads them).
__device__ void forward_dynamics(EnvData &e) {
    const int nv = c_m.nv;
    unsigned long long inst = 0ull, act;
    double Dc[MAXV], ar[MAXV], sg[MAXV];
    if (c_m.limits) {
        for (int i = 6; i < nv; i++) {
            const double lo = c_m.lim_lo[i], hi = c_m.lim_hi[i], q = e.q[i + 1];
            if (!(lo < hi)) continue;
            double dist, s;
            if (q - lo < 0.0) { dist = q - lo; s = 1.0; }
            else if (hi - q < 0.0) { dist = hi - q; s = -1.0; }
            else continue;
            limit_row(c_m.lim_iw[i], dist, s * e.v[i], Dc[i], ar[i]);
            sg[i] = s;
            inst |= 1ull << i;
        }
    }
    const int nrow = 4 * e.ncon;
    if (!inst && !nrow) {
        pass_backward<0>(e);
        pass_accel(e, e.qacc);
        return;
    }
    act = inst;
    for (int k = 0; k < nrow; k++) e.row_act[k] = 1;
    for (int it = 0; it < 100; it++) {
        for (int i = 0; i < nv; i++) {
            const bool on = (act >> i) & 1ull;
            e.dadd[i] = on ? Dc[i] : 0.0;
            e.radd[i] = on ? Dc[i] * sg[i] * ar[i] : 0.0;
        }
        if (it == 0) pass_backward<3>(e);
        else pass_backward<2>(e);
        pass_accel_rows(e, e.qacc);
        // all rows that disagree flip at once; if that has not settled after CONS_CAREFUL passes (coupled rows can flip back and
        // forth: about 1 of 20 000 solves), only the first disagreeing row flips per pass
        const bool careful = it >= CONS_CAREFUL;
        bool same = true;
        for (int i = 6; i < nv; i++) {
            if (!((inst >> i) & 1ull)) continue;
            const bool on = sg[i] * e.qacc[i] - ar[i] < 0.0;
            if (on != (bool)((act >> i) & 1ull)) {
                if (!careful || same) act ^= 1ull << i;
                same = false;
            }
        }
        for (int k = 0; k < nrow; k++) {
            if (e.row_new[k] != e.row_act[k]) {
                if (!careful || same) e.row_act[k] = e.row_new[k];
                same = false;
            }
        }
        if (same) break;
        if (it == 99) atomicAdd(&g_cons_cap_hits, 1ull);
    }
}

// sim.forward() at the current state (envs/common/mujoco_env.py:100-101): refresh tree data + bias
__device__ void env_forward(EnvData &e) {
    pass_kinematics(e);
    for (int i = 0; i < c_m.nv; i++) e.tau[i] = 0.0;
    pass_backward<0>(e);
}

// One iteration of do_simulation (humanoid_v1.py:166-174): compute_torque on the stale tree data, then mj_step.
__device__ void env_substep(EnvData &e, const double *ctrl /* per dof, [nv] */, double *torque_out) {
    const int nv = c_m.nv;
    const double h = c_m.h;
    // stable PD (humanoid_v1.py:130-156): (M + Kd h) x = -C - Kp e_q - Kd v ; torque = -kp e_q - kd (v + x h)
    for (int i = 0; i < nv; i++) {
        double eq = i >= 6 ? e.q[i + 1] - ctrl[i] : 0.0;
        e.tau[i] = -e.C[i] - c_m.kp[i] * eq - c_m.kd[i] * e.v[i];
    }
    pass_backward<1>(e);
    pass_accel(e, e.qacc);
    for (int i = 0; i < nv; i++) {
        double t = 0.0;
        if (i >= 6) {
            double eq = e.q[i + 1] - ctrl[i];
            t = -c_m.kp[i] * eq - c_m.kd[i] * (e.v[i] + e.qacc[i] * h);
            double lim = c_m.tlim[i];
            t = t < -lim ? -lim : (t > lim ? lim : t);     // np.clip (humanoid_v1.py:172)
        }
        e.tau[i] = t;
        if (torque_out && i >= 6) torque_out[i - 6] = t;
    }
    // mj_step: forward at (q, v) then semi-implicit Euler
    pass_kinematics(e);
    forward_dynamics(e);
    for (int i = 0; i < nv; i++) e.v[i] += h * e.qacc[i];
    for (int k = 0; k < 3; k++) e.q[k] += h * e.v[k];
    {
        double w[3] = {e.v[3], e.v[4], e.v[5]};
        double n = sqrt(dot3(w, w)), ax[3] = {1.0, 0.0, 0.0};
        if (n > 1e-15) { ax[0] = w[0] / n; ax[1] = w[1] / n; ax[2] = w[2] / n; }
        double sn, cs;
        sincos(0.5 * h * n, &sn, &cs);
        double qr[4] = {cs, ax[0] * sn, ax[1] * sn, ax[2] * sn};
        double qn = sqrt(e.q[3] * e.q[3] + e.q[4] * e.q[4] + e.q[5] * e.q[5] + e.q[6] * e.q[6]);
        double qq[4] = {e.q[3] / qn, e.q[4] / qn, e.q[5] / qn, e.q[6] / qn}, o[4];
        quat_mul(qq, qr, o);
        e.q[3] = o[0]; e.q[4] = o[1]; e.q[5] = o[2]; e.q[6] = o[3];
    }
    for (int i = 6; i < nv; i++) e.q[i + 1] += h * e.v[i];
}

// humanoid_v1.py:73-96 get_full_obs (obs_coord 'heading', root_deheading, obs_vel 'full')
__device__ void env_obs(const EnvData &e, double *obs) {
    double dq[4], vl[3];
    de_heading(e.q + 3, dq);
    obs[0] = e.q[2];
    obs[1] = dq[0]; obs[2] = dq[1]; obs[3] = dq[2]; obs[4] = dq[3];
    const int nq = c_m.nq, nv = c_m.nv;
    for (int k = 7; k < nq; k++) obs[k - 2] = e.q[k];
    to_heading(e.v, e.q + 3, vl);
    double *ov = obs + (nq - 2);
    ov[0] = vl[0]; ov[1] = vl[1]; ov[2] = vl[2];
    for (int k = 3; k < nv; k++) ov[k] = e.v[k];
}

// humanoid_v1.py:113-125 get_body_quat
__device__ void env_body_quat(const double *q, double *bq) {
    bq[0] = q[3]; bq[1] = q[4]; bq[2] = q[5]; bq[3] = q[6];
    for (int b = 1; b < c_m.nbody; b++) {
        const int qa = c_m.body_qposadr[b], nd = c_m.body_dofnum[b];
        double e0 = q[qa], e1 = nd > 1 ? q[qa + 1] : 0.0, e2 = nd > 2 ? q[qa + 2] : 0.0;
        quat_from_euler(e0, e1, e2, bq + 4 * b);
    }
}

// humanoid_v1.py:98-111 get_ee_pos('heading') from the (stale) body positions
__device__ void env_ee_pos(const EnvData &e, double *ee) {
    for (int k = 0; k < EGP_NEE; k++) {
        const double *x = e.xp[c_m.ee_body[k]];
        double v[3] = {x[0] - e.q[0], x[1] - e.q[1], x[2] - e.q[2]};
        to_heading(v, e.q + 3, ee + 3 * k);
    }
}

// reward_function.py:4-60 quat_space_reward_v3
__device__ double env_reward(const EnvData &e, const double *prev_root /*7*/, const double *prev_bq, const double *cur_bq,
                             const double *row, double dt, int t, int episode_len, int end, double end_reward,
                             double *info5) {
    const int nb = c_m.nbody;
    // get_qvel_fd(prev_qpos, cur_qpos, dt, 'heading')[:6]  (utils/math.py:20-35)
    double lin[3], qi[4], qrel[4], axis[3], ang, rv[3], fdl[3], fda[3];
    for (int k = 0; k < 3; k++) lin[k] = (e.q[k] - prev_root[k]) / dt;
    quat_inv(prev_root + 3, qi);
    quat_mul(e.q + 3, qi, qrel);
    rot_from_quat(qrel, axis, &ang);
    if (1.0 - qrel[0] < 1e-8) { axis[0] = 1.0; }
    if (ang > M_PI) ang -= 2 * M_PI; else if (ang < -M_PI) ang += 2 * M_PI;
    for (int k = 0; k < 3; k++) rv[k] = (axis[k] * ang) / dt;
    to_root(rv, prev_root + 3, fda);
    to_heading(lin, prev_root + 3, fdl);
    double rq[4], ee[3 * EGP_NEE];
    de_heading(e.q + 3, rq);
    env_ee_pos(e, ee);
    const double *e_bquat = row + EGP_X_BQUAT, *e_bangvel = row + EGP_X_BANGVEL;
    double pose2 = 0.0, vd = 0.0;
    for (int b = 1; b < nb; b++) {
        double q1[4], qd[4];
        quat_inv(e_bquat + 4 * b, q1);
        quat_mul(cur_bq + 4 * b, q1, qd);
        double w = fmin(fmax(qd[0], -1.0), 1.0);
        double a = acos(w) * c_m.b_diffw[b - 1];
        pose2 += a * a;
        // get_angvel_fd (utils/math.py:38-44) for this body
        double p1[4], pd[4], av[3], aang;
        quat_inv(prev_bq + 4 * b, p1);
        quat_mul(cur_bq + 4 * b, p1, pd);
        rot_from_quat(pd, av, &aang);
        if (1.0 - pd[0] < 1e-8) av[0] = 1.0;
        for (int k = 0; k < 3; k++) {
            double df = fabs(av[k] * aang / dt - e_bangvel[3 * b + k]);
            vd += c_m.v_ord == 1 ? df : df * df;
        }
    }
    double pose_dist = sqrt(pose2);
    double pose_reward = exp(-c_m.k_p * (pose_dist * pose_dist));
    double vel_dist = c_m.v_ord == 1 ? vd : sqrt(vd);
    double vel_reward = exp(-c_m.k_v * (vel_dist * vel_dist));
    double e2 = 0.0;
    for (int k = 0; k < 3 * EGP_NEE; k++) { double df = ee[k] - row[EGP_X_EE_POS + k]; e2 += df * df; }
    double ee_dist = sqrt(e2);
    double ee_reward = exp(-c_m.k_e * (ee_dist * ee_dist));
    double hd = e.q[2] - row[EGP_X_QPOS + 2], q1[4], qd[4];
    quat_inv(row + EGP_X_RQ_RMH, q1);
    quat_mul(rq, q1, qd);
    double rqd = acos(fmin(fmax(qd[0], -1.0), 1.0));
    double root_pose_reward = exp(-c_m.k_rh * (hd * hd) - c_m.k_rq * (rqd * rqd));
    double l2 = 0.0, a2 = 0.0;
    for (int k = 0; k < 3; k++) {
        double dl = fdl[k] - row[EGP_X_RLINV_LOCAL + k], da = fda[k] - row[EGP_X_RANGV + k];
        l2 += dl * dl; a2 += da * da;
    }
    double ld = sqrt(l2), ad = sqrt(a2);
    double root_vel_reward = exp(-c_m.k_rl * (ld * ld) - c_m.k_ra * (ad * ad));
    double reward = c_m.w_p * pose_reward + c_m.w_v * vel_reward + c_m.w_e * ee_reward + c_m.w_rp * root_pose_reward +
                    c_m.w_rv * root_vel_reward;
    reward /= c_m.w_p + c_m.w_v + c_m.w_e + c_m.w_rp + c_m.w_rv;
    if (c_m.decay) reward *= 1.0 - (double)t / episode_len;
    if (end) reward += end_reward;
    info5[0] = pose_reward; info5[1] = vel_reward; info5[2] = ee_reward; info5[3] = root_pose_reward; info5[4] = root_vel_reward;
    return reward;
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator (Salmon et al. 2011), used for perf-mode noise and reset draws
__device__ __forceinline__ void philox4x32(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
__device__ __forceinline__ double u01(uint32_t a, uint32_t b) {     // (0, 1]
    unsigned long long x = ((unsigned long long)a << 21) ^ (unsigned long long)(b >> 11);   // 53 bits
    return ((double)(x & ((1ULL << 53) - 1)) + 1.0) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ void normal2(uint64_t seed, uint64_t stream, uint32_t a, uint32_t b, double *z0, double *z1) {
    uint32_t c[4] = {a, b, (uint32_t)stream, (uint32_t)(stream >> 32)};
    philox4x32(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    double u1 = u01(c[0], c[1]), u2 = u01(c[2], c[3]);
    double r = sqrt(-2.0 * log(u1)), sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    *z0 = r * cs; *z1 = r * sn;
}

// ------------------------------------------------------------------------------------------------
struct RolloutArgs {
    EgpRolloutCfg cfg;
    EgpRolloutIn in;
    EgpTrajOut out;
    // expert tables
    const int32_t *take_off;
    const double *rows, *head_lb, *ctx;
    const int32_t *win_off;
    int n_takes, ctx_dim, ctx_mode, ctx_T;
    // policy (transposed + padded): W1t [D][H1p], W2t [H1][H2p], W3t [H2][Ap]
    const double *W1t, *b1, *W2t, *b2, *W3t, *b3, *log_std;
    int D, H1, H2, A, H1p, H2p, Ap;
    int K1p, K2p, K3p;      // T4: k padded to the tile depth (weights packed as tiles)
    int chunk23;            // T4: 1 = chunked layer-2/3 path for wide policies
    int xrows, hrows, xs;   // T4: rows of the activation buffers xs / h1s and their row stride in doubles
    double *log_part;       // [CTAs][EGP_LOG_SIZE] per-CTA logger partials (merged by logger_merge_kernel)
    double *cons;           // T4 with joint limits / floor contact: [CTAs][CS_PER_BODY * nbody][32] constraint scratch
    // value net of the 'valuefs' evaluation rule (same plan as the policy, head padded to vAp rows), its context table
    const double *vW1t, *vb1, *vW2t, *vb2, *vW3t, *vb3, *vctx;
    int vAp;
    // state LSTM (VideoForecastNet.s_net, step mode): tile-packed [4H][S + H] weights, bias, per-CTA h | c scratch
    const double *sn_Wp, *sn_b;
    double *sn_state;
    int sn_H, sn_Kp;
};

// video-context row of (take, start, t): per-frame table, or one row block per (take, start) episode window
__device__ __forceinline__ const double *ctx_row(const RolloutArgs &A, int take, int start, int t) {
    const size_t r = A.ctx_mode == 0 ? (size_t)(A.take_off[take] + start + t)
                   : A.ctx_mode == 1 ? (size_t)(A.win_off[take] + start - A.cfg.fr_margin) * A.ctx_T + t
                                     : (size_t)(A.win_off[take] + start - A.cfg.fr_margin);
    return A.ctx + r * A.ctx_dim;
}

// one dense layer for the 32 environments of the CTA: ys[j][lane] = act(b[j] + sum_k Wt[k][j] xs[k][lane])
template <bool RELU>
__device__ __forceinline__ void mlp_layer(const double *__restrict__ Wt, const double *__restrict__ bias, int K, int Np,
                                          const double *xs, double *ys, int lane) {
    const int EP = blockDim.x;                  // environments of this CTA = row stride of the activation tiles
    for (int j0 = 0; j0 < Np; j0 += JB) {
        double acc[JB];
#pragma unroll
        for (int jj = 0; jj < JB; jj++) acc[jj] = bias[j0 + jj];
        for (int k = 0; k < K; k++) {
            const double xv = xs[k * EP + lane];
            const double2 *w = reinterpret_cast<const double2 *>(Wt + (size_t)k * Np + j0);
#pragma unroll
            for (int jj = 0; jj < JB / 2; jj++) {
                double2 ww = __ldg(w + jj);
                acc[2 * jj] += ww.x * xv;
                acc[2 * jj + 1] += ww.y * xv;
            }
        }
#pragma unroll
        for (int jj = 0; jj < JB; jj++) ys[(j0 + jj) * EP + lane] = RELU ? fmax(acc[jj], 0.0) : acc[jj];
    }
}

__device__ void zfilter(const double *x, double *y, int n, const double *mean, const double *sd, double clip) {
    for (int k = 0; k < n; k++) {
        double v = x[k];
        if (mean) {
            v = (v - mean[k]) / (sd[k] + 1e-8);
            if (clip > 0.0) v = fmin(fmax(v, -clip), clip);
        }
        y[k] = v;
    }
}

__device__ void load_reset(EnvData &e, const RolloutArgs &A, int take, int start) {
    const double *row = A.rows + (size_t)(A.take_off[take] + start) * EGP_X_STRIDE;
    for (int k = 0; k < c_m.nq; k++) e.q[k] = row[EGP_X_QPOS + k];
    for (int k = 0; k < c_m.nv; k++) e.v[k] = row[EGP_X_QVEL + k];
    env_forward(e);
}

__global__ void __launch_bounds__(ENVS_PER_CTA)
rollout_kernel(const RolloutArgs A) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x;
    const int S = c_m.nq - 2 + c_m.nv, nu = c_m.nu, nv = c_m.nv, nb = c_m.nbody;
    // The CTA is one (possibly partial) warp of EP <= 32 environments (32 unless EGP_V1_ENVS_PER_CTA says otherwise)
    const int EP = blockDim.x;
    double *xs = smem;                                  // [max(D, H2p)][EP]
    const int xrows = A.D > A.H2p ? A.D : A.H2p;
    double *h1s = smem + (size_t)xrows * EP;            // [max(H1p, Ap)][EP] (also receives the action means)
    const int T = A.cfg.horizon, E = A.cfg.n_env;
    const double dt = c_m.h * c_m.frame_skip;
    const int env = blockIdx.x * EP + lane;
    const bool live = env < E;
    const int eid = live ? env : E - 1;                 // idle lanes shadow the last env and write nothing

    EnvData e;
    double obs[MAXV * 2], state[MAXV * 2], nstate[MAXV * 2], ctrl[MAXV], act[MAXV];
    double bq_a[4 * MAXB], bq_b[4 * MAXB];
    double *cur_bq = bq_a, *prev_bq = bq_b;
    double log_acc[EGP_LOG_SIZE];
    for (int k = 0; k < EGP_LOG_SIZE; k++) log_acc[k] = 0.0;
    log_acc[EGP_LOG_MIN_C_REWARD] = INFINITY; log_acc[EGP_LOG_MAX_C_REWARD] = -INFINITY;
    log_acc[EGP_LOG_MIN_EPISODE_REWARD] = INFINITY; log_acc[EGP_LOG_MAX_EPISODE_REWARD] = -INFINITY;
    double ep_reward = 0.0;

    int n_reset = 0, take, start, cur_t = 0;
    auto draw_reset = [&](int r) {
        if (A.in.d_reset_take) {
            int rr = r < A.cfg.max_resets ? r : A.cfg.max_resets - 1;
            take = A.in.d_reset_take[(size_t)eid * A.cfg.max_resets + rr];
            start = A.in.d_reset_start[(size_t)eid * A.cfg.max_resets + rr];
        } else {
            uint32_t c[4] = {(uint32_t)eid, (uint32_t)r, (uint32_t)A.cfg.iteration, 0x52535421u ^ (uint32_t)(A.cfg.iteration >> 32)};
            philox4x32(c, (uint32_t)A.cfg.seed, (uint32_t)(A.cfg.seed >> 32));
            take = (int)(c[0] % (uint32_t)A.n_takes);                              // humanoid_v1.py:210
            int len = A.take_off[take + 1] - A.take_off[take];
            int span = len - A.cfg.episode_len - 2 * A.cfg.fr_margin;              // humanoid_v1.py:214
            start = A.cfg.fr_margin + (span > 0 ? (int)(c[1] % (uint32_t)span) : 0);
        }
    };
    draw_reset(0);
    load_reset(e, A, take, start);
    env_body_quat(e.q, cur_bq);
    env_obs(e, obs);
    zfilter(obs, state, S, A.in.d_zf_mean, A.in.d_zf_std, A.cfg.zf_clip);
    for (int i = 0; i < 6; i++) ctrl[i] = 0.0;

    for (int t = 0; t < T; t++) {
        const size_t n = (size_t)eid * T + t;
        // ---- policy input cat(ctx[frame], state) -> shared, feature-major (video_state_net.py:62-64)
        int off = 0;
        if (A.ctx) {
            const double *cx = ctx_row(A, take, start, cur_t);
            for (int k = 0; k < A.ctx_dim; k++) xs[k * EP + lane] = cx[k];
            off = A.ctx_dim;
        }
        for (int k = 0; k < S; k++) xs[(off + k) * EP + lane] = state[k];
        __syncthreads();
        // ---- PolicyGaussian forward (policy_gaussian.py:19-24, mlp.py:22-25)
        mlp_layer<true>(A.W1t, A.b1, A.D, A.H1p, xs, h1s, lane);
        mlp_layer<true>(A.W2t, A.b2, A.H1, A.H2p, h1s, xs, lane);
        mlp_layer<false>(A.W3t, A.b3, A.H2, A.Ap, xs, h1s, lane);
        // ---- sample (distributions.py / policy.py:12-15): a = mu + exp(log_std) eps, or the mean
        bool mean_flag = A.cfg.mean_action != 0;
        if (A.in.d_mean_flag) mean_flag = mean_flag || A.in.d_mean_flag[n] != 0;
        else if (A.cfg.noise_rate < 1.0) {
            uint32_t c[4] = {(uint32_t)eid, (uint32_t)t, (uint32_t)A.cfg.iteration, 0x4d45414eu ^ (uint32_t)(A.cfg.iteration >> 32)};
            philox4x32(c, (uint32_t)A.cfg.seed, (uint32_t)(A.cfg.seed >> 32));
            mean_flag = mean_flag || (u01(c[0], c[1]) <= 1.0 - A.cfg.noise_rate);   // binomial(1, 1 - noise_rate)
        }
        for (int a = 0; a < nu; a += 2) {
            double z0 = 0.0, z1 = 0.0;
            if (!mean_flag) {
                if (A.in.d_eps) { z0 = A.in.d_eps[n * nu + a]; z1 = a + 1 < nu ? A.in.d_eps[n * nu + a + 1] : 0.0; }
                else normal2(A.cfg.seed, A.cfg.iteration, (uint32_t)eid, (uint32_t)(t * 64 + a), &z0, &z1);
            }
            act[a] = h1s[a * EP + lane] + exp(A.log_std[a]) * z0;
            if (a + 1 < nu) act[a + 1] = h1s[(a + 1) * EP + lane] + exp(A.log_std[a + 1]) * z1;
        }
        if (live) {
            for (int k = 0; k < S; k++) A.out.d_states[n * S + k] = state[k];
            for (int a = 0; a < nu; a++) A.out.d_actions[n * nu + a] = act[a];
            if (A.out.d_raw_obs) for (int k = 0; k < S; k++) A.out.d_raw_obs[n * S + k] = obs[k];
        }
        // ---- env.step (humanoid_v1.py:179-199)
        double prev_root[7];
        for (int k = 0; k < 7; k++) prev_root[k] = e.q[k];
        { double *tmp = prev_bq; prev_bq = cur_bq; cur_bq = tmp; }
        for (int a = 0; a < nu; a++) ctrl[6 + a] = c_m.a_ref[6 + a] + act[a] * c_m.a_scale[6 + a];
        for (int s = 0; s < c_m.frame_skip; s++) env_substep(e, ctrl, nullptr);
        cur_t += 1;
        env_body_quat(e.q, cur_bq);
        const double head_z = e.xp[c_m.head_body][2];
        const double lb = isnan(A.cfg.fix_head_lb) ? A.head_lb[take] - 0.1 : A.cfg.fix_head_lb;
        bool fail = head_z < lb;
        const bool end = cur_t >= A.cfg.episode_len;
        env_obs(e, obs);
        zfilter(obs, nstate, S, A.in.d_zf_mean, A.in.d_zf_std, A.cfg.zf_clip);
        double info5[5];
        const double *row = A.rows + (size_t)(A.take_off[take] + start + cur_t) * EGP_X_STRIDE;
        double rew = env_reward(e, prev_root, prev_bq, cur_bq, row, dt, cur_t, A.cfg.episode_len, end, A.cfg.end_reward, info5);
        bool bad = !(isfinite(rew) && isfinite(head_z));
        for (int k = 0; k < S && !bad; k++) bad = !(fabs(obs[k]) < 1e10);
        if (bad) {      // mj_checkPos/Vel analogue: terminate, reset, count
            rew = 0.0; fail = true;
            for (int k = 0; k < 5; k++) info5[k] = 0.0;
            for (int k = 0; k < S; k++) nstate[k] = 0.0;
            log_acc[EGP_LOG_NUM_NAN_RESETS] += 1.0;
        }
        const bool done = fail || end;
        if (live) {
            A.out.d_rewards[n] = rew;
            A.out.d_masks[n] = (done || t == T - 1) ? 0.0 : 1.0;
            A.out.d_exps[n] = mean_flag ? 0.0 : 1.0;
            A.out.d_v_metas[2 * n] = take;
            A.out.d_v_metas[2 * n + 1] = start;
            if (A.out.d_next_states) for (int k = 0; k < S; k++) A.out.d_next_states[n * S + k] = nstate[k];
            if (A.out.d_c_info) for (int k = 0; k < 5; k++) A.out.d_c_info[n * 5 + k] = info5[k];
        }
        // LoggerRL.step / end_episode (core/logger_rl.py:24-36)
        ep_reward += 1.0;
        log_acc[EGP_LOG_NUM_STEPS] += 1.0;
        log_acc[EGP_LOG_TOTAL_C_REWARD] += rew;
        log_acc[EGP_LOG_MIN_C_REWARD] = fmin(log_acc[EGP_LOG_MIN_C_REWARD], rew);
        log_acc[EGP_LOG_MAX_C_REWARD] = fmax(log_acc[EGP_LOG_MAX_C_REWARD], rew);
        for (int k = 0; k < 5; k++) log_acc[EGP_LOG_C_INFO + k] += info5[k];
        if (done || t == T - 1) {
            log_acc[EGP_LOG_NUM_EPISODES] += 1.0;
            log_acc[EGP_LOG_TOTAL_REWARD] += ep_reward;
            log_acc[EGP_LOG_MIN_EPISODE_REWARD] = fmin(log_acc[EGP_LOG_MIN_EPISODE_REWARD], ep_reward);
            log_acc[EGP_LOG_MAX_EPISODE_REWARD] = fmax(log_acc[EGP_LOG_MAX_EPISODE_REWARD], ep_reward);
            ep_reward = 0.0;
        }
        if (done && t < T - 1) {
            n_reset++;
            draw_reset(n_reset);
            cur_t = 0;
            load_reset(e, A, take, start);
            env_body_quat(e.q, cur_bq);
            env_obs(e, obs);
            zfilter(obs, state, S, A.in.d_zf_mean, A.in.d_zf_std, A.cfg.zf_clip);
        } else {
            for (int k = 0; k < S; k++) state[k] = nstate[k];
        }
        __syncthreads();
    }
    if (live) {
        if (A.out.d_final_qpos) for (int k = 0; k < c_m.nq; k++) A.out.d_final_qpos[(size_t)env * c_m.nq + k] = e.q[k];
        if (A.out.d_final_qvel) for (int k = 0; k < nv; k++) A.out.d_final_qvel[(size_t)env * nv + k] = e.v[k];
        if (A.out.d_logger) {
            double *L = A.out.d_logger;
            atomicAdd(L + EGP_LOG_NUM_STEPS, log_acc[EGP_LOG_NUM_STEPS]);
            atomicAdd(L + EGP_LOG_NUM_EPISODES, log_acc[EGP_LOG_NUM_EPISODES]);
            atomicAdd(L + EGP_LOG_TOTAL_REWARD, log_acc[EGP_LOG_TOTAL_REWARD]);
            atomicAdd(L + EGP_LOG_TOTAL_C_REWARD, log_acc[EGP_LOG_TOTAL_C_REWARD]);
            atomicAdd(L + EGP_LOG_NUM_NAN_RESETS, log_acc[EGP_LOG_NUM_NAN_RESETS]);
            for (int k = 0; k < 5; k++) atomicAdd(L + EGP_LOG_C_INFO + k, log_acc[EGP_LOG_C_INFO + k]);
            // min / max through ordered-int tricks are avoided: rewards are >= 0 except NaN resets; use CAS loops
            auto amin = [](double *addr, double v) {
                unsigned long long *a = (unsigned long long *)addr, old = *a, assumed;
                do { assumed = old; if (__longlong_as_double(assumed) <= v) break;
                     old = atomicCAS(a, assumed, __double_as_longlong(v)); } while (assumed != old);
            };
            auto amax = [](double *addr, double v) {
                unsigned long long *a = (unsigned long long *)addr, old = *a, assumed;
                do { assumed = old; if (__longlong_as_double(assumed) >= v) break;
                     old = atomicCAS(a, assumed, __double_as_longlong(v)); } while (assumed != old);
            };
            amin(L + EGP_LOG_MIN_C_REWARD, log_acc[EGP_LOG_MIN_C_REWARD]);
            amax(L + EGP_LOG_MAX_C_REWARD, log_acc[EGP_LOG_MAX_C_REWARD]);
            amin(L + EGP_LOG_MIN_EPISODE_REWARD, log_acc[EGP_LOG_MIN_EPISODE_REWARD]);
            amax(L + EGP_LOG_MAX_EPISODE_REWARD, log_acc[EGP_LOG_MAX_EPISODE_REWARD]);
        }
    }
    (void)nb;
}


// ================================================================================================
// T4 variant: CTA = 4 warps x 32 environments.  Lane = environment, warp = owner of a set of kinematic
// chains; the tree sweeps run level-synchronously (root | spine, legs | arms, head) so the serial depth per
// pass is 28 of the 58 DoFs, and the policy MLP is split 4 ways.  Hot per-DoF data (joint axes, body anchors,
// U = I^A S, q, v) live in shared memory [row][env] (bank-conflict free); junction records pass between warps
// through shared memory; everything else is owner-private thread-local.
constexpr int T4_WARPS = 8;                  // warps per CTA: 4 chain warps (tree sweeps) + 4 helper warps (policy MLP, obs, reward)
constexpr int T4_THREADS = T4_WARPS * 32;

struct T4Local {
    double sav_ax[MAXV][3], sav_anc[MAXB][3];
    double bqp[MAXB][4], bqc[MAXB][4];
};

// ---- Tensor Memory as a per-thread scratchpad.  The kernel issues no tcgen05.mma, so the SM's 256 KB of
// TMEM is free: each warp owns its 32-lane quarter, each thread gets 512 x 32-bit private columns (256
// doubles) at ~12 cycles load latency - the values that round-trip between the tree sweeps (pivots, solve
// right-hand sides, bias forces, body inertias) live there instead of in L2-backed thread-local memory.
// All accesses are warp-uniform (tcgen05.ld/st are .sync.aligned).
__device__ __forceinline__ void tm_st1(uint32_t a, double v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(a), "r"(__double2loint(v)), "r"(__double2hiint(v)));
}
__device__ __forceinline__ double tm_ld1(uint32_t a) {
    int lo, hi;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];\n\ttcgen05.wait::ld.sync.aligned;" : "=r"(lo), "=r"(hi) : "r"(a) : "memory");
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_st2(uint32_t a, double v0, double v1) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(__double2loint(v0)), "r"(__double2hiint(v0)),
                 "r"(__double2loint(v1)), "r"(__double2hiint(v1)));
}
__device__ __forceinline__ void tm_ld2(uint32_t a, double &v0, double &v1) {
    int r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a) : "memory");
    v0 = __hiloint2double(r[1], r[0]); v1 = __hiloint2double(r[3], r[2]);
}
__device__ __forceinline__ void tm_st4(uint32_t a, const double *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(a),
                 "r"(__double2loint(v[0])), "r"(__double2hiint(v[0])), "r"(__double2loint(v[1])), "r"(__double2hiint(v[1])),
                 "r"(__double2loint(v[2])), "r"(__double2hiint(v[2])), "r"(__double2loint(v[3])), "r"(__double2hiint(v[3])));
}
__device__ __forceinline__ void tm_ld4(uint32_t a, double *v) {
    int r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a) : "memory");
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = __hiloint2double(r[2 * k + 1], r[2 * k]);
}
__device__ __forceinline__ void tm_st8(uint32_t a, const double *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(a),
                 "r"(__double2loint(v[0])), "r"(__double2hiint(v[0])), "r"(__double2loint(v[1])), "r"(__double2hiint(v[1])),
                 "r"(__double2loint(v[2])), "r"(__double2hiint(v[2])), "r"(__double2loint(v[3])), "r"(__double2hiint(v[3])),
                 "r"(__double2loint(v[4])), "r"(__double2hiint(v[4])), "r"(__double2loint(v[5])), "r"(__double2hiint(v[5])),
                 "r"(__double2loint(v[6])), "r"(__double2hiint(v[6])), "r"(__double2loint(v[7])), "r"(__double2hiint(v[7])));
}
__device__ __forceinline__ void tm_ld8(uint32_t a, double *v) {
    int r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
                 "tcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(a) : "memory");
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = __hiloint2double(r[2 * k + 1], r[2 * k]);
}

// ---- asynchronous scratch loads: tcgen05.ld is issued without the wait (no "memory" clobber either: these accesses do not
// touch ordinary memory, and volatile asm statements keep their mutual order), the destination registers become valid
// after tm_wait_ld, whose "+r" operands make that dependency visible to the compiler
template <int NR> struct TmLd;
template <> struct TmLd<8> {
    static __device__ __forceinline__ void issue(uint32_t a, int (&r)[8]) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a));
    }
    static __device__ __forceinline__ void wait(int (&r)[8]) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]));
    }
};
template <> struct TmLd<16> {
    static __device__ __forceinline__ void issue(uint32_t a, int (&r)[16]) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(a));
    }
    static __device__ __forceinline__ void wait(int (&r)[16]) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                     "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
    }
};
template <> struct TmLd<32> {
    static __device__ __forceinline__ void issue(uint32_t a, int (&r)[32]) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                     "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                       "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                       "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(a));
    }
    static __device__ __forceinline__ void wait(int (&r)[32]) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                     "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]),
                     "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),
                     "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31]));
    }
};

// storage context of the tree sweeps (csrc/tree.cuh) on the device: shared rows [row][env] + Tensor Memory scratch
extern __shared__ double t4_smem[];          // the CTA's dynamic shared memory (named here so that non-inlined sweeps
                                             // address it as shared memory, not through a generic pointer)
struct T4Ctx {
    static constexpr bool CONS = false;
    int lane, w;
    uint32_t tm;        // TMEM base of this warp's lane quarter
    T4Off o;
    __device__ __forceinline__ double &at(int off, int idx) const { return t4_smem[(off + idx) * 32 + lane]; }
    // TMEM column addresses of a dof's ctrl / tau / C entry and of a body's record
    __device__ __forceinline__ uint32_t a_ctrl(int i) const { return tm + c_m.dof_col_ctrl[i]; }
    __device__ __forceinline__ uint32_t a_tau(int i) const { return tm + c_m.dof_col_tau[i]; }
    __device__ __forceinline__ uint32_t a_c(int i) const { return tm + c_m.dof_col_c[i]; }
    __device__ __forceinline__ uint32_t a_cin(int b) const { return tm + c_m.bk[b].rec + R_CIN; }
    __device__ __forceinline__ void st_cin(int b, const double *ci) const { tm_st8(a_cin(b), ci); tm_st2(a_cin(b) + 16, ci[8], ci[9]); }
    __device__ __forceinline__ void ld_cin(int b, double *ci) const { tm_ld8(a_cin(b), ci); tm_ld2(a_cin(b) + 16, ci[8], ci[9]); }
    template <int NR> __device__ __forceinline__ void tld_issue(int col, int (&r)[NR]) const { TmLd<NR>::issue(tm + col, r); }
    template <int NR> __device__ __forceinline__ void tld_wait(int (&r)[NR]) const { TmLd<NR>::wait(r); }
    static __device__ __forceinline__ double unpack(const int *r, int k) { return __hiloint2double(r[2 * k + 1], r[2 * k]); }
    template <int N> __device__ __forceinline__ void tst(int col, const double *v) const {
        uint32_t a = tm + col;
        int k = 0;
        if (N >= 16) { tm_st8(a, v); tm_st8(a + 16, v + 8); k = 16; }
        else if (N >= 8) { tm_st8(a, v); k = 8; }
        if (N - k >= 4) { tm_st4(a + 2 * k, v + k); k += 4; }
        if (N - k >= 2) { tm_st2(a + 2 * k, v[k], v[k + 1]); k += 2; }
        if (N - k >= 1) tm_st1(a + 2 * k, v[k]);
    }
    __device__ __forceinline__ void twait_st() const { tm_wait_st(); }
};

// the same context with the constraint rows of joint limits / floor contact (csrc/tree.cuh, cons_*): scratch in L2-resident
// global memory ([slot][lane] per CTA) + the per-thread state of the active-set iteration
struct T4CtxC : T4Ctx {
    static constexpr bool CONS = true;
    double *csb;
    mutable unsigned ccnt, lim_inst, lim_act, lim_pinst, lim_pact;
    mutable int cchg, cflip;
    int cit;
    __device__ __forceinline__ double &cs(int slot) const { return csb[(size_t)slot * 32 + lane]; }
};

// barrier of the four chain warps (the helper warps never enter the sweeps)
__device__ __forceinline__ void t4_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// level-synchronous tree sweeps (root | spine, legs | arms, head): one chain per warp and level, junction records
// cross warps through shared memory (csrc/tree.cuh holds the per-chain work)
template <int MODE, class X>
__device__ __forceinline__ void t4_forward(const X &x) {
#pragma unroll 1
    for (int L = 0; L < c_m.nlevel; L++) {
        const int c = c_m.lvl_chain[L][x.w];
        if (c >= 0) t5_fwd_chain<MODE>(x, c);
        if (MODE == 0) { EGP_CLK_MARK(11) }
        t4_bar();
        if (MODE == 0) { EGP_CLK_MARK(12) }
    }
}

template <class X>
__device__ __forceinline__ void t4_backward(const X &x, const int MODE) {
#pragma unroll 1
    for (int L = c_m.nlevel - 1; L >= 0; L--) {
        const int c = c_m.lvl_chain[L][x.w];
        Bwd w;
        t5_bwd_gather(x, c, w);
        t4_bar();                   // children records consumed before this level overwrites the slots
        if (c >= 0) t5_bwd_chain(x, c, w, MODE);
        t4_bar();
    }
}

// loop helper: for every dof i owned by warp w
#define T4_FOR_OWN_DOFS(i, b)                                                                       \
    for (int _c = 0; _c < c_m.nchain; _c++)                                                          \
        if (c_m.chain_warp[_c] == x.w)                                                               \
            for (int b = c_m.chain_lo[_c]; b <= c_m.chain_hi[_c]; b++)                               \
                for (int i = c_m.body_dofadr[b]; i < c_m.body_dofadr[b] + c_m.body_dofnum[b]; i++)

template <class X>
__device__ __noinline__ void t4_forward_only(const X x0) {     // sim.forward(); chain warps only; own register allocation
    X x = x0;
    if constexpr (X::CONS) { x.ccnt = 0u; x.lim_act = 0u; x.lim_pinst = 0u; x.lim_pact = 0u; x.cit = 1; }      // no constraint rows: only the tree data and the bias are used
    t4_forward<2>(x);
    T4_FOR_OWN_DOFS(i, b) if (i >= 6) tm_st1(x.a_tau(i), 0.0);
    tm_wait_st();
    t4_backward(x, 0);
}

// One iteration of do_simulation (humanoid_v1.py:166-174): stable-PD torque on the stale tree data, then mj_step.
#ifdef EGP_T4_CLK
__device__ unsigned long long g_t4_clk[8];      // build with -DEGP_T4_CLK: cycles of warp 0 of CTA 0 per sweep kind
#endif
template <class X>
__device__ __forceinline__ void t4_substep(const X &x) {
#ifdef EGP_T4_CLK
    const bool pr = blockIdx.x == 0 && x.w == 0 && x.lane == 0;
    long long t0 = clock64();
#define T4_CLK(slot) { long long t1 = clock64(); if (pr) g_t4_clk[slot] += t1 - t0; t0 = t1; }
#else
#define T4_CLK(slot)
#endif
    t4_backward(x, 1);          // (M_stale + Kd h) factor + reduce, rhs from the current q, v
    T4_CLK(0)
    t4_forward<1>(x);           // desired accel -> clipped torque ; kinematics / velocities / body forces at (q, v)
    T4_CLK(1)
    if constexpr (X::CONS) {
        // mj_step with constraint rows: the kinematics pass above has collected the floor contacts; the
        // factor / solve sweeps carry the active rows (pivots, articulated inertias, bias forces), the solve-only forward
        // sweep re-evaluates them, and the pair is repeated until the active set of every environment of the CTA is at its
        // fixed point (first pass: every row active; bias C and factor share that pass); then the Euler step.
        X &xm = const_cast<X &>(x);
        x.lim_pinst = x.lim_inst; x.lim_pact = x.lim_act;
        x.lim_inst = 0u; x.lim_act = 0u;
#pragma unroll 1
        for (xm.cit = 0; xm.cit < 100; xm.cit++) {
            x.cchg = 0; x.cflip = 0;
            t4_backward(x, 0);
            t4_forward<3>(x);
            const bool any = __any_sync(0xffffffffu, x.cchg != 0);
            if (x.lane == 0) t4_smem[(x.o.red + x.w) * 32] = any ? 1.0 : 0.0;
            t4_bar();
            const double tot = t4_smem[x.o.red * 32] + t4_smem[(x.o.red + 1) * 32] + t4_smem[(x.o.red + 2) * 32] + t4_smem[(x.o.red + 3) * 32];
            if (tot == 0.0) break;
            if (xm.cit == 99 && x.cchg) atomicAdd(&g_cons_cap_hits, 1ull);
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) { g_cons_passes[0] += 1; g_cons_passes[1] += (unsigned long long)(xm.cit < 100 ? xm.cit + 1 : 100); }
        T4_CLK(2)
#pragma unroll 1
        for (int L = 0; L < c_m.nlevel; L++) {
            const int c = c_m.lvl_chain[L][x.w];
            if (c >= 0) t5_integrate_chain(x, c);
        }
        t4_bar();
        T4_CLK(3)
    } else {
        t4_backward(x, 0);          // bias C, M factor + reduce with rhs = torque - C
        T4_CLK(2)
        t4_forward<0>(x);           // qacc, semi-implicit Euler
        T4_CLK(3)
    }
#ifdef EGP_T4_CLK
    if (pr) g_t4_clk[4] += 1;
#endif
}

// do_simulation (humanoid_v1.py:158-177): all sub-steps of one env step.  NOT inlined into the kernel: the sweeps get a
// register allocation of their own (the kernel's per-step state - filtered observation shares, logger sums - is
// saved / restored once per env step at the call, not carried through every sweep)
template <class X>
__device__ __forceinline__ void t4_do_simulation(const X &x, const int n_sub) {
#pragma unroll 1
    for (int s = 0; s < n_sub; s++) t4_substep(x);
}

// observation entry k (humanoid_v1.py:73-96) from the shared q / v rows; hd* = de-headed root quaternion,
// vl* = root linear velocity in the heading frame (both computed redundantly by every thread and passed as
// scalars: no dynamically indexed thread-local arrays)
__device__ __forceinline__ double t4_obs_entry(const T4Ctx &x, int k, double hd0, double hd1, double hd2, double hd3,
                                               double vl0, double vl1, double vl2) {
    const int nq = c_m.nq;
    if (k == 0) return x.at(x.o.q, 2);
    if (k < 5) return k == 1 ? hd0 : (k == 2 ? hd1 : (k == 3 ? hd2 : hd3));
    if (k < nq - 2) return x.at(x.o.q, k + 2);
    const int j = k - (nq - 2);
    if (j < 3) return j == 0 ? vl0 : (j == 1 ? vl1 : vl2);
    return x.at(x.o.v, j);
}

// ---- policy MLP on the FP64 tensor cores (csrc/dmma.cuh) ------------------------------------------------------------
// PolicyGaussian trunk + head for the CTA's 32 environments (policy_gaussian.py:19-24, mlp.py:22-25):
// xs [D rows, zero rows up to the padded K] -> action means in h1s rows [0, A).  Normal path: three full layers.
// Chunked path (second hidden layer too wide for shared memory): layer 2 is produced MLP_C2 neurons at a time into the
// (dead) input rows and immediately folded into the head's accumulators (Ap / 16 tiles x 2 halves <= 8 items: one per warp).
constexpr int MLP_C2 = 64;
template <bool CHUNK>
__device__ __forceinline__ void t4_policy_forward(const RolloutArgs &A, double *xs, double *h1s, int lane, int w) {
    const int XS = A.xs;
    t4_mlp_layer<true>(A.W1t, A.b1, A.K1p / 4, 0, A.H1p / MLP_NT, 0, xs, h1s, XS, lane, w);
    __syncthreads();
    if (!CHUNK) {
        t4_mlp_layer<true>(A.W2t, A.b2, A.K2p / 4, 0, A.H2p / MLP_NT, 0, h1s, xs, XS, lane, w);
        __syncthreads();
        t4_mlp_layer<false>(A.W3t, A.b3, A.K3p / 4, 0, A.Ap / MLP_NT, 0, xs, h1s, XS, lane, w);
        __syncthreads();
        return;
    }
    const int nitem = 2 * (A.Ap / MLP_NT);                  // <= T4_WARPS, checked on the host
    const int eh = w & 1, nt = w >> 1;
    double acc3[2][2][2];
    if (w < nitem) t4_dmma_init(acc3, A.b3, nt, lane);
    for (int c2 = 0; c2 < A.H2p; c2 += MLP_C2) {
        const int hi = c2 + MLP_C2 < A.H2p ? c2 + MLP_C2 : A.H2p;
        t4_mlp_layer<true>(A.W2t, A.b2, A.K2p / 4, c2 / MLP_NT, hi / MLP_NT, 0, h1s, xs, XS, lane, w);
        __syncthreads();
        if (w < nitem) t4_dmma_acc(acc3, A.W3t, A.K3p / 4, nt, c2 / 4, hi / 4, c2, xs, XS, eh, lane);
        __syncthreads();
    }
    if (w < nitem) t4_dmma_store<false>(acc3, nt * MLP_NT, h1s, XS, eh, lane);
    __syncthreads();
}

template <bool CHUNK, bool SNET, bool VALFS = false, bool CONS = false>
__global__ void __launch_bounds__(T4_THREADS, 1)
rollout_kernel_t4(const RolloutArgs A, const T4Off O) {
    extern __shared__ double smem[];
    typename std::conditional<CONS, T4CtxC, T4Ctx>::type x;
    x.lane = threadIdx.x & 31; x.w = threadIdx.x >> 5; x.o = O;
    if constexpr (CONS) {
        x.csb = A.cons + (size_t)blockIdx.x * CS_PER_BODY * c_m.nbody * 32;
        x.ccnt = 0u; x.lim_inst = 0u; x.lim_act = 0u; x.lim_pinst = 0u; x.lim_pact = 0u; x.cchg = 0; x.cflip = 0; x.cit = 0;
    }
    const int lane = x.lane, w = x.w;
    // allocate the whole Tensor Memory of this SM (1 CTA per SM) as scratch; base address comes back via smem
    __shared__ uint32_t s_tmem_base;
    if (w == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = s_tmem_base;
    x.tm = tmem_base + ((uint32_t)((w & 3) * 32) << 16);     // warps w and w + 4 share a lane quarter
    const bool cw = w < T4_CW;                             // chain warp: owns tree sweeps (helper warps: MLP / obs / reward only)
    const int S = c_m.nq - 2 + c_m.nv, nu = c_m.nu, nv = c_m.nv, nb = c_m.nbody, nq = c_m.nq;
    // MLP activations alias the axis / anchor / U rows: xs [A.xrows][XS], h1s [A.hrows][XS] (plan made on the host)
    const int XS = A.xs;
    double *xs = smem + (size_t)O.ax * 32;
    double *h1s = xs + (size_t)A.xrows * XS;
    const int T = A.cfg.horizon, E = A.cfg.n_env;
    const double dt = c_m.h * c_m.frame_skip;
    const int env = blockIdx.x * 32 + lane;
    const bool live = env < E;
    const int eid = live ? env : E - 1;
    const bool w0 = w == 0;

    T4Local l;
    double st[(2 * MAXV + T4_WARPS - 1) / T4_WARPS];     // this thread's strided share of the filtered state
    double raw[(2 * MAXV + T4_WARPS - 1) / T4_WARPS];
    double log_acc[EGP_LOG_SIZE];
    for (int k = 0; k < EGP_LOG_SIZE; k++) log_acc[k] = 0.0;
    log_acc[EGP_LOG_MIN_C_REWARD] = INFINITY; log_acc[EGP_LOG_MAX_C_REWARD] = -INFINITY;
    log_acc[EGP_LOG_MIN_EPISODE_REWARD] = INFINITY; log_acc[EGP_LOG_MAX_EPISODE_REWARD] = -INFINITY;
    double ep_reward = 0.0;
    int n_reset = 0, take, start, cur_t = 0;
    // 'valuefs' (ego_mimic_eval.py:156-159,167): running mean of the value-net outputs, continued across launches
    double vs_n = 0.0, vs_mean = 0.0, value = 0.0;
    if (VALFS && A.in.d_value_stat) { vs_n = A.in.d_value_stat[2 * eid]; vs_mean = A.in.d_value_stat[2 * eid + 1]; }

    auto draw_reset = [&](int r) {
        if (A.in.d_reset_take) {
            int rr = r < A.cfg.max_resets ? r : A.cfg.max_resets - 1;
            take = A.in.d_reset_take[(size_t)eid * A.cfg.max_resets + rr];
            start = A.in.d_reset_start[(size_t)eid * A.cfg.max_resets + rr];
        } else {
            uint32_t c[4] = {(uint32_t)eid, (uint32_t)r, (uint32_t)A.cfg.iteration, 0x52535421u ^ (uint32_t)(A.cfg.iteration >> 32)};
            philox4x32(c, (uint32_t)A.cfg.seed, (uint32_t)(A.cfg.seed >> 32));
            take = (int)(c[0] % (uint32_t)A.n_takes);
            int len = A.take_off[take + 1] - A.take_off[take];
            int span = len - A.cfg.episode_len - 2 * A.cfg.fr_margin;
            start = A.cfg.fr_margin + (span > 0 ? (int)(c[1] % (uint32_t)span) : 0);
        }
    };
    // reset: owners load their dofs of the expert frame, then sim.forward() (humanoid_v1.py:201-226)
    auto do_reset = [&]() {
        const double *row = A.rows + (size_t)(A.take_off[take] + start) * EGP_X_STRIDE;
        T4_FOR_OWN_DOFS(i, b) {
            x.at(O.v, i) = row[EGP_X_QVEL + i];
            if (i >= 6) x.at(O.q, i + 1) = row[EGP_X_QPOS + i + 1];
        }
        if (c_m.chain_warp[0] == w) for (int k = 0; k < 7; k++) x.at(O.q, k) = row[EGP_X_QPOS + k];
        __syncthreads();
        if (cw) t4_forward_only(x);
        for (int b = 1 + w; b < nb; b += T4_WARPS) {        // body quaternions of this thread's bodies
            const int qa = c_m.body_qposadr[b], nd = c_m.body_dofnum[b];
            quat_from_euler(x.at(O.q, qa), nd > 1 ? x.at(O.q, qa + 1) : 0.0, nd > 2 ? x.at(O.q, qa + 2) : 0.0, l.bqc[b]);
        }
    };
    // observation -> strided raw / filtered shares
    auto make_state = [&](double *dst_raw, double *dst) {
        double hd[4], vl[3];
        double q4[4] = {x.at(O.q, 3), x.at(O.q, 4), x.at(O.q, 5), x.at(O.q, 6)};
        double v3[3] = {x.at(O.v, 0), x.at(O.v, 1), x.at(O.v, 2)};
        de_heading(q4, hd);
        to_heading(v3, q4, vl);
        int s = 0;
        for (int k = w; k < S; k += T4_WARPS, s++) {
            double v = t4_obs_entry(x, k, hd[0], hd[1], hd[2], hd[3], vl[0], vl[1], vl[2]);
            dst_raw[s] = v;
            if (A.in.d_zf_mean) {
                v = (v - A.in.d_zf_mean[k]) / (A.in.d_zf_std[k] + 1e-8);
                if (A.cfg.zf_clip > 0.0) v = fmin(fmax(v, -A.cfg.zf_clip), A.cfg.zf_clip);
            }
            dst[s] = v;
        }
    };

    // evaluation roll-out: reset_env_state(state_pred[frame], env.data.qpos) (ego_mimic_eval.py:93-99): joint angles,
    // height and velocities from the predicted observation, root xy / heading kept from the simulated state
    // (align_human_state, utils/tools.py:71-75); the caller runs sim.forward() afterwards
    auto load_state_pred = [&]() {
        const double *sp = A.in.d_state_pred + (size_t)(A.take_off[take] + start + cur_t) * S;
        T4_FOR_OWN_DOFS(i, b) {
            if (i >= 3) x.at(O.v, i) = sp[nq - 2 + i];
            if (i >= 6) x.at(O.q, i + 1) = sp[i - 1];
        }
        if (c_m.chain_warp[0] == w) {
            const double qw = x.at(O.q, 3), qz = x.at(O.q, 6), hn = sqrt(qw * qw + qz * qz);
            const double hq[4] = {qw / hn, 0.0, 0.0, qz / hn};              // get_heading_q, utils/math.py:60-65
            double o4[4];
            quat_mul(hq, sp + 1, o4);
            x.at(O.q, 2) = sp[0];
            for (int k = 0; k < 4; k++) x.at(O.q, 3 + k) = o4[k];
            const double cz = hq[0] * hq[0] - hq[3] * hq[3], sz = 2.0 * hq[0] * hq[3];     // quat_mul_vec(hq, v)
            const double vx = sp[nq - 2], vy = sp[nq - 1];
            x.at(O.v, 0) = cz * vx - sz * vy;
            x.at(O.v, 1) = sz * vx + cz * vy;
            x.at(O.v, 2) = sp[nq];
        }
    };

    // Trajbatch rows are written row-contiguous: a [feature][env] tile in shared memory (the activation buffers, free at
    // that point of the step) is read transposed, warp = environment, lane = feature, so every row of 115 / 52 doubles
    // leaves as full 128-byte lines instead of one 8-byte store per environment 276 KB apart.
    auto write_rows = [&](double *dst, const double *tile, int width, size_t step_n) {       // dst row of env e: (e_global * T + t) * width
        if (!dst) return;
        for (int e = w; e < 32; e += T4_WARPS) {
            const long long ge = (long long)blockIdx.x * 32 + e;
            if (ge >= E) break;
            double *row = dst + ((size_t)ge * T + step_n) * width;
            for (int k = lane; k < width; k += 32) __stcs(row + k, tile[(size_t)k * XS + e]);   // streaming: must not evict the CTAs' L2-resident scratch
        }
    };
    auto stage_share = [&](double *tile, const double *share) {     // this thread's strided share -> tile[k][env]
        int s = 0;
        for (int k = w; k < S; k += T4_WARPS, s++) tile[(size_t)k * XS + lane] = share[s];
    };
    double pend[(2 * MAXV + T4_WARPS - 1) / T4_WARPS];   // next_state of the previous step, written with the next step's rows
    draw_reset(0);
    do_reset();
    if (A.cfg.eval_mode || A.in.d_init_qpos) {
        __syncthreads();
        if (A.in.d_init_qpos) {                             // env.set_state(qpos, qvel) right after reset (ego_forecast_eval.py:119-121)
            T4_FOR_OWN_DOFS(i, b) {
                x.at(O.v, i) = A.in.d_init_qvel[(size_t)eid * nv + i];
                if (i >= 6) x.at(O.q, i + 1) = A.in.d_init_qpos[(size_t)eid * nq + i + 1];
            }
            if (c_m.chain_warp[0] == w) for (int k = 0; k < 7; k++) x.at(O.q, k) = A.in.d_init_qpos[(size_t)eid * nq + k];
        } else {
            load_state_pred();
        }
        __syncthreads();
        if (cw) t4_forward_only(x);
    }
    make_state(raw, st);
    T4_FOR_OWN_DOFS(i, b) if (i >= 6) tm_st1(x.a_ctrl(i), 0.0);
    tm_wait_st();
    // state-LSTM h | c rows of this CTA, [2H][32] (s_net.initialize() at pre_episode, rnn.py:22-26)
    double *sn_g = SNET ? A.sn_state + (size_t)blockIdx.x * 2 * A.sn_H * 32 : nullptr;
    if (sn_g) for (int j = w; j < 2 * A.sn_H; j += T4_WARPS) sn_g[j * 32 + lane] = 0.0;

    for (int t = 0; t < T; t++) {
        const size_t n = (size_t)eid * T + t;
        if (A.out.d_qpos_traj && live) {                    // env.data.qpos / qvel before the step (ego_mimic_eval.py:136-138)
            for (int k = w; k < nq; k += T4_WARPS) A.out.d_qpos_traj[n * nq + k] = x.at(O.q, k);
            for (int k = w; k < nv; k += T4_WARPS) A.out.d_qvel_traj[n * nv + k] = x.at(O.v, k);
        }
        // ---- save the tree rows the MLP buffers alias (needed stale by the next sub-step's PD solve)
        T4_FOR_OWN_DOFS(i, b) for (int r = 0; r < 3; r++) l.sav_ax[i][r] = x.at(O.ax, 3 * i + r);
        for (int c = 0; c < c_m.nchain; c++)
            if (c_m.chain_warp[c] == w)
                for (int b = c_m.chain_lo[c]; b <= c_m.chain_hi[c]; b++)
                    for (int r = 0; r < 3; r++) l.sav_anc[b][r] = x.at(O.anc, 3 * b + r);
        __syncthreads();
        // ---- trajbatch rows of this step (raw observation, previous step's next_state) through the hidden-layer tile
        if (A.out.d_raw_obs) {
            stage_share(h1s, raw);
            __syncthreads();
            write_rows(A.out.d_raw_obs, h1s, S, t);
            __syncthreads();
        }
        if (A.out.d_next_states && t > 0) {
            stage_share(h1s, pend);
            __syncthreads();
            write_rows(A.out.d_next_states, h1s, S, t - 1);
            __syncthreads();
        }
        // ---- policy input cat(ctx[frame], state), feature-major; rows up to the padded K are zero
        int off = 0;
        if (A.ctx && !sn_g) {
            const double *cx = ctx_row(A, take, start, cur_t);
            for (int k = w; k < A.ctx_dim; k += T4_WARPS) xs[(size_t)k * XS + lane] = cx[k];
            off = A.ctx_dim;
        }
        stage_share(xs + (size_t)off * XS, st);
        for (int k = (sn_g ? S + A.sn_H : A.D) + w; k < A.xrows; k += T4_WARPS) xs[(size_t)k * XS + lane] = 0.0;
        __syncthreads();
        write_rows(A.out.d_states, xs + (size_t)off * XS, S, t);
        if (sn_g) {
            // ---- s_net LSTMCell step on the filtered state (video_forecast_net.py:89-93, rnn.py:36-43):
            // gates = W [state | h_prev] + b, 64 units (256 packed gate rows 4u + g) at a time through the dense-layer
            // routine, then c = sig(f) c + sig(i) tanh(g), h = sig(o) tanh(c); policy input = cat(ctx row, h)
            const int H = A.sn_H;
            for (int j = w; j < H; j += T4_WARPS) xs[(size_t)(S + j) * XS + lane] = sn_g[j * 32 + lane];
            __syncthreads();
            for (int c0 = 0; c0 < H; c0 += 64) {
                const int hi = c0 + 64 < H ? c0 + 64 : H;
                t4_mlp_layer<false>(A.sn_Wp, A.sn_b, A.sn_Kp / 4, c0 * 4 / MLP_NT, (hi * 4 + MLP_NT - 1) / MLP_NT, 0, xs, h1s, XS, lane, w);
                __syncthreads();
                for (int u = c0 + w; u < hi; u += T4_WARPS) {
                    const double *gr = h1s + (size_t)(u - c0) * 4 * XS + lane;
                    const double ig = 1.0 / (1.0 + exp(-gr[0])), fg = 1.0 / (1.0 + exp(-gr[XS]));
                    const double gg = tanh(gr[2 * XS]), og = 1.0 / (1.0 + exp(-gr[3 * XS]));
                    const double cn = fg * sn_g[(H + u) * 32 + lane] + ig * gg;
                    sn_g[(H + u) * 32 + lane] = cn;
                    sn_g[u * 32 + lane] = og * tanh(cn);
                }
                __syncthreads();
            }
            if (A.ctx) {
                const double *cx = ctx_row(A, take, start, cur_t);
                for (int k = w; k < A.ctx_dim; k += T4_WARPS) xs[(size_t)k * XS + lane] = cx[k];
            }
            for (int j = w; j < H; j += T4_WARPS) xs[(size_t)(A.ctx_dim + j) * XS + lane] = sn_g[j * 32 + lane];
            for (int k = A.D + w; k < A.xrows; k += T4_WARPS) xs[(size_t)k * XS + lane] = 0.0;
        }
        __syncthreads();
        if (VALFS) {
            // value = value_net(value_vs_net(state)) on cat(value context row, state); value_stat.push(value)
            if (A.vctx) {
                const double *cx = A.vctx + (ctx_row(A, take, start, cur_t) - A.ctx);
                for (int k = w; k < A.ctx_dim; k += T4_WARPS) xs[(size_t)k * XS + lane] = cx[k];
                __syncthreads();
            }
            RolloutArgs V = A;
            V.W1t = A.vW1t; V.b1 = A.vb1; V.W2t = A.vW2t; V.b2 = A.vb2; V.W3t = A.vW3t; V.b3 = A.vb3; V.Ap = A.vAp;
            t4_policy_forward<CHUNK>(V, xs, h1s, lane, w);
            value = h1s[lane];
            vs_n += 1.0;
            vs_mean += (value - vs_mean) / vs_n;
            if (live && w0 && A.out.d_values) A.out.d_values[n] = value;
            __syncthreads();
            // the dense layers overwrote the input rows: rebuild the policy input
            const double *cx = ctx_row(A, take, start, cur_t);
            for (int k = w; k < A.ctx_dim; k += T4_WARPS) xs[(size_t)k * XS + lane] = cx[k];
            stage_share(xs + (size_t)A.ctx_dim * XS, st);
            for (int k = A.D + w; k < A.xrows; k += T4_WARPS) xs[(size_t)k * XS + lane] = 0.0;
            __syncthreads();
        }
        t4_policy_forward<CHUNK>(A, xs, h1s, lane, w);
        bool mean_flag = A.cfg.mean_action != 0;
        if (A.in.d_mean_flag) mean_flag = mean_flag || A.in.d_mean_flag[n] != 0;
        else if (A.cfg.noise_rate < 1.0) {
            uint32_t c[4] = {(uint32_t)eid, (uint32_t)t, (uint32_t)A.cfg.iteration, 0x4d45414eu ^ (uint32_t)(A.cfg.iteration >> 32)};
            philox4x32(c, (uint32_t)A.cfg.seed, (uint32_t)(A.cfg.seed >> 32));
            mean_flag = mean_flag || (u01(c[0], c[1]) <= 1.0 - A.cfg.noise_rate);
        }
        // ---- sample (distributions.py / policy.py:12-15): a = mu + exp(log_std) eps, in place in the mean rows, all warps
        for (int a = w; a < nu; a += T4_WARPS) {
            double z = 0.0;
            if (!mean_flag) {
                if (A.in.d_eps) z = A.in.d_eps[n * nu + a];
                else {
                    double z0, z1;
                    normal2(A.cfg.seed, A.cfg.iteration, (uint32_t)eid, (uint32_t)(t * 64 + (a & ~1)), &z0, &z1);
                    z = (a & 1) ? z1 : z0;
                }
            }
            h1s[(size_t)a * XS + lane] += exp(A.log_std[a]) * z;
        }
        __syncthreads();
        write_rows(A.out.d_actions, h1s, nu, t);
        T4_FOR_OWN_DOFS(i, b) {
            if (i < 6) continue;
            tm_st1(x.a_ctrl(i), c_m.a_ref[i] + h1s[(size_t)(i - 6) * XS + lane] * c_m.a_scale[i]);
        }
        __syncthreads();
        // ---- restore the aliased tree rows
        T4_FOR_OWN_DOFS(i, b) for (int r = 0; r < 3; r++) x.at(O.ax, 3 * i + r) = l.sav_ax[i][r];
        for (int c = 0; c < c_m.nchain; c++)
            if (c_m.chain_warp[c] == w)
                for (int b = c_m.chain_lo[c]; b <= c_m.chain_hi[c]; b++)
                    for (int r = 0; r < 3; r++) x.at(O.anc, 3 * b + r) = l.sav_anc[b][r];
        // ---- env.step
        double prev_root[7];
        for (int k = 0; k < 7; k++) prev_root[k] = x.at(O.q, k);
        for (int b = 1 + w; b < nb; b += T4_WARPS) for (int k = 0; k < 4; k++) l.bqp[b][k] = l.bqc[b][k];
        __syncthreads();
        if (cw) {
            tm_wait_st();
            t4_do_simulation(x, c_m.frame_skip);
        }
        __syncthreads();
        cur_t += 1;
        const double head_z = x.at(O.xp, 3 * c_m.head_xp_slot + 2);
        const double lb = isnan(A.cfg.fix_head_lb) ? A.head_lb[take] - 0.1 : A.cfg.fix_head_lb;
        bool fail = head_z < lb;
        if (VALFS && A.cfg.eval_mode == 2) fail = value < 0.6 * vs_mean;        // ego_mimic_eval.py:167
        const bool end = cur_t >= (A.in.d_fix_len ? A.in.d_fix_len[eid] : A.cfg.episode_len);
        // expert frame of the reward, clamped to the take (a window may end on the take's last frame in evaluation)
        const int xfr = min(A.take_off[take] + start + cur_t, A.take_off[take + 1] - 1);
        const double *row = A.rows + (size_t)xfr * EGP_X_STRIDE;
        // body-quaternion terms of the reward for this thread's bodies (reward_function.py:35-41)
        double pose2 = 0.0, vd = 0.0;
        for (int b = 1 + w; b < nb; b += T4_WARPS) {
            const int qa = c_m.body_qposadr[b], nd = c_m.body_dofnum[b];
            quat_from_euler(x.at(O.q, qa), nd > 1 ? x.at(O.q, qa + 1) : 0.0, nd > 2 ? x.at(O.q, qa + 2) : 0.0, l.bqc[b]);
            double q1[4], qd[4];
            quat_inv(row + EGP_X_BQUAT + 4 * b, q1);
            quat_mul(l.bqc[b], q1, qd);
            double aa = acos(fmin(fmax(qd[0], -1.0), 1.0)) * c_m.b_diffw[b - 1];
            pose2 += aa * aa;
            double p1[4], pd[4], av[3], aang;
            quat_inv(l.bqp[b], p1);
            quat_mul(l.bqc[b], p1, pd);
            rot_from_quat(pd, av, &aang);
            if (1.0 - pd[0] < 1e-8) av[0] = 1.0;
            for (int k = 0; k < 3; k++) {
                double df = fabs(av[k] * aang / dt - row[EGP_X_BANGVEL + 3 * b + k]);
                vd += c_m.v_ord == 1 ? df : df * df;
            }
        }
        x.at(O.red, 2 * w) = pose2;
        x.at(O.red, 2 * w + 1) = vd;
        double nraw[(2 * MAXV + T4_WARPS - 1) / T4_WARPS], nst[(2 * MAXV + T4_WARPS - 1) / T4_WARPS];
        make_state(nraw, nst);
        __syncthreads();
        double rew = 0.0, info5[5] = {0, 0, 0, 0, 0};
        if (w0) {
            // sums in body order b = 1.. so the result does not depend on the warp split: re-add per warp share
            pose2 = 0.0; vd = 0.0;
#pragma unroll
            for (int k = 0; k < T4_WARPS; k++) { pose2 += x.at(O.red, 2 * k); vd += x.at(O.red, 2 * k + 1); }
            double q7[7];
            for (int k = 0; k < 7; k++) q7[k] = x.at(O.q, k);
            double lin[3], qi[4], qrel[4], axis[3], ang, rv[3], fdl[3], fda[3];
            for (int k = 0; k < 3; k++) lin[k] = (q7[k] - prev_root[k]) / dt;
            quat_inv(prev_root + 3, qi);
            quat_mul(q7 + 3, qi, qrel);
            rot_from_quat(qrel, axis, &ang);
            if (1.0 - qrel[0] < 1e-8) axis[0] = 1.0;
            if (ang > M_PI) ang -= 2 * M_PI; else if (ang < -M_PI) ang += 2 * M_PI;
            for (int k = 0; k < 3; k++) rv[k] = (axis[k] * ang) / dt;
            to_root(rv, prev_root + 3, fda);
            to_heading(lin, prev_root + 3, fdl);
            double rq[4];
            de_heading(q7 + 3, rq);
            double e2 = 0.0;
            for (int k = 0; k < EGP_NEE; k++) {
                const int xk = c_m.ee_xp_slot[k];
                double vv[3] = {x.at(O.xp, 3 * xk) - q7[0], x.at(O.xp, 3 * xk + 1) - q7[1], x.at(O.xp, 3 * xk + 2) - q7[2]}, ee[3];
                to_heading(vv, q7 + 3, ee);
                for (int r = 0; r < 3; r++) { double df = ee[r] - row[EGP_X_EE_POS + 3 * k + r]; e2 += df * df; }
            }
            double pose_dist = sqrt(pose2), vel_dist = c_m.v_ord == 1 ? vd : sqrt(vd), ee_dist = sqrt(e2);
            info5[0] = exp(-c_m.k_p * (pose_dist * pose_dist));
            info5[1] = exp(-c_m.k_v * (vel_dist * vel_dist));
            info5[2] = exp(-c_m.k_e * (ee_dist * ee_dist));
            double hd2 = q7[2] - row[EGP_X_QPOS + 2], q1[4], qd[4];
            quat_inv(row + EGP_X_RQ_RMH, q1);
            quat_mul(rq, q1, qd);
            double rqd = acos(fmin(fmax(qd[0], -1.0), 1.0));
            info5[3] = exp(-c_m.k_rh * (hd2 * hd2) - c_m.k_rq * (rqd * rqd));
            double l2 = 0.0, a2 = 0.0;
            for (int k = 0; k < 3; k++) {
                double dl = fdl[k] - row[EGP_X_RLINV_LOCAL + k], da = fda[k] - row[EGP_X_RANGV + k];
                l2 += dl * dl; a2 += da * da;
            }
            double ld = sqrt(l2), ad = sqrt(a2);
            info5[4] = exp(-c_m.k_rl * (ld * ld) - c_m.k_ra * (ad * ad));
            rew = c_m.w_p * info5[0] + c_m.w_v * info5[1] + c_m.w_e * info5[2] + c_m.w_rp * info5[3] + c_m.w_rv * info5[4];
            rew /= c_m.w_p + c_m.w_v + c_m.w_e + c_m.w_rp + c_m.w_rv;
            if (c_m.decay) rew *= 1.0 - (double)cur_t / A.cfg.episode_len;
            if (end) rew += A.cfg.end_reward;
        }
        // NaN guard (mj_checkPos/Vel analogue): every warp checks its share, warp 0 adds the reward
        bool bad = !isfinite(head_z) || (w0 && !isfinite(rew));
        { int s = 0; for (int k = w; k < S && !bad; k += T4_WARPS, s++) bad = !(fabs(nraw[s]) < 1e10); }
        __syncthreads();
        x.at(O.red, w) = bad ? 1.0 : 0.0;
        __syncthreads();
        {
            double nbad = 0.0;
#pragma unroll
            for (int k = 0; k < T4_WARPS; k++) nbad += x.at(O.red, k);
            bad = nbad > 0.0;
        }
        if (bad) {
            rew = 0.0; fail = true;
            for (int k = 0; k < 5; k++) info5[k] = 0.0;
            int s = 0;
            for (int k = w; k < S; k += T4_WARPS, s++) nst[s] = 0.0;
            if (w0) log_acc[EGP_LOG_NUM_NAN_RESETS] += 1.0;
        }
        const bool done = fail || end;
        if (A.out.d_next_states) { int s = 0; for (int k = w; k < S; k += T4_WARPS, s++) pend[s] = nst[s]; }
        if (live) {
            if (w0) {
                A.out.d_rewards[n] = rew;
                A.out.d_masks[n] = (done || t == T - 1) ? 0.0 : 1.0;
                A.out.d_exps[n] = mean_flag ? 0.0 : 1.0;
                A.out.d_v_metas[2 * n] = take;
                A.out.d_v_metas[2 * n + 1] = start;
                if (A.out.d_c_info) for (int k = 0; k < 5; k++) A.out.d_c_info[n * 5 + k] = info5[k];
            }
        }
        if (w0) {
            ep_reward += 1.0;
            log_acc[EGP_LOG_NUM_STEPS] += 1.0;
            log_acc[EGP_LOG_TOTAL_C_REWARD] += rew;
            log_acc[EGP_LOG_MIN_C_REWARD] = fmin(log_acc[EGP_LOG_MIN_C_REWARD], rew);
            log_acc[EGP_LOG_MAX_C_REWARD] = fmax(log_acc[EGP_LOG_MAX_C_REWARD], rew);
            for (int k = 0; k < 5; k++) log_acc[EGP_LOG_C_INFO + k] += info5[k];
            if (done || t == T - 1) {
                log_acc[EGP_LOG_NUM_EPISODES] += 1.0;
                log_acc[EGP_LOG_TOTAL_REWARD] += ep_reward;
                log_acc[EGP_LOG_MIN_EPISODE_REWARD] = fmin(log_acc[EGP_LOG_MIN_EPISODE_REWARD], ep_reward);
                log_acc[EGP_LOG_MAX_EPISODE_REWARD] = fmax(log_acc[EGP_LOG_MAX_EPISODE_REWARD], ep_reward);
                ep_reward = 0.0;
            }
        }
        // per-lane reset decisions differ across lanes but are identical across the 4 warps of a lane;
        // the tree sweeps contain CTA barriers, so every lane runs the reset sweep and commits selectively
        const bool need_reset = done && t < T - 1;
        const int any_reset = __syncthreads_or(need_reset ? 1 : 0);
        if (any_reset) {
            // Lanes that continue their episode must keep their (stale) tree data across the reset sweep that
            // the whole CTA runs: snapshot it (TMEM accesses stay warp-uniform), and give those lanes their own
            // current state as sweep input so no garbage is produced; restore with a per-lane select afterwards.
            double k_cin[MAXB][10], k_C[MAXV], k_ax[MAXV][3], k_anc[MAXB][3], k_bqc[MAXB][4], k_xp[3 * (EGP_NEE + 1)];
            T4_FOR_OWN_DOFS(i, b) { k_C[i] = tm_ld1(x.a_c(i)); for (int r = 0; r < 3; r++) k_ax[i][r] = x.at(O.ax, 3 * i + r); }
            for (int c = 0; c < c_m.nchain; c++)
                if (c_m.chain_warp[c] == w)
                    for (int b = c_m.chain_lo[c]; b <= c_m.chain_hi[c]; b++) {
                        x.ld_cin(b, k_cin[b]);
                        for (int r = 0; r < 3; r++) k_anc[b][r] = x.at(O.anc, 3 * b + r);
                    }
            for (int b = 1 + w; b < nb; b += T4_WARPS) for (int k = 0; k < 4; k++) k_bqc[b][k] = l.bqc[b][k];
            if (w0) for (int k = 0; k < 3 * (EGP_NEE + 1); k++) k_xp[k] = x.at(O.xp, k);
            const bool tele = A.cfg.eval_mode && need_reset && !end;   // fail-safe: replace the state, keep the clock
            if (need_reset && !tele) {
                n_reset++; draw_reset(n_reset); cur_t = 0;
                if (sn_g) for (int j = w; j < 2 * A.sn_H; j += T4_WARPS) sn_g[j * 32 + lane] = 0.0;
            }
            if (tele && w0 && n_reset == 0) log_acc[EGP_LOG_NUM_FAILSAFE_RESETS] += 1.0;   // first episode of the environment only
            __syncthreads();
            if (tele) load_state_pred();
            else if (need_reset) {
                const double *r0 = A.rows + (size_t)(A.take_off[take] + start) * EGP_X_STRIDE;
                T4_FOR_OWN_DOFS(i, b) {
                    x.at(O.v, i) = r0[EGP_X_QVEL + i];
                    if (i >= 6) x.at(O.q, i + 1) = r0[EGP_X_QPOS + i + 1];
                }
                if (c_m.chain_warp[0] == w) for (int k = 0; k < 7; k++) x.at(O.q, k) = r0[EGP_X_QPOS + k];
            }
            __syncthreads();
            if (cw) t4_forward_only(x);
            // restore parked lanes (uniform TMEM traffic, per-lane select of the value)
            T4_FOR_OWN_DOFS(i, b) {
                const double cur = tm_ld1(x.a_c(i));
                tm_st1(x.a_c(i), need_reset ? cur : k_C[i]);
                if (!need_reset) for (int r = 0; r < 3; r++) x.at(O.ax, 3 * i + r) = k_ax[i][r];
            }
            for (int c = 0; c < c_m.nchain; c++)
                if (c_m.chain_warp[c] == w)
                    for (int b = c_m.chain_lo[c]; b <= c_m.chain_hi[c]; b++) {
                        double cur[10];
                        x.ld_cin(b, cur);
                        if (!need_reset) for (int r = 0; r < 10; r++) cur[r] = k_cin[b][r];
                        x.st_cin(b, cur);
                        if (!need_reset) for (int r = 0; r < 3; r++) x.at(O.anc, 3 * b + r) = k_anc[b][r];
                    }
            tm_wait_st();
            if (need_reset) {
                if (!tele)                                  // set_state() alone leaves env.bquat stale (ego_mimic_eval.py:98)
                    for (int b = 1 + w; b < nb; b += T4_WARPS) {
                        const int qa = c_m.body_qposadr[b], nd = c_m.body_dofnum[b];
                        quat_from_euler(x.at(O.q, qa), nd > 1 ? x.at(O.q, qa + 1) : 0.0, nd > 2 ? x.at(O.q, qa + 2) : 0.0, l.bqc[b]);
                    }
                make_state(raw, st);
            } else {
                for (int b = 1 + w; b < nb; b += T4_WARPS) for (int k = 0; k < 4; k++) l.bqc[b][k] = k_bqc[b][k];
                if (w0) for (int k = 0; k < 3 * (EGP_NEE + 1); k++) x.at(O.xp, k) = k_xp[k];
                int s2 = 0;
                for (int k = w; k < S; k += T4_WARPS, s2++) { st[s2] = nst[s2]; raw[s2] = nraw[s2]; }
            }
        } else {
            int s = 0;
            for (int k = w; k < S; k += T4_WARPS, s++) { st[s] = nst[s]; raw[s] = nraw[s]; }
        }
        __syncthreads();
    }
    if (VALFS && live && w0 && A.in.d_value_stat) { A.in.d_value_stat[2 * env] = vs_n; A.in.d_value_stat[2 * env + 1] = vs_mean; }
    if (live) {
        if (A.out.d_final_qpos) for (int k = w; k < nq; k += T4_WARPS) A.out.d_final_qpos[(size_t)env * nq + k] = x.at(O.q, k);
        if (A.out.d_final_qvel) for (int k = w; k < nv; k += T4_WARPS) A.out.d_final_qvel[(size_t)env * nv + k] = x.at(O.v, k);
    }
    if (A.out.d_next_states) {                              // next_state of the last step (tree rows are dead now)
        __syncthreads();
        stage_share(h1s, pend);
        __syncthreads();
        write_rows(A.out.d_next_states, h1s, S, T - 1);
    }
    // LoggerRL sums / extrema: fixed-order butterfly over the CTA's environments, one partial row per CTA, merged in CTA
    // order by logger_merge_kernel - the same bits on every run (no floating-point atomics)
    if (A.log_part && w0) {
        if (!live) {
            for (int k = 0; k < EGP_LOG_SIZE; k++) log_acc[k] = 0.0;
            log_acc[EGP_LOG_MIN_C_REWARD] = INFINITY; log_acc[EGP_LOG_MAX_C_REWARD] = -INFINITY;
            log_acc[EGP_LOG_MIN_EPISODE_REWARD] = INFINITY; log_acc[EGP_LOG_MAX_EPISODE_REWARD] = -INFINITY;
        }
#pragma unroll
        for (int k = 0; k < EGP_LOG_SIZE; k++) {
            double v = log_acc[k];
            const bool mn = k == EGP_LOG_MIN_C_REWARD || k == EGP_LOG_MIN_EPISODE_REWARD;
            const bool mx = k == EGP_LOG_MAX_C_REWARD || k == EGP_LOG_MAX_EPISODE_REWARD;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double u = __shfl_xor_sync(0xffffffffu, v, o);
                v = mn ? fmin(v, u) : (mx ? fmax(v, u) : v + u);
            }
            if (lane == 0) A.log_part[(size_t)blockIdx.x * EGP_LOG_SIZE + k] = v;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

// d_logger[k] = ordered reduction of the per-CTA partial rows (sum; min / max for the extrema slots)
__global__ void logger_merge_kernel(const double *__restrict__ part, int nblk, double *__restrict__ out) {
    const int k = threadIdx.x;
    if (k >= EGP_LOG_SIZE) return;
    const bool mn = k == EGP_LOG_MIN_C_REWARD || k == EGP_LOG_MIN_EPISODE_REWARD;
    const bool mx = k == EGP_LOG_MAX_C_REWARD || k == EGP_LOG_MAX_EPISODE_REWARD;
    double v = mn ? INFINITY : (mx ? -INFINITY : 0.0);
    for (int b = 0; b < nblk; b++) {
        const double u = part[(size_t)b * EGP_LOG_SIZE + k];
        v = mn ? fmin(v, u) : (mx ? fmax(v, u) : v + u);
    }
    out[k] = v;
}

// transposes W [out][in] -> Wt [in][outp] (zero padded), biases padded (one-warp V1 kernel)
__global__ void transpose_pad_kernel(const double *__restrict__ W, const double *__restrict__ b, int out, int in, int outp,
                                     double *__restrict__ Wt, double *__restrict__ bp) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < in * outp) {
        int k = idx / outp, j = idx % outp;
        Wt[idx] = j < out ? W[(size_t)j * in + k] : 0.0;
    }
    if (idx < outp) bp[idx] = idx < out ? b[idx] : 0.0;
}

// ---- debug / parity kernels ---------------------------------------------------------------------
__global__ void forward_debug_kernel(int n, const double *qpos, const double *qvel, const double *ctrl, double *bias,
                                     double *xpos, double *qacc) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    EnvData e;
    const int nq = c_m.nq, nv = c_m.nv, nb = c_m.nbody;
    for (int k = 0; k < nq; k++) e.q[k] = qpos[(size_t)i * nq + k];
    for (int k = 0; k < nv; k++) e.v[k] = qvel[(size_t)i * nv + k];
    pass_kinematics(e);
    for (int k = 0; k < nv; k++) e.tau[k] = (k >= 6 && ctrl) ? ctrl[(size_t)i * c_m.nu + k - 6] : 0.0;
    forward_dynamics(e);
    for (int k = 0; k < nv; k++) { bias[(size_t)i * nv + k] = e.C[k]; qacc[(size_t)i * nv + k] = e.qacc[k]; }
    for (int b = 0; b < nb; b++) for (int r = 0; r < 3; r++) xpos[((size_t)i * nb + b) * 3 + r] = e.xp[b][r];
}

__global__ void env_step_debug_kernel(int n, double *qpos, double *qvel, const double *action, double *obs_out,
                                      double *head_z, double *torque0) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    EnvData e;
    const int nq = c_m.nq, nv = c_m.nv, nu = c_m.nu, S = nq - 2 + nv;
    for (int k = 0; k < nq; k++) e.q[k] = qpos[(size_t)i * nq + k];
    for (int k = 0; k < nv; k++) e.v[k] = qvel[(size_t)i * nv + k];
    env_forward(e);
    double ctrl[MAXV], obs[2 * MAXV];
    for (int k = 0; k < 6; k++) ctrl[k] = 0.0;
    for (int a = 0; a < nu; a++) ctrl[6 + a] = c_m.a_ref[6 + a] + action[(size_t)i * nu + a] * c_m.a_scale[6 + a];
    for (int s = 0; s < c_m.frame_skip; s++) env_substep(e, ctrl, s == 0 ? torque0 + (size_t)i * nu : nullptr);
    for (int k = 0; k < nq; k++) qpos[(size_t)i * nq + k] = e.q[k];
    for (int k = 0; k < nv; k++) qvel[(size_t)i * nv + k] = e.v[k];
    env_obs(e, obs);
    for (int k = 0; k < S; k++) obs_out[(size_t)i * S + k] = obs[k];
    head_z[i] = e.xp[c_m.head_body][2];
}

// gen_expert.py:28-83: one thread per frame; velocities by finite differences against the previous frame
__global__ void expert_features_kernel(int L, const double *qpos, double *rows, double *head_z, double *extras) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L) return;
    const int nq = c_m.nq, nb = c_m.nbody;
    const double dt = c_m.h * c_m.frame_skip;
    EnvData e;
    double *row = rows + (size_t)i * EGP_X_STRIDE;
    for (int k = 0; k < EGP_X_STRIDE; k++) row[k] = 0.0;
    for (int k = 0; k < nq; k++) { e.q[k] = qpos[(size_t)i * nq + k]; row[EGP_X_QPOS + k] = e.q[k]; }
    for (int k = 0; k < c_m.nv; k++) e.v[k] = 0.0;
    pass_kinematics(e);
    de_heading(e.q + 3, row + EGP_X_RQ_RMH);
    env_ee_pos(e, row + EGP_X_EE_POS);
    double bq[4 * MAXB];
    env_body_quat(e.q, bq);
    for (int k = 0; k < 4 * nb; k++) row[EGP_X_BQUAT + k] = bq[k];
    head_z[i] = e.xp[c_m.head_body][2];
    if (extras) {       // gen_expert.py:49-50,44: head_pos = get_body_com('Head'), com = subtree_com[0], ee_wpos = get_ee_pos(None)
        double *x = extras + (size_t)i * EGP_XE_STRIDE;
        double mt = 0.0, mc[3] = {0.0, 0.0, 0.0};
        for (int b = 0; b < nb; b++) { mt += e.cin[b][0]; for (int k = 0; k < 3; k++) mc[k] += e.cin[b][1 + k]; }
        for (int k = 0; k < 3; k++) {
            x[EGP_XE_HEAD_POS + k] = e.xp[c_m.head_body][k];
            x[EGP_XE_COM + k] = mc[k] / mt + e.q[k];            // first moments are kept about the root position
        }
        for (int j = 0; j < EGP_NEE; j++)
            for (int k = 0; k < 3; k++) x[EGP_XE_EE_WPOS + 3 * j + k] = e.xp[c_m.ee_body[j]][k];   // world frame (humanoid_v1.py:106-110)
    }
    // frame 0 copies the finite differences of frame 1 (gen_expert.py:67-70,76)
    const int ic = i > 0 ? i : (L > 1 ? 1 : 0);
    if (L > 1) {
        const double *cur = qpos + (size_t)ic * nq, *prev = qpos + (size_t)(ic - 1) * nq;
        double lin[3], qi[4], qrel[4], axis[3], ang, rv[3];
        for (int k = 0; k < 3; k++) lin[k] = (cur[k] - prev[k]) / dt;
        quat_inv(prev + 3, qi);
        quat_mul(cur + 3, qi, qrel);
        rot_from_quat(qrel, axis, &ang);
        if (1.0 - qrel[0] < 1e-8) axis[0] = 1.0;
        if (ang > M_PI) ang -= 2 * M_PI; else if (ang < -M_PI) ang += 2 * M_PI;
        for (int k = 0; k < 3; k++) rv[k] = axis[k] * ang / dt;
        double *qv = row + EGP_X_QVEL;
        qv[0] = lin[0]; qv[1] = lin[1]; qv[2] = lin[2];
        to_root(rv, prev + 3, qv + 3);
        for (int k = 7; k < nq; k++) qv[k - 1] = (cur[k] - prev[k]) / dt;
        to_heading(lin, cur + 3, row + EGP_X_RLINV_LOCAL);
        for (int k = 0; k < 3; k++) row[EGP_X_RANGV + k] = qv[3 + k];
        double bqp[4 * MAXB], bqc[4 * MAXB];
        env_body_quat(prev, bqp);
        env_body_quat(cur, bqc);
        for (int b = 0; b < nb; b++) {
            double p1[4], pd[4], av[3], aang;
            quat_inv(bqp + 4 * b, p1);
            quat_mul(bqc + 4 * b, p1, pd);
            rot_from_quat(pd, av, &aang);
            if (1.0 - pd[0] < 1e-8) av[0] = 1.0;
            for (int k = 0; k < 3; k++) row[EGP_X_BANGVEL + 3 * b + k] = av[k] * aang / dt;
        }
    }
}

// out[n][ctx_dim + S] = cat(ctx[frame(n)], states[n]); t within the episode is recovered from the masks:
// rows are env-major with `horizon` rows per env, an episode starts at t = 0 or after a mask == 0 row.
__global__ void build_input_kernel(const double *__restrict__ states, const int32_t *__restrict__ v_metas,
                                   const double *__restrict__ masks, long long n_env, int horizon, int S,
                                   const int32_t *__restrict__ take_off, const double *__restrict__ ctx, int ctx_dim,
                                   double *__restrict__ x) {
    // one warp per env row-block: lane-parallel copy, serial walk over t for the episode-relative index
    const int lane = threadIdx.x & 31;
    const long long env = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (env >= n_env) return;
    const int D = ctx_dim + S;
    int cur_t = 0;
    for (int t = 0; t < horizon; t++) {
        const long long n = env * horizon + t;
        const long long frame = (long long)take_off[v_metas[2 * n]] + v_metas[2 * n + 1] + cur_t;
        const double *cx = ctx + frame * ctx_dim;
        double *xr = x + n * D;
        for (int k = lane; k < ctx_dim; k += 32) xr[k] = cx[k];
        for (int k = lane; k < S; k += 32) xr[ctx_dim + k] = states[n * S + k];
        cur_t = masks[n] == 0.0 ? 0 : cur_t + 1;
    }
}

}  // namespace egp

using namespace egp;

extern "C" {

int egp_model_create(const EgpModelDesc *s, int device, EgpModel **out) {
    if (!s || !out) { set_error("egp_model_create: null argument"); return EGP_EINVAL; }
    EgpModel *m = new EgpModel();
    memset(m, 0, sizeof(*m));
    const char *why = nullptr;
    const int frc = fill_dev_model(s, m->host, &why);
    if (frc != EGP_OK) {
        delete m;
        set_error("egp_model_create: %s (nq=%d nv=%d nu=%d nbody=%d)", why ? why : "bad model", s->nq, s->nv, s->nu, s->nbody);
        return frc;
    }
    m->device = device;
    *out = m;
    return EGP_OK;
}

// solref (timeconst, dampratio) / solimp (d0, dwidth, width, midpoint, power) -> the constants of limit_row; one set shared
// by limit and contact rows (the XML sets neither: MuJoCo's defaults)
static int set_solparams(DevModel &d, const double *solref, const double *solimp, const char *who) {
    const double sr[2] = {solref ? solref[0] : 0.02, solref ? solref[1] : 1.0};
    const double si[5] = {solimp ? solimp[0] : 0.9, solimp ? solimp[1] : 0.95, solimp ? solimp[2] : 0.001, solimp ? solimp[3] : 0.5,
                          solimp ? solimp[4] : 2.0};
    if (!(sr[0] > 0.0) || !(sr[1] > 0.0)) { set_error("%s: only the (timeconst, dampratio) form of solref is supported", who); return EGP_EINVAL; }
    auto clampi = [](double x) { return x < 1e-4 ? 1e-4 : (x > 0.9999 ? 0.9999 : x); };
    const double tc = sr[0] < 2.0 * d.h ? 2.0 * d.h : sr[0];                                 // refsafe
    d.lim_d0 = clampi(si[0]); d.lim_dw = clampi(si[1]); d.lim_width = si[2] < 0.0 ? 0.0 : si[2]; d.lim_mid = clampi(si[3]);
    d.lim_pow = si[4] < 1.0 ? 1.0 : si[4];
    d.lim_k = 1.0 / (d.lim_dw * d.lim_dw * tc * tc * sr[1] * sr[1]);
    d.lim_b = 2.0 / (d.lim_dw * tc);
    return EGP_OK;
}

static void unbind(EgpModel *m) {
    if (m->device >= 0 && m->device < 64 && g_bound_model[m->device] == m) g_bound_model[m->device] = nullptr;   // re-upload on next use
}

int egp_model_set_joint_limits(EgpModel *m, const double *range, const double *invweight0, const double *solref,
                               const double *solimp) {
    if (!m) { set_error("egp_model_set_joint_limits: null model"); return EGP_EINVAL; }
    DevModel &d = m->host;
    unbind(m);
    if (!range) { d.limits = 0; return EGP_OK; }
    if (!invweight0) { set_error("egp_model_set_joint_limits: invweight0 is required with a range table"); return EGP_EINVAL; }
    for (int i = 0; i < d.nv; i++) {
        d.lim_lo[i] = range[2 * i]; d.lim_hi[i] = range[2 * i + 1]; d.lim_iw[i] = invweight0[i];
        if (i < 6 && d.body_dofnum[0] == 6) { d.lim_lo[i] = 0.0; d.lim_hi[i] = 0.0; }      // the free root has no range
        if (d.lim_lo[i] < d.lim_hi[i] && !(invweight0[i] > 0.0)) { set_error("egp_model_set_joint_limits: invweight0[%d] must be positive", i); return EGP_EINVAL; }
    }
    const int rc = set_solparams(d, solref, solimp, "egp_model_set_joint_limits");
    if (rc != EGP_OK) return rc;
    d.limits = 1;
    return EGP_OK;
}

int egp_model_set_contacts(EgpModel *m, const int32_t *geom_type, const double *geom_size, const double *geom_p0,
                           const double *geom_p1, const double *body_invweight0, double margin, double friction,
                           const double *solref, const double *solimp) {
    if (!m) { set_error("egp_model_set_contacts: null model"); return EGP_EINVAL; }
    DevModel &d = m->host;
    unbind(m);
    if (!geom_type) { d.contacts = 0; return EGP_OK; }
    if (!geom_size || !geom_p0 || !geom_p1 || !body_invweight0 || !(friction > 0.0) || margin < 0.0) {
        set_error("egp_model_set_contacts: geom tables, body_invweight0, friction > 0 and margin >= 0 are required");
        return EGP_EINVAL;
    }
    for (int b = 0; b < d.nbody; b++) {
        if (geom_type[b] < 0 || geom_type[b] > 2) { set_error("egp_model_set_contacts: geom_type[%d] = %d (0 sphere, 1 capsule, 2 box)", b, geom_type[b]); return EGP_EINVAL; }
        d.geom_type[b] = geom_type[b];
        for (int k = 0; k < 3; k++) { d.geom_size[b][k] = geom_size[3 * b + k]; d.geom_p0[b][k] = geom_p0[3 * b + k]; d.geom_p1[b][k] = geom_p1[3 * b + k]; }
        d.body_iw[b] = body_invweight0[2 * b];
    }
    const int rc = set_solparams(d, solref, solimp, "egp_model_set_contacts");
    if (rc != EGP_OK) return rc;
    d.con_margin = margin; d.con_mu = friction;
    d.contacts = 1;
    return EGP_OK;
}

/* number of constrained solves (per environment and sub-step) since the last reset that ran into the iteration cap (100) of
 * the active-set loop without reaching its fixed point; synchronises the device */
/* block-sweep kernel: out[0] = sub-steps, out[1] = active-set passes they took (first CTA of every launch); synchronises */
int egp_cons_passes(int64_t *out2, int reset) {
    unsigned long long v[2] = {0, 0}, z[2] = {0, 0};
    if (!out2 || cudaMemcpyFromSymbol(v, g_cons_passes, sizeof v) != cudaSuccess) return EGP_ECUDA;
    if (reset && cudaMemcpyToSymbol(g_cons_passes, z, sizeof z) != cudaSuccess) return EGP_ECUDA;
    out2[0] = (int64_t)v[0]; out2[1] = (int64_t)v[1];
    return EGP_OK;
}

int64_t egp_cons_cap_hits(int reset) {
    unsigned long long v = 0, z = 0;
    if (cudaMemcpyFromSymbol(&v, g_cons_cap_hits, sizeof v) != cudaSuccess) return -1;
    if (reset && cudaMemcpyToSymbol(g_cons_cap_hits, &z, sizeof z) != cudaSuccess) return -1;
    return (int64_t)v;
}

void egp_model_destroy(EgpModel *m) {
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->device >= 0 && m->device < 64 && g_bound_model[m->device] == m) g_bound_model[m->device] = nullptr;
    cudaFree(m->d_take_off); cudaFree(m->d_rows); cudaFree(m->d_head_lb); cudaFree(m->d_ctx); cudaFree(m->d_wbuf); cudaFree(m->d_cons);
    delete m;
}

int egp_expert_upload(EgpModel *m, int n_takes, const int32_t *take_off, const double *rows, const double *head_lb,
                      const double *ctx, int ctx_dim) {
    if (!m || n_takes < 1 || !take_off || !rows || !head_lb) { set_error("egp_expert_upload: bad argument"); return EGP_EINVAL; }
    EGP_CUDA(cudaSetDevice(m->device));
    cudaFree(m->d_take_off); cudaFree(m->d_rows); cudaFree(m->d_head_lb); cudaFree(m->d_ctx);
    m->d_take_off = nullptr; m->d_rows = nullptr; m->d_head_lb = nullptr; m->d_ctx = nullptr;
    long long total = take_off[n_takes];
    EGP_CUDA(cudaMalloc(&m->d_take_off, sizeof(int32_t) * (n_takes + 1)));
    EGP_CUDA(cudaMalloc(&m->d_rows, sizeof(double) * total * EGP_X_STRIDE));
    EGP_CUDA(cudaMalloc(&m->d_head_lb, sizeof(double) * n_takes));
    EGP_CUDA(cudaMemcpy(m->d_take_off, take_off, sizeof(int32_t) * (n_takes + 1), cudaMemcpyHostToDevice));
    EGP_CUDA(cudaMemcpy(m->d_rows, rows, sizeof(double) * total * EGP_X_STRIDE, cudaMemcpyHostToDevice));
    EGP_CUDA(cudaMemcpy(m->d_head_lb, head_lb, sizeof(double) * n_takes, cudaMemcpyHostToDevice));
    if (ctx && ctx_dim > 0) {
        EGP_CUDA(cudaMalloc(&m->d_ctx, sizeof(double) * total * ctx_dim));
        EGP_CUDA(cudaMemcpy(m->d_ctx, ctx, sizeof(double) * total * ctx_dim, cudaMemcpyHostToDevice));
    }
    m->n_takes = n_takes; m->ctx_dim = ctx ? ctx_dim : 0; m->total_frames = total;
    return EGP_OK;
}

int egp_expert_features_f64(EgpModel *m, int L, const double *d_qpos, double *d_rows, double *d_head_z, void *stream) {
    return egp_expert_features_ex_f64(m, L, d_qpos, d_rows, d_head_z, nullptr, stream);
}

int egp_expert_features_ex_f64(EgpModel *m, int L, const double *d_qpos, double *d_rows, double *d_head_z, double *d_extras,
                               void *stream) {
    if (!m || L < 1 || !d_qpos || !d_rows || !d_head_z) { set_error("egp_expert_features_f64: bad argument"); return EGP_EINVAL; }
    int rc = bind_model(m);
    if (rc) return rc;
    expert_features_kernel<<<(L + 63) / 64, 64, 0, (cudaStream_t)stream>>>(L, d_qpos, d_rows, d_head_z, d_extras);
    EGP_CHECK_LAUNCH("expert_features_kernel");
    return EGP_OK;
}

int egp_forward_debug_f64(EgpModel *m, int n, const double *d_qpos, const double *d_qvel, const double *d_ctrl,
                          double *d_bias, double *d_xpos, double *d_qacc, void *stream) {
    if (!m || n < 1 || !d_qpos || !d_qvel || !d_bias || !d_xpos || !d_qacc) { set_error("egp_forward_debug_f64: bad argument"); return EGP_EINVAL; }
    int rc = bind_model(m);
    if (rc) return rc;
    forward_debug_kernel<<<(n + 31) / 32, 32, 0, (cudaStream_t)stream>>>(n, d_qpos, d_qvel, d_ctrl, d_bias, d_xpos, d_qacc);
    EGP_CHECK_LAUNCH("forward_debug_kernel");
    return EGP_OK;
}

int egp_env_step_debug_f64(EgpModel *m, int n, double *d_qpos, double *d_qvel, const double *d_action, double *d_obs,
                           double *d_head_z, double *d_torque0, void *stream) {
    if (!m || n < 1 || !d_qpos || !d_qvel || !d_action || !d_obs || !d_head_z || !d_torque0) {
        set_error("egp_env_step_debug_f64: bad argument");
        return EGP_EINVAL;
    }
    int rc = bind_model(m);
    if (rc) return rc;
    env_step_debug_kernel<<<(n + 31) / 32, 32, 0, (cudaStream_t)stream>>>(n, d_qpos, d_qvel, d_action, d_obs, d_head_z, d_torque0);
    EGP_CHECK_LAUNCH("env_step_debug_kernel");
    return EGP_OK;
}

int egp_build_input_f64(EgpModel *m, const double *d_states, const int32_t *d_v_metas, const double *d_masks, int64_t n,
                        int horizon, double *d_x, void *stream) {
    if (!m || !m->d_ctx || n < 1 || horizon < 1 || n % horizon || !d_states || !d_v_metas || !d_masks || !d_x) {
        set_error("egp_build_input_f64: bad argument (needs context rows uploaded and n %% horizon == 0)");
        return EGP_EINVAL;
    }
    long long n_env = n / horizon;
    int S = m->host.nq - 2 + m->host.nv;
    long long threads = n_env * 32;
    build_input_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_states, d_v_metas, d_masks, n_env, horizon, S, m->d_take_off, m->d_ctx, m->ctx_dim, d_x);
    EGP_CHECK_LAUNCH("build_input_kernel");
    return EGP_OK;
}

int egp_rollout_f64(EgpModel *m, const EgpPolicyWeights *pol, const EgpRolloutCfg *cfg, const EgpRolloutIn *in,
                    const EgpTrajOut *out, void *stream) {
    if (!m || !pol || !cfg || !out || !m->d_rows) { set_error("egp_rollout_f64: null argument or experts not uploaded"); return EGP_EINVAL; }
    const DevModel &d = m->host;
    const int S = d.nq - 2 + d.nv;
    if (cfg->n_env < 1 || cfg->horizon < 1 || cfg->episode_len < 1) { set_error("egp_rollout_f64: bad sizes"); return EGP_EINVAL; }
    const bool ctx_override = in && in->d_ctx;
    const int ctx_dim = ctx_override ? in->ctx_dim : m->ctx_dim;
    if (ctx_override && (in->ctx_dim < 1 || (in->ctx_mode == 1 && (!in->d_win_off || in->ctx_T < cfg->episode_len)) ||
                         (in->ctx_mode == 2 && !in->d_win_off) || in->ctx_mode < 0 || in->ctx_mode > 2)) {
        set_error("egp_rollout_f64: bad context table (dim %d mode %d T %d)", in->ctx_dim, in->ctx_mode, in->ctx_T);
        return EGP_EINVAL;
    }
    const int snH = in ? in->snet_hdim : 0;
    if (snH < 0 || (snH & 1) || (snH > 0 && (!in->d_snet_W || !in->d_snet_b || !in->d_snet_state))) {
        set_error("egp_rollout_f64: bad state-LSTM arguments (hdim %d must be even, W / b / state scratch required)", snH);
        return EGP_EINVAL;
    }
    if (pol->in_dim != (snH ? snH : S) + ctx_dim || pol->out_dim != d.nu) {
        set_error("egp_rollout_f64: policy dims (in %d out %d) do not match obs %d + ctx %d / nu %d", pol->in_dim, pol->out_dim, S, ctx_dim, d.nu);
        return EGP_ESIZE;
    }
    if (!out->d_states || !out->d_actions || !out->d_masks || !out->d_rewards || !out->d_exps || !out->d_v_metas) {
        set_error("egp_rollout_f64: required trajbatch outputs missing");
        return EGP_EINVAL;
    }
    if (in && ((in->d_reset_take == nullptr) != (in->d_reset_start == nullptr))) { set_error("egp_rollout_f64: reset lists must come in pairs"); return EGP_EINVAL; }
    if (in && in->d_reset_take && cfg->max_resets < 1) { set_error("egp_rollout_f64: max_resets < 1"); return EGP_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    int rc = bind_model(m);
    if (rc) return rc;
    RolloutArgs A;
    memset(&A, 0, sizeof A);
    A.cfg = *cfg;
    if (in) A.in = *in;
    A.out = *out;
    A.take_off = m->d_take_off; A.rows = m->d_rows; A.head_lb = m->d_head_lb;
    A.n_takes = m->n_takes;
    A.ctx = ctx_override ? in->d_ctx : m->d_ctx;
    A.ctx_dim = ctx_dim;
    A.ctx_mode = ctx_override ? in->ctx_mode : 0;
    A.ctx_T = ctx_override ? in->ctx_T : 0;
    A.win_off = ctx_override ? in->d_win_off : nullptr;
    A.D = pol->in_dim; A.H1 = pol->h1; A.H2 = pol->h2; A.A = pol->out_dim;
    const int sn_in = snH ? S + snH : 0;
    A.sn_H = snH;
    int blocks = (cfg->n_env + ENVS_PER_CTA - 1) / ENVS_PER_CTA;
    // variant selection: T4 (8 warps per 32 envs, tree data in shared memory, policy MLP on the FP64 tensor cores) when
    // the model and the policy width fit, else the one-warp V1 kernel; EGP_ROLLOUT_VARIANT=1 forces V1 (A/B parity runs)
    T4Off O = t4_offsets(d);
    // T4 MLP plan inside the alias window [O.ax, limit): input rows xs [xrows][XS] + hidden rows h1s [hrows][XS], every
    // width padded to 16 (one work item = 16 neurons; K = 4 x a multiple of 4 fragments).  Full layers, else the chunked
    // layer-2/3 path in which the second hidden layer never exists in full.
    auto pad16 = [](int x) { return (x + 15) / 16 * 16; };
    const int limit_rows = 227 * 1024 / 256;
    const char *force = getenv("EGP_ROLLOUT_VARIANT");
    bool use_t4 = false;
    A.chunk23 = 0;
    if (d.t4_ok && !(force && force[0] == '1') && 2 * (pad16(A.A) / MLP_NT) <= T4_WARPS) {
        const int H1p = pad16(A.H1), H2p = pad16(A.H2), Ap = pad16(A.A), Dp = pad16(A.D);
        for (int pi = 0; pi < 4 && !use_t4; pi++) {        // (stride 36 | 32) x (full | chunked layer 2/3)
            const int ch = pi & 1, xs_stride = pi < 2 ? XS_WIDE : 32;
            int xr = Dp > (ch ? MLP_C2 : H2p) ? Dp : (ch ? MLP_C2 : H2p);
            int hr = H1p > Ap ? H1p : Ap;
            if (snH) {                  // LSTM input rows [S + H] and one 64-unit gate chunk (256 rows)
                if (xr < pad16(sn_in)) xr = pad16(sn_in);
                if (hr < 256) hr = 256;
            }
            const int need_rows = ((xr + hr) * xs_stride + 31) / 32;
            if (O.ax + need_rows <= limit_rows) {
                use_t4 = true; A.chunk23 = ch; A.xrows = xr; A.hrows = hr; A.xs = xs_stride;
                if (O.total - O.ax < need_rows) O.total = O.ax + need_rows;
            }
        }
    }
    // joint limits / floor contact: the block sweeps carry the rows (kernel instantiation CONS) for the plain policy plan; the
    // state LSTM, the value rule and the chunked wide-policy plan fall back to the one-warp kernel (EGP_CONS_VARIANT=1 forces it)
    bool cons = false;
    if (d.limits || d.contacts) {
        const char *cforce = getenv("EGP_CONS_VARIANT");
        if (use_t4 && !snH && !(in && in->value_net) && !A.chunk23 && !(cforce && cforce[0] == '1')) cons = true;
        else { use_t4 = false; A.chunk23 = 0; }
    }
    auto pad = [use_t4](int x) { return use_t4 ? (x + 15) / 16 * 16 : (x + JB - 1) / JB * JB; };
    A.H1p = pad(A.H1); A.H2p = pad(A.H2); A.Ap = pad(A.A);
    A.K1p = use_t4 ? pad(A.D) : A.D; A.K2p = use_t4 ? pad(A.H1) : A.H1; A.K3p = use_t4 ? pad(A.H2) : A.H2;
    size_t smem4 = sizeof(double) * 32 * (size_t)O.total;
    if (snH && !use_t4) { set_error("egp_rollout_f64: the state LSTM needs the T4 rollout variant (policy too wide?)"); return EGP_ESIZE; }
    const bool ev = cfg->eval_mode || (in && (in->d_fix_len || in->d_init_qpos || in->d_init_qvel)) || out->d_qpos_traj || out->d_qvel_traj;
    if (in && ((in->d_init_qpos == nullptr) != (in->d_init_qvel == nullptr))) { set_error("egp_rollout_f64: d_init_qpos / d_init_qvel come in pairs"); return EGP_EINVAL; }
    if (ev && !use_t4) { set_error("egp_rollout_f64: evaluation roll-outs (eval_mode / fix_len / qpos_traj) need the T4 rollout variant"); return EGP_ESIZE; }
    if (cfg->eval_mode && !(in && in->d_state_pred)) { set_error("egp_rollout_f64: eval_mode needs d_state_pred"); return EGP_EINVAL; }
    if ((out->d_qpos_traj == nullptr) != (out->d_qvel_traj == nullptr)) { set_error("egp_rollout_f64: d_qpos_traj / d_qvel_traj come in pairs"); return EGP_EINVAL; }
    A.sn_Kp = snH ? pad(sn_in) : 0;
    const int snN = snH ? pad(4 * snH) : 0;             // gate rows, padded
    const EgpPolicyWeights *vn = in ? in->value_net : nullptr;
    if (cfg->eval_mode == 2 && !vn) { set_error("egp_rollout_f64: eval_mode 2 ('valuefs') needs value_net"); return EGP_EINVAL; }
    if (vn) {
        if (!use_t4 || snH || A.chunk23) { set_error("egp_rollout_f64: value_net needs the plain T4 rollout variant (no state LSTM, no chunked wide policy)"); return EGP_ESIZE; }
        if (vn->in_dim != pol->in_dim || vn->h1 != pol->h1 || vn->h2 != pol->h2 || vn->out_dim != 1 || !vn->d_W1 || !vn->d_b1 ||
            !vn->d_W2 || !vn->d_b2 || !vn->d_W3 || !vn->d_b3) {
            set_error("egp_rollout_f64: value_net must be a [%d, %d, %d, 1] MLP like the policy trunk", pol->in_dim, pol->h1, pol->h2);
            return EGP_EINVAL;
        }
        if (in->d_vctx && !A.ctx) { set_error("egp_rollout_f64: d_vctx without a policy context table"); return EGP_EINVAL; }
        A.vAp = pad(1);
        A.vctx = in->d_vctx;
    }
    const size_t log_elems = (size_t)blocks * EGP_LOG_SIZE;
    size_t need = (size_t)A.K1p * A.H1p + A.H1p + (size_t)A.K2p * A.H2p + A.H2p + (size_t)A.K3p * A.Ap + A.Ap +
                  (size_t)A.sn_Kp * snN + snN + log_elems;
    const size_t vneed = vn ? (size_t)A.K1p * A.H1p + A.H1p + (size_t)A.K2p * A.H2p + A.H2p + (size_t)A.K3p * A.vAp + A.vAp : 0;
    need += vneed;
    if (need > m->wbuf_elems) {
        cudaFree(m->d_wbuf);
        m->d_wbuf = nullptr; m->wbuf_elems = 0;
        EGP_CUDA(cudaMalloc(&m->d_wbuf, sizeof(double) * need));
        m->wbuf_elems = need;
    }
    double *w = m->d_wbuf;
    double *W1t = w; w += (size_t)A.K1p * A.H1p;
    double *b1 = w; w += A.H1p;
    double *W2t = w; w += (size_t)A.K2p * A.H2p;
    double *b2 = w; w += A.H2p;
    double *W3t = w; w += (size_t)A.K3p * A.Ap;
    double *b3 = w; w += A.Ap;
    double *snW = w; w += (size_t)A.sn_Kp * snN;
    double *snb = w; w += snN;
    double *logp = w; w += log_elems;
    auto pack = [&](const double *W, const double *bsrc, int out_n, int in_n, int outp, int Kp, double *Wf, double *bp) {
        const long long total = (long long)outp * Kp;       // = (outp / 8) * (Kp / 4) * 32
        pack_frag_kernel_t<false><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(W, bsrc, out_n, in_n, outp, Kp / 4, Wf, bp);
    };
    if (vn) {
        double *v1 = w; w += (size_t)A.K1p * A.H1p;
        double *vb1 = w; w += A.H1p;
        double *v2 = w; w += (size_t)A.K2p * A.H2p;
        double *vb2 = w; w += A.H2p;
        double *v3 = w; w += (size_t)A.K3p * A.vAp;
        double *vb3 = w;
        pack(vn->d_W1, vn->d_b1, A.H1, A.D, A.H1p, A.K1p, v1, vb1);
        pack(vn->d_W2, vn->d_b2, A.H2, A.H1, A.H2p, A.K2p, v2, vb2);
        pack(vn->d_W3, vn->d_b3, 1, A.H2, A.vAp, A.K3p, v3, vb3);
        A.vW1t = v1; A.vb1 = vb1; A.vW2t = v2; A.vb2 = vb2; A.vW3t = v3; A.vb3 = vb3;
    }
    if (snH) {
        pack(in->d_snet_W, in->d_snet_b, 4 * snH, sn_in, snN, A.sn_Kp, snW, snb);
        A.sn_Wp = snW; A.sn_b = snb; A.sn_state = in->d_snet_state;
    }
    if (use_t4) {
        pack(pol->d_W1, pol->d_b1, A.H1, A.D, A.H1p, A.K1p, W1t, b1);
        pack(pol->d_W2, pol->d_b2, A.H2, A.H1, A.H2p, A.K2p, W2t, b2);
        pack(pol->d_W3, pol->d_b3, A.A, A.H2, A.Ap, A.K3p, W3t, b3);
    } else {
        transpose_pad_kernel<<<(A.D * A.H1p + 255) / 256, 256, 0, st>>>(pol->d_W1, pol->d_b1, A.H1, A.D, A.H1p, W1t, b1);
        transpose_pad_kernel<<<(A.H1 * A.H2p + 255) / 256, 256, 0, st>>>(pol->d_W2, pol->d_b2, A.H2, A.H1, A.H2p, W2t, b2);
        transpose_pad_kernel<<<(A.H2 * A.Ap + 255) / 256, 256, 0, st>>>(pol->d_W3, pol->d_b3, A.A, A.H2, A.Ap, W3t, b3);
    }
    EGP_CHECK_LAUNCH("weight packing");
    A.W1t = W1t; A.b1 = b1; A.W2t = W2t; A.b2 = b2; A.W3t = W3t; A.b3 = b3; A.log_std = pol->d_log_std;
    A.log_part = out->d_logger ? logp : nullptr;
    int rc2 = EGP_OK;
    if (use_t4) {
        auto launch = [&](auto kern) -> int {
            EGP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4));
            kern<<<blocks, T4_THREADS, smem4, st>>>(A, O);
            EGP_CHECK_LAUNCH("rollout_kernel_t4");
            return EGP_OK;
        };
        if (cons) {
            // joint limits / floor contact: the block sweeps with constraint rows (plain policy plan only)
            const size_t need = (size_t)blocks * CS_PER_BODY * d.nbody * 32;
            if (m->cons_elems < need) {
                cudaFree(m->d_cons);
                m->d_cons = nullptr; m->cons_elems = 0;
                EGP_CUDA(cudaMalloc(&m->d_cons, need * sizeof(double)));
                m->cons_elems = need;
            }
            A.cons = m->d_cons;
            rc2 = launch(rollout_kernel_t4<false, false, false, true>);
        } else if (snH) {
            if (A.chunk23) { set_error("egp_rollout_f64: state LSTM with the chunked wide-policy plan is not supported"); return EGP_ESIZE; }
            rc2 = launch(rollout_kernel_t4<false, true>);
        } else if (vn) rc2 = launch(rollout_kernel_t4<false, false, true>);
        else if (!A.chunk23) rc2 = launch(rollout_kernel_t4<false, false>);
        else rc2 = launch(rollout_kernel_t4<true, false>);
    } else {
        const int xrows = A.D > A.H2p ? A.D : A.H2p, hrows = A.H1p > A.Ap ? A.H1p : A.Ap;
        // environments per CTA: 32.  Narrower CTAs (more warps per SM at 4096 environments) were measured SLOWER on B200
        // (full-physics rollout 4.28 s at 32, 5.18 s at 16, 6.06 s at 8, 6.73 s at 4): the kernel is bound by its thread-local
        // tree data moving through L1 / L2, where a partial warp wastes most of every line; EGP_V1_ENVS_PER_CTA keeps the A/B
        int ep = ENVS_PER_CTA;
        const char *epf = getenv("EGP_V1_ENVS_PER_CTA");
        if (epf && atoi(epf) >= 1 && atoi(epf) <= 32) ep = atoi(epf);
        blocks = (cfg->n_env + ep - 1) / ep;
        size_t smem = sizeof(double) * ep * ((size_t)xrows + hrows);
        if (smem > 227 * 1024) { set_error("egp_rollout_f64: policy too wide for shared memory (%zu bytes)", smem); return EGP_ESIZE; }
        if (out->d_logger) {            // the one-warp kernel accumulates with atomics
            double init[EGP_LOG_SIZE] = {0};
            init[EGP_LOG_MIN_C_REWARD] = INFINITY; init[EGP_LOG_MAX_C_REWARD] = -INFINITY;
            init[EGP_LOG_MIN_EPISODE_REWARD] = INFINITY; init[EGP_LOG_MAX_EPISODE_REWARD] = -INFINITY;
            EGP_CUDA(cudaMemcpyAsync(out->d_logger, init, sizeof init, cudaMemcpyHostToDevice, st));
        }
        EGP_CUDA(cudaFuncSetAttribute(rollout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rollout_kernel<<<blocks, ep, smem, st>>>(A);
        EGP_CHECK_LAUNCH("rollout_kernel");
    }
    if (rc2 != EGP_OK) return rc2;
    if (out->d_logger && use_t4) {
        logger_merge_kernel<<<1, 32, 0, st>>>(logp, blocks, out->d_logger);
        EGP_CHECK_LAUNCH("logger_merge_kernel");
    }
    return EGP_OK;
}

#ifdef EGP_T4_CLK
int egp_debug_t4_clk(unsigned long long *out8) {
    EGP_CUDA(cudaDeviceSynchronize());
    EGP_CUDA(cudaMemcpyFromSymbol(out8, g_t4_clk, sizeof(unsigned long long) * 8));
    EGP_CUDA(cudaMemcpyFromSymbol(out8 + 8, g_t4_clk2, sizeof(unsigned long long) * 16));
    unsigned long long z2[16] = {0};
    EGP_CUDA(cudaMemcpyToSymbol(g_t4_clk2, z2, sizeof z2));
    unsigned long long z[8] = {0};
    EGP_CUDA(cudaMemcpyToSymbol(g_t4_clk, z, sizeof z));
    return EGP_OK;
}
#endif

}  // extern "C"
