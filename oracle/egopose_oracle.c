/* TEST INFRASTRUCTURE ONLY.  CPU float64 restatement of EgoPose's rollout path; see egopose_oracle.h.
 *
 * Each function cites the reference file:line it follows (paths relative to /root/reference).
 * The smooth-dynamics functions restate MuJoCo's documented mj_forward/mj_step pipeline, which the
 * reference reaches through mujoco_py (envs/common/mujoco_env.py:22-23,100-101;
 * ego_pose/envs/humanoid_v1.py:134,174): PARITY UNPINNED for those (no MuJoCo binary available).
 */
#include "egopose_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* ---------------------------------------------------------------- small vector helpers */
static void cross3(const double *a, const double *b, double *o) {
    double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
static double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double dot6(const double *a, const double *b) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}

/* quaternion (w,x,y,z) product, utils/transformation.py:1379-1393 (quaternion_multiply(q1, q0)) */
void eo_quat_mul(const double *q1, const double *q0, double *out) {
    double w0 = q0[0], x0 = q0[1], y0 = q0[2], z0 = q0[3];
    double w1 = q1[0], x1 = q1[1], y1 = q1[2], z1 = q1[3];
    double r0 = -x1 * x0 - y1 * y0 - z1 * z0 + w1 * w0;
    double r1 = x1 * w0 + y1 * z0 - z1 * y0 + w1 * x0;
    double r2 = -x1 * z0 + y1 * w0 + z1 * x0 + w1 * y0;
    double r3 = x1 * y0 - y1 * x0 + z1 * w0 + w1 * z0;
    out[0] = r0; out[1] = r1; out[2] = r2; out[3] = r3;
}

/* utils/transformation.py:1410-1421 quaternion_inverse: conjugate / dot(q,q) */
void eo_quat_inv(const double *q, double *out) {
    double n = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    out[0] = q[0] / n; out[1] = -q[1] / n; out[2] = -q[2] / n; out[3] = -q[3] / n;
}

/* utils/transformation.py:1194-1248 quaternion_from_euler(ai, aj, ak, 'sxyz'):
 * firstaxis 0, parity 0, repetition 0, frame 0 -> i,j,k = 1,2,3 */
void eo_quat_from_euler(double ai, double aj, double ak, double *q) {
    ai /= 2.0; aj /= 2.0; ak /= 2.0;
    double ci = cos(ai), si = sin(ai), cj = cos(aj), sj = sin(aj), ck = cos(ak), sk = sin(ak);
    double cc = ci * ck, cs = ci * sk, sc = si * ck, ss = si * sk;
    q[0] = cj * cc + sj * ss;
    q[1] = cj * sc - sj * cs;
    q[2] = cj * ss + sj * cc;
    q[3] = cj * cs - sj * sc;
}

/* rotation matrix of a quaternion, utils/transformation.py:1267-1291 quaternion_matrix (normalising) */
static void quat_matrix(const double *quat, double R[9]) {
    double n = quat[0] * quat[0] + quat[1] * quat[1] + quat[2] * quat[2] + quat[3] * quat[3];
    if (n < 2.220446049250313e-16 * 4.0) { /* _EPS = finfo(float).eps * 4 */
        R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
        return;
    }
    double s = sqrt(2.0 / n);
    double q0 = quat[0] * s, q1 = quat[1] * s, q2 = quat[2] * s, q3 = quat[3] * s;
    R[0] = 1.0 - q2 * q2 - q3 * q3; R[1] = q1 * q2 - q3 * q0;       R[2] = q1 * q3 + q2 * q0;
    R[3] = q1 * q2 + q3 * q0;       R[4] = 1.0 - q1 * q1 - q3 * q3; R[5] = q2 * q3 - q1 * q0;
    R[6] = q1 * q3 - q2 * q0;       R[7] = q2 * q3 + q1 * q0;       R[8] = 1.0 - q1 * q1 - q2 * q2;
}
static void mat_vec(const double R[9], const double *v, double *o) {
    double x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
    double y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
    double z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
static void matT_vec(const double R[9], const double *v, double *o) {
    double x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
    double y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
    double z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}

/* utils/math.py:62-67 get_heading_q */
static void heading_q(const double *q, double *hq) {
    double n = sqrt(q[0] * q[0] + q[3] * q[3]);
    hq[0] = q[0] / n; hq[1] = 0.0; hq[2] = 0.0; hq[3] = q[3] / n;
}

/* utils/math.py:80-81 de_heading */
void eo_de_heading(const double *q, double *out) {
    double hq[4], ihq[4];
    heading_q(q, hq);
    eo_quat_inv(hq, ihq);
    eo_quat_mul(ihq, q, out);
}

/* utils/math.py:47-59 transform_vec: rot(q or heading(q))^T v */
void eo_transform_vec(const double *v, const double *q, int heading, double *out) {
    double R[9], hq[4];
    if (heading) { heading_q(q, hq); quat_matrix(hq, R); } else quat_matrix(q, R);
    matT_vec(R, v, out);
}

/* utils/transformation.py:348-356 rotation_from_quaternion */
void eo_rotation_from_quat(const double *q, double *axis, double *angle) {
    if (1.0 - q[0] < 1e-8) { axis[0] = 1.0; axis[1] = 0.0; axis[2] = 0.0; *angle = 0.0; return; }
    double s = sqrt(1.0 - q[0] * q[0]);
    axis[0] = q[1] / s; axis[1] = q[2] / s; axis[2] = q[3] / s;
    *angle = 2.0 * acos(q[0]);
}

/* utils/math.py:20-35 get_qvel_fd, first six entries (root linear + angular velocity) */
void eo_qvel_fd6(const double *cur_qpos, const double *next_qpos, double dt, int heading, double *out6) {
    double v[3], qi[4], qrel[4], axis[3], angle, rv[3];
    for (int k = 0; k < 3; k++) v[k] = (next_qpos[k] - cur_qpos[k]) / dt;
    eo_quat_inv(cur_qpos + 3, qi);
    eo_quat_mul(next_qpos + 3, qi, qrel);
    eo_rotation_from_quat(qrel, axis, &angle);
    if (angle > M_PI) angle -= 2 * M_PI; else if (angle < -M_PI) angle += 2 * M_PI;
    for (int k = 0; k < 3; k++) rv[k] = (axis[k] * angle) / dt;
    eo_transform_vec(rv, cur_qpos + 3, 0, out6 + 3);
    if (heading) eo_transform_vec(v, cur_qpos + 3, 1, out6);
    else { out6[0] = v[0]; out6[1] = v[1]; out6[2] = v[2]; }
}

/* full get_qvel_fd (utils/math.py:20-35); transform: 0 none, 1 'heading' */
void eo_qvel_fd(int nq, const double *cur_qpos, const double *next_qpos, double dt, int transform, double *out) {
    eo_qvel_fd6(cur_qpos, next_qpos, dt, transform, out);
    for (int k = 7; k < nq; k++) out[k - 1] = (next_qpos[k] - cur_qpos[k]) / dt;
}

/* utils/math.py:38-44 get_angvel_fd over nquat stacked quaternions */
void eo_angvel_fd(int nquat, const double *prev, const double *cur, double dt, double *out) {
    for (int i = 0; i < nquat; i++) {
        double qi[4], qd[4], axis[3], angle;
        eo_quat_inv(prev + 4 * i, qi);
        eo_quat_mul(cur + 4 * i, qi, qd);
        eo_rotation_from_quat(qd, axis, &angle);
        for (int k = 0; k < 3; k++) out[3 * i + k] = axis[k] * angle / dt;
    }
}

/* ---------------------------------------------------------------- MuJoCo restatement */

/* rotate v by unit quaternion q (MuJoCo mju_rotVecQuat) */
static void rot_vec_quat(const double *q, const double *v, double *o) {
    double R[9];
    /* unit-quaternion rotation matrix without renormalisation */
    double w = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = w * w + x * x - y * y - z * z; R[1] = 2 * (x * y - w * z);           R[2] = 2 * (x * z + w * y);
    R[3] = 2 * (x * y + w * z);           R[4] = w * w - x * x + y * y - z * z; R[5] = 2 * (y * z - w * x);
    R[6] = 2 * (x * z - w * y);           R[7] = 2 * (y * z + w * x);           R[8] = w * w - x * x - y * y + z * z;
    mat_vec(R, v, o);
}
static void quat_to_mat(const double *q, double R[9]) {
    double w = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = w * w + x * x - y * y - z * z; R[1] = 2 * (x * y - w * z);           R[2] = 2 * (x * z + w * y);
    R[3] = 2 * (x * y + w * z);           R[4] = w * w - x * x + y * y - z * z; R[5] = 2 * (y * z - w * x);
    R[6] = 2 * (x * z - w * y);           R[7] = 2 * (y * z + w * x);           R[8] = w * w - x * x - y * y + z * z;
}

/* mj_kinematics: body frames from qpos.  Outputs world axis/anchor per dof. */
void eo_kinematics(const EoModel *m, EoData *d, double *dof_axis_w, double *dof_anchor_w) {
    for (int b = 0; b < m->nbody; b++) {
        int p = m->body_parent[b];
        int da = m->body_dofadr[b], qa = m->body_qposadr[b];
        double pos[3], quat[4];
        if (m->body_dofnum[b] == 6) {           /* free joint: pose straight from qpos, quat normalised */
            double *q = d->qpos + qa + 3;
            double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
            for (int k = 0; k < 3; k++) pos[k] = d->qpos[qa + k];
            for (int k = 0; k < 4; k++) quat[k] = q[k] / n;
            for (int k = 0; k < 3; k++) {       /* translations: world axes */
                double *ax = dof_axis_w + 3 * (da + k);
                ax[0] = ax[1] = ax[2] = 0.0; ax[k] = 1.0;
                memcpy(dof_anchor_w + 3 * (da + k), pos, sizeof pos);
            }
            for (int k = 0; k < 3; k++) {       /* rotations: body axes (angular velocity is body-local) */
                double e[3] = {0, 0, 0};
                e[k] = 1.0;
                rot_vec_quat(quat, e, dof_axis_w + 3 * (da + 3 + k));
                memcpy(dof_anchor_w + 3 * (da + 3 + k), pos, sizeof pos);
            }
        } else {
            double off[3];
            if (p >= 0) {
                rot_vec_quat(d->xquat[p], m->body_pos + 3 * b, off);
                for (int k = 0; k < 3; k++) pos[k] = d->xpos[p][k] + off[k];
                memcpy(quat, d->xquat[p], sizeof quat);
            } else {
                memcpy(pos, m->body_pos + 3 * b, sizeof pos);
                quat[0] = 1; quat[1] = quat[2] = quat[3] = 0;
            }
            for (int j = 0; j < m->body_dofnum[b]; j++) {   /* hinges, XML order, about the fixed anchor */
                int i = da + j;
                double anchor[3], axis[3], qloc[4], qn[4], back[3];
                rot_vec_quat(quat, m->dof_anchor + 3 * i, anchor);
                for (int k = 0; k < 3; k++) anchor[k] += pos[k];
                rot_vec_quat(quat, m->dof_axis + 3 * i, axis);
                memcpy(dof_axis_w + 3 * i, axis, sizeof axis);
                memcpy(dof_anchor_w + 3 * i, anchor, sizeof anchor);
                double half = 0.5 * d->qpos[qa + j], s = sin(half);
                qloc[0] = cos(half);
                for (int k = 0; k < 3; k++) qloc[k + 1] = m->dof_axis[3 * i + k] * s;
                eo_quat_mul(quat, qloc, qn);
                memcpy(quat, qn, sizeof quat);
                rot_vec_quat(quat, m->dof_anchor + 3 * i, back);
                for (int k = 0; k < 3; k++) pos[k] = anchor[k] - back[k];
            }
        }
        memcpy(d->xpos[b], pos, sizeof pos);
        memcpy(d->xquat[b], quat, sizeof quat);
        double ip[3];
        rot_vec_quat(quat, m->body_ipos + 3 * b, ip);
        for (int k = 0; k < 3; k++) d->xipos[b][k] = pos[k] + ip[k];
    }
}

/* spatial inertia about the world origin, world axes: mass, h = m*c, I_O (xx yy zz xy xz yz) */
typedef struct { double m, h[3], I[6]; } SpI;

static void spi_mul(const SpI *s, const double *v /* [w; lin] */, double *f /* [torque; force] */) {
    const double *w = v, *l = v + 3;
    double hxl[3], hxw[3];
    cross3(s->h, l, hxl);
    cross3(s->h, w, hxw);
    f[0] = s->I[0] * w[0] + s->I[3] * w[1] + s->I[4] * w[2] + hxl[0];
    f[1] = s->I[3] * w[0] + s->I[1] * w[1] + s->I[5] * w[2] + hxl[1];
    f[2] = s->I[4] * w[0] + s->I[5] * w[1] + s->I[2] * w[2] + hxl[2];
    f[3] = s->m * l[0] - hxw[0];
    f[4] = s->m * l[1] - hxw[1];
    f[5] = s->m * l[2] - hxw[2];
}
static void cross_motion(const double *v, const double *s, double *o) {     /* v x s (motion vectors) */
    double a[3], b[3], c[3];
    cross3(v, s, a);
    cross3(v, s + 3, b);
    cross3(v + 3, s, c);
    o[0] = a[0]; o[1] = a[1]; o[2] = a[2];
    o[3] = b[0] + c[0]; o[4] = b[1] + c[1]; o[5] = b[2] + c[2];
}
static void cross_force(const double *v, const double *f, double *o) {      /* v x* f */
    double a[3], b[3], c[3];
    cross3(v, f, a);
    cross3(v + 3, f + 3, b);
    cross3(v, f + 3, c);
    o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2];
    o[3] = c[0]; o[4] = c[1]; o[5] = c[2];
}

/* in-place Cholesky solve of the dense SPD system A x = b (scipy cho_factor/cho_solve,
 * ego_pose/envs/humanoid_v1.py:143); returns -1 if not positive definite */
int eo_chol_solve(int n, double *A, double *b) {
    for (int j = 0; j < n; j++) {
        double s = A[j * n + j];
        for (int k = 0; k < j; k++) s -= A[j * n + k] * A[j * n + k];
        if (!(s > 0.0)) return -1;
        double l = sqrt(s);
        A[j * n + j] = l;
        for (int i = j + 1; i < n; i++) {
            double t = A[i * n + j];
            for (int k = 0; k < j; k++) t -= A[i * n + k] * A[j * n + k];
            A[i * n + j] = t / l;
        }
    }
    for (int i = 0; i < n; i++) {
        double t = b[i];
        for (int k = 0; k < i; k++) t -= A[i * n + k] * b[k];
        b[i] = t / A[i * n + i];
    }
    for (int i = n - 1; i >= 0; i--) {
        double t = b[i];
        for (int k = i + 1; k < n; k++) t -= A[k * n + i] * b[k];
        b[i] = t / A[i * n + i];
    }
    return 0;
}

/* mj_forward, smooth part: kinematics -> subtree COM -> CRBA (qM + armature)
 * -> RNE bias (gravity + Coriolis/centrifugal) -> qacc = M^-1 (ctrl - bias).  No collision, joint
 * limits or constraint solver (north-star scope). */
/* one limit row: impedance d(dist) (solimp: d0, dwidth, width, midpoint, power), R = (1 - d)/d * invweight, D = 1/R,
 * aref = -b vel - k d dist with k = 1/(dmax^2 tc^2 dr^2), b = 2/(dmax tc), tc >= 2 dt (solref: timeconst, dampratio) */
void eo_limit_row(const EoModel *m, double dist, double vel, double invweight, double *Dout, double *aref) {
    const double MINVAL = 1e-15, MINIMP = 1e-4, MAXIMP = 0.9999;
    double d0 = m->solimp[0], dw = m->solimp[1], width = m->solimp[2], mid = m->solimp[3], power = m->solimp[4];
    d0 = d0 < MINIMP ? MINIMP : (d0 > MAXIMP ? MAXIMP : d0);
    dw = dw < MINIMP ? MINIMP : (dw > MAXIMP ? MAXIMP : dw);
    mid = mid < MINIMP ? MINIMP : (mid > MAXIMP ? MAXIMP : mid);
    if (power < 1.0) power = 1.0;
    double imp;
    if (d0 == dw || width <= MINVAL) imp = 0.5 * (d0 + dw);
    else {
        double x = fabs(dist) / width, y;
        if (x >= 1.0) y = 1.0;
        else if (x <= 0.0) y = 0.0;
        else if (power == 1.0) y = x;
        else if (x <= mid) y = pow(x, power) / pow(mid, power - 1.0);
        else y = 1.0 - pow(1.0 - x, power) / pow(1.0 - mid, power - 1.0);
        imp = d0 + y * (dw - d0);
    }
    double tc = m->solref[0], dr = m->solref[1];
    if (tc < 2.0 * m->timestep) tc = 2.0 * m->timestep;
    double k = 1.0 / (dw * dw * tc * tc * dr * dr), bb = 2.0 / (dw * tc);
    double R = (1.0 - imp) / imp * invweight;
    if (R < MINVAL) R = MINVAL;
    *Dout = 1.0 / R;
    *aref = -bb * vel - k * imp * dist;
}

/* ---- constraint rows: joint limits and floor contacts -------------------------------------------------------------
 * Every row i is unilateral with Jacobian J_i, weight D_i and reference acceleration aref_i; MuJoCo's solver minimises
 *     1/2 (a - a0)^T M (a - a0) + sum_i s_i(J_i a - aref_i),   s(r) = 1/2 D r^2 for r < 0, else 0
 * (a0 = smooth acceleration).  For a fixed active set the minimiser solves
 *     (M + sum_act D_i J_i^T J_i) a = qfrc_smooth + sum_act D_i aref_i J_i^T;
 * the active set {i: J_i a - aref_i < 0} is iterated to its fixed point (Newton on the piecewise-quadratic cost with unit
 * steps; the optimum is unique, MuJoCo's Newton solver converges to the same point up to its tolerance). */
#define EO_MAXCON 40
#define EO_CAREFUL 12
#define EO_MAXROW (EO_MAXV + 4 * EO_MAXCON)

typedef struct { int n; double J[EO_MAXROW][EO_MAXV], D[EO_MAXROW], aref[EO_MAXROW]; } EoRows;

/* Jacobian row of direction dir at world point pt of body b: cdof is about the world origin, [axis; anchor x axis] */
static void point_jac_row(const EoModel *m, const EoData *d, int b, const double *pt, const double *dir, double *J) {
    double p[6];
    cross3(pt, dir, p);
    p[3] = dir[0]; p[4] = dir[1]; p[5] = dir[2];
    for (int i = 0; i < m->nv; i++) J[i] = 0.0;
    for (int i = m->body_dofadr[b] + m->body_dofnum[b] - 1; i >= 0; i = m->dof_parent[i]) J[i] = dot6(d->cdof[i], p);
}

static void add_contact(const EoModel *m, const EoData *d, EoRows *R, int b, const double *pt, double dist) {
    /* frame: normal +z (plane), tangents y and -x (mju_makeFrame of (0,0,1)); pyramid edges n +- mu t (condim 3) */
    static const double nrm[3] = {0, 0, 1}, t1[3] = {0, 1, 0}, t2[3] = {-1, 0, 0};
    if (R->n + 4 > EO_MAXROW) return;
    const double mu = m->contact_mu, margin = m->contact_margin;
    double Jn[EO_MAXV], J1[EO_MAXV], J2[EO_MAXV];
    point_jac_row(m, d, b, pt, nrm, Jn);
    point_jac_row(m, d, b, pt, t1, J1);
    point_jac_row(m, d, b, pt, t2, J2);
    const double tran = m->body_invweight0[2 * b];          /* world body: 0 */
    for (int e = 0; e < 4; e++) {
        double *J = R->J[R->n];
        const double *Jt = e < 2 ? J1 : J2;
        const double sg = (e & 1) ? -mu : mu;
        double vel = 0.0;
        for (int i = 0; i < m->nv; i++) { J[i] = Jn[i] + sg * Jt[i]; vel += J[i] * d->qvel[i]; }
        double D0, aref;
        eo_limit_row(m, dist - margin, vel, tran + mu * mu * tran, &D0, &aref);
        R->D[R->n] = D0 / (2.0 * mu * mu);                  /* pyramidal: R = 2 mu^2 R_0 */
        R->aref[R->n] = aref;
        R->n++;
    }
}

static void collect_rows(const EoModel *m, const EoData *d, EoRows *R) {
    R->n = 0;
    if (m->dof_range) {
        for (int i = 0; i < m->nv; i++) {
            if (m->body_dofnum[m->dof_body[i]] == 6) continue;
            double lo = m->dof_range[2 * i], hi = m->dof_range[2 * i + 1];
            if (!(lo < hi)) continue;
            double q = d->qpos[i + 1], dist, s;             /* hinge dof i reads qpos[i + 1] (free root: 7 qpos, 6 dofs) */
            if (q - lo < 0.0) { dist = q - lo; s = 1.0; }
            else if (hi - q < 0.0) { dist = hi - q; s = -1.0; }
            else continue;
            for (int k = 0; k < m->nv; k++) R->J[R->n][k] = 0.0;
            R->J[R->n][i] = s;
            eo_limit_row(m, dist, s * d->qvel[i], m->dof_invweight0[i], &R->D[R->n], &R->aref[R->n]);
            R->n++;
        }
    }
    if (m->geom_type) {
        const double margin = m->contact_margin;
        for (int b = 0; b < m->nbody; b++) {
            double Rm[9], c0[3], c1[3];
            quat_to_mat(d->xquat[b], Rm);
            mat_vec(Rm, m->geom_p0 + 3 * b, c0);
            for (int k = 0; k < 3; k++) c0[k] += d->xpos[b][k];
            const double *sz = m->geom_size + 3 * b;
            if (m->geom_type[b] == 0 || m->geom_type[b] == 1) {
                /* sphere; capsule = the spheres at its two ends (second end point first: mjc_PlaneCapsule tests +axis) */
                int ne = m->geom_type[b] == 1 ? 2 : 1;
                if (ne == 2) {
                    mat_vec(Rm, m->geom_p1 + 3 * b, c1);
                    for (int k = 0; k < 3; k++) c1[k] += d->xpos[b][k];
                }
                for (int e = 0; e < ne; e++) {
                    const double *c = (ne == 2 && e == 0) ? c1 : c0;
                    double dist = c[2] - sz[0];
                    if (dist >= margin) continue;
                    double pt[3] = {c[0], c[1], c[2] - (sz[0] + 0.5 * dist)};
                    add_contact(m, d, R, b, pt, dist);
                }
            } else {
                int cnt = 0;
                for (int v = 0; v < 8 && cnt < 4; v++) {    /* box corners, at most 4 contacts (mjc_PlaneBox) */
                    double loc[3] = {(v & 1 ? sz[0] : -sz[0]), (v & 2 ? sz[1] : -sz[1]), (v & 4 ? sz[2] : -sz[2])}, w[3];
                    mat_vec(Rm, loc, w);
                    for (int k = 0; k < 3; k++) w[k] += c0[k];
                    double dist = w[2];
                    if (dist > margin) continue;
                    double pt[3] = {w[0], w[1], w[2] - 0.5 * dist};
                    add_contact(m, d, R, b, pt, dist);
                    cnt++;
                }
            }
        }
    }
}

/* diagnostics of the active-set iteration over all solves since the last reset: [0] solves with rows, [1] largest number of
 * passes, [2] solves that hit the cap without reaching a fixed point (not thread-safe counters: test use only) */
static long g_solver_stats[3];
void eo_solver_stats(long *out3, int reset) {
    for (int k = 0; k < 3; k++) { out3[k] = g_solver_stats[k]; if (reset) g_solver_stats[k] = 0; }
}

int eo_constraint_solve(const EoModel *m, EoData *d, const double *smooth) {
    const int nv = m->nv;
    static _Thread_local EoRows R;
    double A[EO_MAXV * EO_MAXV], rhs[EO_MAXV];
    int act[EO_MAXROW];
    collect_rows(m, d, &R);
    for (int i = 0; i < R.n; i++) act[i] = 1;
    int it;
    for (it = 0; it < 100; it++) {
        memcpy(A, d->qM, sizeof(double) * nv * nv);
        memcpy(rhs, smooth, sizeof(double) * nv);
        for (int i = 0; i < R.n; i++) {
            if (!act[i]) continue;
            const double *J = R.J[i];
            for (int r = 0; r < nv; r++) {
                if (J[r] == 0.0) continue;
                rhs[r] += R.D[i] * R.aref[i] * J[r];
                for (int c = 0; c < nv; c++) A[r * nv + c] += R.D[i] * J[r] * J[c];
            }
        }
        eo_chol_solve(nv, A, rhs);
        /* all rows that disagree flip at once; if that has not settled after EO_CAREFUL passes (coupled rows can flip back and
         * forth: seen in about 1 of 20 000 solves), only the first disagreeing row flips per pass */
        int changed = 0;
        for (int i = 0; i < R.n; i++) {
            double r = -R.aref[i];
            for (int k = 0; k < nv; k++) r += R.J[i][k] * rhs[k];
            int a1 = r < 0.0;
            if (a1 != act[i]) {
                if (it < EO_CAREFUL || !changed) act[i] = a1;
                changed = 1;
            }
        }
        if (!changed) break;
    }
    memcpy(d->qacc, rhs, sizeof(double) * nv);
    d->n_efc = R.n;
    d->solver_iter = it;
    if (R.n) { g_solver_stats[0]++; if (it + 1 > g_solver_stats[1]) g_solver_stats[1] = it + 1; if (it >= 100) g_solver_stats[2]++; }
    return R.n;
}

void eo_forward(const EoModel *m, EoData *d) {
    int nv = m->nv, nb = m->nbody;
    double axis_w[EO_MAXV * 3], anchor_w[EO_MAXV * 3];
    eo_kinematics(m, d, axis_w, anchor_w);

    /* cdof about the world origin: [axis; anchor x axis]; free translations [0; e_k] */
    for (int i = 0; i < nv; i++) {
        int b = m->dof_body[i];
        int j = i - m->body_dofadr[b];
        double *c = d->cdof[i];
        if (m->body_dofnum[b] == 6 && j < 3) {
            c[0] = c[1] = c[2] = 0.0;
            c[3] = axis_w[3 * i]; c[4] = axis_w[3 * i + 1]; c[5] = axis_w[3 * i + 2];
        } else {
            c[0] = axis_w[3 * i]; c[1] = axis_w[3 * i + 1]; c[2] = axis_w[3 * i + 2];
            cross3(anchor_w + 3 * i, axis_w + 3 * i, c + 3);
        }
    }

    /* body spatial inertias in world axes about the world origin */
    SpI bi[EO_MAXB], crb[EO_MAXB];
    double mtot = 0, com[3] = {0, 0, 0};
    for (int b = 0; b < nb; b++) {
        double R[9], Ib[9], T[9], Iw[9];
        const double *in = m->body_inertia + 6 * b;
        quat_to_mat(d->xquat[b], R);
        Ib[0] = in[0]; Ib[4] = in[1]; Ib[8] = in[2];
        Ib[1] = Ib[3] = in[3]; Ib[2] = Ib[6] = in[4]; Ib[5] = Ib[7] = in[5];
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) {
            T[3 * r + c] = R[3 * r] * Ib[c] + R[3 * r + 1] * Ib[3 + c] + R[3 * r + 2] * Ib[6 + c];
        }
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) {
            Iw[3 * r + c] = T[3 * r] * R[3 * c] + T[3 * r + 1] * R[3 * c + 1] + T[3 * r + 2] * R[3 * c + 2];
        }
        double mass = m->body_mass[b];
        const double *c = d->xipos[b];
        double cc = dot3(c, c);
        SpI *s = &bi[b];
        s->m = mass;
        for (int k = 0; k < 3; k++) s->h[k] = mass * c[k];
        s->I[0] = Iw[0] + mass * (cc - c[0] * c[0]);
        s->I[1] = Iw[4] + mass * (cc - c[1] * c[1]);
        s->I[2] = Iw[8] + mass * (cc - c[2] * c[2]);
        s->I[3] = Iw[1] - mass * c[0] * c[1];
        s->I[4] = Iw[2] - mass * c[0] * c[2];
        s->I[5] = Iw[5] - mass * c[1] * c[2];
        crb[b] = *s;
        mtot += mass;
        for (int k = 0; k < 3; k++) com[k] += mass * c[k];
    }
    for (int k = 0; k < 3; k++) d->subtree_com[k] = com[k] / mtot;

    /* CRBA: composite inertias then M[i][j] = cdof_j . (crb_body(i) cdof_i), j ancestor-or-self of i */
    for (int b = nb - 1; b > 0; b--) {
        int p = m->body_parent[b];
        if (p < 0) continue;
        crb[p].m += crb[b].m;
        for (int k = 0; k < 3; k++) crb[p].h[k] += crb[b].h[k];
        for (int k = 0; k < 6; k++) crb[p].I[k] += crb[b].I[k];
    }
    memset(d->qM, 0, sizeof(double) * nv * nv);
    for (int i = 0; i < nv; i++) {
        double f[6];
        spi_mul(&crb[m->dof_body[i]], d->cdof[i], f);
        for (int j = i; j >= 0; j = m->dof_parent[j]) {
            double v = dot6(d->cdof[j], f);
            d->qM[i * nv + j] = v;
            d->qM[j * nv + i] = v;
        }
        d->qM[i * nv + i] += m->dof_armature[i];
    }

    /* velocities / bias accelerations (mj_comVel) then RNE with zero qacc (mj_rne flg_acc=0) */
    double cvel[EO_MAXB][6], cacc[EO_MAXB][6], cfrc[EO_MAXB][6];
    for (int b = 0; b < nb; b++) {
        int p = m->body_parent[b];
        double v[6], a[6];
        if (p >= 0) { memcpy(v, cvel[p], sizeof v); memcpy(a, cacc[p], sizeof a); }
        else {
            memset(v, 0, sizeof v);
            a[0] = a[1] = a[2] = 0.0;
            a[3] = -m->gravity[0]; a[4] = -m->gravity[1]; a[5] = -m->gravity[2];
        }
        int da = m->body_dofadr[b], n = m->body_dofnum[b], j = 0;
        if (n == 6) {                       /* free joint: translations have zero cdof_dot */
            for (; j < 3; j++) for (int k = 0; k < 6; k++) v[k] += d->cdof[da + j][k] * d->qvel[da + j];
            double vb[6], dd[6];
            memcpy(vb, v, sizeof vb);       /* ball part: all three cdof_dot from the pre-rotation velocity */
            for (; j < 6; j++) {
                cross_motion(vb, d->cdof[da + j], dd);
                for (int k = 0; k < 6; k++) {
                    a[k] += dd[k] * d->qvel[da + j];
                    v[k] += d->cdof[da + j][k] * d->qvel[da + j];
                }
            }
        } else {
            for (; j < n; j++) {
                double dd[6];
                cross_motion(v, d->cdof[da + j], dd);
                for (int k = 0; k < 6; k++) {
                    a[k] += dd[k] * d->qvel[da + j];
                    v[k] += d->cdof[da + j][k] * d->qvel[da + j];
                }
            }
        }
        memcpy(cvel[b], v, sizeof v);
        memcpy(cacc[b], a, sizeof a);
        double Ia[6], Iv[6], vxIv[6];
        spi_mul(&bi[b], a, Ia);
        spi_mul(&bi[b], v, Iv);
        cross_force(v, Iv, vxIv);
        for (int k = 0; k < 6; k++) cfrc[b][k] = Ia[k] + vxIv[k];
    }
    for (int b = nb - 1; b > 0; b--) {
        int p = m->body_parent[b];
        if (p >= 0) for (int k = 0; k < 6; k++) cfrc[p][k] += cfrc[b][k];
    }
    for (int i = 0; i < nv; i++) d->qfrc_bias[i] = dot6(d->cdof[i], cfrc[m->dof_body[i]]);

    /* qacc = M^-1 (qfrc_actuator - qfrc_bias); unit-gear motors on dofs 6.. (xml:139-192) */
    double A[EO_MAXV * EO_MAXV], rhs[EO_MAXV], smooth[EO_MAXV];
    int first = nv - m->nu;
    for (int i = 0; i < nv; i++) smooth[i] = (i >= first ? d->ctrl[i - first] : 0.0) - d->qfrc_bias[i];
    if (!m->dof_range && !m->geom_type) {
        memcpy(A, d->qM, sizeof(double) * nv * nv);
        memcpy(rhs, smooth, sizeof(double) * nv);
        eo_chol_solve(nv, A, rhs);
        memcpy(d->qacc, rhs, sizeof(double) * nv);
        return;
    }
    eo_constraint_solve(m, d, smooth);
}

/* mj_step = mj_forward + semi-implicit Euler; position-dependent outputs
 * (xpos, qM, qfrc_bias) are NOT refreshed after integration. */
void eo_step(const EoModel *m, EoData *d) {
    double h = m->timestep;
    eo_forward(m, d);
    for (int i = 0; i < m->nv; i++) d->qvel[i] += h * d->qacc[i];
    for (int b = 0; b < m->nbody; b++) {
        int qa = m->body_qposadr[b], da = m->body_dofadr[b];
        if (m->body_dofnum[b] == 6) {
            for (int k = 0; k < 3; k++) d->qpos[qa + k] += h * d->qvel[da + k];
            double *q = d->qpos + qa + 3, *w = d->qvel + da + 3;
            double n = sqrt(dot3(w, w)), ax[3] = {1, 0, 0};
            if (n > 1e-15) { ax[0] = w[0] / n; ax[1] = w[1] / n; ax[2] = w[2] / n; }
            double ang = h * n, s = sin(0.5 * ang), qr[4] = {cos(0.5 * ang), ax[0] * s, ax[1] * s, ax[2] * s};
            double qn = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]), qq[4], out[4];
            for (int k = 0; k < 4; k++) qq[k] = q[k] / qn;
            eo_quat_mul(qq, qr, out);
            memcpy(q, out, sizeof out);
        } else {
            for (int j = 0; j < m->body_dofnum[b]; j++) d->qpos[qa + j] += h * d->qvel[da + j];
        }
    }
}

/* ---------------------------------------------------------------- HumanoidEnv restatement */

/* ego_pose/envs/humanoid_v1.py:130-156 compute_desired_accel + compute_torque (stable PD).
 * Reads d->qM / d->qfrc_bias as they are (stale by one sub-step). */
void eo_compute_torque(const EoModel *m, const EoCfg *c, const EoData *d, const double *ctrl, double *torque) {
    int nv = m->nv, nu = m->nu, first = nv - nu;
    double dt = m->timestep;
    double A[EO_MAXV * EO_MAXV], rhs[EO_MAXV], qerr[EO_MAXV];
    memcpy(A, d->qM, sizeof(double) * nv * nv);
    for (int i = 0; i < nv; i++) {
        double kp = i >= first ? c->jkp[i - first] : 0.0, kd = i >= first ? c->jkd[i - first] : 0.0;
        qerr[i] = i >= first ? d->qpos[i + 1] - ctrl[i - first] : 0.0;
        A[i * nv + i] += kd * dt;
        rhs[i] = -d->qfrc_bias[i] - kp * qerr[i] - kd * d->qvel[i];
    }
    eo_chol_solve(nv, A, rhs);
    for (int a = 0; a < nu; a++) {
        int i = first + a;
        double verr = d->qvel[i] + rhs[i] * dt;
        torque[a] = -c->jkp[a] * qerr[i] - c->jkd[a] * verr;
    }
}

/* ego_pose/envs/humanoid_v1.py:113-125 get_body_quat */
void eo_body_quat(const EoModel *m, const double *qpos, double *bquat) {
    memcpy(bquat, qpos + 3, 4 * sizeof(double));
    for (int b = 1; b < m->nbody; b++) {
        double e[3] = {0, 0, 0};
        for (int j = 0; j < m->body_dofnum[b]; j++) e[j] = qpos[m->body_qposadr[b] + j];
        eo_quat_from_euler(e[0], e[1], e[2], bquat + 4 * b);
    }
}

/* ego_pose/envs/humanoid_v1.py:98-111 get_ee_pos(transform): uses data.body_xpos as it is (stale) */
void eo_ee_pos(const EoModel *m, const EoData *d, int heading, double *ee) {
    for (int k = 0; k < EO_NEE; k++) {
        const double *x = d->xpos[m->ee_body[k]];
        if (heading) {
            double v[3] = {x[0] - d->qpos[0], x[1] - d->qpos[1], x[2] - d->qpos[2]};
            eo_transform_vec(v, d->qpos + 3, 1, ee + 3 * k);
        } else memcpy(ee + 3 * k, x, 3 * sizeof(double));
    }
}

/* ego_pose/envs/humanoid_v1.py:73-96 get_full_obs with obs_coord='heading', root_deheading=True,
 * obs_heading=False, obs_vel='full', no phase (egomimic_config.py:99-103) */
void eo_env_obs(const EoModel *m, const EoEnv *e, double *obs) {
    const EoData *d = &e->d;
    double dq[4], v[3];
    eo_de_heading(d->qpos + 3, dq);
    obs[0] = d->qpos[2];
    memcpy(obs + 1, dq, sizeof dq);
    for (int k = 7; k < m->nq; k++) obs[k - 2] = d->qpos[k];
    eo_transform_vec(d->qvel, d->qpos + 3, 1, v);
    double *ov = obs + (m->nq - 2);
    ov[0] = v[0]; ov[1] = v[1]; ov[2] = v[2];
    for (int k = 3; k < m->nv; k++) ov[k] = d->qvel[k];
}

/* envs/common/mujoco_env.py:95-101 set_state (+ sim.forward()) */
void eo_env_set_state(const EoModel *m, EoEnv *e, const double *qpos, const double *qvel) {
    memcpy(e->d.qpos, qpos, sizeof(double) * m->nq);
    memcpy(e->d.qvel, qvel, sizeof(double) * m->nv);
    eo_forward(m, &e->d);
}

/* envs/common/mujoco_env.py:84-93 reset + humanoid_v1.py:201-226 reset_model with the sampled
 * (expert_ind, start_ind) supplied by the caller and env_init_noise = 0 */
void eo_env_reset(const EoModel *m, const EoCfg *c, const EoExpert *x, EoEnv *e, int take, int start) {
    (void)c;
    memset(e->d.ctrl, 0, sizeof e->d.ctrl);     /* sim.reset() */
    e->cur_t = 0;
    e->take = take;
    e->start_ind = start;
    const double *row = x->rows + (size_t)(x->take_off[take] + start) * EO_X_STRIDE;
    eo_env_set_state(m, e, row + EO_X_QPOS, row + EO_X_QVEL);
    eo_body_quat(m, e->d.qpos, e->bquat);
}

/* ego_pose/envs/humanoid_v1.py:158-199 do_simulation + step */
void eo_env_step(const EoModel *m, const EoCfg *c, const EoExpert *x, EoEnv *e, const double *action,
                 int *fail, int *end) {
    EoData *d = &e->d;
    memcpy(e->prev_qpos, d->qpos, sizeof(double) * m->nq);
    memcpy(e->prev_qvel, d->qvel, sizeof(double) * m->nv);
    memcpy(e->prev_bquat, e->bquat, sizeof(double) * 4 * m->nbody);
    double ctrl[EO_MAXV], torque[EO_MAXV];
    for (int a = 0; a < m->nu; a++) ctrl[a] = c->a_ref[a] + action[a] * c->a_scale[a];
    for (int i = 0; i < c->frame_skip; i++) {
        eo_compute_torque(m, c, d, ctrl, torque);
        for (int a = 0; a < m->nu; a++) {
            double t = torque[a], lim = c->torque_lim[a];
            d->ctrl[a] = t < -lim ? -lim : (t > lim ? lim : t);
        }
        eo_step(m, d);
    }
    e->cur_t += 1;
    eo_body_quat(m, d->qpos, e->bquat);
    double head_z = d->xpos[m->head_body][2];
    double lb = isnan(c->fix_head_lb) ? x->head_height_lb[e->take] - 0.1 : c->fix_head_lb;
    *fail = head_z < lb;
    *end = e->cur_t >= c->episode_len;
}

/* ego_pose/core/reward_function.py:4-60 quat_space_reward_v3 */
double eo_reward(const EoModel *m, const EoCfg *c, const EoExpert *x, const EoEnv *e, int end, double *info5) {
    const EoData *d = &e->d;
    double dt = m->timestep * c->frame_skip;
    int t = e->cur_t, nb = m->nbody;
    const double *row = x->rows + (size_t)(x->take_off[e->take] + e->start_ind + t) * EO_X_STRIDE;
    double fd[6], rq[4], ee[3 * EO_NEE], cur_bquat[4 * EO_MAXB], bangvel[3 * EO_MAXB];
    eo_qvel_fd6(e->prev_qpos, d->qpos, dt, 1, fd);
    eo_de_heading(d->qpos + 3, rq);
    eo_ee_pos(m, d, 1, ee);
    eo_body_quat(m, d->qpos, cur_bquat);
    eo_angvel_fd(nb, e->prev_bquat, cur_bquat, dt, bangvel);
    const double *e_bquat = row + EO_X_BQUAT, *e_bangvel = row + EO_X_BANGVEL;
    /* pose */
    double pose2 = 0;
    for (int b = 1; b < nb; b++) {
        double qi[4], qd[4];
        eo_quat_inv(e_bquat + 4 * b, qi);
        eo_quat_mul(cur_bquat + 4 * b, qi, qd);
        double w = qd[0] < -1.0 ? -1.0 : (qd[0] > 1.0 ? 1.0 : qd[0]);
        double a = acos(w) * c->b_diffw[b - 1];
        pose2 += a * a;
    }
    double pose_dist = sqrt(pose2);
    double pose_reward = exp(-c->k_p * (pose_dist * pose_dist));
    /* velocity (body angular velocity, root ignored) */
    double vd = 0;
    for (int k = 3; k < 3 * nb; k++) {
        double df = fabs(bangvel[k] - e_bangvel[k]);
        vd += c->v_ord == 1 ? df : df * df;
    }
    double vel_dist = c->v_ord == 1 ? vd : sqrt(vd);
    double vel_reward = exp(-c->k_v * (vel_dist * vel_dist));
    /* end effectors */
    double e2 = 0;
    for (int k = 0; k < 3 * EO_NEE; k++) { double df = ee[k] - row[EO_X_EE_POS + k]; e2 += df * df; }
    double ee_dist = sqrt(e2);
    double ee_reward = exp(-c->k_e * (ee_dist * ee_dist));
    /* root pose */
    double hd = d->qpos[2] - row[EO_X_QPOS + 2], qi[4], qd[4];
    eo_quat_inv(row + EO_X_RQ_RMH, qi);
    eo_quat_mul(rq, qi, qd);
    double w = qd[0] < -1.0 ? -1.0 : (qd[0] > 1.0 ? 1.0 : qd[0]);
    double rqd = acos(w);
    double root_pose_reward = exp(-c->k_rh * (hd * hd) - c->k_rq * (rqd * rqd));
    /* root velocity */
    double l2 = 0, a2 = 0;
    for (int k = 0; k < 3; k++) {
        double dl = fd[k] - row[EO_X_RLINV_LOCAL + k], da = fd[3 + k] - row[EO_X_RANGV + k];
        l2 += dl * dl; a2 += da * da;
    }
    double ld = sqrt(l2), ad = sqrt(a2);
    double root_vel_reward = exp(-c->k_rl * (ld * ld) - c->k_ra * (ad * ad));
    double reward = c->w_p * pose_reward + c->w_v * vel_reward + c->w_e * ee_reward + c->w_rp * root_pose_reward
                    + c->w_rv * root_vel_reward;
    reward /= c->w_p + c->w_v + c->w_e + c->w_rp + c->w_rv;
    if (c->decay) reward *= 1.0 - (double)t / c->episode_len;
    if (end) reward += c->end_reward;
    info5[0] = pose_reward; info5[1] = vel_reward; info5[2] = ee_reward;
    info5[3] = root_pose_reward; info5[4] = root_vel_reward;
    return reward;
}

/* ego_pose/data_process/gen_expert.py:28-83 get_expert for one take (lb=0, ub=L): packed rows */
void eo_expert_features(const EoModel *m, int L, const double *qpos_in, double dt, double *rows,
                        double *head_z_min) {
    EoEnv *e = (EoEnv *)calloc(1, sizeof(EoEnv));
    double *bq = (double *)malloc(sizeof(double) * 4 * m->nbody * L);
    double axis_w[EO_MAXV * 3], anchor_w[EO_MAXV * 3];
    double hmin = INFINITY;
    int nq = m->nq, nv = m->nv;
    for (int i = 0; i < L; i++) {
        double *row = rows + (size_t)i * EO_X_STRIDE;
        memset(row, 0, sizeof(double) * EO_X_STRIDE);
        memcpy(row + EO_X_QPOS, qpos_in + (size_t)i * nq, sizeof(double) * nq);
        memcpy(e->d.qpos, row, sizeof(double) * nq);
        eo_kinematics(m, &e->d, axis_w, anchor_w);      /* env.sim.forward(): fresh body_xpos */
        eo_de_heading(row + 3, row + EO_X_RQ_RMH);
        eo_ee_pos(m, &e->d, 1, row + EO_X_EE_POS);
        eo_body_quat(m, row, row + EO_X_BQUAT);
        memcpy(bq + (size_t)i * 4 * m->nbody, row + EO_X_BQUAT, sizeof(double) * 4 * m->nbody);
        double hz = e->d.xpos[m->head_body][2];
        if (hz < hmin) hmin = hz;
        if (i > 0) {
            const double *prev = qpos_in + (size_t)(i - 1) * nq;
            double qv[EO_MAXV];
            eo_qvel_fd(nq, prev, row, dt, 0, qv);
            memcpy(row + EO_X_QVEL, qv, sizeof(double) * nv);
            eo_transform_vec(qv, row + 3, 1, row + EO_X_RLINV_LOCAL);   /* heading of the CURRENT frame */
            memcpy(row + EO_X_RANGV, qv + 3, 3 * sizeof(double));
            eo_angvel_fd(m->nbody, bq + (size_t)(i - 1) * 4 * m->nbody, row + EO_X_BQUAT, dt, row + EO_X_BANGVEL);
        }
    }
    if (L > 1) {        /* frame 0 copies frame 1 (gen_expert.py:67-70,76) */
        double *r0 = rows, *r1 = rows + EO_X_STRIDE;
        memcpy(r0 + EO_X_QVEL, r1 + EO_X_QVEL, sizeof(double) * nv);
        memcpy(r0 + EO_X_RLINV_LOCAL, r1 + EO_X_RLINV_LOCAL, 3 * sizeof(double));
        memcpy(r0 + EO_X_RANGV, r1 + EO_X_RANGV, 3 * sizeof(double));
        memcpy(r0 + EO_X_BANGVEL, r1 + EO_X_BANGVEL, sizeof(double) * 3 * m->nbody);
    }
    *head_z_min = hmin;
    free(bq);
    free(e);
}

/* ---------------------------------------------------------------- policy + rollout driver */

/* core/policy_gaussian.py:19-24 + models/mlp.py:22-25 (relu trunk) -> action mean.
 * scratch: h1 + h2 doubles */
void eo_policy_mean(const EoPolicy *p, const double *x, double *mean, double *scratch) {
    double *h1 = scratch, *h2 = scratch + p->h1;
    for (int j = 0; j < p->h1; j++) {
        double s = p->b1[j];
        const double *w = p->W1 + (size_t)j * p->in_dim;
        for (int k = 0; k < p->in_dim; k++) s += w[k] * x[k];
        h1[j] = s > 0 ? s : 0;
    }
    for (int j = 0; j < p->h2; j++) {
        double s = p->b2[j];
        const double *w = p->W2 + (size_t)j * p->h1;
        for (int k = 0; k < p->h1; k++) s += w[k] * h1[k];
        h2[j] = s > 0 ? s : 0;
    }
    for (int j = 0; j < p->out_dim; j++) {
        double s = p->b3[j];
        const double *w = p->W3 + (size_t)j * p->h2;
        for (int k = 0; k < p->h2; k++) s += w[k] * h2[k];
        mean[j] = s;
    }
}

/* utils/zfilter.py:58-67 ZFilter.__call__(x, update=False) with frozen statistics */
static void zfilter_apply(int n, const double *x, const double *mean, const double *std, double clip, double *y) {
    for (int k = 0; k < n; k++) {
        double v = x[k];
        if (mean) {
            v = (v - mean[k]) / (std[k] + 1e-8);
            if (clip > 0) v = v < -clip ? -clip : (v > clip ? clip : v);
        }
        y[k] = v;
    }
}

/* agents/agent.py:29-76 sample_worker, batched: n_env independent environments each record exactly T
 * steps (env-major rows n = e*T + t), auto-resetting on done from the caller's pre-drawn
 * (take, start) list; the last recorded step of every env gets mask 0 (SURVEY 7 "episode packing").
 * Noise eps[e][t][nu] is supplied by the caller (parity mode, SURVEY 7 "RNG parity"). */
typedef struct {
    const EoModel *m; const EoCfg *c; const EoExpert *x; const EoPolicy *p;
    int n_env, T, max_resets; const int *reset_take, *reset_start;
    const double *eps; const unsigned char *mean_flag;
    const double *zf_mean, *zf_std; double zf_clip;
    double *states, *actions, *rewards, *masks, *next_states, *exps; int *v_metas;
    double *c_info, *raw_obs, *final_qpos, *final_qvel;
    int next_env, status;
    pthread_mutex_t mu;
} RolloutJob;

static void rollout_env(RolloutJob *J, int e) {
    const EoModel *m = J->m; const EoCfg *c = J->c; const EoExpert *x = J->x; const EoPolicy *p = J->p;
    int T = J->T, max_resets = J->max_resets;
    int S = m->nq - 2 + m->nv, nu = m->nu, D = p->in_dim, cd = x->ctx_dim;
    EoEnv *env = (EoEnv *)calloc(1, sizeof(EoEnv));
    double *xin = (double *)malloc(sizeof(double) * (D + p->h1 + p->h2 + 4 * S + 2 * nu));
    double *scratch = xin + D, *obs = scratch + p->h1 + p->h2, *state = obs + S, *nobs = state + S,
           *nstate = nobs + S, *mean = nstate + S, *act = mean + nu;
    int r = 0;
    eo_env_reset(m, c, x, env, J->reset_take[(size_t)e * max_resets], J->reset_start[(size_t)e * max_resets]);
    eo_env_obs(m, env, obs);
    zfilter_apply(S, obs, J->zf_mean, J->zf_std, J->zf_clip, state);
    for (int t = 0; t < T; t++) {
        size_t n = (size_t)e * T + t;
        int off = 0;
        if (x->ctx) {
            const double *cx = x->ctx + (size_t)(x->take_off[env->take] + env->start_ind + env->cur_t) * cd;
            memcpy(xin, cx, sizeof(double) * cd);
            off = cd;
        }
        memcpy(xin + off, state, sizeof(double) * S);
        eo_policy_mean(p, xin, mean, scratch);
        int mf = J->mean_flag ? J->mean_flag[n] : 0;
        for (int a = 0; a < nu; a++)
            act[a] = mf ? mean[a] : mean[a] + exp(p->log_std[a]) * J->eps[n * nu + a];
        int fail, end;
        if (J->raw_obs) memcpy(J->raw_obs + n * S, obs, sizeof(double) * S);
        eo_env_step(m, c, x, env, act, &fail, &end);
        eo_env_obs(m, env, nobs);
        zfilter_apply(S, nobs, J->zf_mean, J->zf_std, J->zf_clip, nstate);
        double info5[5];
        double rew = eo_reward(m, c, x, env, end, info5);
        int done = fail || end;
        memcpy(J->states + n * S, state, sizeof(double) * S);
        memcpy(J->actions + n * nu, act, sizeof(double) * nu);
        if (J->next_states) memcpy(J->next_states + n * S, nstate, sizeof(double) * S);
        J->rewards[n] = rew;
        J->masks[n] = (done || t == T - 1) ? 0.0 : 1.0;
        J->exps[n] = mf ? 0.0 : 1.0;
        J->v_metas[2 * n] = env->take;
        J->v_metas[2 * n + 1] = env->start_ind;
        if (J->c_info) memcpy(J->c_info + n * 5, info5, sizeof info5);
        if (done && t < T - 1) {
            r++;
            if (r >= max_resets) {
                pthread_mutex_lock(&J->mu);
                J->status = -3;
                pthread_mutex_unlock(&J->mu);
                r = max_resets - 1;
            }
            eo_env_reset(m, c, x, env, J->reset_take[(size_t)e * max_resets + r],
                         J->reset_start[(size_t)e * max_resets + r]);
            eo_env_obs(m, env, obs);
            zfilter_apply(S, obs, J->zf_mean, J->zf_std, J->zf_clip, state);
        } else {
            memcpy(state, nstate, sizeof(double) * S);
            memcpy(obs, nobs, sizeof(double) * S);
        }
    }
    if (J->final_qpos) memcpy(J->final_qpos + (size_t)e * m->nq, env->d.qpos, sizeof(double) * m->nq);
    if (J->final_qvel) memcpy(J->final_qvel + (size_t)e * m->nv, env->d.qvel, sizeof(double) * m->nv);
    free(xin);
    free(env);
}

static void *rollout_worker(void *arg) {
    RolloutJob *J = (RolloutJob *)arg;
    for (;;) {
        pthread_mutex_lock(&J->mu);
        int e = J->next_env++;
        pthread_mutex_unlock(&J->mu);
        if (e >= J->n_env) break;
        rollout_env(J, e);
    }
    return NULL;
}

int eo_rollout(const EoModel *m, const EoCfg *c, const EoExpert *x, const EoPolicy *p,
               int n_env, int T, int max_resets, const int *reset_take, const int *reset_start,
               const double *eps, const unsigned char *mean_flag,
               const double *zf_mean, const double *zf_std, double zf_clip,
               double *states, double *actions, double *rewards, double *masks, double *next_states,
               double *exps, int *v_metas, double *c_info, double *raw_obs, double *final_qpos,
               double *final_qvel, int n_threads) {
    int S = m->nq - 2 + m->nv;
    if (p->in_dim != S + (x->ctx ? x->ctx_dim : 0)) return -2;
    RolloutJob J = {m, c, x, p, n_env, T, max_resets, reset_take, reset_start, eps, mean_flag,
                    zf_mean, zf_std, zf_clip, states, actions, rewards, masks, next_states, exps, v_metas,
                    c_info, raw_obs, final_qpos, final_qvel, 0, 0, PTHREAD_MUTEX_INITIALIZER};
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    if (n_threads == 1) { rollout_worker(&J); return J.status; }
    pthread_t th[256];
    for (int i = 0; i < n_threads; i++) pthread_create(&th[i], NULL, rollout_worker, &J);
    for (int i = 0; i < n_threads; i++) pthread_join(th[i], NULL);
    return J.status;
}
