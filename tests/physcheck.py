"""Independent check of the oracle's M(q) qacc + C(q, v): per-body Newton-Euler with accelerations and
partial velocities obtained ONLY by finite differences of the forward kinematics (no spatial algebra,
no CRBA/RNE).  Shares nothing with oracle/egopose_oracle.c except eo_kinematics."""
import numpy as np


def quat_mul(a, b):
    w1, x1, y1, z1 = a
    w0, x0, y0, z0 = b
    return np.array([w1 * w0 - x1 * x0 - y1 * y0 - z1 * z0, w1 * x0 + x1 * w0 + y1 * z0 - z1 * y0,
                     w1 * y0 - x1 * z0 + y1 * w0 + z1 * x0, w1 * z0 + x1 * y0 - y1 * x0 + z1 * w0])


def quat_exp(v):
    a = np.linalg.norm(v)
    if a < 1e-300:
        return np.array([1.0, 0, 0, 0])
    return np.concatenate([[np.cos(a / 2)], np.sin(a / 2) * v / a])


def quat_to_mat(q):
    w, x, y, z = q
    return np.array([[w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])


def rot_log(R):
    c = np.clip((np.trace(R) - 1) / 2, -1, 1)
    ang = np.arccos(c)
    v = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2
    s = np.linalg.norm(v)
    if s < 1e-300:
        return np.zeros(3)
    return v / s * ang


def displace(q0, v, a, t):
    """configuration at time t for initial velocity v and constant generalized acceleration a"""
    q = q0.copy()
    q[:3] = q0[:3] + v[:3] * t + 0.5 * a[:3] * t * t
    w = v[3:6] * t + 0.5 * a[3:6] * t * t + np.cross(v[3:6], a[3:6]) * t ** 3 / 12.0
    q[3:7] = quat_mul(q0[3:7] / np.linalg.norm(q0[3:7]), quat_exp(w))
    q[7:] = q0[7:] + v[6:] * t + 0.5 * a[6:] * t * t
    return q


def body_poses(orc, q):
    xpos, xquat, xipos, _, _ = orc.kinematics(q)
    return xipos, np.stack([quat_to_mat(x) for x in xquat])


def newton_euler_tau(orc, q, v, a, delta=2e-4, eps=1e-6):
    md = orc.md
    nb, nv = md['nbody'], md['nv']
    g = np.array(md['gravity'])
    cm, Rm = body_poses(orc, displace(q, v, a, -delta))
    c0, R0 = body_poses(orc, q)
    cp, Rp = body_poses(orc, displace(q, v, a, +delta))
    F = np.zeros((nb, 3))
    N = np.zeros((nb, 3))
    for b in range(nb):
        acc = (cp[b] - 2 * c0[b] + cm[b]) / delta ** 2
        w_plus = rot_log(Rp[b] @ R0[b].T) / delta
        w_minus = rot_log(R0[b] @ Rm[b].T) / delta
        w = 0.5 * (w_plus + w_minus)
        alpha = (w_plus - w_minus) / delta
        inn = md['body_inertia'][b]
        Ib = np.array([[inn[0], inn[3], inn[4]], [inn[3], inn[1], inn[5]], [inn[4], inn[5], inn[2]]])
        Iw = R0[b] @ Ib @ R0[b].T
        F[b] = md['body_mass'][b] * (acc - g)
        N[b] = Iw @ alpha + np.cross(w, Iw @ w)
    tau = np.zeros(nv)
    zero = np.zeros(nv)
    for i in range(nv):
        e = np.zeros(nv)
        e[i] = 1.0
        c1, R1 = body_poses(orc, displace(q, e, zero, eps))
        c2, R2 = body_poses(orc, displace(q, e, zero, -eps))
        for b in range(nb):
            jv = (c1[b] - c2[b]) / (2 * eps)
            jw = rot_log(R1[b] @ R2[b].T) / (2 * eps)
            tau[i] += F[b] @ jv + N[b] @ jw
    return tau
