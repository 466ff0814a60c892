"""Pose metrics of ego_pose/eval_pose.py:31-66 ('stats' mode) over the (results, meta) pairs egopose_b200.evaluate writes:
joint-angle distance, joint-velocity distance and smoothness (mean |acceleration|), as defined in
ego_pose/utils/metrics.py.  Host numpy, vectorised over frames (the reference loops frame by frame); not a hot path.
"""
import numpy as np


def _quat_to_mat(q):
    """utils/transformation.py:1267-1291 quaternion_matrix for unit-length-normalised q, batched [n, 4] -> [n, 3, 3]"""
    q = np.asarray(q, dtype=np.float64)
    n2 = np.einsum('ij,ij->i', q, q)
    s = np.sqrt(2.0 / n2)[:, None] * q
    w, x, y, z = s[:, 0], s[:, 1], s[:, 2], s[:, 3]
    R = np.empty((q.shape[0], 3, 3))
    R[:, 0, 0] = 1.0 - y * y - z * z; R[:, 0, 1] = x * y - z * w;       R[:, 0, 2] = x * z + y * w      # noqa: E702
    R[:, 1, 0] = x * y + z * w;       R[:, 1, 1] = 1.0 - x * x - z * z; R[:, 1, 2] = y * z - x * w      # noqa: E702
    R[:, 2, 0] = x * z - y * w;       R[:, 2, 1] = y * z + x * w;       R[:, 2, 2] = 1.0 - x * x - y * y  # noqa: E702
    return R


def get_joint_angles(poses):
    """metrics.py:5-13: root Euler angles ('sxyz', utils/transformation.py:1125-1180) with the yaw zeroed | hinge angles"""
    poses = np.asarray(poses, dtype=np.float64)
    M = _quat_to_mat(poses[:, 3:7])
    cy = np.sqrt(M[:, 0, 0] ** 2 + M[:, 1, 0] ** 2)
    ok = cy > np.finfo(float).eps * 4.0
    ax = np.where(ok, np.arctan2(M[:, 2, 1], M[:, 2, 2]), np.arctan2(-M[:, 1, 2], M[:, 1, 1]))
    ay = np.arctan2(-M[:, 2, 0], cy)
    return np.hstack([ax[:, None], ay[:, None], np.zeros((poses.shape[0], 1)), poses[:, 7:]])


def _quat_mul(q1, q0):
    w0, x0, y0, z0 = q0.T
    w1, x1, y1, z1 = q1.T
    return np.stack([-x1 * x0 - y1 * y0 - z1 * z0 + w1 * w0, x1 * w0 + y1 * z0 - z1 * y0 + w1 * x0,
                     -x1 * z0 + y1 * w0 + z1 * x0 + w1 * y0, x1 * y0 - y1 * x0 + z1 * w0 + w1 * z0], axis=1)


def get_joint_vels(poses, dt):
    """metrics.py:16-22: get_qvel_fd(poses[i], poses[i + 1], dt, 'heading') (utils/math.py:20-35) for every frame pair"""
    poses = np.asarray(poses, dtype=np.float64)
    cur, nxt = poses[:-1], poses[1:]
    v = (nxt[:, :3] - cur[:, :3]) / dt
    qc = cur[:, 3:7]
    qi = qc * np.array([1.0, -1.0, -1.0, -1.0]) / np.einsum('ij,ij->i', qc, qc)[:, None]     # quaternion_inverse
    qrel = _quat_mul(nxt[:, 3:7], qi)
    small = 1.0 - qrel[:, 0] < 1e-8                                                          # rotation_from_quaternion
    s = np.sqrt(np.where(small, 1.0, 1.0 - qrel[:, 0] ** 2))
    axis = np.where(small[:, None], np.array([1.0, 0.0, 0.0]), qrel[:, 1:4] / s[:, None])
    angle = np.where(small, 0.0, 2.0 * np.arccos(np.clip(qrel[:, 0], -1.0, 1.0)))
    angle = np.where(angle > np.pi, angle - 2 * np.pi, np.where(angle < -np.pi, angle + 2 * np.pi, angle))
    rv = axis * (angle / dt)[:, None]
    rv = np.einsum('nji,nj->ni', _quat_to_mat(qc), rv)                                       # transform_vec(., q, 'root'): R^T v
    hq = qc * np.array([1.0, 0.0, 0.0, 1.0])
    vl = np.einsum('nji,nj->ni', _quat_to_mat(hq), v)                                        # 'heading'
    return np.hstack([vl, rv, (nxt[:, 7:] - cur[:, 7:]) / dt])


def get_joint_accels(vels, dt):
    return np.diff(vels, axis=0) / dt


def get_mean_dist(x, y):
    return np.linalg.norm(x - y, axis=1).mean()


def get_mean_abs(x):
    return np.abs(x).mean()


def remove_noisy_hands(results):
    """ego_pose/utils/tools.py:35-40 (qpos slices 32:35 and 42:45), in place"""
    if results is None:
        return
    for traj in results.values():
        for take in traj:
            traj[take][..., 32:35] = 0
            traj[take][..., 42:45] = 0


def compute_metrics(results, dt=1.0 / 30.0):
    """eval_pose.py:31-66 -> dict(pose_dist, vel_dist, smoothness, per_take={take: (pose, vel, accel)})"""
    per = {}
    for take, traj in results['traj_pred'].items():
        gt = results['traj_orig'][take]
        angs_gt, vels_gt = get_joint_angles(gt), get_joint_vels(gt, dt)
        angs, vels = get_joint_angles(traj), get_joint_vels(traj, dt)
        per[take] = (get_mean_dist(angs, angs_gt), get_mean_dist(vels, vels_gt), get_mean_abs(get_joint_accels(vels, dt)))
    n = len(per)
    return dict(pose_dist=sum(p[0] for p in per.values()) / n, vel_dist=sum(p[1] for p in per.values()) / n,
                smoothness=sum(p[2] for p in per.values()) / n, per_take=per)
