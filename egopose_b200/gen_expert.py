"""Expert feature files: ego_pose/data_process/gen_expert.py:28-100 on the GPU.

``get_expert`` turns one take's mocap ``qpos`` trajectory into the reference's expert dict (13 arrays + 3 scalars,
one thread per frame in ``egp_expert_features_ex_f64``); ``write_expert_file`` pickles ``{take: expert}`` exactly as
gen_expert.py:99-100 does, so the file loads in the reference's ``HumanoidEnv.load_experts`` (humanoid_v1.py:45-54) and
in ``egopose_b200.env.HumanoidEnv.load_experts`` alike.

Bug-compatible details kept: hand joints are zeroed in place (:38-39); frame 0 copies the finite-difference
velocities of frame 1 (:67-70,76); ``obs`` is taken while the simulator's qvel is still zero (the script only writes
qpos before ``sim.forward()``, :40-43), so its velocity half is zero.
"""
import pickle

import numpy as np

from . import lib


def get_expert(env, expert_qpos, lb=0, ub=None):
    """gen_expert.py:28-83 get_expert(expert_qpos, lb, ub) -> dict"""
    X = lib.X
    q = np.array(expert_qpos, dtype=np.float64, copy=True)
    for hand in ('LeftHand', 'RightHand'):
        a, b = env.body_qposaddr[hand]
        q[:, a:b] = 0.0
    rows, _, extras = env.kernel.expert_features(q, extras=True)
    rows, extras = rows.cpu().numpy(), extras.cpu().numpy()
    nq, nv, nb = env.md.nq, env.md.nv, env.md.nbody
    ub = q.shape[0] if ub is None else ub
    col = lambda name, w: rows[lb:ub, X[name]:X[name] + w].copy()  # noqa: E731
    ex = {'qpos': q[lb:ub].copy(), 'qvel': col('QVEL', nv), 'rlinv': col('QVEL', 3), 'rlinv_local': col('RLINV_LOCAL', 3),
          'rangv': col('RANGV', 3), 'rq_rmh': col('RQ_RMH', 4), 'ee_pos': col('EE_POS', 15), 'bquat': col('BQUAT', 4 * nb),
          'bangvel': col('BANGVEL', 3 * nb), 'head_pos': extras[lb:ub, 0:3].copy(), 'com': extras[lb:ub, 3:6].copy(),
          'ee_wpos': extras[lb:ub, 6:21].copy()}
    ex['obs'] = np.concatenate([q[lb:ub, 2:3], ex['rq_rmh'], q[lb:ub, 7:], np.zeros((ub - lb, nv))], axis=1)
    ex['len'] = ex['qpos'].shape[0]
    ex['height_lb'] = ex['qpos'][:, 2].min()
    ex['head_height_lb'] = ex['head_pos'][:, 2].min()
    return ex


def gen_expert_dict(env, takes, trajs, msync=None):
    """the loop of gen_expert.py:86-95: trajs[i] = mocap qpos of takes[i]; msync[take] = (_, lb, ub) frame window"""
    out = {}
    for take, traj in zip(takes, trajs):
        lb, ub = (0, None) if msync is None else msync[take][1:3]
        out[take] = get_expert(env, traj, lb, ub)
    return out


def write_expert_file(path, expert_dict):
    """gen_expert.py:99-100"""
    with open(path, 'wb') as f:
        pickle.dump(expert_dict, f)
