"""TEST INFRASTRUCTURE ONLY - never imported by the product path.

CPU restatement of the evaluation roll-out of ego_pose/ego_mimic_eval.py:93-177 (eval_expert) on top of the C
oracle's env primitives: one take from frame fr_margin, mean action, ZFilter frozen (update=False), simulator
state recorded BEFORE every step, and the 'naivefs' fail-safe (fix_head_lb, :52-53,167-173): a failing humanoid is
not restarted, its state is replaced by the predicted observation of the next frame aligned to the simulated root
(reset_env_state :93-99 + align_human_state utils/tools.py:71-75) and the episode clock keeps running.

Pinned by tests/golden/eval_traj.npz (the reference's own HumanoidEnv / align_human_state / PolicyGaussian driven in
the script's call order on the restated physics, tests/golden/make_golden.py gen_eval).  The state-regression net
that produces ``state_pred`` in the reference (models/video_reg_net.py) is outside the hot path: the table is an
input here.  'valuefs' (value < 0.6 x running mean over all takes, :156-159,167) is restated with the Value MLP.
"""
import numpy as np

from . import cphys


def quat_mul(q1, q0):
    """utils/transformation.py:1379-1391 quaternion_multiply(q1, q0), (w, x, y, z)"""
    w0, x0, y0, z0 = q0
    w1, x1, y1, z1 = q1
    return np.array([-x1 * x0 - y1 * y0 - z1 * z0 + w1 * w0, x1 * w0 + y1 * z0 - z1 * y0 + w1 * x0,
                     -x1 * z0 + y1 * w0 + z1 * x0 + w1 * y0, x1 * y0 - y1 * x0 + z1 * w0 + w1 * z0])


def heading_q(q):
    """utils/math.py:60-65 get_heading_q"""
    hq = np.array([q[0], 0.0, 0.0, q[3]])
    return hq / np.linalg.norm(hq)


def reset_env_state(orc, env, state, nq):
    """ego_mimic_eval.py:93-99: qpos[2:], qvel from the predicted observation; align_human_state (utils/tools.py:71-75)
    keeps the simulated root xy and heading; env.set_state -> sim.forward()"""
    ref_qpos = np.array(env.d.qpos[:nq])
    qpos = ref_qpos.copy()
    qpos[2:] = state[:nq - 2]
    qvel = np.array(state[nq - 2:], dtype=np.float64)
    hq = heading_q(ref_qpos[3:7])
    qpos[3:7] = quat_mul(hq, qpos[3:7])
    c, s = hq[0] * hq[0] - hq[3] * hq[3], 2.0 * hq[0] * hq[3]          # quat_mul_vec(hq, v): rotation about z
    qvel[:3] = [c * qvel[0] - s * qvel[1], s * qvel[0] + c * qvel[1], qvel[2]]
    orc.env_set_state(env, qpos, qvel)
    return orc.env_obs(env)


def eval_take(orc, policy, take, fr_margin, test_len, state_pred, ctx=None, zf_mean=None, zf_std=None, zf_clip=5.0,
              fail_safe='naivefs', value_policy=None, vctx=None, value_stat=None):
    """state_pred [L, S] and ctx [L, ctx_dim] are indexed by the take's frame (frame = fr_margin + t).
    fail_safe 'valuefs' (:156-159,167): ``value_policy`` = the Value MLP as an EoPolicy with out_dim 1, ``vctx`` its
    context table, ``value_stat`` = [n, mean] list shared across takes (RunningStat(1), utils/zfilter.py:18-27).
    Returns dict(traj_pred [n, nq], vel_pred [n, nv], states [n, S], actions, rewards, values, num_reset)."""
    nq, nv = orc.nq, orc.nv

    def zf(x):
        if zf_mean is None:
            return x
        y = (x - zf_mean) / (zf_std + 1e-8)
        return np.clip(y, -zf_clip, zf_clip) if zf_clip > 0 else y

    orc.cfg.episode_len = int(test_len)                 # env.set_fix_sampling(expert_ind, fr_margin, test_len)
    env = cphys.EoEnv()
    orc.env_reset(env, int(take), int(fr_margin))
    state = zf(reset_env_state(orc, env, state_pred[fr_margin], nq))
    out = dict(traj_pred=[], vel_pred=[], states=[], actions=[], rewards=[], values=[], num_reset=0)
    for t in range(test_len):
        out['traj_pred'].append(np.array(env.d.qpos[:nq]))
        out['vel_pred'].append(np.array(env.d.qvel[:nv]))
        x = state if ctx is None else np.concatenate([ctx[fr_margin + t], state])
        value = 0.0
        if value_policy is not None:
            vt = vctx if vctx is not None else ctx
            value = float(orc.policy_mean(value_policy, state if vt is None else np.concatenate([vt[fr_margin + t], state]))[0])
            value_stat[0] += 1                                      # RunningStat.push
            value_stat[1] += (value - value_stat[1]) / value_stat[0]
            out['values'].append(value)
        action = orc.policy_mean(policy, x)
        fail, end = orc.env_step(env, action)
        next_state = zf(orc.env_obs(env))
        rew, _ = orc.env_reward(env, end)
        out['states'].append(state)
        out['actions'].append(action)
        out['rewards'].append(rew)
        if end:
            break
        if (fail_safe == 'naivefs' and fail) or (fail_safe == 'valuefs' and value < 0.6 * value_stat[1]):
            out['num_reset'] += 1
            state = zf(reset_env_state(orc, env, state_pred[fr_margin + t + 1], nq))
        else:
            state = next_state
    for k in ('traj_pred', 'vel_pred', 'states', 'actions'):
        out[k] = np.array(out[k])
    out['rewards'] = np.array(out['rewards'])
    out['values'] = np.array(out['values'])
    return out


def sync_traj(qpos_traj, qvel_traj, ref_qpos):
    """ego_pose/utils/tools.py:18-32"""
    h0 = heading_q(qpos_traj[0, 3:7])
    rel = quat_mul(heading_q(ref_qpos[3:7]), np.array([h0[0], -h0[1], -h0[2], -h0[3]]))     # quaternion_inverse of a unit q
    c, s = rel[0] * rel[0] - rel[3] * rel[3], 2.0 * rel[0] * rel[3]
    R = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])                                # quat_mul_vec(rel, .)
    start_pos = np.array([qpos_traj[0, 0], qpos_traj[0, 1], ref_qpos[2]])
    qp, qv = [], []
    for qpos, qvel in zip(qpos_traj, qvel_traj):
        nq_, nv_ = qpos.copy(), qvel.copy()
        nq_[:2] = (R @ (qpos[:3] - start_pos))[:2] + ref_qpos[:2]
        nq_[3:7] = quat_mul(rel, qpos[3:7])
        nv_[:3] = R @ qvel[:3]
        qp.append(nq_)
        qv.append(nv_)
    return np.vstack(qp), np.vstack(qv)


def forecast_init(expert_qpos, em_traj, em_vel, start, fm, test_len, em_offset):
    """ego_forecast_eval.py:107-136 without --gt-init -> (qpos0, qvel0, past rows of traj_pred [fm, nq])"""
    lo = max(0, start - fm - em_offset)
    state_pred = em_traj[lo: start + test_len - em_offset]
    vel_pred = em_vel[lo: start + test_len - em_offset]
    miss_len = fm + test_len - state_pred.shape[0]
    if start - fm - em_offset >= 0:
        state_pred, vel_pred = sync_traj(state_pred, vel_pred, expert_qpos[start - fm])
    ind = fm - miss_len
    past = []
    for t in range(-fm, 0):
        past.append(expert_qpos[start + t] if t + fm < miss_len else state_pred[t + fm - miss_len])
    return state_pred[ind].copy(), vel_pred[ind].copy(), np.array(past)


def forecast_window(orc, policy, take, start, test_len, ctx=None, zf_mean=None, zf_std=None, zf_clip=5.0, init=None):
    """One window of ego_pose/ego_forecast_eval.py:95-180 with --gt-init: reset to the expert state of frame ``start``,
    ``test_len`` mean-action steps, simulator qpos recorded before every step; a fall does not stop the window
    (:171-176).  ``ctx`` [L, ctx_dim] is a per-frame context table (identity video net); ``init`` = (qpos, qvel) replaces
    the expert start state (no --gt-init)."""
    nq = orc.nq

    def zf(x):
        if zf_mean is None:
            return x
        y = (x - zf_mean) / (zf_std + 1e-8)
        return np.clip(y, -zf_clip, zf_clip) if zf_clip > 0 else y

    orc.cfg.episode_len = int(test_len)
    env = cphys.EoEnv()
    orc.env_reset(env, int(take), int(start))
    if init is not None:                                # env.set_state(qpos, qvel) of the ego-mimic prediction (:119-121)
        orc.env_set_state(env, init[0], init[1])
    state = zf(orc.env_obs(env))
    traj = []
    for t in range(test_len):
        traj.append(np.array(env.d.qpos[:nq]))
        x = state if ctx is None else np.concatenate([ctx[start + t], state])
        orc.env_step(env, orc.policy_mean(policy, x))
        state = zf(orc.env_obs(env))
    return np.array(traj)
