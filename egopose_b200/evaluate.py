"""Evaluation roll-outs on the fused rollout kernel.

Mirrors ego_pose/ego_mimic_eval.py:102-195: every take of the loaded expert list is rolled out once from frame
``fr_margin`` for ``len - 2 * fr_margin`` steps under the mean action with frozen ZFilter statistics, the simulator
state is recorded before every step, and the result is the ``(results, meta)`` pair the script pickles
(``results = {'traj_pred', 'traj_orig', 'vel_pred'}`` keyed by take, ``meta = {'algo', 'num_reset'}``, :186-193) and
ego_pose/eval_pose.py:31-66 consumes.  All takes run in ONE kernel launch (environment e = take e, per-environment
episode length); the reference loops over takes and steps in Python.

Fail-safe (:167-173): 'naivefs' (head below ``fix_head_lb`` = 0.3, :52-53) replaces the state of a fallen humanoid in
place by ``state_pred`` of the next frame, aligned to the simulated root (reset_env_state :93-99), inside the kernel;
'none' never replaces it.  ``state_pred`` is what the reference obtains from its state-regression net
(models/video_reg_net.py, outside the hot path): pass its per-take predictions, or leave it None to use the
expert's own observation of each frame.  'valuefs' (the script's default) evaluates the value net inside the step
loop and replaces the state when value < 0.6 x the running mean of all values so far.
"""
import pickle

import numpy as np
import torch

from . import lib


def expert_obs_table(model):
    """[total_frames, S]: HumanoidEnv.get_obs() (humanoid_v1.py:73-96) of every expert frame (qpos, qvel), read off the packed
    expert rows: qpos[2] | de-headed root quaternion | hinge angles | root velocity in the heading frame | qvel[3:].
    The heading transform is applied here (utils/math.py:47-59) rather than taken from ``rlinv_local``: a take's frame 0
    copies frame 1's finite differences (gen_expert.py:67-70), so its stored rlinv_local belongs to frame 1's heading."""
    X, r = lib.X, model.rows_host
    nq, nv = model.nq, model.nv
    q, v = r[:, X['QPOS']:X['QPOS'] + nq], r[:, X['QVEL']:X['QVEL'] + nv]
    hn = np.sqrt(q[:, 3] ** 2 + q[:, 6] ** 2)
    hw, hz = q[:, 3] / hn, q[:, 6] / hn
    c, s = hw * hw - hz * hz, 2.0 * hw * hz                     # rotation about z by the heading; transform_vec applies R^T
    vl = np.stack([c * v[:, 0] + s * v[:, 1], -s * v[:, 0] + c * v[:, 1], v[:, 2]], axis=1)
    return np.ascontiguousarray(np.concatenate([q[:, 2:3], r[:, X['RQ_RMH']:X['RQ_RMH'] + 4], q[:, 7:], vl, v[:, 3:]], axis=1))


def _context_table(env, policy_vs_net, dev):
    """per-frame context rows for whole-take episodes: test-mode VideoStateNet.initialize(cnn_feat) over the full
    take (ego_mimic_eval.py:122-124), placed at the frames it describes; a tensor / array is taken as the table itself;
    anything else (FrameContext, None) -> the table uploaded with the experts"""
    from .nets import VideoStateNet
    if isinstance(policy_vs_net, (np.ndarray, torch.Tensor)):      # a ready per-frame table [total_frames, dim]
        return torch.as_tensor(policy_vs_net, dtype=torch.float64, device=dev).contiguous()
    if not isinstance(policy_vs_net, VideoStateNet):
        return None
    m = policy_vs_net.v_margin
    p = policy_vs_net._p()
    rows = []
    with torch.no_grad():
        for feat in env.cnn_feat:
            f = torch.as_tensor(feat, dtype=p.dtype, device=p.device)
            t = torch.zeros((f.shape[0], policy_vs_net.v_hdim), dtype=p.dtype, device=p.device)
            t[m:f.shape[0] - m] = policy_vs_net.forward_v_net(f.unsqueeze(1)).squeeze(1)[m:-m]
            rows.append(t)
    return torch.cat(rows).to(dev).contiguous()


def _mlp_weights(net, head):
    L = net.net.affine_layers
    w = dict(W1=L[0].weight.data, b1=L[0].bias.data, W2=L[1].weight.data, b2=L[1].bias.data, W3=head.weight.data, b3=head.bias.data)
    return {k: t.contiguous() for k, t in w.items()}


def eval_takes(env, policy_net, policy_vs_net=None, running_state=None, state_pred=None, fail_safe='naivefs',
               fix_head_lb=0.3, show_noise=False, seed=1, algo='ego_mimic', value_net=None, value_vs_net=None,
               sequential=True):
    """-> (results, meta, info).  ``state_pred``: None or a per-take list of [len, S] arrays indexed by frame.
    fail_safe 'valuefs' (the script's default, :29,167) needs ``value_net`` (+ ``value_vs_net``): the value net runs
    inside the step loop and a state is replaced when value < 0.6 x the running mean of all values so far.  The script
    shares that mean across takes in list order: ``sequential=True`` reproduces it with one launch per take (the
    statistic is carried on the device); ``sequential=False`` runs all takes in one launch with a mean per take."""
    if fail_safe not in ('naivefs', 'none', 'valuefs'):
        raise ValueError('fail_safe %r' % (fail_safe,))
    if fail_safe == 'valuefs':
        if value_net is None:
            raise ValueError("fail_safe 'valuefs' needs value_net")
        return _eval_takes_valuefs(env, policy_net, policy_vs_net, running_state, state_pred, show_noise, seed, algo,
                                   value_net, value_vs_net, sequential)
    model, cfg = env.kernel, env.cfg
    fm = int(cfg.fr_margin)
    off = np.asarray(model.take_off, dtype=np.int64)
    lens = np.diff(off)
    test_len = lens - 2 * fm
    if test_len.min() < 1:
        raise ValueError('a take is shorter than 2 * fr_margin + 1 frames')
    E, T = len(lens), int(test_len.max())
    dev = policy_net.action_mean.weight.device
    L = policy_net.net.affine_layers
    w = dict(W1=L[0].weight.data, b1=L[0].bias.data, W2=L[1].weight.data, b2=L[1].bias.data,
             W3=policy_net.action_mean.weight.data, b3=policy_net.action_mean.bias.data,
             log_std=policy_net.action_log_std.data.view(-1))
    w = {k: t.contiguous() for k, t in w.items()}
    sp = expert_obs_table(model) if state_pred is None else np.concatenate([np.asarray(a, dtype=np.float64) for a in state_pred])
    cu = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)  # noqa: E731
    zm = zs = None
    clip = 0.0
    if running_state is not None:
        zm, zs = cu(running_state.rs.mean, torch.float64), cu(running_state.rs.std, torch.float64)
        clip = running_state.clip or 0.0
    head_lb = fix_head_lb if fail_safe == 'naivefs' else -1e30          # 'none': the fail rule never fires
    out = model.rollout(
        w, E, T, episode_len=T, fr_margin=fm, fix_head_lb=head_lb, mean_action=not show_noise, zf_mean=zm, zf_std=zs,
        zf_clip=clip, seed=seed, reset_take=cu(np.arange(E)[:, None], torch.int32),
        reset_start=cu(np.full((E, 1), fm), torch.int32), want_next=False, want_raw=False,
        ctx=_context_table(env, policy_vs_net, dev), eval_mode=True, fix_len=cu(test_len, torch.int32),
        state_pred=cu(sp, torch.float64), want_traj=True)
    qpos = out['qpos_traj'].view(E, T, model.nq).cpu().numpy()
    qvel = out['qvel_traj'].view(E, T, model.nv).cpu().numpy()
    rewards = out['rewards'].view(E, T).cpu().numpy()
    X = lib.X
    results = {'traj_pred': {}, 'traj_orig': {}, 'vel_pred': {}}
    for e, take in enumerate(env.expert_list):
        n = int(test_len[e])
        results['traj_pred'][take] = qpos[e, :n].copy()
        results['vel_pred'][take] = qvel[e, :n].copy()
        results['traj_orig'][take] = model.rows_host[off[e] + fm:off[e] + fm + n, X['QPOS']:X['QPOS'] + model.nq].copy()
    num_reset = int(round(float(out['logger'][lib.LOG['NUM_FAILSAFE_RESETS']].item())))
    meta = {'algo': algo, 'num_reset': num_reset}
    info = {'rewards': {take: rewards[e, :int(test_len[e])].copy() for e, take in enumerate(env.expert_list)},
            'test_len': {take: int(test_len[e]) for e, take in enumerate(env.expert_list)}}
    return results, meta, info


def _eval_takes_valuefs(env, policy_net, policy_vs_net, running_state, state_pred, show_noise, seed, algo, value_net,
                        value_vs_net, sequential):
    model, cfg = env.kernel, env.cfg
    fm = int(cfg.fr_margin)
    off = np.asarray(model.take_off, dtype=np.int64)
    lens = np.diff(off)
    test_len = lens - 2 * fm
    if test_len.min() < 1:
        raise ValueError('a take is shorter than 2 * fr_margin + 1 frames')
    n_takes = len(lens)
    dev = policy_net.action_mean.weight.device
    w = _mlp_weights(policy_net, policy_net.action_mean)
    w['log_std'] = policy_net.action_log_std.data.view(-1).contiguous()
    vw = _mlp_weights(value_net, value_net.value_head)
    sp = expert_obs_table(model) if state_pred is None else np.concatenate([np.asarray(a, dtype=np.float64) for a in state_pred])
    cu = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)  # noqa: E731
    zm = zs = None
    clip = 0.0
    if running_state is not None:
        zm, zs = cu(running_state.rs.mean, torch.float64), cu(running_state.rs.std, torch.float64)
        clip = running_state.clip or 0.0
    ctx, vctx, sp_d = _context_table(env, policy_vs_net, dev), _context_table(env, value_vs_net, dev), cu(sp, torch.float64)
    if ctx is not None and vctx is None and model.ctx_dim == 0:
        raise ValueError('value_vs_net: no context table for the value net')
    if ctx is not None and vctx is None:
        vctx = torch.as_tensor(np.concatenate(env.cnn_feat), dtype=torch.float64, device=dev).contiguous()
    groups = [[e] for e in range(n_takes)] if sequential else [list(range(n_takes))]
    stat = torch.zeros((1, 2), dtype=torch.float64, device=dev)        # (n, mean) of value_stat, shared in script order
    X = lib.X
    results = {'traj_pred': {}, 'traj_orig': {}, 'vel_pred': {}}
    info = {'rewards': {}, 'values': {}, 'test_len': {}}
    num_reset = 0
    for g in groups:
        E, T = len(g), int(test_len[g].max())
        vs = stat if sequential else torch.zeros((E, 2), dtype=torch.float64, device=dev)
        out = model.rollout(
            w, E, T, episode_len=T, fr_margin=fm, fix_head_lb=-1e30, mean_action=not show_noise, zf_mean=zm, zf_std=zs,
            zf_clip=clip, seed=seed, reset_take=cu(np.asarray(g)[:, None], torch.int32),
            reset_start=cu(np.full((E, 1), fm), torch.int32), want_next=False, want_raw=False, ctx=ctx, eval_mode=2,
            fix_len=cu(test_len[g], torch.int32), state_pred=sp_d, want_traj=True, value_weights=vw, vctx=vctx, value_stat=vs)
        qpos = out['qpos_traj'].view(E, T, model.nq).cpu().numpy()
        qvel = out['qvel_traj'].view(E, T, model.nv).cpu().numpy()
        rew, val = out['rewards'].view(E, T).cpu().numpy(), out['values'].view(E, T).cpu().numpy()
        num_reset += int(round(float(out['logger'][lib.LOG['NUM_FAILSAFE_RESETS']].item())))
        for i, e in enumerate(g):
            take, n = env.expert_list[e], int(test_len[e])
            results['traj_pred'][take], results['vel_pred'][take] = qpos[i, :n].copy(), qvel[i, :n].copy()
            results['traj_orig'][take] = model.rows_host[off[e] + fm:off[e] + fm + n, X['QPOS']:X['QPOS'] + model.nq].copy()
            info['rewards'][take], info['values'][take], info['test_len'][take] = rew[i, :n].copy(), val[i, :n].copy(), n
    return results, {'algo': algo, 'num_reset': num_reset}, info


def _quat_mul(q1, q0):
    """utils/transformation.py:1379-1391 quaternion_multiply, (w, x, y, z)"""
    w0, x0, y0, z0 = q0
    w1, x1, y1, z1 = q1
    return np.array([-x1 * x0 - y1 * y0 - z1 * z0 + w1 * w0, x1 * w0 + y1 * z0 - z1 * y0 + w1 * x0,
                     -x1 * z0 + y1 * w0 + z1 * x0 + w1 * y0, x1 * y0 - y1 * x0 + z1 * w0 + w1 * z0])


def _heading_q(q):
    hq = np.array([q[0], 0.0, 0.0, q[3]])
    return hq / np.linalg.norm(hq)


def sync_traj(qpos_traj, qvel_traj, ref_qpos):
    """ego_pose/utils/tools.py:18-32: rotate / translate a predicted trajectory so that its first frame has the heading
    and xy position of ``ref_qpos`` (both heading quaternions are rotations about z)"""
    h0 = _heading_q(qpos_traj[0, 3:7])
    rel = _quat_mul(_heading_q(ref_qpos[3:7]), np.array([h0[0], -h0[1], -h0[2], -h0[3]]))
    c, s = rel[0] * rel[0] - rel[3] * rel[3], 2.0 * rel[0] * rel[3]
    start = np.array([qpos_traj[0, 0], qpos_traj[0, 1], ref_qpos[2]])
    qp, qv = qpos_traj.copy(), qvel_traj.copy()
    d = qpos_traj[:, :3] - start
    qp[:, 0] = c * d[:, 0] - s * d[:, 1] + ref_qpos[0]
    qp[:, 1] = s * d[:, 0] + c * d[:, 1] + ref_qpos[1]
    for i in range(qp.shape[0]):
        qp[i, 3:7] = _quat_mul(rel, qpos_traj[i, 3:7])
    qv[:, 0] = c * qvel_traj[:, 0] - s * qvel_traj[:, 1]
    qv[:, 1] = s * qvel_traj[:, 0] + c * qvel_traj[:, 1]
    return qp, qv


def forecast_window_init(expert_qpos, em_traj, em_vel, start, fm, test_len, em_offset):
    """ego_forecast_eval.py:107-136 without --gt-init: the window starts from the ego-mimic prediction of frame
    ``start`` (synced to the expert pose fm frames earlier when the prediction reaches that far back).
    -> (qpos0, qvel0, past [fm, nq]): initial simulator state and the fm 'past' rows of traj_pred"""
    lo = max(0, start - fm - em_offset)
    sp, vp = em_traj[lo:start + test_len - em_offset], em_vel[lo:start + test_len - em_offset]
    miss = fm + test_len - sp.shape[0]
    if start - fm - em_offset >= 0:
        sp, vp = sync_traj(sp, vp, expert_qpos[start - fm])
    ind = fm - miss
    past = np.stack([expert_qpos[start - fm + j] if j < miss else sp[j - miss] for j in range(fm)])
    return sp[ind].copy(), vp[ind].copy(), past


def eval_forecast(env, policy_net, policy_vs_net, running_state=None, test_len=None, show_noise=False, seed=1, em_res=None,
                  em_fr_margin=None):
    """ego_pose/ego_forecast_eval.py:95-204 ('save' mode with --gt-init): every take is cut into windows starting every
    ``fr_margin`` frames (:188-196); each window is one environment of ONE kernel launch, started from the expert state
    of its first frame and rolled out ``test_len`` steps under the mean action.  A fall does not end a window (the
    script only logs it, :171-176).  -> (results, meta): results['traj_pred' | 'traj_orig'][take] =
    [n_windows, fr_margin + test_len, nq], the first fr_margin rows being the past (:126-136).
    ``em_res`` = the ``results`` dict of an ego-mimic evaluation (``eval_takes`` / ego_mimic_eval.py:186) and
    ``em_fr_margin`` its fr_margin: windows then start from the ego-mimic prediction (no --gt-init, :107-121)."""
    from .nets import VideoForecastNet
    model, cfg = env.kernel, env.cfg
    fm = int(cfg.fr_margin)
    T = int(test_len or cfg.env_episode_len)
    off = np.asarray(model.take_off, dtype=np.int64)
    lens = np.diff(off)
    win = [(k, s0) for k in range(len(lens)) for s0 in range(fm, int(lens[k]) - T + 1, fm)]
    if not win:
        raise ValueError('no take holds fr_margin + test_len frames')
    E = len(win)
    dev = policy_net.action_mean.weight.device
    L = policy_net.net.affine_layers
    w = dict(W1=L[0].weight.data, b1=L[0].bias.data, W2=L[1].weight.data, b2=L[1].bias.data,
             W3=policy_net.action_mean.weight.data, b3=policy_net.action_mean.bias.data,
             log_std=policy_net.action_log_std.data.view(-1))
    w = {k: t.contiguous() for k, t in w.items()}
    cu = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)  # noqa: E731
    extra = {}
    if isinstance(policy_vs_net, VideoForecastNet):
        # test-mode v_out of every window (video_forecast_net.py:58-59): causal LSTM over the fr_margin frames before
        # its start; one constant row per window, rows ordered like ``win``; the state LSTM runs inside the kernel
        p = policy_vs_net._p()
        rows, win_off = [], [0]
        with torch.no_grad():
            for k in range(len(lens)):
                f = torch.as_tensor(env.cnn_feat[k], dtype=p.dtype, device=p.device)
                nst = max(0, int(lens[k]) - T - fm + 1)                       # starts fm .. len - T, every frame
                if nst:
                    idx = torch.arange(fm, device=p.device)[:, None] + torch.arange(nst, device=p.device)[None, :]
                    rows.append(policy_vs_net.forward_v_net(f[idx])[-1])
                win_off.append(win_off[-1] + nst)
        extra = dict(ctx=torch.cat(rows).contiguous(), win_off=torch.tensor(win_off, dtype=torch.int32, device=dev),
                     ctx_const=True, snet=policy_vs_net.snet_packed())
    zm = zs = None
    clip = 0.0
    if running_state is not None:
        zm, zs = cu(running_state.rs.mean, torch.float64), cu(running_state.rs.std, torch.float64)
        clip = running_state.clip or 0.0
    X = lib.X
    eq = model.rows_host[:, X['QPOS']:X['QPOS'] + model.nq]
    pasts = None
    if em_res is not None:
        if em_fr_margin is None:
            raise ValueError('em_fr_margin (fr_margin of the ego-mimic evaluation) is required with em_res')
        q0, v0, pasts = [], [], []
        for k, s0 in win:
            take = env.expert_list[k]
            a, b, c = forecast_window_init(eq[off[k]:off[k + 1]], np.asarray(em_res['traj_pred'][take]),
                                           np.asarray(em_res['vel_pred'][take]), s0, fm, T, int(em_fr_margin))
            q0.append(a), v0.append(b), pasts.append(c)
        extra.update(init_qpos=cu(np.stack(q0), torch.float64), init_qvel=cu(np.stack(v0), torch.float64))
    out = model.rollout(
        w, E, T, episode_len=T, fr_margin=fm, fix_head_lb=-1e30, mean_action=not show_noise, zf_mean=zm, zf_std=zs,
        zf_clip=clip, seed=seed, reset_take=cu([[k] for k, _ in win], torch.int32),
        reset_start=cu([[s0] for _, s0 in win], torch.int32), want_next=False, want_raw=False, want_traj=True, **extra)
    qpos = out['qpos_traj'].view(E, T, model.nq).cpu().numpy()
    results = {'traj_pred': {}, 'traj_orig': {}}
    for k, take in enumerate(env.expert_list):
        ids = [e for e, (kk, _) in enumerate(win) if kk == k]
        if not ids:
            continue
        past = [eq[off[k] + win[e][1] - fm:off[k] + win[e][1]] if pasts is None else pasts[e] for e in ids]
        results['traj_pred'][take] = np.stack([np.concatenate([pa, qpos[e]]) for pa, e in zip(past, ids)])
        results['traj_orig'][take] = np.stack([eq[off[k] + win[e][1] - fm:off[k] + win[e][1] + T] for e in ids])
    return results, {'algo': 'ego_forecast'}


def save_results(results, meta, path):
    """the pickle ego_mimic_eval.py:192 writes and eval_pose.py:31 reads"""
    with open(path, 'wb') as f:
        pickle.dump((results, meta), f)
