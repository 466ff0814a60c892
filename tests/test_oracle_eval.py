"""Oracle evaluation roll-out (oracle/evalloop.py: mean action, frozen ZFilter, 'naivefs' in-place state replacement,
recorded simulator states) vs tests/golden/eval_traj.npz, produced by the reference's unmodified HumanoidEnv,
align_human_state, PolicyGaussian and ZFilter driven in the call order of ego_pose/ego_mimic_eval.py:93-177."""
import numpy as np

from oracle import cphys, evalloop


def setup_eval(g):
    fm = int(g['fr_margin'])
    orc = cphys.Oracle()
    orc.cfg.fr_margin = fm
    takes = [g['take%d.qpos' % i] for i in range(2)]
    orc.make_expert(takes)
    pol = orc.make_policy(g['policy.net.affine_layers.0.weight'], g['policy.net.affine_layers.0.bias'],
                          g['policy.net.affine_layers.1.weight'], g['policy.net.affine_layers.1.bias'],
                          g['policy.action_mean.weight'], g['policy.action_mean.bias'], g['policy.action_log_std'])
    return orc, pol, fm


def test_eval_rollout_matches_reference(golden):
    g = golden('eval_traj')
    orc, pol, fm = setup_eval(g)
    for ti in range(2):
        pre = 'take%d.' % ti
        L = g[pre + 'qpos'].shape[0]
        orc.cfg.fix_head_lb = float(g[pre + 'head_lb'])
        out = evalloop.eval_take(orc, pol, ti, fm, L - 2 * fm, g[pre + 'state_pred'], ctx=g[pre + 'cnn'], zf_mean=g['zf_mean'],
                                 zf_std=g['zf_std'], zf_clip=5.0, fail_safe='naivefs')
        assert out['num_reset'] == int(g[pre + 'num_reset']) and out['num_reset'] > 3
        assert out['traj_pred'].shape == g[pre + 'traj_pred'].shape
        assert np.allclose(out['traj_pred'], g[pre + 'traj_pred'], rtol=1e-9, atol=1e-10)
        assert np.allclose(out['vel_pred'], g[pre + 'vel_pred'], rtol=1e-8, atol=1e-8)
        assert np.allclose(out['states'], g[pre + 'states'], rtol=1e-8, atol=1e-8)
        assert np.allclose(out['actions'], g[pre + 'actions'], rtol=1e-9, atol=1e-10)
        assert np.allclose(out['rewards'], g[pre + 'rewards'], rtol=1e-8, atol=1e-10)


def test_reset_env_state_keeps_root_xy_and_heading(golden):
    g = golden('eval_traj')
    orc, _, fm = setup_eval(g)
    env = cphys.EoEnv()
    orc.env_reset(env, 0, fm)
    before = np.array(env.d.qpos[:orc.nq])
    sp = g['take0.state_pred'][fm + 3]
    obs = evalloop.reset_env_state(orc, env, sp, orc.nq)
    after = np.array(env.d.qpos[:orc.nq])
    assert np.array_equal(after[:2], before[:2]) and after[2] == sp[0]
    assert np.allclose(evalloop.heading_q(after[3:7]), evalloop.heading_q(before[3:7]), atol=1e-12)
    # the observation of the replaced state is the prediction itself (de-heading undoes the alignment)
    assert np.allclose(obs, sp, atol=1e-12)


def setup_forecast(g):
    fm = int(g['fr_margin'])
    orc = cphys.Oracle()
    orc.cfg.fr_margin = fm
    orc.cfg.fix_head_lb = -1e30
    orc.make_expert([g['qpos']])
    pol = orc.make_policy(g['policy.net.affine_layers.0.weight'], g['policy.net.affine_layers.0.bias'],
                          g['policy.net.affine_layers.1.weight'], g['policy.net.affine_layers.1.bias'],
                          g['policy.action_mean.weight'], g['policy.action_mean.bias'], g['policy.action_log_std'])
    return orc, pol, fm


def test_forecast_windows_match_reference(golden):
    """ego_forecast_eval.py windows with --gt-init and from an ego-mimic result (sync_traj, missing-past handling)"""
    g = golden('eval_forecast')
    orc, pol, fm = setup_forecast(g)
    T, emo = int(g['test_len']), int(g['em_offset'])
    assert int(g['starts'][0]) - fm - emo < 0 <= int(g['starts'][1]) - fm - emo      # both branches of :111-115
    for wi, s0 in enumerate(int(x) for x in g['starts']):
        ref = evalloop.forecast_window(orc, pol, 0, s0, T, ctx=g['cnn'], zf_mean=g['zf_mean'], zf_std=g['zf_std'])
        assert np.allclose(ref, g['traj_pred_gt'][wi, fm:], rtol=1e-9, atol=1e-10)
        assert np.array_equal(g['traj_pred_gt'][wi, :fm], g['qpos'][s0 - fm:s0])
        q0, v0, past = evalloop.forecast_init(g['qpos'], g['em_traj'], g['em_vel'], s0, fm, T, emo)
        ref = evalloop.forecast_window(orc, pol, 0, s0, T, ctx=g['cnn'], zf_mean=g['zf_mean'], zf_std=g['zf_std'], init=(q0, v0))
        assert np.allclose(past, g['traj_pred_em'][wi, :fm], rtol=1e-12, atol=1e-12)
        assert np.allclose(ref, g['traj_pred_em'][wi, fm:], rtol=1e-9, atol=1e-10)


def value_policy(orc, g):
    return orc.make_policy(g['value.net.affine_layers.0.weight'], g['value.net.affine_layers.0.bias'],
                           g['value.net.affine_layers.1.weight'], g['value.net.affine_layers.1.bias'],
                           g['value.value_head.weight'], g['value.value_head.bias'], np.zeros(1))


def test_eval_rollout_valuefs_matches_reference(golden):
    """the script's default fail-safe: value net in the loop, one running mean shared by the takes in list order"""
    g = golden('eval_traj')
    orc, pol, fm = setup_eval(g)
    orc2 = cphys.Oracle()                       # second handle: make_policy keeps one weight set alive per Oracle
    vpol = value_policy(orc2, g)
    orc.cfg.fix_head_lb = float('nan')
    stat = [0, 0.0]
    for ti in range(2):
        pre, vpre = 'take%d.' % ti, 'vfs.take%d.' % ti
        L = g[pre + 'qpos'].shape[0]
        out = evalloop.eval_take(orc, pol, ti, fm, L - 2 * fm, g[pre + 'state_pred'], ctx=g[pre + 'cnn'], zf_mean=g['zf_mean'],
                                 zf_std=g['zf_std'], zf_clip=5.0, fail_safe='valuefs', value_policy=vpol, vctx=g[vpre + 'vcnn'],
                                 value_stat=stat)
        assert out['num_reset'] == int(g[vpre + 'num_reset']) >= 3
        assert np.allclose(out['values'], g[vpre + 'values'], rtol=1e-8, atol=1e-9)
        assert np.allclose(out['traj_pred'], g[vpre + 'traj_pred'], rtol=1e-8, atol=1e-9)
        assert np.allclose(out['vel_pred'], g[vpre + 'vel_pred'], rtol=1e-7, atol=1e-7)
    assert stat[0] == int(g['vfs.value_stat'][0]) and abs(stat[1] - g['vfs.value_stat'][1]) < 1e-10
