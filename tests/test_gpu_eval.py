"""Evaluation roll-outs on the fused kernel (eval_mode: in-place fail-safe state replacement, per-environment episode
length, recorded simulator states) vs the reference golden (tests/golden/eval_traj.npz, reference env code driven in
ego_mimic_eval.py's call order) and vs the CPU oracle's eval loop through the public egopose_b200.evaluate API."""
import pickle

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')

from oracle import cphys, evalloop  # noqa: E402
import helpers  # noqa: E402
from test_oracle_eval import setup_eval  # noqa: E402


def cu(a, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device='cuda')


def test_eval_rollout_matches_reference_golden(golden):
    g = golden('eval_traj')
    orc, _, fm = setup_eval(g)
    lens = [int(x) for x in g['lens']]
    cnn = np.concatenate([g['take%d.cnn' % i] for i in range(2)])
    sp = np.concatenate([g['take%d.state_pred' % i] for i in range(2)])
    model = helpers.make_model()
    model.upload_experts(orc._keep['x_rows'], orc._keep['x_off'], orc._keep['x_lb'], cnn)
    w = dict(W1=g['policy.net.affine_layers.0.weight'], b1=g['policy.net.affine_layers.0.bias'],
             W2=g['policy.net.affine_layers.1.weight'], b2=g['policy.net.affine_layers.1.bias'],
             W3=g['policy.action_mean.weight'], b3=g['policy.action_mean.bias'], log_std=g['policy.action_log_std'].ravel())
    wd = {k: cu(v) for k, v in w.items()}
    # the two takes use different fail lines in the golden: one launch per take (the kernel's fix_head_lb is global)
    for ti in range(2):
        pre = 'take%d.' % ti
        n = lens[ti] - 2 * fm
        out = model.rollout(wd, 1, n, episode_len=n, fr_margin=fm, fix_head_lb=float(g[pre + 'head_lb']), mean_action=True,
                            zf_mean=cu(g['zf_mean']), zf_std=cu(g['zf_std']), zf_clip=5.0,
                            reset_take=cu([[ti]], torch.int32), reset_start=cu([[fm]], torch.int32), eval_mode=True,
                            fix_len=cu([n], torch.int32), state_pred=cu(sp), want_traj=True)
        torch.cuda.synchronize()
        nrec = g[pre + 'traj_pred'].shape[0]
        assert nrec == n
        assert int(out['logger'][14].item()) == int(g[pre + 'num_reset']) > 3
        # every replacement re-anchors the state, so errors do not accumulate over the take
        assert np.allclose(out['qpos_traj'].cpu().numpy(), g[pre + 'traj_pred'], rtol=1e-7, atol=1e-8)
        assert np.allclose(out['qvel_traj'].cpu().numpy(), g[pre + 'vel_pred'], rtol=1e-6, atol=1e-6)
        assert np.allclose(out['states'].cpu().numpy(), g[pre + 'states'], rtol=1e-6, atol=1e-6)
        assert np.allclose(out['actions'].cpu().numpy(), g[pre + 'actions'], rtol=1e-6, atol=1e-8)
        assert np.allclose(out['rewards'].cpu().numpy(), g[pre + 'rewards'], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize('fail_safe', ['naivefs', 'none'])
def test_eval_takes_api_matches_oracle(fail_safe, tmp_path):
    from egopose_b200 import evaluate
    from egopose_b200.config import Config
    from egopose_b200.env import HumanoidEnv
    from egopose_b200.nets import MLP, FrameContext, PolicyGaussian
    from egopose_b200.zfilter import ZFilter
    torch.set_default_dtype(torch.float64)
    torch.manual_seed(3)
    cfg = Config('subject_03')
    cfg.fr_margin = 4
    env = HumanoidEnv(cfg, device=0)
    lens, ctxd = [30, 22, 41], 8
    takes = [cphys.synthetic_takes(cphys.Oracle().md, 1, L, seed=50 + i)[0] for i, L in enumerate(lens)]
    rng = np.random.RandomState(8)
    cnn = [rng.randn(L, ctxd) for L in lens]
    names = ['t%d' % i for i in range(3)]
    env.set_expert_qpos(names, takes, cnn)
    S, nu = env.obs_dim, env.md.nu
    pol = PolicyGaussian(MLP(S + ctxd, (32, 16), 'relu'), nu, log_std=-2.3, fix_std=True).cuda()
    rs = ZFilter((S,), clip=5)
    for _ in range(30):
        rs(0.5 * rng.randn(S))
    head_lb = 0.75                       # above the default 0.3 so that the short test takes do fall below it
    results, meta, info = evaluate.eval_takes(env, pol, FrameContext(ctxd), rs, fail_safe=fail_safe, fix_head_lb=head_lb)
    # ---- oracle: same takes (hands zeroed as set_expert_qpos / gen_expert.py:38-39 do), same tables
    orc = cphys.Oracle()
    orc.cfg.fr_margin = 4
    orc.cfg.fix_head_lb = head_lb if fail_safe == 'naivefs' else -1e30
    tz = []
    for q in takes:
        q = q.copy()
        for hand in ('LeftHand', 'RightHand'):
            a, b = env.body_qposaddr[hand]
            q[:, a:b] = 0.0
        tz.append(q)
    orc.make_expert(tz)
    sd = {k: v.detach().cpu().numpy() for k, v in pol.state_dict().items()}
    opol = orc.make_policy(sd['net.affine_layers.0.weight'], sd['net.affine_layers.0.bias'], sd['net.affine_layers.1.weight'],
                           sd['net.affine_layers.1.bias'], sd['action_mean.weight'], sd['action_mean.bias'], sd['action_log_std'])
    sp_all = evaluate.expert_obs_table(env.kernel)
    off = np.asarray(env.kernel.take_off)
    total = 0
    for ti, name in enumerate(names):
        ref = evalloop.eval_take(orc, opol, ti, 4, lens[ti] - 8, sp_all[off[ti]:off[ti + 1]], ctx=cnn[ti], zf_mean=rs.rs.mean,
                                 zf_std=rs.rs.std, zf_clip=5.0, fail_safe=fail_safe)
        total += ref['num_reset']
        assert results['traj_pred'][name].shape == (lens[ti] - 8, 59)
        if fail_safe == 'naivefs':
            assert np.allclose(results['traj_pred'][name], ref['traj_pred'], rtol=1e-7, atol=1e-8)
            assert np.allclose(results['vel_pred'][name], ref['vel_pred'], rtol=1e-6, atol=1e-6)
        else:       # free fall for the whole take: chaotic growth of rounding differences, compare the early part tightly
            assert np.allclose(results['traj_pred'][name][:12], ref['traj_pred'][:12], rtol=1e-6, atol=1e-7)
            assert np.all(np.isfinite(results['traj_pred'][name]))
        assert np.array_equal(results['traj_orig'][name], tz[ti][4:lens[ti] - 4])
    assert meta == {'algo': 'ego_mimic', 'num_reset': total}
    assert (total > 3) == (fail_safe == 'naivefs')
    path = tmp_path / 'iter_0000_test_naivefs.p'
    evaluate.save_results(results, meta, str(path))
    r2, m2 = pickle.load(open(path, 'rb'))      # the (results, meta) pair eval_pose.py:31 unpacks
    assert set(r2) == {'traj_pred', 'traj_orig', 'vel_pred'} and m2['num_reset'] == total
    env.close()


def test_eval_forecast_windows():
    """ego_forecast_eval.py 'save' mode with --gt-init: windows every fr_margin frames, one environment each"""
    from egopose_b200 import evaluate
    from egopose_b200.config import Config
    from egopose_b200.env import HumanoidEnv
    from egopose_b200.nets import MLP, FrameContext, PolicyGaussian, VideoForecastNet
    from egopose_b200.zfilter import ZFilter
    torch.set_default_dtype(torch.float64)
    torch.manual_seed(5)
    cfg = Config('subject_03', task='egoforecast')
    cfg.fr_margin, cfg.env_episode_len = 6, 9
    env = HumanoidEnv(cfg, device=0)
    lens, ctxd, T, fm = [33, 27], 8, 9, 6
    md = cphys.Oracle().md
    takes = [cphys.synthetic_takes(md, 1, L, seed=70 + i)[0] for i, L in enumerate(lens)]
    rng = np.random.RandomState(9)
    cnn = [rng.randn(L, ctxd) for L in lens]
    env.set_expert_qpos(['a', 'b'], takes, cnn)
    S, nu = env.obs_dim, env.md.nu
    rs = ZFilter((S,), clip=5)
    for _ in range(30):
        rs(0.5 * rng.randn(S))
    # ---- per-frame context (identity video net) vs the oracle loop
    pol = PolicyGaussian(MLP(S + ctxd, (32, 16), 'relu'), nu, log_std=-2.3, fix_std=True).cuda()
    results, meta = evaluate.eval_forecast(env, pol, FrameContext(ctxd), rs)
    assert meta == {'algo': 'ego_forecast'}
    orc = cphys.Oracle()
    orc.cfg.fr_margin = fm
    orc.cfg.fix_head_lb = -1e30
    tz = []
    for q in takes:
        q = q.copy()
        for hand in ('LeftHand', 'RightHand'):
            a, b = env.body_qposaddr[hand]
            q[:, a:b] = 0.0
        tz.append(q)
    orc.make_expert(tz)
    sd = {k: v.detach().cpu().numpy() for k, v in pol.state_dict().items()}
    opol = orc.make_policy(sd['net.affine_layers.0.weight'], sd['net.affine_layers.0.bias'], sd['net.affine_layers.1.weight'],
                           sd['net.affine_layers.1.bias'], sd['action_mean.weight'], sd['action_mean.bias'], sd['action_log_std'])
    for k, name in enumerate(['a', 'b']):
        starts = list(range(fm, lens[k] - T + 1, fm))
        assert results['traj_pred'][name].shape == (len(starts), fm + T, 59) == results['traj_orig'][name].shape
        assert starts[-1] + T == lens[k] or starts[-1] + T + fm > lens[k]
        for wi, s0 in enumerate(starts):
            ref = evalloop.forecast_window(orc, opol, k, s0, T, ctx=cnn[k], zf_mean=rs.rs.mean, zf_std=rs.rs.std)
            assert np.array_equal(results['traj_pred'][name][wi, :fm], tz[k][s0 - fm:s0])
            assert np.allclose(results['traj_pred'][name][wi, fm:], ref, rtol=1e-6, atol=1e-7)
            assert np.array_equal(results['traj_orig'][name][wi], tz[k][s0 - fm:s0 + T])
    # ---- VideoForecastNet: constant per-window context + in-kernel state LSTM (numerics of that path are pinned by
    #      test_gpu_forecast.py); here: the plumbing of the evaluation windows
    pvs = VideoForecastNet(ctxd, S, 16, fm, 'lstm', None, 20, 'lstm').cuda()
    pol2 = PolicyGaussian(MLP(pvs.out_dim, (32, 16), 'relu'), nu, log_std=-2.3, fix_std=True).cuda()
    r2, _ = evaluate.eval_forecast(env, pol2, pvs, rs)
    for k, name in enumerate(['a', 'b']):
        assert r2['traj_pred'][name].shape == results['traj_pred'][name].shape
        assert np.all(np.isfinite(r2['traj_pred'][name]))
        for wi, s0 in enumerate(range(fm, lens[k] - T + 1, fm)):       # first simulated frame = expert state of the window start
            assert np.allclose(r2['traj_pred'][name][wi, fm], tz[k][s0], atol=1e-14)
    env.close()


def test_eval_forecast_from_ego_mimic_result_matches_reference_golden(golden):
    """ego_forecast_eval.py without --gt-init: window start states from an ego-mimic result file (sync_traj, missing
    past), vs the golden produced by the reference's own env / sync_traj in the script's call order"""
    from egopose_b200 import evaluate
    from egopose_b200.config import Config
    from egopose_b200.env import HumanoidEnv
    from egopose_b200.nets import MLP, FrameContext, PolicyGaussian
    from egopose_b200.zfilter import ZFilter
    torch.set_default_dtype(torch.float64)
    g = golden('eval_forecast')
    fm, T, emo = int(g['fr_margin']), int(g['test_len']), int(g['em_offset'])
    cfg = Config('subject_03', task='egoforecast')
    cfg.fr_margin, cfg.env_episode_len = fm, T
    env = HumanoidEnv(cfg, device=0)
    env.set_expert_qpos(['t'], [g['qpos']], [g['cnn']])
    S, nu, ctxd = env.obs_dim, env.md.nu, g['cnn'].shape[1]
    pol = PolicyGaussian(MLP(S + ctxd, (32, 16), 'relu'), nu, log_std=-2.3, fix_std=True)
    pol.load_state_dict({k[len('policy.'):]: torch.from_numpy(g[k]) for k in g.files if k.startswith('policy.')})
    pol = pol.cuda()
    rs = ZFilter((S,), clip=5)
    rs.rs._n, rs.rs._M = 50, g['zf_mean'].copy()
    rs.rs._S = (g['zf_std'] ** 2) * 49
    assert np.allclose(rs.rs.std, g['zf_std']) and np.allclose(rs.rs.mean, g['zf_mean'])
    # the synthetic take keeps the hand joints at zero, so set_expert_qpos' hand zeroing leaves it unchanged
    r_gt, _ = evaluate.eval_forecast(env, pol, FrameContext(ctxd), rs)
    assert np.allclose(r_gt['traj_pred']['t'], g['traj_pred_gt'], rtol=1e-6, atol=1e-7)
    em_res = {'traj_pred': {'t': g['em_traj']}, 'vel_pred': {'t': g['em_vel']}}
    r_em, meta = evaluate.eval_forecast(env, pol, FrameContext(ctxd), rs, em_res=em_res, em_fr_margin=emo)
    assert meta == {'algo': 'ego_forecast'}
    assert r_em['traj_pred']['t'].shape == g['traj_pred_em'].shape
    assert np.allclose(r_em['traj_pred']['t'][:, :fm], g['traj_pred_em'][:, :fm], rtol=1e-12, atol=1e-12)
    assert np.allclose(r_em['traj_pred']['t'][:, fm:], g['traj_pred_em'][:, fm:], rtol=1e-6, atol=1e-7)
    assert not np.allclose(r_em['traj_pred']['t'][:, fm], r_gt['traj_pred']['t'][:, fm], atol=1e-3)
    env.close()


def test_eval_takes_valuefs_matches_reference_golden(golden):
    """fail_safe='valuefs' (the script's default): value net inside the kernel's step loop, running mean carried across
    the per-take launches; vs the golden produced by the reference's env / Value / RunningStat in script order"""
    from egopose_b200 import evaluate
    from egopose_b200.config import Config
    from egopose_b200.env import HumanoidEnv
    from egopose_b200.nets import MLP, FrameContext, PolicyGaussian, Value
    from egopose_b200.zfilter import ZFilter
    torch.set_default_dtype(torch.float64)
    g = golden('eval_traj')
    fm = int(g['fr_margin'])
    cfg = Config('subject_03')
    cfg.fr_margin = fm
    env = HumanoidEnv(cfg, device=0)
    takes = [g['take%d.qpos' % i] for i in range(2)]
    env.set_expert_qpos(['a', 'b'], takes, [g['take%d.cnn' % i] for i in range(2)])
    S, nu, ctxd = env.obs_dim, env.md.nu, g['take0.cnn'].shape[1]
    pol = PolicyGaussian(MLP(S + ctxd, (32, 16), 'relu'), nu, log_std=-2.3, fix_std=True)
    pol.load_state_dict({k[len('policy.'):]: torch.from_numpy(g[k]) for k in g.files if k.startswith('policy.')})
    val = Value(MLP(S + ctxd, (32, 16), 'relu'))
    val.load_state_dict({k[len('value.'):]: torch.from_numpy(g[k]) for k in g.files if k.startswith('value.')})
    pol, val = pol.cuda(), val.cuda()
    rs = ZFilter((S,), clip=5)
    rs.rs._n, rs.rs._M = 50, g['zf_mean'].copy()
    rs.rs._S = (g['zf_std'] ** 2) * 49

    # the value net has its own per-frame context table (value_vs_net of the script)
    vtab = np.concatenate([g['vfs.take%d.vcnn' % i] for i in range(2)])
    sp = [g['take%d.state_pred' % i] for i in range(2)]
    results, meta, info = evaluate.eval_takes(env, pol, FrameContext(ctxd), rs, fail_safe='valuefs', value_net=val,
                                              value_vs_net=vtab, sequential=True, state_pred=sp)
    batched = evaluate.eval_takes(env, pol, FrameContext(ctxd), rs, fail_safe='valuefs', value_net=val, value_vs_net=vtab,
                                  sequential=False, state_pred=sp)
    total = 0
    for i, name in enumerate(['a', 'b']):
        vpre = 'vfs.take%d.' % i
        total += int(g[vpre + 'num_reset'])
        assert np.allclose(info['values'][name], g[vpre + 'values'], rtol=1e-6, atol=1e-7)
        assert np.allclose(results['traj_pred'][name], g[vpre + 'traj_pred'], rtol=1e-6, atol=1e-7)
        assert np.allclose(results['vel_pred'][name], g[vpre + 'vel_pred'], rtol=1e-5, atol=1e-5)
    assert meta == {'algo': 'ego_mimic', 'num_reset': total} and total >= 6
    # one launch for all takes: the first take sees the same statistic, later takes start their own
    assert np.allclose(batched[0]['traj_pred']['a'], results['traj_pred']['a'], rtol=1e-9, atol=1e-10)
    assert batched[0]['traj_pred']['b'].shape == results['traj_pred']['b'].shape
    env.close()


def test_eval_script_synthetic(tmp_path):
    """examples/eval_egomimic.py: the command line of ego_mimic_eval.py end to end (ragged synthetic takes, reference
    layer sizes [300, 200], both fail-safes), result file names and pickle layout of ego_mimic_eval.py:186-192"""
    import os
    import sys
    sys.path.insert(0, os.path.join(helpers.ROOT, 'examples'))
    import eval_egomimic
    p = eval_egomimic.main(['--synthetic', '--takes', '2', '--len', '36', '--fail-safe', 'valuefs', '--out', str(tmp_path)])
    assert p.endswith('iter_0000_test.p')
    results, meta = pickle.load(open(p, 'rb'))
    assert set(results) == {'traj_pred', 'traj_orig', 'vel_pred'} and meta['algo'] == 'ego_mimic' and meta['num_reset'] >= 0
    assert results['traj_pred']['take_0'].shape == (16, 59) and results['traj_pred']['take_1'].shape == (23, 59)
    assert results['vel_pred']['take_1'].shape == (23, 58) and np.all(np.isfinite(results['traj_pred']['take_1']))
    assert np.allclose(results['traj_pred']['take_0'][0], results['traj_orig']['take_0'][0], atol=1e-12)   # starts on the expert
    p2 = eval_egomimic.main(['--synthetic', '--takes', '2', '--len', '36', '--fail-safe', 'naivefs', '--out', str(tmp_path)])
    assert p2.endswith('iter_0000_test_naivefs.p')
    r2, m2 = pickle.load(open(p2, 'rb'))
    assert r2['traj_pred']['take_1'].shape == (23, 59) and m2['num_reset'] >= 0 and np.all(np.isfinite(r2['traj_pred']['take_1']))
