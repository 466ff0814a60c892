"""Host-side mirrors (no GPU): ZFilter vs the reference golden, batched moment merge, Config constants and
schedules, LoggerRL, TrajBatch laziness, flat parameter aliasing, compat import paths, MJCF compiler."""
import os
import pickle
import sys

import numpy as np
import pytest
import torch

import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_zfilter_mirror_matches_reference_golden(golden):
    from egopose_b200.zfilter import ZFilter
    g = golden('zfilter')
    zf = ZFilter((9,), clip=5)
    ys = np.stack([zf(x) for x in g['xs']])
    assert np.allclose(ys, g['ys'], rtol=1e-13, atol=1e-13)
    assert np.allclose(np.stack([zf(x, update=False) for x in g['xs'][:5]]), g['frozen'], rtol=1e-13, atol=1e-13)
    assert zf.rs.n == int(g['n']) and np.allclose(zf.rs.mean, g['mean']) and np.allclose(zf.rs.std, g['std'])
    # batched merge == sequential pushes
    zb = ZFilter((9,), clip=5)
    for lo, hi in ((0, 7), (7, 8), (8, 40)):
        xs = g['xs'][lo:hi]
        shift = zb.rs.mean.copy() if zb.rs.n else np.zeros(9)
        zb.rs.merge_moments(hi - lo, (xs - shift).sum(0), ((xs - shift) ** 2).sum(0), shift)
    assert zb.rs.n == 40 and np.allclose(zb.rs.mean, g['mean'], rtol=1e-12) and np.allclose(zb.rs._S, g['S'], rtol=1e-10)
    assert isinstance(pickle.loads(pickle.dumps(zb)), ZFilter)


def test_config_constants_and_schedules():
    from egopose_b200.config import Config
    cfg = Config('subject_03')
    assert cfg.policy_hsize == [300, 200] and cfg.env_episode_len == 200 and cfg.fr_margin == 10
    assert cfg.jkp.shape == (52,) and cfg.jkp[0] == 500.0 and cfg.jkd[0] == 50.0      # jkp_multiplier 0.5 scales both
    assert abs(cfg.a_ref[22] - np.deg2rad(-80.0)) < 1e-15 and cfg.torque_lim[24] == 60.0
    assert cfg.reward_weights['w_v'] == 0.0 and cfg.b_diffw.shape == (20,)
    cfg.adp_iter_cp = np.array([0, 100, 200])
    cfg.adp_noise_rate_cp = np.array([1.0, 0.5, 0.5])
    cfg.adp_log_std_cp = np.array([-2.3, -2.3, -3.0])
    cfg.adp_policy_lr_cp = np.array([5e-5, 5e-5, 1e-5])
    cfg.update_adaptive_params(50)
    assert abs(cfg.adp_noise_rate - 0.75) < 1e-15
    cfg.update_adaptive_params(150)
    assert abs(cfg.adp_log_std + 2.65) < 1e-12 and abs(cfg.adp_policy_lr - 3e-5) < 1e-18
    cfg.update_adaptive_params(500)
    assert cfg.adp_log_std == -3.0


def test_logger_from_device_and_merge():
    from egopose_b200.lib import LOG
    from egopose_b200.logger_rl import LoggerRL
    v = np.zeros(LOG['SIZE'])
    v[LOG['NUM_STEPS']], v[LOG['NUM_EPISODES']], v[LOG['TOTAL_REWARD']], v[LOG['TOTAL_C_REWARD']] = 100, 20, 100, 37.5
    v[LOG['MIN_C_REWARD']], v[LOG['MAX_C_REWARD']] = 0.01, 0.9
    v[LOG['C_INFO']:LOG['C_INFO'] + 5] = [10, 20, 30, 40, 50]
    lg = LoggerRL.from_device(v)
    assert lg.avg_c_reward == 0.375 and lg.avg_episode_reward == 5.0 and np.allclose(lg.avg_c_info, [.1, .2, .3, .4, .5])
    m = LoggerRL.merge([lg, lg])
    assert m.num_steps == 200 and m.avg_c_reward == 0.375 and m.min_c_reward == 0.01


def test_trajbatch_contract():
    from egopose_b200.trajbatch import TrajBatchEgo
    dev = dict(states=torch.randn(6, 115, dtype=torch.float64), actions=torch.randn(6, 52, dtype=torch.float64),
               masks=torch.tensor([1., 1, 0, 1, 1, 0], dtype=torch.float64), next_states=None,
               rewards=torch.rand(6, dtype=torch.float64), exps=torch.ones(6, dtype=torch.float64),
               v_metas=torch.zeros(6, 2, dtype=torch.int32))
    b = TrajBatchEgo(dev=dev, horizon=3)
    assert b.states.shape == (6, 115) and b.states.dtype == np.float64 and b.masks.dtype == np.int64
    assert b.v_metas.shape == (6, 2) and len(b) == 6
    with pytest.raises(AttributeError):
        b.next_states
    h = TrajBatchEgo.from_numpy(states=np.zeros((2, 115)), rewards=np.zeros(2))
    assert h.states.shape == (2, 115) and len(h) == 2
    # reference-format copy: host arrays only, nothing pending when no side stream was involved
    ho = b.host_only()
    assert not ho.dev and ho.horizon == 3 and np.array_equal(ho.states, b.states) and np.array_equal(ho.masks, b.masks)
    assert ho.host_event('states') is None and ho.host_chunks('states') == [] and ho.wait_host() is ho
    assert 'next_states' not in ho._host


def test_flat_storage_aliases_modules_and_adam_state():
    from egopose_b200.agent import _FlatNet
    from egopose_b200.nets import MLP, PolicyGaussian
    torch.manual_seed(0)
    pol = PolicyGaussian(MLP(10, (8, 6), 'relu'), 4, log_std=-2.3, fix_std=True).double()
    before = {k: v.clone() for k, v in pol.state_dict().items()}
    opt = torch.optim.Adam(pol.parameters(), lr=1e-3)
    flat = _FlatNet([(n, p) for n, p in pol.named_parameters() if p.requires_grad], opt, torch.device('cpu'))
    assert 'action_log_std' not in flat.names and flat.flat.numel() == 10 * 8 + 8 + 8 * 6 + 6 + 6 * 4 + 4
    for k, v in pol.state_dict().items():
        assert torch.equal(v, before[k])
    flat.flat.add_(1.0)                                   # kernels write the flat buffer -> modules see it
    assert torch.equal(pol.net.affine_layers[0].weight.data, before['net.affine_layers.0.weight'] + 1.0)
    flat.m.fill_(0.5)
    assert float(opt.state[pol.action_mean.bias]['exp_avg'][0]) == 0.5
    assert set(opt.state_dict()['state'][1].keys()) >= {'step', 'exp_avg', 'exp_avg_sq'}
    assert list(pol.state_dict().keys()) == ['action_log_std', 'net.affine_layers.0.weight', 'net.affine_layers.0.bias',
                                             'net.affine_layers.1.weight', 'net.affine_layers.1.bias',
                                             'action_mean.weight', 'action_mean.bias']


def test_nets_match_reference_init_contract():
    from egopose_b200.nets import MLP, PolicyGaussian, Value
    torch.manual_seed(1)
    p = PolicyGaussian(MLP(243, (300, 200), 'relu'), 52, log_std=-2.3, fix_std=True)
    v = Value(MLP(243, (300, 200), 'relu'))
    assert sum(x.numel() for x in p.parameters()) == 143904 + 0 and sum(x.numel() for x in v.parameters()) == 133601
    assert p.type == 'gaussian' and not p.action_log_std.requires_grad and p.action_log_std.shape == (1, 52)
    assert float(p.action_mean.bias.abs().max()) == 0.0 and float(v.value_head.bias.abs().max()) == 0.0
    x = torch.randn(3, 243)
    a = p.select_action(x, mean_action=True)
    assert a.shape == (3, 52) and p.get_log_prob(x, a).shape == (3, 1) and v(x).shape == (3, 1)


def test_compat_import_paths():
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from utils import *\n"
            "from core.policy_gaussian import PolicyGaussian\nfrom core.critic import Value\nfrom models.mlp import MLP\n"
            "from models.video_state_net import VideoStateNet\nfrom ego_pose.envs.humanoid_v1 import HumanoidEnv\n"
            "from ego_pose.core.agent_ego import AgentEgo\nfrom ego_pose.utils.egomimic_config import Config\n"
            "from ego_pose.core.reward_function import reward_func\nfrom core import estimate_advantages, LoggerRL, TrajBatch\n"
            "assert 'quat_v3' in reward_func and ZFilter.__name__ == 'ZFilter'\n"
            "set_optimizer_lr\nprint('ok')") % (os.path.join(ROOT, 'egopose_b200', 'compat'), ROOT)
    import subprocess
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True)
    assert out.stdout.strip().endswith('ok'), out.stderr[-2000:]


def test_mjcf_compiler_on_builtin_json():
    from egopose_b200.mjcf import load_builtin
    md = load_builtin()
    assert md.nq == 59 and md.nv == 58 and md.nu == 52 and md.nbody == 21
    assert md.body_qposaddr()['RightForeArm'] == (31, 32) and md.body_qposaddr()['LeftHand'] == (42, 45)
    assert md.actuator_names[0] == 'Spine_x' and md.actuator_dof[0] == 6 and md.actuator_dof[-1] == 57
    assert abs(md.total_mass() - 28.455) < 2e-3
    ref = '/root/reference/assets/mujoco_models/humanoid_1205_v1.xml'
    if os.path.exists(ref):
        from egopose_b200.mjcf import compile_mjcf
        m2 = compile_mjcf(ref)
        assert m2.body_mass == md.body_mass and m2.dof_anchor == md.dof_anchor


def test_synthetic_generators_agree():
    """product-side generator == the copy the goldens were produced with"""
    from egopose_b200.mjcf import load_builtin
    from egopose_b200.synthetic import synthetic_takes
    from oracle import cphys
    import dataclasses
    md = load_builtin()
    a = synthetic_takes(md, 2, 30, seed=5)
    b = cphys.synthetic_takes(dataclasses.asdict(md), 2, 30, seed=5)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_checkpoint_roundtrip_and_reference_names(tmp_path):
    """ego_mimic.py:133-139 format; the stream names utils.zfilter.ZFilter like the reference's checkpoints"""
    from egopose_b200 import checkpoint, zfilter
    from egopose_b200.nets import MLP, PolicyGaussian, Value
    torch.manual_seed(0)
    zf = zfilter.ZFilter((5,), clip=5)
    for x in np.random.RandomState(0).randn(10, 5):
        zf(x)
    pol = PolicyGaussian(MLP(8, (6, 4), 'relu'), 3, log_std=-2.3, fix_std=True).double()
    val = Value(MLP(8, (6, 4), 'relu')).double()
    path = str(tmp_path / 'iter_0100.p')
    checkpoint.save_checkpoint(path, pol, None, val, None, zf)
    raw = open(path, 'rb').read()
    assert b'utils.zfilter' in raw and b'egopose_b200' not in raw
    assert 'utils.zfilter' not in sys.modules or not hasattr(sys.modules['utils.zfilter'], '_RefNamedZFilter')
    pol2 = PolicyGaussian(MLP(8, (6, 4), 'relu'), 3, log_std=0.0, fix_std=True).double()
    val2 = Value(MLP(8, (6, 4), 'relu')).double()
    w_before = pol2.net.affine_layers[0].weight.data_ptr()
    cp, rs = checkpoint.load_checkpoint(path, pol2, None, val2, None)
    assert set(cp) == {'policy_dict', 'policy_vs_dict', 'value_dict', 'value_vs_dict', 'running_state'}
    assert pol2.net.affine_layers[0].weight.data_ptr() == w_before            # in-place load keeps aliasing
    for k, v in pol.state_dict().items():
        assert torch.equal(v, pol2.state_dict()[k])
    assert rs.rs.n == 10 and np.allclose(rs.rs.mean, zf.rs.mean) and np.allclose(rs.rs.std, zf.rs.std) and rs.clip == 5
    # the reference itself can read it (build container only)
    if os.path.isdir('/root/reference/utils'):
        code = ("import sys, pickle; sys.path.insert(0, %r)\nfrom oracle import refimport; refimport.install()\n"
                "from utils.zfilter import ZFilter\ncp = pickle.load(open(%r, 'rb'))\n"
                "assert type(cp['running_state']) is ZFilter and cp['running_state'].rs.n == 10\n"
                "from core.policy_gaussian import PolicyGaussian\nfrom models.mlp import MLP\nimport torch\n"
                "p = PolicyGaussian(MLP(8, (6, 4), 'relu'), 3).double(); p.load_state_dict(cp['policy_dict']); print('ok')") % (ROOT, path)
        import subprocess
        out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True)
        assert out.stdout.strip().endswith('ok'), out.stderr[-1500:]


def test_checkpoint_optimizer_and_rng_state(tmp_path):
    """optimizer / RNG state travel under keys the reference ignores; the base format is unchanged"""
    import pickle
    import types
    import torch
    from egopose_b200 import checkpoint, zfilter
    from egopose_b200.nets import MLP, PolicyGaussian, Value
    torch.set_default_dtype(torch.float64)
    pol, val = PolicyGaussian(MLP(6, (8, 4), 'relu'), 3), Value(MLP(6, (8, 4), 'relu'))
    opt_p, opt_v = torch.optim.Adam(pol.parameters(), lr=1e-3), torch.optim.Adam(val.parameters(), lr=1e-3)
    for opt, net in ((opt_p, pol), (opt_v, val)):
        out = net(torch.randn(5, 6))
        (out.loc if hasattr(out, 'loc') else out).sum().backward()
        opt.step()
    agent = types.SimpleNamespace(env=types.SimpleNamespace(_seed=7), iteration=42)
    path = str(tmp_path / 'iter_0001.p')
    checkpoint.save_checkpoint(path, pol, None, val, None, zfilter.ZFilter((4,), clip=5), opt_p, opt_v, agent)
    raw = checkpoint._Unpickler(open(path, 'rb')).load()
    assert {'policy_dict', 'policy_vs_dict', 'value_dict', 'value_vs_dict', 'running_state'} <= set(raw)     # reference keys
    assert raw['rng']['seed'] == 7 and raw['rng']['iteration'] == 42
    pol2 = PolicyGaussian(MLP(6, (8, 4), 'relu'), 3)
    opt_p2 = torch.optim.Adam(pol2.parameters(), lr=1e-3)
    agent2 = types.SimpleNamespace(env=types.SimpleNamespace(_seed=0), iteration=0)
    cp, _ = checkpoint.load_checkpoint(path, pol2, None, None, None)
    checkpoint.restore_training_state(cp, optimizer_policy=opt_p2, agent=agent2)
    st, st2 = opt_p.state[pol.action_mean.weight], opt_p2.state[pol2.action_mean.weight]
    assert torch.equal(st['exp_avg'], st2['exp_avg']) and float(st2['step']) == 1.0
    assert agent2.iteration == 42 and agent2.env._seed == 7


def test_cnn_feature_file_roundtrip(tmp_path):
    """gen_cnn_feature.py:68-70: pickle of (dict take -> [L, F], meta); HumanoidEnv.load_experts reads element 0"""
    import pickle
    import numpy as np
    from egopose_b200 import dataset
    feats = {'take_a': np.random.RandomState(0).randn(12, 128), 'take_b': np.random.RandomState(1).randn(9, 128)}
    path = str(tmp_path / 'features' / 'cnn_feat_x.p')
    dataset.write_cnn_feat_file(path, feats, cfg='subject_03', it=100, meta_id='meta_subject_03')
    got, meta = pickle.load(open(path, 'rb'))       # exactly what humanoid_v1.py:49 does
    assert set(got) == set(feats) and all(np.array_equal(got[k], feats[k]) for k in feats)
    assert meta['cfg'] == 'subject_03' and meta['iter'] == 100 and meta['meta'] == 'meta_subject_03' and 'time' in meta
    g2, _ = dataset.read_cnn_feat_file(path)
    assert np.array_equal(g2['take_b'], feats['take_b'])
