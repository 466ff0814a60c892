"""Agent-level parity through the public API: AgentPPO.update_params vs the golden run of the reference's
own AgentPPO (tests/golden/ppo_small.npz), and AgentEgo.sample + update_params vs the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')

from oracle import cphys, ppo as oppo  # noqa: E402
import helpers  # noqa: E402


def _nets_from(g, prefix_p, prefix_v, D, H, A):
    from egopose_b200.nets import MLP, PolicyGaussian, Value
    torch.set_default_dtype(torch.float64)
    pol = PolicyGaussian(MLP(D, H, 'relu'), A, log_std=-2.3, fix_std=True)
    val = Value(MLP(D, H, 'relu'))
    pol.load_state_dict({k[len(prefix_p):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(prefix_p)})
    val.load_state_dict({k[len(prefix_v):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(prefix_v)})
    return pol.cuda(), val.cuda()


@pytest.mark.parametrize('gemm', ['ozaki', 'cublas'])
def test_agent_ppo_update_matches_reference_golden(golden, gemm):
    from egopose_b200.agent import AgentPPO
    from egopose_b200.trajbatch import TrajBatch
    g = golden('ppo_small')
    gamma, tau, clip, lr_p, lr_v, max_norm = g['hyper']
    pol, val = _nets_from(g, 'p0.', 'v0.', 24, (32, 16), 6)
    opt_p = torch.optim.Adam(pol.parameters(), lr=lr_p)
    opt_v = torch.optim.Adam(val.parameters(), lr=lr_v)
    agent = AgentPPO(env=None, dtype=torch.float64, device=torch.device('cuda'), policy_net=pol, value_net=val,
                     optimizer_policy=opt_p, optimizer_value=opt_v, opt_num_epochs=3, gamma=gamma, tau=tau,
                     clip_epsilon=clip, policy_grad_clip=[(list(pol.parameters()), max_norm)], gemm=gemm)
    batch = TrajBatch.from_numpy(states=g['states'], actions=g['actions'], rewards=g['rewards'], masks=g['masks'],
                                 exps=g['exps'])
    agent.update_params(batch)
    losses = agent.losses()
    # north-star tolerance on the PPO loss: 1e-5; float64 kernels give ~1e-12
    assert np.allclose(losses['surr_loss'], g['surr_loss'], rtol=1e-9, atol=1e-12)
    assert np.allclose(losses['value_loss'], g['value_loss'], rtol=1e-10)
    for k, v in pol.state_dict().items():
        assert np.allclose(v.cpu().numpy(), g['p3.' + k], rtol=1e-8, atol=1e-10), k
    for k, v in val.state_dict().items():
        assert np.allclose(v.cpu().numpy(), g['v3.' + k], rtol=1e-8, atol=1e-10), k
    # Adam state is exposed through the caller's optimizer (checkpoint format)
    st = opt_p.state[pol.action_mean.weight]
    assert int(st['step']) == 3 and st['exp_avg'].abs().sum() > 0


@pytest.mark.parametrize('gemm', ['ozaki', 'cublas'])
def test_agent_ppo_minibatch_matches_reference_golden(golden, gemm):
    """agents/agent_ppo.py:24-43 mini-batch branch (BASELINE config 5): 2 epochs x 4 slices of <= 200 rows"""
    from egopose_b200.agent import AgentPPO
    from egopose_b200.trajbatch import TrajBatch
    g, m = golden('ppo_small'), golden('ppo_minibatch')
    gamma, tau, clip, lr_p, lr_v, max_norm = g['hyper']
    pol, val = _nets_from(g, 'p0.', 'v0.', 24, (32, 16), 6)
    opt_p = torch.optim.Adam(pol.parameters(), lr=lr_p)
    opt_v = torch.optim.Adam(val.parameters(), lr=lr_v)
    agent = AgentPPO(env=None, dtype=torch.float64, device=torch.device('cuda'), policy_net=pol, value_net=val,
                     optimizer_policy=opt_p, optimizer_value=opt_v, opt_num_epochs=int(m['epochs']), gamma=gamma, tau=tau,
                     clip_epsilon=clip, policy_grad_clip=[(list(pol.parameters()), max_norm)], use_mini_batch=True,
                     opt_batch_size=int(m['opt_batch_size']), gemm=gemm)
    batch = TrajBatch.from_numpy(states=g['states'], actions=g['actions'], rewards=g['rewards'], masks=g['masks'],
                                 exps=g['exps'])
    np.random.seed(int(m['seed']))
    agent.update_params(batch)
    assert np.allclose(agent.losses()['surr_loss'], m['surr_loss'], rtol=1e-8, atol=1e-11)
    for k, v in pol.state_dict().items():
        assert np.allclose(v.cpu().numpy(), m['p.' + k], rtol=1e-8, atol=1e-10), k
    for k, v in val.state_dict().items():
        assert np.allclose(v.cpu().numpy(), m['v.' + k], rtol=1e-8, atol=1e-10), k


def test_agent_ego_sample_and_update_vs_oracle():
    from egopose_b200.agent import AgentEgo
    from egopose_b200.config import Config
    from egopose_b200.env import HumanoidEnv
    from egopose_b200.nets import MLP, FrameContext, PolicyGaussian, Value
    from egopose_b200.mjcf import load_builtin
    from egopose_b200.synthetic import synthetic_cnn_feat, synthetic_takes
    torch.set_default_dtype(torch.float64)
    torch.manual_seed(3)
    E, T, EPL, CD = 24, 10, 8, 16
    cfg = Config('subject_03')
    cfg.env_episode_len = EPL
    env = HumanoidEnv(cfg)
    md = load_builtin()
    takes = synthetic_takes(md, 3, 60, seed=4)
    cnn = synthetic_cnn_feat(3, 60, dim=CD)
    env.set_expert_qpos(['a', 'b', 'c'], takes, cnn)
    S, nu = env.obs_dim, md.nu
    pol = PolicyGaussian(MLP(S + CD, (48, 32), 'relu'), nu, log_std=-2.3, fix_std=True).cuda()
    val = Value(MLP(S + CD, (48, 32), 'relu')).cuda()
    p0 = {k: v.cpu().numpy().copy() for k, v in pol.state_dict().items()}
    v0 = {k: v.cpu().numpy().copy() for k, v in val.state_dict().items()}
    opt_p = torch.optim.Adam(pol.parameters(), lr=5e-4)
    opt_v = torch.optim.Adam(val.parameters(), lr=3e-3)
    agent = AgentEgo(env=env, dtype=torch.float64, device=torch.device('cuda'), running_state=None, custom_reward=None,
                     num_threads=12, policy_net=pol, policy_vs_net=FrameContext(CD), value_net=val,
                     value_vs_net=FrameContext(CD), optimizer_policy=opt_p, optimizer_value=opt_v, opt_num_epochs=2,
                     gamma=0.95, tau=0.95, clip_epsilon=0.2, policy_grad_clip=[(list(pol.parameters()), 40)],
                     num_envs=E, horizon=T)
    rng = np.random.RandomState(8)
    rt, rs = rng.randint(0, 3, size=(E, T)), rng.randint(10, 60 - EPL - 10, size=(E, T))
    eps = rng.randn(E * T, nu)
    cu = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device='cuda')  # noqa: E731
    batch, log = agent.sample(E * T, to_host=True, parity=dict(eps=cu(eps), reset_take=cu(rt, torch.int32),
                                                                reset_start=cu(rs, torch.int32)))
    # oracle rollout on the same experts (GPU gen_expert rows are checked separately) / noise / resets
    orc = cphys.Oracle(episode_len=EPL)
    orc.make_expert([np.array(t) for t in takes], np.concatenate(cnn))
    pw = orc.make_policy(p0['net.affine_layers.0.weight'], p0['net.affine_layers.0.bias'], p0['net.affine_layers.1.weight'],
                         p0['net.affine_layers.1.bias'], p0['action_mean.weight'], p0['action_mean.bias'], p0['action_log_std'])
    ref = orc.rollout(pw, E, T, rt, rs, eps)
    assert batch.states.shape == (E * T, S) and batch.masks.dtype == np.int64
    assert np.array_equal(batch.masks, ref['masks'].astype(np.int64))
    assert helpers.relerr(batch.states, ref['states']) < 1e-6
    assert np.allclose(batch.rewards, ref['rewards'], rtol=1e-6, atol=1e-9)
    assert np.array_equal(batch.v_metas, ref['v_metas'])
    assert abs(log.avg_c_reward - ref['rewards'].mean()) < 1e-8 and log.num_steps == E * T
    assert abs(log.avg_episode_reward - E * T / (ref['masks'] == 0).sum()) < 1e-12
    # update from the HOST batch (reference-format call) vs the oracle's PPO update on the oracle's batch
    from egopose_b200.trajbatch import TrajBatchEgo
    hb = TrajBatchEgo(host={k: getattr(batch, k) for k in batch.fields}, horizon=T)
    agent.update_params(hb)
    frames = np.array(orc._keep['x_off'])[ref['v_metas'][:, 0]] + ref['v_metas'][:, 1]
    t_in_ep = np.zeros(E * T, dtype=np.int64)
    for e in range(E):
        c = 0
        for t in range(T):
            t_in_ep[e * T + t] = c
            c = 0 if ref['masks'][e * T + t] == 0 else c + 1
    x = np.concatenate([np.concatenate(cnn)[frames + t_in_ep], ref['states']], axis=1)
    with torch.no_grad():
        values = oppo.value_forward(torch.from_numpy(x), {k: torch.from_numpy(v) for k, v in v0.items()}).numpy()
    adv, ret = oppo.gae(ref['rewards'], ref['masks'], values, 0.95, 0.95)
    new_p, new_v, info = oppo.ppo_update(p0, v0, x, ref['actions'], ret, adv, ref['exps'], 0.2, 5e-4, 3e-3, 40.0, epochs=2)
    losses = agent.losses()
    assert np.allclose(losses['surr_loss'], info['surr_loss'], rtol=1e-5, atol=1e-9)      # north star: 1e-5
    assert np.allclose(losses['value_loss'], info['value_loss'], rtol=1e-5)
    for k, v in pol.state_dict().items():
        assert np.allclose(v.cpu().numpy(), new_p[k], rtol=1e-5, atol=1e-8), k
    for k, v in val.state_dict().items():
        assert np.allclose(v.cpu().numpy(), new_v[k], rtol=1e-5, atol=1e-8), k
    env.close()


def test_agent_ego_with_video_state_net_matches_reference_golden(golden):
    """AgentEgo.update_params with real BiLSTM VideoStateNets (agent_ego.py:34-57, video_state_net.py train mode)
    vs the golden run of the reference's own classes; then the all-windows test-mode table feeds the rollout."""
    import types
    from egopose_b200.agent import AgentEgo
    from egopose_b200.nets import MLP, PolicyGaussian, Value, VideoStateNet
    from egopose_b200.trajbatch import TrajBatchEgo
    torch.set_default_dtype(torch.float64)
    g = golden('vsnet_small')
    F, VH, M, S, A, T = [int(x) for x in g['dims']]
    gamma, tau, clip, lr_p, lr_v, max_norm = g['hyper']

    def load(net, prefix):
        net.load_state_dict({k[len(prefix) + 1:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(prefix + '.')})
        return net.cuda()
    pvs, vvs = load(VideoStateNet(F, VH, M), 'pvs0'), load(VideoStateNet(F, VH, M), 'vvs0')
    pol = load(PolicyGaussian(MLP(S + VH, (16, 12), 'relu'), A, log_std=-1.0, fix_std=True), 'p0')
    val = load(Value(MLP(S + VH, (16, 12), 'relu')), 'v0')
    pparams = list(pol.parameters()) + list(pvs.parameters())
    vparams = list(val.parameters()) + list(vvs.parameters())
    opt_p = torch.optim.Adam(pparams, lr=lr_p)
    opt_v = torch.optim.Adam(vparams, lr=lr_v)
    env = types.SimpleNamespace(cnn_feat=[c for c in g['cnn_feat']], kernel=types.SimpleNamespace(ctx_dim=0))
    agent = AgentEgo(env=env, dtype=torch.float64, device=torch.device('cuda'), running_state=None, custom_reward=None,
                     policy_net=pol, policy_vs_net=pvs, value_net=val, value_vs_net=vvs, optimizer_policy=opt_p,
                     optimizer_value=opt_v, opt_num_epochs=3, gamma=gamma, tau=tau, clip_epsilon=clip,
                     policy_grad_clip=[(pparams, max_norm)])
    batch = TrajBatchEgo.from_numpy(**{k: g['batch.' + k] for k in ('states', 'actions', 'rewards', 'masks', 'exps', 'v_metas')})
    agent.update_params(batch)
    assert np.allclose(agent.losses()['surr_loss'], g['surr_loss'], rtol=1e-8, atol=1e-11)   # north star: 1e-5
    for prefix, net in (('p3', pol), ('v3', val), ('pvs3', pvs), ('vvs3', vvs)):
        for k, v in net.state_dict().items():
            assert np.allclose(v.cpu().numpy(), g[prefix + '.' + k], rtol=1e-7, atol=1e-9), (prefix, k)
    # Adam state of the BiLSTM parameters is exposed through the caller's optimizer as well
    assert int(opt_p.state[pvs.v_net.rnn_f.weight_ih]['step']) == 3


def test_rollout_with_window_context_table():
    """ctx_mode 1 (one row block per (take, start) window) reproduces the per-frame table when it is built from it,
    and AgentEgo.sample feeds the VideoStateNet table to the kernel"""
    from oracle import cphys
    import test_gpu_rollout as tr
    orc, model = tr._setup(3, 64, 8, seed=21)
    S, nu, T_ep, m = orc.S, orc.nu, 12, 10
    w = helpers.policy_weights(S + 8, 48, 32, nu, seed=4)
    wd = {k: tr.cu(v.ravel() if k == 'log_std' else v) for k, v in w.items()}
    a = {k: v.clone() for k, v in model.rollout(wd, 50, 20, episode_len=T_ep, seed=9, iteration=1).items()}
    frame_ctx = orc._keep['x_ctx']                         # [3 * 64, 8] per frame
    nwin = 64 - T_ep - 2 * m
    table = np.zeros((3 * nwin * T_ep, 8))
    for k in range(3):
        for wi in range(nwin):
            start = m + wi
            table[(k * nwin + wi) * T_ep:(k * nwin + wi + 1) * T_ep] = frame_ctx[k * 64 + start: k * 64 + start + T_ep]
    win_off = torch.tensor([0, nwin, 2 * nwin, 3 * nwin], dtype=torch.int32, device='cuda')
    b = model.rollout(wd, 50, 20, episode_len=T_ep, seed=9, iteration=1, ctx=tr.cu(table), win_off=win_off)
    for k in ('states', 'actions', 'rewards', 'masks'):
        assert torch.equal(a[k], b[k]), k
    model.close()


def test_agent_ego_sample_with_video_state_net():
    """sample() hands the all-windows VideoStateNet table to the kernel: identical to a manual rollout with that table"""
    from egopose_b200.agent import AgentEgo
    from egopose_b200.config import Config
    from egopose_b200.env import HumanoidEnv
    from egopose_b200.nets import MLP, PolicyGaussian, Value, VideoStateNet
    from egopose_b200.mjcf import load_builtin
    from egopose_b200.synthetic import synthetic_cnn_feat, synthetic_takes
    torch.set_default_dtype(torch.float64)
    torch.manual_seed(11)
    E, T, EPL, F, VH = 20, 9, 8, 12, 10
    cfg = Config('subject_03')
    cfg.env_episode_len = EPL
    env = HumanoidEnv(cfg)
    md = load_builtin()
    env.set_expert_qpos(['a', 'b'], synthetic_takes(md, 2, 50, seed=2), synthetic_cnn_feat(2, 50, dim=F))
    S, nu = env.obs_dim, md.nu
    pvs, vvs = VideoStateNet(F, VH, cfg.fr_margin).cuda(), VideoStateNet(F, VH, cfg.fr_margin).cuda()
    pol = PolicyGaussian(MLP(S + VH, (32, 24), 'relu'), nu, log_std=-2.3, fix_std=True).cuda()
    val = Value(MLP(S + VH, (32, 24), 'relu')).cuda()
    pparams = list(pol.parameters()) + list(pvs.parameters())
    agent = AgentEgo(env=env, dtype=torch.float64, device=torch.device('cuda'), running_state=None, custom_reward=None,
                     policy_net=pol, policy_vs_net=pvs, value_net=val, value_vs_net=vvs,
                     optimizer_policy=torch.optim.Adam(pparams, lr=1e-3),
                     optimizer_value=torch.optim.Adam(list(val.parameters()) + list(vvs.parameters()), lr=1e-3),
                     opt_num_epochs=2, gamma=0.95, tau=0.95, clip_epsilon=0.2, policy_grad_clip=[(pparams, 40)],
                     num_envs=E, horizon=T)
    batch, log = agent.sample(E * T, to_host=False)
    states = batch.dev['states'].clone()
    table, win_off = pvs.context_table(env.cnn_feat, EPL)
    w = agent._policy_weights()
    ref = env.kernel.rollout(w, E, T, EPL, cfg.fr_margin, seed=env._seed, iteration=0, ctx=table, win_off=win_off)
    assert torch.equal(states, ref['states']) and log.num_steps == E * T
    p_before = pvs.v_net.rnn_b.weight_hh.detach().clone()
    agent.update_params(batch)                     # full sample -> update cycle with BiLSTM gradients
    assert not torch.equal(p_before, pvs.v_net.rnn_b.weight_hh) and torch.isfinite(pvs.v_net.rnn_b.weight_hh).all()
    losses = agent.losses()
    assert np.isfinite(losses['surr_loss']).all() and losses['value_loss'][-1] < losses['value_loss'][0]
    env.close()
