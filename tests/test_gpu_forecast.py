"""egoforecast path (BASELINE config 4): VideoForecastNet on the fused kernels.

  * AgentEgo.update_params with forecast vs-nets vs the golden run of the reference's own AgentEgo +
    models/video_forecast_net.py (tests/golden/fcnet_small.npz)
  * the rollout kernel's in-loop state LSTM (egp_rollout_f64, EgpRolloutIn.d_snet_*): recorded actions must equal
    policy(cat(v_out[window], h_t)) + sigma eps with h_t from the reference-format net stepped over the recorded
    states of each episode (teacher-forced check against the torch mirror that test_fcnet_host.py pins to the
    reference)."""
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')


def test_agent_ego_update_with_forecast_nets_matches_reference_golden(golden):
    from egopose_b200.agent import AgentEgo
    from egopose_b200.nets import MLP, PolicyGaussian, Value, VideoForecastNet
    from egopose_b200.trajbatch import TrajBatchEgo
    torch.set_default_dtype(torch.float64)
    g = golden('fcnet_small')
    F, VH, M, S, A, T, SH = [int(x) for x in g['dims']]
    gamma, tau, clip, lr_p, lr_v, max_norm = g['hyper']

    def load(net, prefix):
        net.load_state_dict({k[len(prefix) + 1:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(prefix + '.')})
        return net.cuda()
    pvs = load(VideoForecastNet(F, S, VH, M, 'lstm', None, SH, 'lstm'), 'pvs0')
    vvs = load(VideoForecastNet(F, S, VH, M, 'lstm', None, SH, 'lstm'), 'vvs0')
    pol = load(PolicyGaussian(MLP(VH + SH, (16, 12), 'relu'), A, log_std=-1.0, fix_std=True), 'p0')
    val = load(Value(MLP(VH + SH, (16, 12), 'relu')), 'v0')
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
    assert int(opt_p.state[pvs.s_net.rnn_f.weight_hh]['step']) == 3


def _forecast_agent(E, T, EPL, F=12, VH=16, SH=128, hidden=(48, 40), seed=11):
    from egopose_b200.agent import AgentEgo
    from egopose_b200.config import Config
    from egopose_b200.env import HumanoidEnv
    from egopose_b200.mjcf import load_builtin
    from egopose_b200.nets import MLP, PolicyGaussian, Value, VideoForecastNet
    from egopose_b200.synthetic import synthetic_cnn_feat, synthetic_takes
    torch.set_default_dtype(torch.float64)
    torch.manual_seed(seed)
    cfg = Config('subject_03', task='egoforecast')
    cfg.env_episode_len = EPL
    env = HumanoidEnv(cfg)
    md = load_builtin()
    L = EPL + 2 * cfg.fr_margin + 20
    env.set_expert_qpos(['a', 'b'], synthetic_takes(md, 2, L, seed=2), synthetic_cnn_feat(2, L, dim=F))
    S, nu = env.obs_dim, md.nu
    mk = lambda: VideoForecastNet(F, S, VH, cfg.fr_margin, 'lstm', None, SH, 'lstm').cuda()  # noqa: E731
    pvs, vvs = mk(), mk()
    pol = PolicyGaussian(MLP(pvs.out_dim, hidden, 'relu'), nu, log_std=-2.3, fix_std=True).cuda()
    val = Value(MLP(vvs.out_dim, hidden, 'relu')).cuda()
    pparams = list(pol.parameters()) + list(pvs.parameters())
    agent = AgentEgo(env=env, dtype=torch.float64, device=torch.device('cuda'), running_state=None, custom_reward=None,
                     policy_net=pol, policy_vs_net=pvs, value_net=val, value_vs_net=vvs,
                     optimizer_policy=torch.optim.Adam(pparams, lr=1e-3),
                     optimizer_value=torch.optim.Adam(list(val.parameters()) + list(vvs.parameters()), lr=1e-3),
                     opt_num_epochs=2, gamma=0.95, tau=0.95, clip_epsilon=0.2, policy_grad_clip=[(pparams, 40)],
                     num_envs=E, horizon=T)
    return agent, env, cfg, pol, pvs, vvs


@pytest.mark.parametrize('SH,hidden', [(128, (48, 40)), (20, (300, 200))])
def test_rollout_state_lstm_teacher_forced(SH, hidden):
    """in-kernel s_net step + policy == the torch mirror stepped over the recorded states of every episode"""
    E, T, EPL = 40, 14, 6
    agent, env, cfg, pol, pvs, vvs = _forecast_agent(E, T, EPL, SH=SH, hidden=hidden)
    assert cfg.fr_margin == 30 and cfg.env_episode_len == EPL
    nu = env.md.nu
    eps = torch.randn(E * T, nu, device='cuda')
    env.set_fix_head_lb(-10.0)          # no falls: every episode runs its EPL steps, so resets hit mid-horizon
    batch, log = agent.sample(E * T, to_host=False, parity=dict(eps=eps))
    b = batch.dev
    states, actions, masks, vm = b['states'], b['actions'], b['masks'].cpu().numpy(), b['v_metas'].cpu().numpy()
    assert log.num_steps == E * T and (masks == 0).sum() >= E * 2
    table, win_off = pvs.context_table(env.cnn_feat, EPL)
    std = torch.exp(pol.action_log_std.detach())
    pvs.set_mode('test')
    worst = 0.0
    with torch.no_grad():
        for e in range(0, E, 7):
            new_ep = True
            for t in range(T):
                n = e * T + t
                if new_ep:
                    pvs.s_net.initialize()
                    take, start = int(vm[n, 0]), int(vm[n, 1])
                    pvs.v_out = table[int(win_off[take]) + start - cfg.fr_margin][None]
                x = pvs(states[n][None])
                mu = pol(x).loc
                want = mu + std * eps[n][None]
                worst = max(worst, float((want - actions[n][None]).abs().max()))
                new_ep = masks[n] == 0
    assert worst < 1e-11, worst
    # train-mode forward over the same batch reproduces the in-kernel hidden states as well: full update cycle
    p_before = pvs.s_net.rnn_f.weight_hh.detach().clone()
    agent.update_params(batch)
    assert not torch.equal(p_before, pvs.s_net.rnn_f.weight_hh) and torch.isfinite(pvs.s_net.rnn_f.weight_hh).all()
    losses = agent.losses()
    assert np.isfinite(losses['surr_loss']).all() and losses['value_loss'][-1] < losses['value_loss'][0]
    assert abs(losses['surr_loss'][0]) < 1e-9        # epoch 0: ratio == 1 -> surrogate = -mean(standardised adv) ~ 0
    env.close()
