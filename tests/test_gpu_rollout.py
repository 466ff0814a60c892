"""Fused rollout kernel vs the CPU oracle's eo_rollout on identical seeds / noise / reset draws, and
size-independent properties at larger sizes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')

from oracle import cphys  # noqa: E402
import helpers  # noqa: E402


def cu(a, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device='cuda')


def _setup(n_takes, L, ctx_dim, seed):
    orc = cphys.Oracle(episode_len=12)
    takes = cphys.synthetic_takes(orc.md, n_takes, L, seed=seed)
    ctx = np.random.RandomState(seed + 1).randn(n_takes * L, ctx_dim) if ctx_dim else None
    orc.make_expert(takes, ctx)
    model = helpers.make_model()
    model.upload_experts(orc._keep['x_rows'], orc._keep['x_off'], orc._keep['x_lb'], ctx)
    return orc, model


@pytest.mark.parametrize('E,T,ctx_dim,hidden,head_lb', [(5, 14, 0, (24, 16), None), (40, 9, 16, (40, 24), None),
                                                       (3, 30, 8, (300, 200), -100.0),
                                                       (4, 5, 128, (376, 376), None),      # tile depth 32 plan
                                                       (4, 5, 128, (512, 512), None)])     # BASELINE config 3: chunked layers 2/3
def test_rollout_matches_oracle(E, T, ctx_dim, hidden, head_lb):
    orc, model = _setup(3, 64, ctx_dim, seed=7)
    orc.cfg.fix_head_lb = float('nan') if head_lb is None else head_lb
    orc.cfg.end_reward = 0.37
    S, nu = orc.S, orc.nu
    w = helpers.policy_weights(S + ctx_dim, hidden[0], hidden[1], nu, seed=3)
    rng = np.random.RandomState(9)
    max_resets = T
    reset_take = rng.randint(0, 3, size=(E, max_resets))
    reset_start = rng.randint(10, 64 - 12 - 10, size=(E, max_resets))
    eps = rng.randn(E * T, nu)
    mean_flag = (rng.rand(E * T) < 0.1).astype(np.uint8)
    zf_mean, zf_std = rng.randn(S) * 0.1, rng.uniform(0.5, 2.0, size=S)
    pol = orc.make_policy(w['W1'], w['b1'], w['W2'], w['b2'], w['W3'], w['b3'], w['log_std'])
    ref = orc.rollout(pol, E, T, reset_take, reset_start, eps, mean_flag, zf_mean, zf_std, 5.0, n_threads=4)
    wd = {k: cu(v.ravel() if k == 'log_std' else v) for k, v in w.items()}
    out = model.rollout(wd, E, T, episode_len=12, end_reward=0.37, fix_head_lb=head_lb, zf_mean=cu(zf_mean),
                        zf_std=cu(zf_std), eps=cu(eps), reset_take=cu(reset_take, torch.int32),
                        reset_start=cu(reset_start, torch.int32), mean_flag=cu(mean_flag, torch.uint8))
    torch.cuda.synchronize()
    assert np.array_equal(out['masks'].cpu().numpy(), ref['masks'])
    assert np.array_equal(out['exps'].cpu().numpy(), ref['exps'])
    assert np.array_equal(out['v_metas'].cpu().numpy(), ref['v_metas'])
    # chaotic dynamics amplify rounding over an episode; per-row tolerance stays far below the 1e-4 north star
    assert helpers.relerr(out['states'].cpu().numpy(), ref['states']) < 1e-6
    assert helpers.relerr(out['actions'].cpu().numpy(), ref['actions']) < 1e-6
    assert helpers.relerr(out['next_states'].cpu().numpy(), ref['next_states']) < 1e-6
    assert np.allclose(out['rewards'].cpu().numpy(), ref['rewards'], rtol=1e-6, atol=1e-9)
    assert np.allclose(out['c_info'].cpu().numpy(), ref['c_info'], rtol=1e-6, atol=1e-9)
    assert helpers.relerr(out['final_qpos'].cpu().numpy(), ref['final_qpos']) < 1e-6
    lg = out['logger'].cpu().numpy()
    assert lg[0] == E * T
    assert lg[1] == (ref['masks'] == 0).sum()
    assert abs(lg[3] - ref['rewards'].sum()) < 1e-6 * max(1.0, ref['rewards'].sum())
    assert abs(lg[4] - ref['rewards'].min()) < 1e-9 and abs(lg[5] - ref['rewards'].max()) < 1e-6
    model.close()


def test_rollout_philox_mode_properties():
    """perf mode (in-kernel Philox): determinism, env-independence of results, masks / episode structure"""
    orc, model = _setup(4, 80, 8, seed=3)
    S, nu = orc.S, orc.nu
    w = helpers.policy_weights(S + 8, 64, 32, nu, seed=5)
    wd = {k: cu(v.ravel() if k == 'log_std' else v) for k, v in w.items()}
    a = {k: v.clone() for k, v in model.rollout(wd, 96, 20, episode_len=12, seed=5, iteration=2).items()}
    b = model.rollout(wd, 96, 20, episode_len=12, seed=5, iteration=2)
    for k in ('states', 'actions', 'rewards', 'masks'):
        assert torch.equal(a[k], b[k]), k
    c = model.rollout(wd, 64, 20, episode_len=12, seed=5, iteration=2)       # fewer envs: same per-env streams
    assert torch.equal(c['states'], a['states'][:64 * 20])
    d = model.rollout(wd, 96, 20, episode_len=12, seed=5, iteration=3)
    assert not torch.equal(d['actions'], a['actions'])
    masks = a['masks'].view(96, 20)
    assert (masks[:, -1] == 0).all()
    assert torch.isfinite(a['states']).all() and torch.isfinite(a['rewards']).all()
    vm = a['v_metas'].view(96, 20, 2)
    assert (vm[..., 0] >= 0).all() and (vm[..., 0] < 4).all()
    assert (vm[..., 1] >= 10).all() and (vm[..., 1] < 80 - 12 - 10).all()
    # noise statistics: (action - mean)/std is standard normal -> check through two rollouts sharing states at t=0
    z = ((a['actions'] - d['actions']).view(96, 20, nu)[:, 0] / (np.exp(-2.3) * np.sqrt(2))).cpu().numpy()
    assert abs(z.mean()) < 0.1 and abs(z.std() - 1.0) < 0.1
    model.close()


def test_rollout_variants_agree():
    """T4 (4 warps / 32 envs, shared-memory tree data) and V1 (1 warp) kernels give the same trajectories"""
    import os
    orc, model = _setup(3, 64, 8, seed=11)
    S, nu = orc.S, orc.nu
    w = helpers.policy_weights(S + 8, 64, 48, nu, seed=2)
    wd = {k: cu(v.ravel() if k == 'log_std' else v) for k, v in w.items()}
    outs = []
    for variant in ('0', '1'):
        os.environ['EGP_ROLLOUT_VARIANT'] = variant
        o = model.rollout(wd, 70, 25, episode_len=12, seed=3, iteration=1, end_reward=0.2)
        torch.cuda.synchronize()
        outs.append({k: v.clone() for k, v in o.items()})
    os.environ.pop('EGP_ROLLOUT_VARIANT')
    a, b = outs
    assert torch.equal(a['masks'], b['masks']) and torch.equal(a['v_metas'], b['v_metas'])
    for k in ('states', 'actions', 'next_states', 'rewards', 'c_info', 'final_qpos', 'final_qvel', 'raw_obs'):
        assert helpers.relerr(a[k].cpu().numpy(), b[k].cpu().numpy()) < 1e-7, k
    la, lb = a['logger'].cpu().numpy(), b['logger'].cpu().numpy()
    assert np.allclose(la, lb, rtol=1e-7, atol=1e-9)
    model.close()


def test_rollout_edge_cases():
    """smallest batch, NaN guard (mj_checkPos analogue: terminate + reset + count, trajbatch stays finite),
    Bernoulli mean-action flags (agents/agent.py:46,61)"""
    orc, model = _setup(2, 64, 0, seed=13)
    S, nu = orc.S, orc.nu
    w = helpers.policy_weights(S, 32, 16, nu, seed=7)
    wd = {k: cu(v.ravel() if k == 'log_std' else v) for k, v in w.items()}
    # E = 1, T = 1 against the oracle
    eps = np.random.RandomState(1).randn(1, nu)
    rt, rs = np.array([[1]]), np.array([[20]])
    ref = orc.rollout(orc.make_policy(w['W1'], w['b1'], w['W2'], w['b2'], w['W3'], w['b3'], w['log_std']), 1, 1, rt, rs, eps)
    out = model.rollout(wd, 1, 1, episode_len=12, eps=cu(eps), reset_take=cu(rt, torch.int32), reset_start=cu(rs, torch.int32))
    assert out['masks'].item() == 0.0 and helpers.relerr(out['next_states'].cpu().numpy(), ref['next_states']) < 1e-9
    assert abs(out['rewards'].item() - ref['rewards'][0]) < 1e-10
    # NaN guard: a poisoned policy head makes every action NaN
    bad = dict(wd)
    bad['W3'] = torch.full_like(wd['W3'], float('nan'))
    o = model.rollout(bad, 40, 6, episode_len=12, seed=3)
    lg = o['logger'].cpu().numpy()
    assert lg[13] == 40 * 6 and (o['masks'] == 0).all() and (o['rewards'] == 0).all()
    assert torch.isfinite(o['states']).all() and torch.isfinite(o['next_states']).all()
    # noise_rate 0.7 -> ~30% mean-action steps, recorded as exps == 0 with zero-noise actions
    o = model.rollout(wd, 64, 40, episode_len=12, seed=5, noise_rate=0.7)
    frac = 1.0 - o['exps'].mean().item()
    assert 0.25 < frac < 0.35
    o2 = model.rollout(wd, 64, 40, episode_len=12, seed=5, noise_rate=0.7, mean_action=True)
    assert (o2['exps'] == 0).all()
    model.close()
