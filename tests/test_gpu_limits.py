"""Joint limits on the GPU (egp_model_set_joint_limits; SURVEY 8f row 1, first half) against the oracle's restatement of
MuJoCo's soft-constraint model.  The oracle side is UNPINNED (tests/test_oracle_limits.py says what it is checked
against); these tests pin the CUDA path to it.  Floor contact is not modelled."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')

from oracle import cphys  # noqa: E402
import helpers  # noqa: E402


def cu(a, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device='cuda')


def _violating_states(orc, n, seed):
    rng = np.random.RandomState(seed)
    r = orc.dof_ranges()
    q = np.tile(np.array(orc.md['qpos0']), (n, 1))
    v = rng.randn(n, orc.nv)
    for e in range(n):
        for i in range(6, orc.nv):
            lo, hi = r[i]
            q[e, i + 1] = lo + (hi - lo) * rng.uniform(0.1, 0.9)
        for i in rng.choice(np.arange(6, orc.nv), rng.randint(0, 10), replace=False):
            lo, hi = r[i]
            q[e, i + 1] = hi + rng.uniform(1e-5, 0.3) if rng.rand() < 0.5 else lo - rng.uniform(1e-5, 0.3)
            v[e, i] = rng.uniform(-8, 8)
        quat = rng.randn(4)
        q[e, 3:7] = quat / np.linalg.norm(quat)
    return q, v


def test_forward_with_limits_vs_oracle_unpinned():
    orc = cphys.Oracle(joint_limits=True)
    model = helpers.make_model()
    model.set_joint_limits(True)
    n = 48
    q, v = _violating_states(orc, n, 3)
    ctrl = 20 * np.random.RandomState(4).randn(n, orc.nu)
    _, _, qacc = model.forward_debug(cu(q), cu(v), cu(ctrl))
    qacc = qacc.cpu().numpy()
    model.set_joint_limits(False)
    _, _, qacc_s = model.forward_debug(cu(q), cu(v), cu(ctrl))
    qacc_s = qacc_s.cpu().numpy()
    n_changed = 0
    for e in range(n):
        d = orc.new_data(q[e], v[e], ctrl[e])
        orc.forward(d)
        ref = np.array(d.qacc[:orc.nv])
        assert helpers.relerr(qacc[e], ref) < 1e-8, (e, helpers.relerr(qacc[e], ref))
        n_changed += helpers.relerr(qacc_s[e], ref) > 1e-3
    assert n_changed > n // 2           # the limit rows did act in most of the states
    model.close()


def test_env_step_with_limits_vs_oracle_unpinned():
    orc = cphys.Oracle(joint_limits=True)
    orc.make_expert(cphys.synthetic_takes(orc.md, 1, 40, seed=2))
    orc.cfg.fix_head_lb = -100.0
    model = helpers.make_model()
    model.set_joint_limits(True)
    n = 16
    q, v = _violating_states(orc, n, 8)
    q[:, 2] = 0.9
    v *= 0.3
    act = 0.5 * np.random.RandomState(9).randn(n, orc.nu)
    qd, vd = cu(q), cu(v)
    model.env_step_debug(qd, vd, cu(act))
    for e in range(n):
        env = cphys.EoEnv()
        orc.env_set_state(env, q[e].copy(), v[e].copy())
        env.take = 0
        orc.env_step(env, act[e])
        assert helpers.relerr(qd[e].cpu().numpy(), np.array(env.d.qpos[:orc.nq])) < 1e-7
        assert helpers.relerr(vd[e].cpu().numpy(), np.array(env.d.qvel[:orc.nv])) < 1e-6
    model.close()


def test_rollout_with_limits_vs_oracle_unpinned():
    """whole fused rollouts (the one-warp kernel carries the limit rows) against the oracle's rollout with the same rows;
    a wide action noise drives joints into their ranges, so the two differ from the smooth roll-out"""
    E, T = 6, 12
    res = {}
    for lim in (True, False):
        orc = cphys.Oracle(episode_len=12, joint_limits=lim)
        takes = cphys.synthetic_takes(orc.md, 3, 64, seed=7)
        orc.make_expert(takes, None)
        S, nu = orc.S, orc.nu
        w = helpers.policy_weights(S, 32, 24, nu, seed=3, log_std=0.3)
        rng = np.random.RandomState(9)
        reset_take = rng.randint(0, 3, size=(E, T))
        reset_start = rng.randint(10, 64 - 12 - 10, size=(E, T))
        eps = rng.randn(E * T, nu)
        mean_flag = np.zeros(E * T, dtype=np.uint8)
        zf_mean, zf_std = np.zeros(S), np.ones(S)
        pol = orc.make_policy(w['W1'], w['b1'], w['W2'], w['b2'], w['W3'], w['b3'], w['log_std'])
        orc.cfg.fix_head_lb = -100.0
        res[lim] = ref = orc.rollout(pol, E, T, reset_take, reset_start, eps, mean_flag, zf_mean, zf_std, 5.0, n_threads=4)
        if not lim:
            continue
        model = helpers.make_model()
        model.upload_experts(orc._keep['x_rows'], orc._keep['x_off'], orc._keep['x_lb'], None)
        model.set_joint_limits(True)
        wd = {k: cu(v.ravel() if k == 'log_std' else v) for k, v in w.items()}
        out = model.rollout(wd, E, T, episode_len=12, fix_head_lb=-100.0, zf_mean=cu(zf_mean), zf_std=cu(zf_std), eps=cu(eps),
                            reset_take=cu(reset_take, torch.int32), reset_start=cu(reset_start, torch.int32),
                            mean_flag=cu(mean_flag, torch.uint8))
        torch.cuda.synchronize()
        assert np.array_equal(out['masks'].cpu().numpy(), ref['masks'])
        assert helpers.relerr(out['states'].cpu().numpy(), ref['states']) < 1e-6
        assert helpers.relerr(out['next_states'].cpu().numpy(), ref['next_states']) < 1e-6
        assert np.allclose(out['rewards'].cpu().numpy(), ref['rewards'], rtol=1e-6, atol=1e-9)
        model.close()
    assert helpers.relerr(res[True]['states'], res[False]['states']) > 1e-3


# ---- floor contact -----------------------------------------------------------------------------------------------
def _floor_states(orc, n, seed):
    """random poses a few centimetres above / into the floor in random orientations, some joints beyond their ranges"""
    rng = np.random.RandomState(seed)
    q, v = _violating_states(orc, n, seed)
    v *= 0.5
    for e in range(n):
        d = orc.new_data(q[e], v[e])
        q[e, 2] = 0.0
        d = orc.new_data(q[e], v[e])
        xp = orc.kinematics(q[e])[0]
        q[e, 2] = -xp[:, 2].min() + rng.uniform(-0.02, 0.08)        # lowest body origin near z = 0
    return q, v


@pytest.mark.parametrize('limits', [False, True])
def test_forward_with_contacts_vs_oracle_unpinned(limits):
    orc = cphys.Oracle(joint_limits=limits, contacts=True)
    model = helpers.make_model()
    model.set_contacts(True)
    model.set_joint_limits(limits)
    n = 64
    q, v = _floor_states(orc, n, 11)
    ctrl = 10 * np.random.RandomState(4).randn(n, orc.nu)
    _, _, qacc = model.forward_debug(cu(q), cu(v), cu(ctrl))
    qacc = qacc.cpu().numpy()
    n_rows, worst = 0, 0.0
    for e in range(n):
        d = orc.new_data(q[e], v[e], ctrl[e])
        orc.forward(d)
        n_rows += d.n_efc
        assert d.solver_iter < 99
        worst = max(worst, helpers.relerr(qacc[e], np.array(d.qacc[:orc.nv])))
    assert n_rows > 8 * n               # contacts everywhere
    assert worst < 1e-7, worst
    model.close()


def test_standing_on_the_floor_vs_oracle_unpinned():
    """T-pose on flat feet, PD target = the pose: the contact rows carry the weight (root height constant to a few mm for
    30 steps = 1 s; without a floor the head-height rule ends the episode after 9 steps), CUDA == oracle step by step"""
    orc = cphys.Oracle(joint_limits=True, contacts=True)
    orc.make_expert(cphys.synthetic_takes(orc.md, 1, 40, seed=2))
    orc.cfg.fix_head_lb = -100.0
    model = helpers.make_model()
    model.set_contacts(True)
    model.set_joint_limits(True)
    q = np.array(orc.md['qpos0'], dtype=np.float64)
    q[2] = 0.8665                                        # soles on z = 0
    v = np.zeros(orc.nv)
    act = (-np.array(orc._keep['a_ref']) / np.array(orc._keep['a_scale']))[None]
    qd, vd = cu(q[None]), cu(v[None])
    for t in range(40):
        env = cphys.EoEnv()
        q0, v0 = qd[0].cpu().numpy().copy(), vd[0].cpu().numpy().copy()
        orc.env_set_state(env, q0, v0)
        env.take = 0
        orc.env_step(env, act[0])
        model.env_step_debug(qd, vd, cu(act))
        assert helpers.relerr(qd[0].cpu().numpy(), np.array(env.d.qpos[:orc.nq])) < 1e-7, t
        assert helpers.relerr(vd[0].cpu().numpy(), np.array(env.d.qvel[:orc.nv])) < 1e-5, t
        assert env.d.n_efc >= 16
        if t < 30:
            assert abs(qd[0, 2].item() - 0.8665) < 0.01, (t, qd[0, 2].item())
    model.close()


def test_rollout_with_contacts_and_limits_vs_oracle_unpinned():
    E, T = 6, 24
    orc = cphys.Oracle(episode_len=20, joint_limits=True, contacts=True)
    takes = cphys.synthetic_takes(orc.md, 3, 64, seed=7)
    orc.make_expert(takes, None)
    S, nu = orc.S, orc.nu
    w = helpers.policy_weights(S, 32, 24, nu, seed=3, log_std=-1.0)
    rng = np.random.RandomState(9)
    reset_take = rng.randint(0, 3, size=(E, T))
    reset_start = rng.randint(10, 64 - 20 - 10, size=(E, T))
    eps = rng.randn(E * T, nu)
    mean_flag = np.zeros(E * T, dtype=np.uint8)
    zf_mean, zf_std = np.zeros(S), np.ones(S)
    pol = orc.make_policy(w['W1'], w['b1'], w['W2'], w['b2'], w['W3'], w['b3'], w['log_std'])
    orc.cfg.fix_head_lb = 0.2
    ref = orc.rollout(pol, E, T, reset_take, reset_start, eps, mean_flag, zf_mean, zf_std, 5.0, n_threads=4)
    model = helpers.make_model()
    model.upload_experts(orc._keep['x_rows'], orc._keep['x_off'], orc._keep['x_lb'], None)
    model.set_joint_limits(True)
    model.set_contacts(True)
    wd = {k: cu(v.ravel() if k == 'log_std' else v) for k, v in w.items()}
    out = model.rollout(wd, E, T, episode_len=20, fix_head_lb=0.2, zf_mean=cu(zf_mean), zf_std=cu(zf_std), eps=cu(eps),
                        reset_take=cu(reset_take, torch.int32), reset_start=cu(reset_start, torch.int32),
                        mean_flag=cu(mean_flag, torch.uint8))
    torch.cuda.synchronize()
    assert np.array_equal(out['masks'].cpu().numpy(), ref['masks'])
    # contacts make the dynamics stiff: rounding differences grow faster along an episode than in free fall
    assert helpers.relerr(out['states'].cpu().numpy(), ref['states']) < 1e-5
    assert np.allclose(out['rewards'].cpu().numpy(), ref['rewards'], rtol=1e-5, atol=1e-8)
    model.close()


def test_block_sweep_kernel_with_rows_matches_the_one_warp_kernel(monkeypatch):
    """joint limits + floor contact: the 8-warp block-sweep kernel (rows folded into its articulated-body sweeps, active
    set iterated per CTA) against the one-warp kernel (per-thread active-set loop) on a perf-mode roll-out: same Philox
    streams, so the trajectories must agree to rounding; and both must differ from the smooth roll-out"""
    from egopose_b200 import lib
    orc = cphys.Oracle(episode_len=20)
    takes = cphys.synthetic_takes(orc.md, 3, 64, seed=7)
    orc.make_expert(takes, None)
    S, nu = orc.S, orc.nu
    w = helpers.policy_weights(S, 64, 48, nu, seed=3, log_std=-1.0)
    wd = {k: cu(v.ravel() if k == 'log_std' else v) for k, v in w.items()}
    E, T = 96, 25
    outs = {}
    for name, lim, force in (('smooth', False, None), ('block', True, None), ('one_warp', True, '1')):
        if force is None:
            monkeypatch.delenv('EGP_CONS_VARIANT', raising=False)
        else:
            monkeypatch.setenv('EGP_CONS_VARIANT', force)
        model = helpers.make_model()
        model.upload_experts(orc._keep['x_rows'], orc._keep['x_off'], orc._keep['x_lb'], None)
        model.set_joint_limits(lim)
        model.set_contacts(lim)
        lib.load().egp_cons_cap_hits(1)
        out = model.rollout(wd, E, T, episode_len=20, fix_head_lb=0.2, seed=11, iteration=3)
        torch.cuda.synchronize()
        # every constrained solve reached the fixed point of its active set (this roll-out contains two solves in which the
        # all-rows-at-once iteration flips two coupled rows back and forth: the one-row-per-pass rule settles them)
        assert lib.load().egp_cons_cap_hits(0) == 0
        outs[name] = {k: out[k].clone() for k in ('states', 'rewards', 'masks', 'logger')}
        model.close()
    a, b = outs['block'], outs['one_warp']
    assert torch.equal(a['masks'], b['masks'])
    assert helpers.relerr(a['states'].cpu().numpy(), b['states'].cpu().numpy()) < 1e-6
    assert np.allclose(a['rewards'].cpu().numpy(), b['rewards'].cpu().numpy(), rtol=1e-6, atol=1e-9)
    assert helpers.relerr(a['states'].cpu().numpy(), outs['smooth']['states'].cpu().numpy()) > 1e-3
