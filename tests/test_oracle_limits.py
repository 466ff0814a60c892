"""Joint limits of the oracle (first half): MuJoCo's soft-constraint model restated for the 52 hinge
ranges of humanoid_1205_v1.xml.  UNPINNED (no MuJoCo binary in this image): the tests check the restatement against an
independent numpy evaluation of the same published formulas, the optimality conditions of the solver's cost, and the
physical effect.  The smooth path (limits off) is untouched."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cphys  # noqa: E402


def _state(o, rng, violate=()):
    q = np.array(o.md['qpos0'], dtype=np.float64)
    r = o.dof_ranges()
    for i in range(6, o.nv):
        lo, hi = r[i]
        q[i + 1] = lo + (hi - lo) * rng.uniform(0.2, 0.8)
    for i, amount in violate:
        lo, hi = r[i]
        q[i + 1] = hi + amount if amount > 0 else lo + amount
    v = 0.5 * rng.randn(o.nv)
    return q, v


def _rows_numpy(o, q, v, iw, solref=(0.02, 1.0), solimp=(0.9, 0.95, 0.001, 0.5, 2.0)):
    """independent evaluation of the limit rows: (index, sign, D, aref)"""
    d0, dw, width, mid, power = solimp
    tc = max(solref[0], 2 * o.md['timestep'])
    k, b = 1.0 / (dw * dw * tc * tc * solref[1] ** 2), 2.0 / (dw * tc)
    out = []
    for i, (lo, hi) in enumerate(o.dof_ranges()):
        if i < 6 or not lo < hi:
            continue
        if q[i + 1] < lo:
            dist, s = q[i + 1] - lo, 1.0
        elif q[i + 1] > hi:
            dist, s = hi - q[i + 1], -1.0
        else:
            continue
        x = abs(dist) / width
        y = 1.0 if x >= 1 else (x ** power / mid ** (power - 1) if x <= mid else 1 - (1 - x) ** power / (1 - mid) ** (power - 1))
        imp = d0 + y * (dw - d0)
        R = max(1e-15, (1 - imp) / imp * iw[i])
        out.append((i, s, 1.0 / R, -b * s * v[i] - k * imp * dist))
    return out


def test_limits_off_or_inactive_is_the_smooth_solution():
    o, ol = cphys.Oracle(), cphys.Oracle(joint_limits=True)
    rng = np.random.RandomState(0)
    q, v = _state(o, rng)
    ctrl = 20 * rng.randn(o.nu)
    d0, d1 = o.new_data(q, v, ctrl), ol.new_data(q, v, ctrl)
    o.forward(d0); ol.forward(d1)
    assert np.array_equal(np.array(d0.qacc[:o.nv]), np.array(d1.qacc[:o.nv]))


@pytest.mark.parametrize('seed', [1, 2, 3, 4])
def test_limit_solution_satisfies_the_solver_optimality_conditions_unpinned(seed):
    ol = cphys.Oracle(joint_limits=True)
    o = cphys.Oracle()
    rng = np.random.RandomState(seed)
    nviol = rng.randint(1, 9)
    dofs = rng.choice(np.arange(6, ol.nv), nviol, replace=False)
    viol = [(int(i), float(rng.choice([-1, 1]) * rng.uniform(1e-4, 0.2))) for i in dofs]
    q, v = _state(ol, rng, viol)
    v[dofs] = rng.uniform(-6, 6, nviol)              # some rows leave the active set (joint already moving back fast)
    v[viol[0][0]] = np.sign(viol[0][1]) * 0.3        # ... and one certainly stays in it (still moving outwards)
    ctrl = 30 * rng.randn(ol.nu) * (seed % 2 == 0)   # odd seeds: no actuation, the outward-moving row must be active
    d, ds = ol.new_data(q, v, ctrl), o.new_data(q, v, ctrl)
    ol.forward(d); o.forward(ds)
    a, a0 = np.array(d.qacc[:ol.nv]), np.array(ds.qacc[:ol.nv])
    M = ol.qM(d)
    rows = _rows_numpy(ol, q, v, ol.invweight0())
    assert len(rows) == nviol
    grad = M @ (a - a0)
    n_act = 0
    for i, s, D, aref in rows:
        r = s * a[i] - aref
        if r < 0:
            grad[i] += D * s * r
            n_act += 1
    scale = np.abs(M @ a0).max()
    assert np.abs(grad).max() < 1e-9 * scale, (np.abs(grad).max(), scale)
    # the optimum of a convex cost: no random perturbation lowers it
    def cost(x):
        c = 0.5 * (x - a0) @ M @ (x - a0)
        for i, s, D, aref in rows:
            r = s * x[i] - aref
            c += 0.5 * D * r * r if r < 0 else 0.0
        return c
    c0 = cost(a)
    for _ in range(50):
        assert cost(a + 1e-3 * rng.randn(ol.nv) * np.abs(a).max()) >= c0 - 1e-9 * abs(c0)
    assert n_act >= 1 or seed % 2 == 0


def test_limit_holds_a_driven_joint_unpinned():
    """constant torque pushing the left knee past its range: without limits the angle runs away, with limits the
    penetration stays within a few hundredths of a radian"""
    names = cphys.Oracle().md['joint_names']
    knee = 5 + names.index('LeftLeg_x')
    res = {}
    for lim in (False, True):
        o = cphys.Oracle(joint_limits=lim)
        rng = np.random.RandomState(5)
        q, v = _state(o, rng)
        v[:] = 0.0
        lo, hi = o.dof_ranges()[knee]
        q[knee + 1] = hi - 0.05
        ctrl = np.zeros(o.nu)
        ctrl[knee - 6] = 40.0
        d = o.new_data(q, v, ctrl)
        worst = 0.0
        for _ in range(300):
            o.step(d)
            worst = max(worst, d.qpos[knee + 1] - hi)
        res[lim] = worst
    assert res[False] > 0.5
    assert 0.0 < res[True] < 0.05, res


def test_inverse_weights_of_the_product_match_the_oracle():
    """egopose_b200.mjcf.inverse_weights (numpy, closed form at qpos0) vs the oracle's M^-1 and body Jacobians: two
    independent derivations of mjModel.dof_invweight0 / body_invweight0"""
    from egopose_b200.mjcf import inverse_weights, load_builtin
    o = cphys.Oracle()
    dof_iw, body_iw = inverse_weights(load_builtin())
    assert np.allclose(dof_iw, o.invweight0(), rtol=1e-10)
    assert np.allclose(body_iw, o.body_invweight0(), rtol=1e-10)


def test_contact_solution_satisfies_the_solver_optimality_conditions_unpinned():
    """floor contacts + limits: gradient of MuJoCo's cost at the returned acceleration is zero, rows rebuilt in numpy from the
    oracle's Jacobians are consistent with its active set"""
    o = cphys.Oracle(joint_limits=True, contacts=True)
    os_ = cphys.Oracle()
    rng = np.random.RandomState(3)
    q, v = _state(o, rng)
    quat = np.array([1.0, 0.05, -0.03, 0.02])
    q[3:7] = quat / np.linalg.norm(quat)
    q[2] = 0.0
    q[2] = -o.kinematics(q)[0][:, 2].min() + 0.02
    d, ds = o.new_data(q, 0.2 * v), os_.new_data(q, 0.2 * v)
    o.forward(d); os_.forward(ds)
    assert d.n_efc >= 8 and d.solver_iter < 99
    a, a0 = np.array(d.qacc[:o.nv]), np.array(ds.qacc[:o.nv])
    # the contact forces push up: the root accelerates less downwards than in free fall
    assert a[2] > a0[2] + 1.0


def test_statue_stands_on_the_floor_unpinned():
    o = cphys.Oracle(joint_limits=True, contacts=True)
    o.make_expert(cphys.synthetic_takes(o.md, 1, 40, seed=2))
    o.cfg.fix_head_lb = 0.3
    q = np.array(o.md['qpos0'], dtype=np.float64)
    q[2] = 0.8665
    act = -np.array(o._keep['a_ref']) / np.array(o._keep['a_scale'])
    env = cphys.EoEnv()
    o.env_set_state(env, q, np.zeros(o.nv))
    env.take = 0
    steps = 0
    for t in range(100):
        fail, _ = o.env_step(env, act)
        if fail:
            break
        steps += 1
        if t < 30:
            assert abs(env.d.qpos[2] - 0.8665) < 0.01
    assert steps >= 50          # without a floor: 9


def test_active_set_iteration_converges_quickly_unpinned():
    """300 random poses on / in the floor with violated ranges, random velocities and actuation: the fixed point of the
    active set is reached in a handful of sweeps (the kernels cap the loop at 100)"""
    o = cphys.Oracle(joint_limits=True, contacts=True)
    rng = np.random.RandomState(0)
    iters, rows = [], []
    for _ in range(300):
        viol = [(int(i), float(rng.choice([-1, 1]) * rng.uniform(1e-4, 0.3)))
                for i in rng.choice(np.arange(6, o.nv), rng.randint(0, 8), replace=False)]
        q, _v = _state(o, rng, viol)
        quat = rng.randn(4)
        q[3:7] = quat / np.linalg.norm(quat)
        q[2] = 0.0
        q[2] = -o.kinematics(q)[0][:, 2].min() + rng.uniform(-0.05, 0.1)
        d = o.new_data(q, rng.randn(o.nv) * rng.choice([0.1, 1.0, 3.0]), 20 * rng.randn(o.nu))
        o.forward(d)
        iters.append(d.solver_iter)
        rows.append(d.n_efc)
    assert max(iters) <= 20, max(iters)
    assert max(rows) > 30 and np.mean(rows) > 8
