"""Self-validation of the restated smooth dynamics (PARITY UNPINNED w.r.t. MuJoCo: no binary available).
Checks that do not share the oracle's spatial-algebra code: per-body Newton-Euler from finite differences
of the kinematics (tests/physcheck.py), momentum balance, energy drift order, SPD mass matrix, masses."""
import numpy as np

import physcheck as pc


def _rand_state(orc, rng, vel=1.0):
    q = np.array(orc.md['qpos0'])
    q[2] = 0.9
    quat = rng.randn(4)
    q[3:7] = quat / np.linalg.norm(quat)
    q[7:] = rng.uniform(-0.5, 0.5, size=52)
    return q, rng.randn(58) * vel


def test_compiled_model_matches_survey(oracle):
    md = oracle.md
    assert (md['nq'], md['nv'], md['nu'], md['nbody']) == (59, 58, 52, 21)
    assert abs(sum(md['body_mass']) - 28.455) < 2e-3
    by = dict(zip(md['body_names'], md['body_mass']))
    for name, m in (('Hips', 3.591), ('Head', 1.767), ('RightUpLeg', 3.042), ('LeftFoot', 1.488), ('Neck', 0.341)):
        assert abs(by[name] - m) < 2e-3
    assert md['dof_parent'][24] == 17 and md['dof_parent'][34] == 17 and md['dof_parent'][44] == 5
    assert md['dof_armature'][:6] == [0.0] * 6 and md['dof_armature'][6] == 0.01


def test_mass_matrix_and_bias_vs_newton_euler(oracle):
    rng = np.random.RandomState(0)
    arm = np.array(oracle.md['dof_armature'])
    for _ in range(2):
        q, v = _rand_state(oracle, rng)
        a = rng.randn(58) * 5.0
        d = oracle.new_data(q, v)
        oracle.forward(d)
        M = oracle.qM(d)
        assert np.abs(M - M.T).max() == 0.0
        assert np.linalg.eigvalsh(M).min() > 0
        tau = M @ a + np.array(d.qfrc_bias[:58]) - arm * a
        ref = pc.newton_euler_tau(oracle, q, v, a)
        assert np.abs(tau - ref).max() / np.abs(ref).max() < 2e-5


def _energy_momentum(orc, d):
    md = orc.md
    q = np.array(d.qpos[:59])
    v = np.array(d.qvel[:58])
    xpos, xquat, xipos, _, _ = orc.kinematics(q)
    M = orc.qM(d) - np.diag(md['dof_armature'])
    mass = np.array(md['body_mass'])
    ke = 0.5 * v @ M @ v
    pe = 9.81 * (mass * xipos[:, 2]).sum()
    return ke + pe


def test_energy_drift_first_order_and_momentum(oracle):
    """zero torque, no armature effect on energy bookkeeping: drift of semi-implicit Euler is O(h)."""
    rng = np.random.RandomState(3)
    q, v = _rand_state(oracle, rng, vel=0.5)
    drifts = []
    for sub in (1, 4):
        import ctypes as C
        from oracle import cphys
        m2 = cphys.Oracle()
        m2.model.timestep = oracle.md['timestep'] / sub
        # remove armature so that kinetic energy is exactly 1/2 v'Mv of the rigid bodies
        m2._keep['dof_armature'][:] = 0.0
        d = m2.new_data(q, v)
        m2.forward(d)
        e0 = _energy_momentum(m2, d)
        for _ in range(45 * sub):
            m2.step(d)
        m2.forward(d)
        drifts.append(abs(_energy_momentum(m2, d) - e0) / abs(e0))
    assert drifts[0] < 2e-2
    assert drifts[1] < drifts[0] * 0.45          # ~4x smaller step -> ~4x smaller drift
    # linear momentum: free fall with internal motion, d/dt (m * v_com) = m g -> COM accel = g
    d = oracle.new_data(q, v)
    oracle.forward(d)
    c0 = np.array(d.subtree_com)
    n, h = 30, oracle.md['timestep']
    coms = [c0]
    for _ in range(n):
        oracle.step(d)
        oracle.forward(d)
        coms.append(np.array(d.subtree_com))
    coms = np.array(coms)
    acc = (coms[2:] - 2 * coms[1:-1] + coms[:-2]) / h ** 2
    assert np.allclose(acc.mean(0), [0, 0, -9.81], atol=2e-2)


def test_stale_quantities_after_step(oracle):
    """mj_step does not refresh xpos / qM / qfrc_bias after integrating."""
    rng = np.random.RandomState(5)
    q, v = _rand_state(oracle, rng)
    d = oracle.new_data(q, v)
    oracle.step(d)
    stale_bias = np.array(d.qfrc_bias[:58])
    d2 = oracle.new_data(q, v)
    oracle.forward(d2)
    assert np.array_equal(stale_bias, np.array(d2.qfrc_bias[:58]))
    assert not np.allclose(np.array(d.qpos[:59]), q)
