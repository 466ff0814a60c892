"""CUDA physics / env step / expert features (through the C ABI) vs the C oracle and the golden
trajectories recorded from the reference's own HumanoidEnv.  North-star tolerance: 1e-4 rel on qpos/qvel;
float64 on both sides lets us demand far tighter."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')

from oracle import cphys  # noqa: E402
import helpers  # noqa: E402


def cu(a, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device='cuda')


@pytest.fixture(scope='module')
def model():
    m = helpers.make_model()
    yield m
    m.close()


def test_forward_bias_xpos_qacc(model, oracle):
    n = 37
    q, v = helpers.rand_states(n, seed=11)
    ctrl = np.random.RandomState(1).randn(n, 52) * 20
    bias, xpos, qacc = model.forward_debug(cu(q), cu(v), cu(ctrl))
    for i in range(n):
        d = oracle.new_data(q[i], v[i], ctrl[i])
        oracle.forward(d)
        assert helpers.relerr(bias[i].cpu().numpy(), np.array(d.qfrc_bias[:58])) < 1e-11
        assert helpers.relerr(xpos[i].cpu().numpy(), np.array(d.xpos).reshape(-1, 3)[:21]) < 1e-13
        # ABA vs dense Cholesky: identical up to conditioning of M
        assert helpers.relerr(qacc[i].cpu().numpy(), np.array(d.qacc[:58])) < 1e-8


def test_env_step_vs_oracle(model, oracle):
    """reset-forward + 15 stable-PD sub-steps from random states: qpos/qvel/obs/head height/first torque"""
    n = 33
    q, v = helpers.rand_states(n, seed=5, vel=0.5)
    act = np.random.RandomState(2).randn(n, 52) * 0.3
    takes = cphys.synthetic_takes(oracle.md, 1, 40, seed=2)
    oracle.make_expert(takes)
    qd, vd = cu(q), cu(v)
    obs, hz, tq = model.env_step_debug(qd, vd, cu(act))
    for i in range(n):
        env = cphys.EoEnv()
        oracle.L.eo_env_set_state(cphys.C.byref(oracle.model), cphys.C.byref(env), cphys._p(q[i].copy()), cphys._p(v[i].copy()))
        ctrl = np.array(oracle._keep['a_ref']) + act[i] * np.array(oracle._keep['a_scale'])
        t0 = np.clip(oracle.compute_torque(env.d, ctrl), -np.array(oracle._keep['torque_lim']), np.array(oracle._keep['torque_lim']))
        assert helpers.relerr(tq[i].cpu().numpy(), t0) < 1e-8
        env.take = 0
        oracle.cfg.fix_head_lb = -100.0
        oracle.env_step(env, act[i])
        assert helpers.relerr(qd[i].cpu().numpy(), np.array(env.d.qpos[:59])) < 1e-7      # north star: 1e-4
        assert helpers.relerr(vd[i].cpu().numpy(), np.array(env.d.qvel[:58])) < 1e-6      # north star: 1e-4
        assert helpers.relerr(obs[i].cpu().numpy(), oracle.env_obs(env)) < 1e-6
        assert abs(hz[i].item() - np.array(env.d.xpos).reshape(-1, 3)[oracle.md['body_names'].index('Head'), 2]) < 1e-8


def test_env_step_vs_reference_golden(model, golden):
    """first env.step of both golden episodes (reference HumanoidEnv on the restated physics)"""
    g = golden('env_traj')
    for ei in range(2):
        q0, v0 = g['ep%d.qpos' % ei][0], g['ep%d.qvel' % ei][0]
        qd, vd = cu(q0[None]), cu(v0[None])
        obs, hz, tq = model.env_step_debug(qd, vd, cu(g['ep%d.action' % ei][:1]))
        # the golden records compute_torque() before np.clip (humanoid_v1.py:171-172): clip with the per-joint limits
        lim = np.array([jp[5] for jp in helpers.cfg_dict()['joint_params']], dtype=np.float64)
        assert helpers.relerr(tq[0].cpu().numpy(), np.clip(g['ep%d.torque0' % ei][0], -lim, lim)) < 1e-7
        assert helpers.relerr(qd[0].cpu().numpy(), g['ep%d.qpos' % ei][1]) < 1e-7
        assert helpers.relerr(vd[0].cpu().numpy(), g['ep%d.qvel' % ei][1]) < 1e-6
        assert helpers.relerr(obs[0].cpu().numpy(), g['ep%d.obs' % ei][1]) < 1e-6
        assert abs(hz[0].item() - g['ep%d.head_z' % ei][0]) < 1e-8


def test_expert_features_vs_reference_golden(model, golden):
    g = golden('env_traj')
    X = cphys.X
    for ti in range(2):
        rows, lb = model.expert_features(g['takes_qpos'][ti])
        r = rows.cpu().numpy()
        for key, col, n in (('qvel', 'QVEL', 58), ('rlinv_local', 'RLINV_LOCAL', 3), ('rangv', 'RANGV', 3),
                            ('rq_rmh', 'RQ_RMH', 4), ('ee_pos', 'EE_POS', 15), ('bquat', 'BQUAT', 84),
                            ('bangvel', 'BANGVEL', 63)):
            ref = g['expert.' + key][ti]
            assert np.allclose(r[:, X[col]:X[col] + n], ref, rtol=1e-9, atol=1e-9), key
        assert abs(lb - g['expert.head_height_lb'][ti]) < 1e-12
