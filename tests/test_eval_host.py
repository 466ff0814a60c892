"""Host-side logic of egopose_b200.evaluate that needs no GPU: sync_traj / window start states (ego_forecast_eval.py:107-136)
against the golden produced with the reference's own sync_traj, and the observation table read off the packed expert rows
(HumanoidEnv.get_obs of every expert frame) against the oracle env."""
import types

import numpy as np

from egopose_b200 import evaluate, lib
from oracle import cphys, evalloop


def test_forecast_window_init_matches_reference_golden(golden):
    g = golden('eval_forecast')
    fm, T, emo = int(g['fr_margin']), int(g['test_len']), int(g['em_offset'])
    for wi, s0 in enumerate(int(x) for x in g['starts']):
        q0, v0, past = evaluate.forecast_window_init(g['qpos'], g['em_traj'], g['em_vel'], s0, fm, T, emo)
        # the golden's first simulated row is the state the reference set with env.set_state(qpos, qvel)
        assert np.allclose(q0, g['traj_pred_em'][wi, fm], rtol=1e-12, atol=1e-12)
        assert np.allclose(past, g['traj_pred_em'][wi, :fm], rtol=1e-12, atol=1e-12)
        oq, ov, op = evalloop.forecast_init(g['qpos'], g['em_traj'], g['em_vel'], s0, fm, T, emo)
        assert np.allclose(v0, ov, rtol=1e-12, atol=1e-12) and np.allclose(q0, oq, rtol=1e-12, atol=1e-12)
    # first window: the prediction does not reach fm frames back -> unsynced prediction, expert rows fill the missing past
    assert np.array_equal(evaluate.forecast_window_init(g['qpos'], g['em_traj'], g['em_vel'], int(g['starts'][0]), fm, T, emo)[2][:emo],
                          g['qpos'][:emo])


def test_sync_traj_aligns_first_frame():
    rng = np.random.RandomState(0)
    qp = rng.randn(6, 59)
    qp[:, 3:7] /= np.linalg.norm(qp[:, 3:7], axis=1, keepdims=True)
    qv = rng.randn(6, 58)
    ref = rng.randn(59)
    ref[3:7] /= np.linalg.norm(ref[3:7])
    sp, sv = evaluate.sync_traj(qp, qv, ref)
    op, ov = evalloop.sync_traj(qp, qv, ref)
    assert np.allclose(sp, op, atol=1e-13) and np.allclose(sv, ov, atol=1e-13)
    assert np.allclose(sp[0, :2], ref[:2], atol=1e-13)                                   # xy of the first frame = reference
    assert np.allclose(evalloop.heading_q(sp[0, 3:7]), evalloop.heading_q(ref[3:7]), atol=1e-12)
    assert np.allclose(sp[:, 2], qp[:, 2]) and np.allclose(sp[:, 7:], qp[:, 7:]) and np.allclose(sv[:, 3:], qv[:, 3:])


def test_expert_obs_table_is_the_env_observation():
    orc = cphys.Oracle()
    takes = cphys.synthetic_takes(orc.md, 2, 12, seed=4)
    orc.make_expert(takes)
    model = types.SimpleNamespace(rows_host=np.asarray(orc._keep['x_rows']), nq=orc.nq, nv=orc.nv)
    tab = evaluate.expert_obs_table(model)
    assert tab.shape == (24, orc.S)
    X = lib.X
    for fr in (0, 5, 13, 23):
        env = cphys.EoEnv()
        row = model.rows_host[fr]
        orc.env_set_state(env, row[X['QPOS']:X['QPOS'] + orc.nq], row[X['QVEL']:X['QVEL'] + orc.nv])
        assert np.allclose(tab[fr], orc.env_obs(env), rtol=1e-12, atol=1e-12)
