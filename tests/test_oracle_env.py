"""C oracle env logic (obs / PD torque / step / done / reward / expert features) vs golden trajectories
produced by the reference's unmodified HumanoidEnv + quat_space_reward_v3 on the same restated physics
(tests/golden/make_golden.py gen_env).  Pins the env half of the oracle; MuJoCo itself stays unpinned."""
import numpy as np

from oracle import cphys

X = cphys.X


def _setup(golden):
    g = golden('env_traj')
    orc = cphys.Oracle(episode_len=int(g['episode_len']))
    takes = list(g['takes_qpos'])
    orc.make_expert(takes)
    return g, orc


def test_expert_features_match_reference_pipeline(golden):
    g, orc = _setup(golden)
    off = np.array(orc._keep['x_off'])
    rows = orc._keep['x_rows']
    for ti in range(2):
        r = rows[off[ti]:off[ti + 1]]
        for key, col, n in (('qvel', 'QVEL', 58), ('rlinv_local', 'RLINV_LOCAL', 3), ('rangv', 'RANGV', 3),
                            ('rq_rmh', 'RQ_RMH', 4), ('ee_pos', 'EE_POS', 15), ('bquat', 'BQUAT', 84),
                            ('bangvel', 'BANGVEL', 63)):
            ref = g['expert.' + key][ti]
            assert np.allclose(r[:, X[col]:X[col] + n], ref, rtol=1e-11, atol=1e-10), key
        assert abs(orc._keep['x_lb'][ti] - g['expert.head_height_lb'][ti]) < 1e-14


def test_env_episodes_match_reference(golden):
    g, orc = _setup(golden)
    for ei in range(2):
        take, start, head_lb, end_reward = g['ep%d.meta' % ei]
        orc.cfg.fix_head_lb = head_lb
        orc.cfg.end_reward = end_reward
        env = cphys.EoEnv()
        orc.env_reset(env, int(take), int(start))
        assert np.allclose(orc.env_obs(env), g['ep%d.obs' % ei][0], rtol=0, atol=1e-13)
        actions = g['ep%d.action' % ei]
        for t in range(actions.shape[0]):
            ctrl = np.array(orc._keep['a_ref']) + actions[t] * np.array(orc._keep['a_scale'])
            tq = orc.compute_torque(env.d, ctrl)
            assert np.allclose(tq, g['ep%d.torque0' % ei][t], rtol=1e-9, atol=1e-9)
            fail, end = orc.env_step(env, actions[t])
            assert fail == bool(g['ep%d.fail' % ei][t]) and end == bool(g['ep%d.end' % ei][t])
            qpos = np.array(env.d.qpos[:59])
            qvel = np.array(env.d.qvel[:58])
            assert np.allclose(qpos, g['ep%d.qpos' % ei][t + 1], rtol=1e-10, atol=1e-11)
            assert np.allclose(qvel, g['ep%d.qvel' % ei][t + 1], rtol=1e-9, atol=1e-9)
            assert np.allclose(orc.env_obs(env), g['ep%d.obs' % ei][t + 1], rtol=1e-9, atol=1e-9)
            rew, info = orc.env_reward(env, end)
            assert np.allclose(info, g['ep%d.c_info' % ei][t], rtol=1e-8, atol=1e-12)
            assert abs(rew - g['ep%d.reward' % ei][t]) < 1e-9 * max(1.0, abs(rew))
        assert fail or end
