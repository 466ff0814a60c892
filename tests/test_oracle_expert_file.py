"""Oracle gen_expert restatement (rows + the keys the rollout never reads) vs tests/golden/expert_file.npz, produced by
the reference's own env methods / math helpers in the order of ego_pose/data_process/gen_expert.py:28-83."""
import numpy as np

from oracle import cphys

X = cphys.X


def oracle_expert(orc, raw_qpos, lb, ub):
    q = np.array(raw_qpos, copy=True)
    q[:, 32:35] = 0.0           # LeftHand / RightHand qpos slices (gen_expert.py:38-39)
    q[:, 42:45] = 0.0
    rows, _ = orc.expert_features(q)
    hp, com, eew = orc.expert_extras(q)
    col = lambda n, w: rows[lb:ub, X[n]:X[n] + w]  # noqa: E731
    ex = {'qpos': q[lb:ub], 'qvel': col('QVEL', 58), 'rlinv': col('QVEL', 3), 'rlinv_local': col('RLINV_LOCAL', 3),
          'rangv': col('RANGV', 3), 'rq_rmh': col('RQ_RMH', 4), 'ee_pos': col('EE_POS', 15), 'bquat': col('BQUAT', 84),
          'bangvel': col('BANGVEL', 63), 'head_pos': hp[lb:ub], 'com': com[lb:ub], 'ee_wpos': eew[lb:ub]}
    ex['obs'] = np.concatenate([q[lb:ub, 2:3], ex['rq_rmh'], q[lb:ub, 7:], np.zeros((ub - lb, 58))], axis=1)
    return ex


KEYS = ['qpos', 'qvel', 'rlinv', 'rlinv_local', 'rangv', 'rq_rmh', 'com', 'head_pos', 'obs', 'ee_pos', 'ee_wpos', 'bquat', 'bangvel']


def test_oracle_expert_dict_matches_reference(golden):
    g = golden('expert_file')
    ex = oracle_expert(cphys.Oracle(), g['raw_qpos'], int(g['lb']), int(g['ub']))
    for k in KEYS:
        assert ex[k].shape == g[k].shape, k
        assert np.allclose(ex[k], g[k], rtol=1e-10, atol=1e-10), k
    assert abs(ex['head_pos'][:, 2].min() - float(g['head_height_lb'])) < 1e-13
