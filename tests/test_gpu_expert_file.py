"""egopose_b200.gen_expert (GPU get_expert + pickle writer) vs the reference golden and the file round trip."""
import pickle

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')

from test_oracle_expert_file import KEYS  # noqa: E402


def test_get_expert_matches_reference_and_file_round_trip(golden, tmp_path):
    from egopose_b200 import gen_expert
    from egopose_b200.config import Config
    from egopose_b200.env import HumanoidEnv
    g = golden('expert_file')
    cfg = Config('subject_03')
    env = HumanoidEnv(cfg, device=0)
    lb, ub = int(g['lb']), int(g['ub'])
    ex = gen_expert.get_expert(env, g['raw_qpos'], lb, ub)
    for k in KEYS:
        assert ex[k].shape == g[k].shape, k
        assert np.allclose(ex[k], g[k], rtol=1e-9, atol=1e-10), k
    assert ex['len'] == int(g['len']) and abs(ex['height_lb'] - float(g['height_lb'])) < 1e-14
    assert abs(ex['head_height_lb'] - float(g['head_height_lb'])) < 1e-12
    # gen_expert.py:86-100 loop + pickle; the file feeds load_experts (humanoid_v1.py:45-54) of both code bases
    d = gen_expert.gen_expert_dict(env, ['take_a', 'take_b'], [g['raw_qpos'], g['raw_qpos'][::-1].copy()],
                                   msync={'take_a': (0, lb, ub), 'take_b': (0, 1, 19)})
    path, cpath = tmp_path / 'expert_subject_03.p', tmp_path / 'cnn_feat_subject_03.p'
    gen_expert.write_expert_file(str(path), d)
    pickle.dump(({'take_a': np.zeros((ub - lb, 4)), 'take_b': np.ones((18, 4))}, {}), open(cpath, 'wb'))
    loaded = pickle.load(open(path, 'rb'))
    assert set(loaded['take_a']) == set(KEYS) | {'len', 'height_lb', 'head_height_lb'} and loaded['take_b']['len'] == 18
    env.load_experts(['take_a', 'take_b'], str(path), str(cpath))
    assert env.kernel.n_takes == 2 and int(env.kernel.take_off[-1]) == (ub - lb) + 18
    assert np.allclose(env.kernel.rows_host[:ub - lb, 59:59 + 58], g['qvel'], rtol=1e-9, atol=1e-10)
    env.close()
