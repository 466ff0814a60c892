"""egopose_b200.metrics (pose / velocity distance, smoothness of ego_pose/eval_pose.py 'stats' mode) vs the reference's own
ego_pose/utils/metrics.py functions (tests/golden/pose_metrics.npz)."""
import numpy as np

from egopose_b200 import metrics


def test_metrics_match_reference(golden):
    g = golden('pose_metrics')
    dt = 1 / 30.0
    res = {'traj_pred': {}, 'traj_orig': {}}
    for i in range(2):
        gt, pred = g['gt%d' % i], g['pred%d' % i]
        assert np.allclose(metrics.get_joint_angles(pred), g['angs%d' % i], rtol=1e-12, atol=1e-12)
        vels = metrics.get_joint_vels(pred, dt)
        assert np.allclose(vels, g['vels%d' % i], rtol=1e-10, atol=1e-10)
        assert np.allclose(metrics.get_joint_accels(vels, dt), g['accels%d' % i], rtol=1e-9, atol=1e-8)
        res['traj_pred']['t%d' % i], res['traj_orig']['t%d' % i] = pred, gt
    m = metrics.compute_metrics(res, dt)
    assert np.allclose([m['pose_dist'], m['vel_dist'], m['smoothness']], g['global'], rtol=1e-10)
    for i in range(2):
        assert np.allclose(m['per_take']['t%d' % i], g['metrics%d' % i], rtol=1e-10)


def test_remove_noisy_hands_and_perfect_prediction():
    rng = np.random.RandomState(0)
    q = rng.randn(9, 59)
    q[:, 3:7] /= np.linalg.norm(q[:, 3:7], axis=1, keepdims=True)
    res = {'traj_pred': {'a': q.copy()}, 'traj_orig': {'a': q.copy()}}
    metrics.remove_noisy_hands(res)
    assert not res['traj_pred']['a'][:, 32:35].any() and not res['traj_orig']['a'][:, 42:45].any()
    m = metrics.compute_metrics(res)
    assert m['pose_dist'] == 0.0 and m['vel_dist'] == 0.0 and m['smoothness'] > 0.0
