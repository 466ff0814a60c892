import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ASSETS = os.path.join(ROOT, 'egopose_b200', 'assets')


def cfg_dict(task='egomimic', cfg_id='subject_03'):
    return json.load(open(os.path.join(ASSETS, '%s_%s.cfg.json' % (task, cfg_id))))


def make_model(device=0, cfg=None):
    from egopose_b200 import lib
    from egopose_b200.mjcf import load_builtin
    return lib.Model.from_cfg_dict(load_builtin(), cfg or cfg_dict(), device=device)


def rand_states(n, seed=0, vel=1.0):
    rng = np.random.RandomState(seed)
    q = np.zeros((n, 59))
    q[:, :2] = rng.randn(n, 2)
    q[:, 2] = 0.9 + 0.05 * rng.randn(n)
    quat = rng.randn(n, 4)
    q[:, 3:7] = quat / np.linalg.norm(quat, axis=1, keepdims=True)
    q[:, 7:] = rng.uniform(-0.5, 0.5, size=(n, 52))
    return q, rng.randn(n, 58) * vel


def policy_weights(D, H1, H2, A, seed=1, log_std=-2.3):
    """nn.Linear default init (uniform +-1/sqrt(fan_in)); heads x0.1, bias 0 (policy_gaussian.py:14-16)"""
    rng = np.random.RandomState(seed)

    def lin(o, i):
        b = 1.0 / np.sqrt(i)
        return rng.uniform(-b, b, size=(o, i)), rng.uniform(-b, b, size=o)
    W1, b1 = lin(H1, D)
    W2, b2 = lin(H2, H1)
    W3, b3 = lin(A, H2)
    return dict(W1=W1, b1=b1, W2=W2, b2=b2, W3=W3 * 0.1, b3=b3 * 0.0, log_std=np.full((1, A), log_std))


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
