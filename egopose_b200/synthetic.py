"""Seeded synthetic inputs for benchmarks / smoke runs (the EgoPose dataset is not redistributable):
smooth expert qpos trajectories inside the joint ranges and N(0,1) CNN features."""
import math

import numpy as np


def synthetic_takes(md, n_takes, length, seed=1, dt=1.0 / 30):
    """md: egopose_b200.mjcf.ModelDesc.  Root xy random walk (<= 1 m/s), z = 0.90 + 0.02 sin, yaw(t) with a
    small tilt, per-joint sinusoids (0.2-1 Hz, amplitude 1/4 of the range, centred), hands zeroed
    (gen_expert.py:38-39)."""
    rng = np.random.RandomState(seed)
    nq = md.nq
    rngs = np.array(md.jnt_range[1:])
    t = np.arange(length) * dt
    takes = []
    for _ in range(n_takes):
        q = np.zeros((length, nq))
        v = rng.uniform(-1, 1, size=2) * 0.5
        ph = rng.uniform(0, 2 * math.pi, size=4)
        q[:, 0] = v[0] * t + 0.1 * np.sin(0.5 * t + ph[0])
        q[:, 1] = v[1] * t + 0.1 * np.sin(0.4 * t + ph[1])
        q[:, 2] = 0.90 + 0.02 * np.sin(2.0 * t + ph[2])
        yaw = rng.uniform(-math.pi, math.pi) + 0.3 * np.sin(0.3 * t + ph[3])
        tilt = 0.05 * np.sin(1.1 * t + ph[0])
        q[:, 3] = np.cos(yaw / 2) * np.cos(tilt / 2)
        q[:, 4] = np.cos(yaw / 2) * np.sin(tilt / 2)
        q[:, 5] = np.sin(yaw / 2) * np.sin(tilt / 2)
        q[:, 6] = np.sin(yaw / 2) * np.cos(tilt / 2)
        freq = rng.uniform(0.2, 1.0, size=nq - 7) * 2 * math.pi
        phase = rng.uniform(0, 2 * math.pi, size=nq - 7)
        mid = 0.5 * (rngs[:, 0] + rngs[:, 1])
        amp = 0.25 * 0.5 * (rngs[:, 1] - rngs[:, 0])
        q[:, 7:] = mid + amp * np.sin(freq * t[:, None] + phase)
        for hand in ('LeftHand', 'RightHand'):
            a = md.body_qposadr[md.body_names.index(hand)]
            q[:, a:a + 3] = 0.0
        takes.append(q)
    return takes


def synthetic_cnn_feat(n_takes, length, dim=128, seed=100):
    return [np.random.RandomState(seed + i).randn(length, dim) for i in range(n_takes)]
