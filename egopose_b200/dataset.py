"""On-disk inputs of the training scripts and a synthetic dataset in that layout.

  datasets/meta/<meta_id>.yml                  take lists read by Config (egomimic_config.py:33-36): keys ``train`` / ``test``
  datasets/features/expert_<id>.p              {take: expert dict}, gen_expert.py:99-100   (egopose_b200.gen_expert)
  datasets/features/cnn_feat_<id>.p            (dict take -> f64[L, F], meta), gen_cnn_feature.py:68-70

The EgoPose dataset is not redistributable, so ``write_synthetic_dataset`` produces seeded stand-ins of the same shapes
(SURVEY 8d): smooth in-range mocap qpos per take, expert features computed from it by the GPU ``gen_expert`` kernel (the same
code path a real dataset takes), N(0, 1) CNN features.  ego_mimic.py / ego_forecast.py then run unchanged on it.
"""
import datetime
import os
import pickle

import numpy as np


def write_cnn_feat_file(path, cnn_features, cfg='synthetic', it=0, meta_id=None):
    """gen_cnn_feature.py:68-70: pickle of (cnn_features, meta)"""
    meta = {'cfg': cfg, 'iter': it, 'meta': meta_id, 'time': datetime.datetime.now()}
    os.makedirs(os.path.dirname(path) or '.', exist_ok=True)
    with open(path, 'wb') as f:
        pickle.dump(({k: np.asarray(v, dtype=np.float64) for k, v in cnn_features.items()}, meta), f)
    return meta


def read_cnn_feat_file(path):
    with open(path, 'rb') as f:
        feats, meta = pickle.load(f)
    return feats, meta


def write_meta_yml(path, train, test, extra=None):
    """the keys Config reads (egomimic_config.py:35-36); real meta files carry more (capture sync, object ids), unused here"""
    import yaml
    os.makedirs(os.path.dirname(path) or '.', exist_ok=True)
    d = {'train': list(train), 'test': list(test)}
    d.update(extra or {})
    with open(path, 'w') as f:
        yaml.safe_dump(d, f)


def write_synthetic_dataset(root, cfg, env=None, n_train=4, n_test=1, length=None, feat_dim=128, seed=1):
    """Writes datasets/{meta,features} under ``root`` for the Config ``cfg`` (its meta_id / expert_feat / cnn_feat ids).
    ``env``: an egopose_b200.env.HumanoidEnv used for the GPU gen_expert pass (created from cfg when None; needs CUDA).
    Returns the dict of written paths."""
    from .gen_expert import gen_expert_dict, write_expert_file
    from .mjcf import load_builtin
    from .synthetic import synthetic_cnn_feat, synthetic_takes
    md = load_builtin(getattr(cfg, 'mujoco_model', 'humanoid_1205_v1'))
    L = length or (cfg.env_episode_len + 2 * cfg.fr_margin + 64)
    names = ['synth_%02d' % i for i in range(n_train + n_test)]
    trajs = synthetic_takes(md, len(names), L, seed=seed)
    feats = synthetic_cnn_feat(len(names), L, dim=feat_dim, seed=100 + seed)
    cid = cfg.cfg_dict
    paths = {'meta': os.path.join(root, 'datasets', 'meta', '%s.yml' % cfg.meta_id),
             'expert': os.path.join(root, 'datasets', 'features', 'expert_%s.p' % cid['expert_feat']),
             'cnn': os.path.join(root, 'datasets', 'features', 'cnn_feat_%s.p' % cid['cnn_feat'])}
    write_meta_yml(paths['meta'], names[:n_train], names[n_train:])
    write_cnn_feat_file(paths['cnn'], dict(zip(names, feats)), meta_id=cfg.meta_id)
    own = env is None
    if own:
        from .env import HumanoidEnv
        env = HumanoidEnv(cfg)
    os.makedirs(os.path.dirname(paths['expert']), exist_ok=True)
    write_expert_file(paths['expert'], gen_expert_dict(env, names, trajs))
    if own:
        env.close()
    return paths
