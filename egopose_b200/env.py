"""Batched, device-resident mirror of ego_pose/envs/humanoid_v1.py HumanoidEnv (+ envs/common/mujoco_env.py).

One object stands for E independent humanoids living inside the fused rollout kernel; the attributes and
methods ego_mimic.py / AgentEgo touch keep their names and meaning:
  env.seed(s), env.load_experts(list, expert_file, cnn_file), env.cnn_feat, env.expert_list,
  env.end_reward, env.observation_space.shape, env.action_space.shape, env.model.actuator_names,
  env.set_fix_head_lb, env.dt, env.cfg
Per-step Python stepping (reset()/step()) is not offered: the step loop is the kernel (egp_rollout_f64).
Scope: smooth dynamics only - no floor contact / joint limits (north-star)."""
import os
import pickle

import numpy as np

from . import lib
from .mjcf import compile_mjcf, load_builtin


class _Space:
    def __init__(self, n):
        self.shape = (n,)
        self.low = -np.inf * np.ones(n)
        self.high = np.inf * np.ones(n)


class _ModelView:
    """the few mujoco_py model attributes callers read (ego_mimic.py:46, humanoid_v1.py:62)"""

    def __init__(self, md):
        self.nq, self.nv, self.nu = md.nq, md.nv, md.nu
        self.actuator_names = tuple(md.actuator_names)
        self.body_names = ('world',) + tuple(md.body_names)
        self.actuator_ctrlrange = np.zeros((md.nu, 2))


class HumanoidEnv:
    def __init__(self, cfg, device=0):
        self.cfg = cfg
        path = getattr(cfg, 'mujoco_model_file', None)
        if path and os.path.exists(path):
            self.md = compile_mjcf(path)                    # envs/common/mujoco_env.py:22
        else:
            self.md = load_builtin(getattr(cfg, 'mujoco_model', 'humanoid_1205_v1'))
        self.frame_skip = 15                                # humanoid_v1.py:16
        if float(getattr(cfg, 'env_init_noise', 0.0) or 0.0) != 0.0:
            # humanoid_v1.py:222 adds N(0, env_init_noise) to qpos[7:] at reset; every shipped yml leaves it at 0
            raise lib.EgpError('env_init_noise != 0 is not implemented by the fused reset (humanoid_v1.py:222)')
        self.device = device
        self.kernel = lib.Model(self.md, cfg.jkp, cfg.jkd, cfg.a_ref, cfg.a_scale, cfg.torque_lim,
                                getattr(cfg, 'b_diffw', np.ones(self.md.nbody - 1)), cfg.reward_weights,
                                frame_skip=self.frame_skip, device=device)
        # joint ranges of the XML as MuJoCo soft constraints: off by default (the fused kernels' scope is the smooth dynamics);
        # cfg.joint_limits = True (or EGP_JOINT_LIMITS=1) switches them on for every roll-out of this environment
        if bool(getattr(cfg, 'joint_limits', False)) or os.environ.get('EGP_JOINT_LIMITS', '0') == '1':
            self.kernel.set_joint_limits(True)
        # floor contact (geom-floor pairs, pyramidal friction cones): cfg.floor_contact = True or EGP_FLOOR_CONTACT=1
        if bool(getattr(cfg, 'floor_contact', False)) or os.environ.get('EGP_FLOOR_CONTACT', '0') == '1':
            self.kernel.set_contacts(True)
        self.model = _ModelView(self.md)
        self.obs_dim = self.md.nq - 2 + self.md.nv
        self.observation_space = _Space(self.obs_dim)
        self.action_space = _Space(self.md.nu)
        self.end_reward = 0.0
        self.fix_head_lb = None
        self.expert_list = None
        self.cnn_feat = None
        self.np_random = np.random.RandomState()
        self._seed = 0
        self.body_qposaddr = self.md.body_qposaddr()

    @property
    def dt(self):
        return self.md.timestep * self.frame_skip          # mujoco_env.py:103-105

    def seed(self, seed=None):
        self._seed = 0 if seed is None else int(seed)
        self.np_random = np.random.RandomState(seed)
        return [seed]

    def set_fix_head_lb(self, fix_head_lb=None):
        self.fix_head_lb = fix_head_lb

    # ---- experts ---------------------------------------------------------------------------------
    def load_experts(self, expert_list, expert_feat_file, cnn_feat_file):
        """humanoid_v1.py:45-54: reads the pickles written by gen_expert.py:99-100 / gen_cnn_feature.py:68-70"""
        expert_dict = pickle.load(open(expert_feat_file, 'rb'))
        cnn_feat_dict, _ = pickle.load(open(cnn_feat_file, 'rb'))
        self.set_experts(expert_list, [expert_dict[x] for x in expert_list], [cnn_feat_dict[x] for x in expert_list])

    def set_experts(self, expert_list, expert_arr, cnn_feat):
        X = lib.X
        rows, off, lbs = [], [0], []
        for ex in expert_arr:
            L = ex['qpos'].shape[0]
            r = np.zeros((L, X['STRIDE']))
            for key, col in (('qpos', 'QPOS'), ('qvel', 'QVEL'), ('rlinv_local', 'RLINV_LOCAL'), ('rangv', 'RANGV'),
                             ('rq_rmh', 'RQ_RMH'), ('ee_pos', 'EE_POS'), ('bquat', 'BQUAT'), ('bangvel', 'BANGVEL')):
                a = np.asarray(ex[key], dtype=np.float64)
                r[:, X[col]:X[col] + a.shape[1]] = a
            rows.append(r)
            off.append(off[-1] + L)
            lbs.append(float(ex['head_height_lb']))
        self.expert_list = list(expert_list)
        self.cnn_feat = [np.asarray(c, dtype=np.float64) for c in cnn_feat] if cnn_feat is not None else None
        ctx = np.concatenate(self.cnn_feat) if self.cnn_feat is not None else None
        self.kernel.upload_experts(np.concatenate(rows), off, lbs, ctx)

    def set_expert_qpos(self, expert_list, takes_qpos, cnn_feat=None):
        """build the expert tables from raw qpos trajectories with the GPU gen_expert kernel
        (egp_expert_features_f64; hands are zeroed as gen_expert.py:38-39 does)"""
        rows, off, lbs = [], [0], []
        for q in takes_qpos:
            q = np.array(q, dtype=np.float64, copy=True)
            for hand in ('LeftHand', 'RightHand'):
                a, b = self.body_qposaddr[hand]
                q[:, a:b] = 0.0
            r, lb = self.kernel.expert_features(q)
            rows.append(r.cpu().numpy())
            off.append(off[-1] + q.shape[0])
            lbs.append(lb)
        self.expert_list = list(expert_list)
        self.cnn_feat = [np.asarray(c, dtype=np.float64) for c in cnn_feat] if cnn_feat is not None else None
        ctx = np.concatenate(self.cnn_feat) if self.cnn_feat is not None else None
        self.kernel.upload_experts(np.concatenate(rows), off, lbs, ctx)

    def close(self):
        self.kernel.close()
