"""ego_pose/utils/egomimic_config.py:9-131 mirror: yml (or dict) -> attribute bag with the same names and
defaults.  Differences: the meta yml (datasets/meta/<id>.yml, absent without the dataset) is optional, and
no result directories are created unless ``create_dirs``."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class Config:
    def __init__(self, cfg_id=None, create_dirs=False, cfg_dict=None, task='egomimic', base_dir='results'):
        self.id = cfg_id
        if cfg_dict is not None:
            cfg = cfg_dict
        else:
            yml = 'config/%s/%s.yml' % (task, cfg_id)
            if os.path.exists(yml):
                import yaml
                cfg = yaml.safe_load(open(yml, 'r'))
            else:       # constants compiled from the reference's yml by tools/compile_model.py
                cfg = json.load(open(os.path.join(HERE, 'assets', '%s_%s.cfg.json' % (task, cfg_id))))
        self.cfg_dict = cfg
        self.base_dir = base_dir
        self.cfg_dir = '%s/%s/%s' % (self.base_dir, task, cfg_id)
        self.model_dir = '%s/models' % self.cfg_dir
        self.result_dir = '%s/results' % self.cfg_dir
        self.log_dir = '%s/log' % self.cfg_dir
        self.tb_dir = '%s/tb' % self.cfg_dir
        if create_dirs:
            for d in (self.model_dir, self.result_dir, self.log_dir, self.tb_dir):
                os.makedirs(d, exist_ok=True)
        # data
        self.meta_id = cfg.get('meta_id')
        self.data_dir = 'datasets'
        meta_file = '%s/meta/%s.yml' % (self.data_dir, self.meta_id)
        if os.path.exists(meta_file):
            import yaml
            self.meta = yaml.safe_load(open(meta_file, 'r'))
            self.takes = {x: self.meta[x] for x in ['train', 'test']}
        else:
            self.meta, self.takes = None, {'train': [], 'test': []}
        self.expert_feat_file = '%s/features/expert_%s.p' % (self.data_dir, cfg['expert_feat']) if 'expert_feat' in cfg else None
        self.cnn_feat_file = '%s/features/cnn_feat_%s.p' % (self.data_dir, cfg['cnn_feat']) if 'cnn_feat' in cfg else None
        self.fr_margin = cfg.get('fr_margin', 10)
        # ego mimic warm start (egoforecast_config.py:38-40)
        self.ego_mimic_cfg = cfg.get('ego_mimic_cfg', None)
        self.ego_mimic_iter = cfg.get('ego_mimic_iter', None)
        # training config
        self.gamma = cfg.get('gamma', 0.95)
        self.tau = cfg.get('tau', 0.95)
        self.causal = cfg.get('causal', False)
        self.policy_htype = cfg.get('policy_htype', 'relu')
        self.policy_hsize = cfg.get('policy_hsize', [300, 200])
        self.policy_v_hdim = cfg.get('policy_v_hdim', 128)
        self.policy_v_net = cfg.get('policy_v_net', 'lstm')
        # egoforecast only (egoforecast_config.py:49-52,62-65); the egomimic defaults select no state net
        self.policy_v_net_param = cfg.get('policy_v_net_param', None)
        self.policy_s_net = cfg.get('policy_s_net', 'id')
        self.policy_s_hdim = cfg.get('policy_s_hdim', None)
        self.policy_dyn_v = cfg.get('policy_dyn_v', False)
        self.value_v_net_param = cfg.get('value_v_net_param', None)
        self.value_s_net = cfg.get('value_s_net', 'id')
        self.value_s_hdim = cfg.get('value_s_hdim', None)
        self.value_dyn_v = cfg.get('value_dyn_v', False)
        self.end_reward = cfg.get('end_reward', True)
        self.obs_phase = cfg.get('obs_phase', False)
        self.random_cur_t = cfg.get('random_cur_t', False)
        self.policy_optimizer = cfg.get('policy_optimizer', 'Adam')
        self.policy_lr = cfg.get('policy_lr', 5e-5)
        self.policy_momentum = cfg.get('policy_momentum', 0.0)
        self.policy_weightdecay = cfg.get('policy_weightdecay', 0.0)
        self.value_htype = cfg.get('value_htype', 'relu')
        self.value_hsize = cfg.get('value_hsize', [300, 200])
        self.value_v_hdim = cfg.get('value_v_hdim', 128)
        self.value_v_net = cfg.get('value_v_net', 'lstm')
        self.value_optimizer = cfg.get('value_optimizer', 'Adam')
        self.value_lr = cfg.get('value_lr', 3e-4)
        self.value_momentum = cfg.get('value_momentum', 0.0)
        self.value_weightdecay = cfg.get('value_weightdecay', 0.0)
        self.adv_clip = cfg.get('adv_clip', np.inf)
        self.clip_epsilon = cfg.get('clip_epsilon', 0.2)
        self.log_std = cfg.get('log_std', -2.3)
        self.fix_std = cfg.get('fix_std', False)
        self.num_optim_epoch = cfg.get('num_optim_epoch', 10)
        self.min_batch_size = cfg.get('min_batch_size', 50000)
        self.max_iter_num = cfg.get('max_iter_num', 1000)
        self.seed = cfg.get('seed', 1)
        self.save_model_interval = cfg.get('save_model_interval', 100)
        self.reward_id = cfg.get('reward_id', 'quat')
        self.reward_weights = cfg.get('reward_weights', None)
        # adaptive parameters (egomimic_config.py:82-91)
        self.adp_iter_cp = np.array(cfg.get('adp_iter_cp', [0]))
        pad = lambda a: np.pad(np.array(a, dtype=np.float64), (0, self.adp_iter_cp.size - len(a)), 'edge')  # noqa: E731
        self.adp_noise_rate_cp = pad(cfg.get('adp_noise_rate_cp', [1.0]))
        self.adp_log_std_cp = pad(cfg.get('adp_log_std_cp', [self.log_std]))
        self.adp_policy_lr_cp = pad(cfg.get('adp_policy_lr_cp', [self.policy_lr]))
        self.adp_init_noise_cp = pad(cfg.get('adp_init_noise_cp', [0.0]))         # egoforecast_config.py:91-92
        self.adp_noise_rate = self.adp_log_std = self.adp_policy_lr = self.adp_init_noise = None
        # env config
        self.mujoco_model = cfg.get('mujoco_model', 'humanoid_1205_v1')
        self.mujoco_model_file = '%s/assets/mujoco_models/%s.xml' % (os.getcwd(), self.mujoco_model)
        self.env_start_first = cfg.get('env_start_first', False)
        self.env_init_noise = cfg.get('env_init_noise', 0.0)
        self.env_episode_len = cfg.get('env_episode_len', 200)
        self.obs_type = cfg.get('obs_type', 'full')
        self.obs_coord = cfg.get('obs_coord', 'heading')
        self.obs_heading = cfg.get('obs_heading', False)
        self.obs_vel = cfg.get('obs_vel', 'full')
        self.root_deheading = cfg.get('root_deheading', True)
        self.sync_exp_interval = cfg.get('sync_exp_interval', 100)
        self.action_type = cfg.get('action_type', 'position')
        if 'joint_params' in cfg:       # egomimic_config.py:105-116
            jp = [np.array(p) for p in zip(*cfg['joint_params'])]
            self.jkp, self.jkd, self.a_ref, self.a_scale, self.torque_lim = [p.astype(np.float64) for p in jp[1:6]]
            self.a_ref = np.deg2rad(self.a_ref)
            mult = cfg.get('jkp_multiplier', 1.0)
            self.jkp = self.jkp * mult
            self.jkd = self.jkd * cfg.get('jkd_multiplier', mult)
        if 'body_params' in cfg:        # egomimic_config.py:119-122
            self.b_diffw = np.array(list(zip(*cfg['body_params']))[1], dtype=np.float64)

    def update_adaptive_params(self, i_iter):
        """egomimic_config.py:124-131 piecewise-linear schedules"""
        cp = self.adp_iter_cp
        ind = np.where(i_iter >= cp)[0][-1]
        nind = ind + int(ind < len(cp) - 1)
        t = (i_iter - cp[ind]) / (cp[nind] - cp[ind]) if nind > ind else 0.0
        self.adp_noise_rate = self.adp_noise_rate_cp[ind] * (1 - t) + self.adp_noise_rate_cp[nind] * t
        self.adp_log_std = self.adp_log_std_cp[ind] * (1 - t) + self.adp_log_std_cp[nind] * t
        self.adp_policy_lr = self.adp_policy_lr_cp[ind] * (1 - t) + self.adp_policy_lr_cp[nind] * t
        inz = np.pad(self.adp_init_noise_cp, (0, max(0, len(cp) - len(self.adp_init_noise_cp))), 'edge')
        self.adp_init_noise = inz[ind] * (1 - t) + inz[nind] * t
