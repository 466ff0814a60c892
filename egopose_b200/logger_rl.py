"""core/logger_rl.py:4-59 mirror; ``from_device`` fills the fields ego_mimic.py:112,120-131 reads from the
rollout kernel's reduction vector (include/egopose_b200.h EGP_LOG_*)."""
import math

import numpy as np

from .lib import LOG


class LoggerRL:
    def __init__(self):
        self.num_steps = 0
        self.num_episodes = 0
        self.total_reward = 0
        self.min_episode_reward = math.inf
        self.max_episode_reward = -math.inf
        self.total_c_reward = 0
        self.min_c_reward = math.inf
        self.max_c_reward = -math.inf
        self.episode_reward = 0
        self.avg_episode_reward = 0
        self.avg_c_reward = 0
        self.total_c_info = 0
        self.avg_c_info = 0
        self.sample_time = 0
        self.num_nan_resets = 0

    def end_sampling(self):
        self.avg_episode_reward = self.total_reward / self.num_episodes
        self.avg_c_reward = self.total_c_reward / self.num_steps
        self.avg_c_info = self.total_c_info / self.num_steps

    @classmethod
    def from_device(cls, vec):
        v = np.asarray(vec, dtype=np.float64)
        lg = cls()
        lg.num_steps = int(v[LOG['NUM_STEPS']])
        lg.num_episodes = int(v[LOG['NUM_EPISODES']])
        lg.total_reward = float(v[LOG['TOTAL_REWARD']])
        lg.total_c_reward = float(v[LOG['TOTAL_C_REWARD']])
        lg.min_c_reward, lg.max_c_reward = float(v[LOG['MIN_C_REWARD']]), float(v[LOG['MAX_C_REWARD']])
        lg.total_c_info = v[LOG['C_INFO']:LOG['C_INFO'] + 5].copy()
        lg.min_episode_reward = float(v[LOG['MIN_EPISODE_REWARD']])
        lg.max_episode_reward = float(v[LOG['MAX_EPISODE_REWARD']])
        lg.num_nan_resets = int(v[LOG['NUM_NAN_RESETS']])
        lg.end_sampling()
        return lg

    @classmethod
    def merge(cls, logger_list):
        """core/logger_rl.py:43-59 (including its min_episode_reward = max(...) quirk)"""
        lg = cls()
        lg.total_reward = sum(x.total_reward for x in logger_list)
        lg.num_episodes = sum(x.num_episodes for x in logger_list)
        lg.num_steps = sum(x.num_steps for x in logger_list)
        lg.avg_episode_reward = lg.total_reward / lg.num_episodes
        lg.max_episode_reward = max(x.max_episode_reward for x in logger_list)
        lg.min_episode_reward = max(x.min_episode_reward for x in logger_list)
        lg.total_c_reward = sum(x.total_c_reward for x in logger_list)
        lg.avg_c_reward = lg.total_c_reward / lg.num_steps
        lg.max_c_reward = max(x.max_c_reward for x in logger_list)
        lg.min_c_reward = min(x.min_c_reward for x in logger_list)
        lg.total_c_info = sum(x.total_c_info for x in logger_list)
        lg.avg_c_info = lg.total_c_info / lg.num_steps
        lg.num_nan_resets = sum(getattr(x, 'num_nan_resets', 0) for x in logger_list)
        return lg
