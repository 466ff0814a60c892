"""ego_pose/core/reward_function.py:78-80 registry.  The rewards are computed inside the rollout kernel;
the registry entries are markers that AgentEgo(custom_reward=...) accepts."""


def quat_space_reward_v3(env, state, action, info):
    raise RuntimeError('quat_v3 is evaluated inside egp_rollout_f64 (csrc/rollout.cu env_reward)')


reward_func = {'quat_v3': quat_space_reward_v3}
