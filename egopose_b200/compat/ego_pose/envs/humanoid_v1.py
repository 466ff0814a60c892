from egopose_b200.env import HumanoidEnv  # noqa: F401
