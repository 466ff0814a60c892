from egopose_b200.torch_utils import *  # noqa: F401,F403
