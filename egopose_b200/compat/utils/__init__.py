"""utils/__init__.py:1-7 star-exports the caller relies on (hot-path subset)"""
from egopose_b200.zfilter import ZFilter, RunningStat  # noqa: F401
from egopose_b200.torch_utils import *  # noqa: F401,F403
from egopose_b200.torch_utils import tensor, zeros, ones, to_cpu, to_device, to_test, to_train, batch_to, set_optimizer_lr  # noqa: F401
import numpy as np  # noqa: F401
import torch  # noqa: F401
from egopose_b200.run_logging import Logger, create_logger  # noqa: F401,E402  (utils/logger.py, utils/tb_logger.py)
from egopose_b200.torch_utils import filter_state_dict  # noqa: F401,E402
import math  # noqa: F401,E402
import os  # noqa: F401,E402
