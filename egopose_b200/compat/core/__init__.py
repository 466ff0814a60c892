from egopose_b200.logger_rl import LoggerRL  # noqa: F401
from egopose_b200.trajbatch import TrajBatch  # noqa: F401
from egopose_b200.common import estimate_advantages  # noqa: F401
