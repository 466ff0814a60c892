from egopose_b200.trajbatch import TrajBatchEgo  # noqa: F401
