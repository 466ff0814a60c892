from egopose_b200.trajbatch import TrajBatch  # noqa: F401
