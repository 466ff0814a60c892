from egopose_b200.zfilter import ZFilter, RunningStat  # noqa: F401
