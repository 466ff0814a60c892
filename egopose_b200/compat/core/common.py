from egopose_b200.common import estimate_advantages  # noqa: F401
