from egopose_b200.nets import Policy  # noqa: F401
