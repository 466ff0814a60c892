from egopose_b200.nets import Value  # noqa: F401
