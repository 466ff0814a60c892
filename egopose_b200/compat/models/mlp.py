from egopose_b200.nets import MLP  # noqa: F401
