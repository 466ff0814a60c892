from egopose_b200.nets import DiagGaussian  # noqa: F401
