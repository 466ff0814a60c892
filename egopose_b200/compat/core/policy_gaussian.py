from egopose_b200.nets import PolicyGaussian  # noqa: F401
