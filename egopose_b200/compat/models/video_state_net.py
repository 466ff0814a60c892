from egopose_b200.nets import VideoStateNet  # noqa: F401
