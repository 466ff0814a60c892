from egopose_b200.nets import VideoForecastNet  # noqa: F401
