from egopose_b200.logger_rl import LoggerRL  # noqa: F401
