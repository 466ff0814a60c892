from egopose_b200.agent import Agent  # noqa: F401
