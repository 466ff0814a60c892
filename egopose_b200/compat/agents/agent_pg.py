from egopose_b200.agent import AgentPG  # noqa: F401
