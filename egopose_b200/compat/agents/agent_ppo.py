from egopose_b200.agent import AgentPPO  # noqa: F401
