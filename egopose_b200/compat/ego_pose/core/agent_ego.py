from egopose_b200.agent import AgentEgo  # noqa: F401
