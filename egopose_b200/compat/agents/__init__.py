from egopose_b200.agent import Agent, AgentPG, AgentPPO  # noqa: F401
