"""marlgrid.agents -> marlgrid_b200.agents (reference: marlgrid/agents.py)."""
from marlgrid_b200.agents import *  # noqa: F401,F403
from marlgrid_b200.agents import GridAgentInterface, IndependentLearners, LearningAgent  # noqa: F401
