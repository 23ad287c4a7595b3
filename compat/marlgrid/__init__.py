"""`marlgrid` import alias for code written against kandouss/marlgrid: put `<repo>/compat` on PYTHONPATH and

    import marlgrid.envs; from marlgrid.agents import GridAgentInterface; from marlgrid import IndependentLearners

resolve to the B200-native package (marlgrid_b200).  Kept out of the repo root so that it can never shadow the real
reference when the oracle harness imports it from /root/reference.
"""
import marlgrid_b200 as _impl
from marlgrid_b200 import GridAgentInterface, IndependentLearners, LearningAgent, agents, objects  # noqa: F401

__version__ = _impl.__version__


def __getattr__(name):
    return getattr(_impl, name)
