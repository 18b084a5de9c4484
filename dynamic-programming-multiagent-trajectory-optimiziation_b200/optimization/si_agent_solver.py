"""SI_AgentSolver -- mirrors SCvx/optimization/si_agent_solver.py:10-105 (3-D single-integrator agents)."""
from .agent_solver import AgentSolver


class SI_AgentSolver(AgentSolver):  # noqa: N801
    _D = 3
