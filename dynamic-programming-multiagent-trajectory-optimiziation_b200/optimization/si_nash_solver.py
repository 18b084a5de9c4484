"""SI_NashSolver -- iterative best response for 3-D single-integrator agents with the reference's interface
(SCvx/optimization/si_nash_solver.py:16-125; `show_progress` is accepted and ignored: there is no progress bar here).
Same sweep machinery as `NashSolver` (optimization/nash_solver.py), positions are the first three states."""
from .nash_solver import _BestResponseSweeps
from .si_agent_best_response import SI_AgentBestResponse


class SI_NashSolver(_BestResponseSweeps):  # noqa: N801
    _RESPONSE = SI_AgentBestResponse
    _D = 3

    def _agent_line(self, i, moved):
        return f"  Agent {i}: dX={moved:.2e}"
