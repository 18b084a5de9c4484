"""SI_ADMMCoordinator -- mirrors SCvx/optimization/si_admm_coordinator.py:13-127 (3-D twin; passes the
INITIAL X_refs/U_refs to setup() every round, si_admm_coordinator.py:80-86)."""
from .admm_coordinator import ADMMCoordinator
from .si_agent_solver import SI_AgentSolver


class SI_ADMMCoordinator(ADMMCoordinator):  # noqa: N801
    _SOLVER = SI_AgentSolver
    _D = 3
    _SI_VARIANT = True
