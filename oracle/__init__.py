"""CPU oracle for the SCvx inner loop -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / the reported CPU baseline.  The product
package (``scvx_b200``) never imports this package and fails loudly when its CUDA
extension is missing.

What is restated here (numpy / scipy, fp64), each function citing the reference
file:line it follows (paths relative to the upstream reference checkout):

* ``oracle.models``      -- unicycle / single-integrator f, A, B; obstacle and inter-agent
                            half-space linearisation (SCvx/models/*.py)
* ``oracle.foh``         -- first-order-hold discretisation of the augmented ODE
                            (SCvx/discretization/first_order_hold.py)
* ``oracle.subproblem``  -- the convex sub-problem of SCProblem / AgentSolver as an explicit
                            LP / QP solved by scipy's vendored HiGHS
                            (SCvx/optimization/sc_problem.py, agent_solver.py)
* ``oracle.scvx``        -- outer SCvx loop and the ADMM consensus coordinator
                            (SCvx/optimization/scvx_solver.py, admm_coordinator.py)
* ``oracle.distopt``     -- Distributed_opt/ADMM_decentralized.py and dist_scvx_3d.py
                            ``x_traj_opt`` restated as QPs

Parity pinning
--------------
Stages 1-2 (FOH, f/A/B, collision linearisation): PINNED.  ``tests/golden/*.npz`` were produced
by importing the reference's own Python modules in the build container
(``tests/golden/make_golden.py`` -- committed), and ``tests/test_oracle_golden.py`` checks this
restatement against them, plus the reference's own known-answer tests
(SCvx/multi_agent_tests/test_multi_agent_model.py:42-59, test_admm_utils.py:7-45).

Stage 3 (convex sub-problem): PARITY UNPINNED against cvxpy+ECOS/CLARABEL -- those packages are
not installed in the build container and the reference's tests at that boundary assert only
shapes.  The restatement follows sc_problem.py:21-83 line by line and is solved by an exact LP/QP
solver (HiGHS); the optimal VALUE of a convex programme is solver independent, which is what the
parity tests compare.
"""
