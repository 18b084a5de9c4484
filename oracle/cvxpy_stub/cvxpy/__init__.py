"""Inert stand-in for cvxpy -- TEST INFRASTRUCTURE ONLY.

The build container has no cvxpy.  Putting this directory on sys.path lets the reference's
SCvx.models.* / SCvx.discretization.* modules import unmodified so that their stage-1/2 code
(f/A/B lambdas, FirstOrderHold, linearize_collision) can generate golden vectors
(tests/golden/make_golden.py).  Nothing here solves anything.
"""


class _Inert:
    def __init__(self, *a, **k):
        self.value = None


class Variable(_Inert):
    pass


class Parameter(_Inert):
    pass


class Constraint(_Inert):
    pass


class Expression(_Inert):
    pass


class Minimize(_Inert):
    pass


class Problem(_Inert):
    pass


class SolverError(Exception):
    pass
