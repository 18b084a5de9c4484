"""Row a5 (BaseModel plug-in contract, SCvx/models/base_model.py:16-88): a model the library has never seen -- given only by
its sympy right-hand side, the way the reference's own models are written (unicycle_model.py:54-63) -- runs on the device.
Stage 1 is checked against the oracle's FOH (the reference's _ode_dVdt restated, tight tolerance) to 1e-9, then the model goes
through the whole loop and its first sub-problem is checked against the exact LP."""
import numpy as np
import pytest
import sympy as sp
import torch

from oracle import foh as ofoh, models as omodels, subproblem as ospb

pytestmark = pytest.mark.gpu


def _lambdas(x_syms, u_syms, f_expr):
    f = sp.Matrix(f_expr)
    A, B = f.jacobian(sp.Matrix(x_syms)), f.jacobian(sp.Matrix(u_syms))
    mk = lambda e: sp.lambdify((sp.Matrix(x_syms), sp.Matrix(u_syms)), e, "numpy")     # noqa: E731
    ff, fA, fB = mk(f), mk(A), mk(B)
    return (lambda x, u: np.asarray(ff(x, u), dtype=float).reshape(-1),
            lambda x, u: np.asarray(fA(x, u), dtype=float), lambda x, u: np.asarray(fB(x, u), dtype=float))


def _drag_unicycle_cls():
    from scvx_b200.models.base_model import BaseModel
    from scvx_b200.models.unicycle_model import UnicycleModel

    class DragUnicycle(UnicycleModel):
        """Unicycle whose forward speed saturates with drag and whose heading is pulled back towards 0: same states, inputs
        and constraints as the shipped unicycle, different dynamics -> compiled at run time."""
        device_model_id = BaseModel.device_model_id      # the generic hook instead of the shipped unicycle's constant
        c_d, k_th = 0.3, 0.2

        def symbolic_dynamics(self):
            x, y, th = sp.symbols("x y theta", real=True)
            v, w = sp.symbols("v w", real=True)
            f = sp.Matrix([v * sp.cos(th) * (1 - self.c_d * v), v * sp.sin(th) * (1 - self.c_d * v), w - self.k_th * sp.sin(th)])
            return [x, y, th], [v, w], f

        def get_equations(self):
            return _lambdas(*self.symbolic_dynamics())

    return DragUnicycle


class _OracleUserModel:
    """What OracleFOH needs from a model: n_x, n_u, f, A, B."""

    def __init__(self, x_syms, u_syms, f_expr):
        self.n_x, self.n_u = len(x_syms), len(u_syms)
        self.f, self.A, self.B = _lambdas(x_syms, u_syms, f_expr)


def _rel(g, w):
    return np.abs(g - w).max() / max(np.abs(w).max(), 1e-300)


def test_drag_unicycle_stage1_parity_and_whole_loop(cuda):
    from scvx_b200 import _lib
    from scvx_b200.discretization.first_order_hold import FirstOrderHold
    from scvx_b200.optimization.scvx_solver import SCVXSolver
    K = 40
    model = _drag_unicycle_cls()()
    mid = model.device_model_id
    assert mid >= _lib.MODEL_USER_BASE and model.device_model_id == mid          # compiled once, cached
    om = _OracleUserModel(*model.symbolic_dynamics())
    rng = np.random.default_rng(7)
    X = np.vstack([rng.uniform(-8, 8, (2, K)), rng.uniform(-3, 3, (1, K))])
    U = np.vstack([rng.uniform(0, 1, (1, K)), rng.uniform(-0.5, 0.5, (1, K))])
    F = FirstOrderHold(model, K)
    O = ofoh.OracleFOH(om, K)
    for sigma in (1.0, 12.0, 40.0):
        got = [np.array(m) for m in F.calculate_discretization(X, U, sigma)]
        want = O.calculate_discretization(X, U, sigma, tol="tight")
        for g, w in zip(got, want):
            assert _rel(g, w) < 1e-9
        assert _rel(F.integrate_nonlinear_piecewise(X, U, sigma), O.integrate_nonlinear_piecewise(X, U, sigma, tol="tight")) < 1e-9
    assert _rel(F.integrate_nonlinear_full(X[:, 0], U, 5.0), O.integrate_nonlinear_full(X[:, 0], U, 5.0, tol="tight")) < 1e-9
    # the dynamics really differ from the shipped unicycle's
    from scvx_b200.models.unicycle_model import UnicycleModel
    plain = [np.array(m) for m in FirstOrderHold(UnicycleModel(), K).calculate_discretization(X, U, 12.0)]
    assert _rel(plain[1], O.calculate_discretization(X, U, 12.0, tol="tight")[1]) > 1e-3

    # whole loop through the reference-facing driver; first sub-problem against the exact LP on the same parameters
    solver = SCVXSolver(model, K); solver.max_iter = 4
    Xs, Us, sig, lg = solver.solve()
    assert Xs.shape == (3, K) and Us.shape == (2, K) and len(lg.records) == 4
    np.testing.assert_allclose(Xs[:, 0], model.x_init, atol=1e-12); np.testing.assert_allclose(Xs[:, -1], model.x_final, atol=1e-12)
    X0, U0 = model.initialize_trajectory(np.zeros((3, K)), np.zeros((2, K)))
    mats = tuple(np.array(m) for m in F.calculate_discretization(X0, U0, 1.0))
    desc = omodels.unicycle()                     # same constraints as the shipped unicycle (only the dynamics differ)
    p = ospb.Params(desc, K, mats, X0, U0, 1.0, 100.0)
    r0 = lg.records[0]
    assert 1e4 * r0["nu_norm"] + 1e6 * r0["slack_norm"] + 100.0 * r0["sigma"] == pytest.approx(ospb.solve(p)["obj"], rel=1e-7)


def test_four_state_model_has_stage1_and_is_refused_by_stage3(cuda):
    """n_x = 4 (unicycle with a speed state, inputs acceleration and turn rate): exercises the generic Phi inverse.  Stage 1
    runs and matches the oracle; the sub-problem kernel has no instantiation of that shape and says so."""
    from scvx_b200 import _device, _lib, codegen
    x, y, th, v = sp.symbols("x y theta v", real=True)
    a, w = sp.symbols("a w", real=True)
    f = sp.Matrix([v * sp.cos(th), v * sp.sin(th), w, a - 0.1 * v * v])
    mid = codegen.register([x, y, th, v], [a, w], f, position_dim=2)
    assert _device.MODEL_DIMS[mid] == (4, 2, 2)
    K = 25
    rng = np.random.default_rng(11)
    X = np.vstack([rng.uniform(-5, 5, (2, K)), rng.uniform(-2, 2, (1, K)), rng.uniform(0, 1, (1, K))])
    U = np.vstack([rng.uniform(-0.5, 0.5, (1, K)), rng.uniform(-0.5, 0.5, (1, K))])
    to = lambda arr: torch.as_tensor(np.ascontiguousarray(arr)[None]).to(cuda)      # noqa: E731
    got = _device.foh(mid, to(X), to(U), torch.full((1,), 9.0, dtype=torch.float64, device=cuda))
    want = ofoh.OracleFOH(_OracleUserModel([x, y, th, v], [a, w], f), K).calculate_discretization(X, U, 9.0, tol="tight")
    for g, wv in zip(got, want):
        assert _rel(g[0].cpu().numpy(), wv) < 1e-9
    assert _lib.load().scvx_solve_workspace_bytes(mid, 1, K, 0, 0) == 0          # no sub-problem kernel of shape (4, 2, 2)


def test_compile_error_is_reported(cuda):
    from scvx_b200 import _lib, codegen
    lib = _lib.load()
    import ctypes
    mid = ctypes.c_int(-1)
    rc = lib.scvx_user_model_register(b"this is not CUDA", 3, 2, 2, b"", ctypes.byref(mid))
    assert rc == -1 and mid.value == -1 and b"error" in lib.scvx_user_model_log()
    s = sp.symbols("s")
    with pytest.raises(ValueError):
        codegen.register([s], [], sp.Matrix([s]), 1)              # a model without inputs
    # a model struct the compiler rejects surfaces as ScvxError-style failure with the NVRTC log available
    prog = codegen.program_source("namespace scvx { struct UserModel { static constexpr int NX = 1, NU = 1, D = 1; }; }")
    rc = lib.scvx_user_model_register(prog.encode(), 1, 1, 1, b"", ctypes.byref(mid))
    assert rc == -1 and b"NVRTC" in lib.scvx_last_error() and len(lib.scvx_user_model_log()) > 0
