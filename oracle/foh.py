"""Oracle restatement of the reference's first-order-hold discretiser (stage 1).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  numpy + scipy.integrate.odeint, fp64.

Follows SCvx/discretization/first_order_hold.py:
  :13-50   augmented state layout V = [x, Phi, B-, B+, S, z], matrices flattened order='F', dt = 1/(K-1)
  :52-87   calculate_discretization (K-1 odeint calls, then Phi @ {B-, B+, S, z})
  :89-125  _ode_dVdt (the augmented right-hand side)
  :127-162 integrate_nonlinear_piecewise / integrate_nonlinear_full / _dx

Two integration modes:
  tol="reference": odeint at its default tolerances -- exactly what the reference runs (its own
                   integration error is ~1e-8 relative);
  tol="tight":     rtol=1e-13, atol=1e-14 -- the parity target for the CUDA RK4 kernel (1e-9 relative).
"""
from __future__ import annotations

import numpy as np
from scipy.integrate import odeint


class OracleFOH:
    def __init__(self, model, K):
        self.model, self.K = model, K
        n_x, n_u = model.n_x, model.n_u
        self.n_x, self.n_u = n_x, n_u
        e = [n_x]
        for w in (n_x * n_x, n_x * n_u, n_x * n_u, n_x, n_x):
            e.append(e[-1] + w)
        self.x_ind = slice(0, e[0])
        self.A_ind = slice(e[0], e[1])
        self.B_ind = slice(e[1], e[2])
        self.C_ind = slice(e[2], e[3])
        self.S_ind = slice(e[3], e[4])
        self.z_ind = slice(e[4], e[5])
        self.nV = e[5]
        self.V0 = np.zeros(self.nV)
        self.V0[self.A_ind] = np.eye(n_x).reshape(-1, order="F")
        self.dt = 1.0 / (K - 1)

    # first_order_hold.py:89-125
    def rhs(self, V, t, u0, u1, sigma):
        n_x = self.n_x
        m = self.model
        alpha = (self.dt - t) / self.dt
        beta = t / self.dt
        x = V[self.x_ind]
        u = u0 + (t / self.dt) * (u1 - u0)
        A_sub = sigma * m.A(x, u)
        B_sub = sigma * m.B(x, u)
        f_sub = m.f(x, u)
        Phi = V[self.A_ind].reshape((n_x, n_x), order="F")
        Phi_inv = np.linalg.inv(Phi)
        dV = np.zeros_like(V)
        dV[self.x_ind] = sigma * f_sub
        dV[self.A_ind] = (A_sub @ Phi).reshape(-1, order="F")
        PiB = (Phi_inv @ B_sub).reshape(-1, order="F")
        dV[self.B_ind] = PiB * alpha
        dV[self.C_ind] = PiB * beta
        dV[self.S_ind] = Phi_inv @ f_sub
        dV[self.z_ind] = Phi_inv @ (-A_sub @ x - B_sub @ u)
        return dV

    # first_order_hold.py:52-87
    def calculate_discretization(self, X, U, sigma, tol="reference"):
        n_x, n_u, K = self.n_x, self.n_u, self.K
        A_bar = np.zeros((n_x * n_x, K - 1))
        B_bar = np.zeros((n_x * n_u, K - 1))
        C_bar = np.zeros((n_x * n_u, K - 1))
        S_bar = np.zeros((n_x, K - 1))
        z_bar = np.zeros((n_x, K - 1))
        kw = {} if tol == "reference" else dict(rtol=1e-13, atol=1e-14, mxstep=100000)
        V0 = self.V0.copy()
        for k in range(K - 1):
            V0[self.x_ind] = X[:, k]
            V = odeint(self.rhs, V0, [0.0, self.dt], args=(U[:, k], U[:, k + 1], sigma), **kw)[1]
            Phi = V[self.A_ind].reshape((n_x, n_x), order="F")
            A_bar[:, k] = Phi.flatten(order="F")
            B_bar[:, k] = (Phi @ V[self.B_ind].reshape((n_x, n_u), order="F")).flatten(order="F")
            C_bar[:, k] = (Phi @ V[self.C_ind].reshape((n_x, n_u), order="F")).flatten(order="F")
            S_bar[:, k] = Phi @ V[self.S_ind]
            z_bar[:, k] = Phi @ V[self.z_ind]
        return A_bar, B_bar, C_bar, S_bar, z_bar

    # first_order_hold.py:157-162
    def _dx(self, x, t, u0, u1, sigma):
        u = u0 + (t / (self.dt * sigma)) * (u1 - u0)
        return self.model.f(x, u)

    # first_order_hold.py:127-140
    def integrate_nonlinear_piecewise(self, X_lin, U, sigma, tol="reference"):
        kw = {} if tol == "reference" else dict(rtol=1e-13, atol=1e-14, mxstep=100000)
        X_nl = np.zeros_like(X_lin)
        X_nl[:, 0] = X_lin[:, 0]
        for k in range(self.K - 1):
            X_nl[:, k + 1] = odeint(self._dx, X_lin[:, k], [0.0, self.dt * sigma],
                                    args=(U[:, k], U[:, k + 1], sigma), **kw)[1]
        return X_nl

    # first_order_hold.py:142-155
    def integrate_nonlinear_full(self, x0, U, sigma, tol="reference"):
        kw = {} if tol == "reference" else dict(rtol=1e-13, atol=1e-14, mxstep=100000)
        X_nl = np.zeros((self.n_x, self.K))
        X_nl[:, 0] = x0
        for k in range(self.K - 1):
            X_nl[:, k + 1] = odeint(self._dx, X_nl[:, k], [0.0, self.dt * sigma],
                                    args=(U[:, k], U[:, k + 1], sigma), **kw)[1]
        return X_nl
