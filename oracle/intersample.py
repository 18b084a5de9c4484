"""CPU restatement of the reference's inter-sample clearance routines -- TEST INFRASTRUCTURE ONLY.

Follows SCvx/utils/intersample_collision.py:7-125 line by line (h_i, find_critical_times, linearize_h, make_segment_f),
with the segment flow integrated by odeint at rtol=1e-12/atol=1e-13 (tol="tight", the parity target of the CUDA RK4) or at
odeint's defaults (tol="reference", what the reference itself runs).  Pinned: tests/golden/intersample_golden.npz holds
outputs of the UNMODIFIED reference functions; at tol="reference" this file reproduces them exactly
(tests/test_oracle_intersample.py).  NOTE the reference's central differences of a default-tolerance odeint are noisy
(phi carries ~1e-8/1e-4 relative noise), so t* agrees between tolerances only to ~1e-4 and gradients to ~1e-3.
"""
import numpy as np
from scipy.integrate import odeint


def make_segment_f(foh, U_ref_k, U_ref_kp1, sigma, tol="tight"):
    dt_phys = foh.dt * sigma
    kw = {"rtol": 1e-12, "atol": 1e-13, "mxstep": 100000} if tol == "tight" else {}

    def f_seg(xk, _u_dummy, t):
        return odeint(foh._dx, xk, [0.0, t * dt_phys], args=(U_ref_k, U_ref_kp1, sigma), **kw)[1]

    return f_seg, dt_phys


def h_i(xk, uk, t, f, T, obstacle):
    p_c, r = obstacle
    return np.linalg.norm(T @ f(xk, uk, t) - p_c) - r


def find_critical_times(xk, uk, f, T, obstacle, dt, num_samples=100, eps=1e-4, tol=1e-6):
    def phi(t):
        return (h_i(xk, uk, t + eps, f, T, obstacle) - h_i(xk, uk, t - eps, f, T, obstacle)) / (2 * eps)

    def phi2(t):
        return (phi(t + eps) - phi(t - eps)) / (2 * eps)

    ts = np.linspace(eps, dt - eps, num_samples)
    phis = np.array([phi(t) for t in ts])
    raw = []
    for i in range(len(ts) - 1):
        if phis[i] == 0 or phis[i] * phis[i + 1] < 0:
            a, b = ts[i], ts[i + 1]
            for _ in range(30):
                c = 0.5 * (a + b)
                if phi(a) * phi(c) <= 0:
                    b = c
                else:
                    a = c
                if abs(b - a) < tol:
                    break
            raw.append(0.5 * (a + b))
    return sorted(r for r in raw if 0 < r < dt and phi2(r) > 0)


def linearize_h(xk, uk, t_star, f, T, obstacle, eps=1e-4):
    h0 = h_i(xk, uk, t_star, f, T, obstacle)
    grad_x = np.zeros_like(xk)
    for j in range(len(xk)):
        xp, xm = xk.copy(), xk.copy()
        xp[j] += eps; xm[j] -= eps
        grad_x[j] = (h_i(xp, uk, t_star, f, T, obstacle) - h_i(xm, uk, t_star, f, T, obstacle)) / (2 * eps)
    grad_u = np.zeros_like(uk)
    for j in range(len(uk)):
        up, um = uk.copy(), uk.copy()
        up[j] += eps; um[j] -= eps
        grad_u[j] = (h_i(xk, up, t_star, f, T, obstacle) - h_i(xk, um, t_star, f, T, obstacle)) / (2 * eps)
    return h0, grad_x, grad_u
