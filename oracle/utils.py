"""CPU restatement of the reference's warm-start generators and analysis metrics -- TEST INFRASTRUCTURE ONLY.

Follows SCvx/utils/initial_guess.py:4-107 (unicycle), SCvx/utils/IS_initial_guess.py:6-126 (single integrator) and
SCvx/utils/analysis.py:10-62.  Pinned: tests/golden/utils_golden.npz holds outputs of the UNMODIFIED reference functions
(tests/golden/make_golden_utils.py); tests/test_oracle_utils.py checks this file against them bit for bit.
Written with explicit scalar arithmetic in the order the CUDA kernels use (csrc/utils.cu), so it doubles as the
statement of that order: sums of squares left to right, numpy.linspace's own formula, Python round-half-even.
"""
import math

import numpy as np


def _norm(v):
    s = 0.0
    for x in v:
        s = s + x * x
    return math.sqrt(s)


def _dot(a, b):
    s = 0.0
    for x, y in zip(a, b):
        s = s + x * y
    return s


def segment_hits_ball(p, q, center, r):
    """initial_guess.py:4-19 / IS_initial_guess.py:6-23."""
    d = [qi - pi for pi, qi in zip(p, q)]
    f = [pi - ci for pi, ci in zip(p, center)]
    a = _dot(d, d)
    b = 2 * _dot(f, d)
    c = _dot(f, f) - r * r
    disc = b * b - 4 * a * c
    if disc < 0:
        return False
    sq = math.sqrt(disc)
    t1 = (-b + sq) / (2 * a)
    t2 = (-b - sq) / (2 * a)
    return (0 < t1 < 1) or (0 < t2 < 1)


def tangent_points(p, center, r):
    """initial_guess.py:22-38."""
    v = [p[0] - center[0], p[1] - center[1]]
    d = _norm(v)
    if d <= r:
        raise ValueError("Point inside/on circle; no tangents.")
    alpha = math.asin(r / d)
    theta = math.atan2(v[1], v[0])
    t1, t2 = theta + alpha, theta - alpha
    return ([center[0] + r * math.cos(t1), center[1] + r * math.sin(t1)],
            [center[0] + r * math.cos(t2), center[1] + r * math.sin(t2)])


def detour_waypoints(p0, p1, center, r):
    """IS_initial_guess.py:26-55."""
    d = [b - a for a, b in zip(p0, p1)]
    nd = _norm(d)
    if nd < 1e-6:
        raise ValueError("p0 and p1 are too close for detour computation.")
    du = [x / nd for x in d]
    tmp = [1.0, 0.0, 0.0] if abs(du[0]) < 0.9 else [0.0, 1.0, 0.0]
    u = [du[1] * tmp[2] - du[2] * tmp[1], du[2] * tmp[0] - du[0] * tmp[2], du[0] * tmp[1] - du[1] * tmp[0]]
    nu = _norm(u)
    u = [x / nu for x in u]
    t = _dot([c - a for a, c in zip(p0, center)], du)
    proj = [a + t * x for a, x in zip(p0, du)]
    return [[pr + r * x for pr, x in zip(proj, u)], [pr - r * x for pr, x in zip(proj, u)]]


def _linspace_point(a, b, num, endpoint, i):
    """Element i of numpy.linspace(a, b, num, endpoint) for vector end points (numpy/core/function_base.py)."""
    div = num - 1 if endpoint else num
    delta = [y - x for x, y in zip(a, b)]
    if div > 0:
        step = [dl / div for dl in delta]
        if any(s == 0 for s in step):
            y = [(i / div) * dl for dl in delta]
        else:
            y = [i * s for s in step]
    else:
        y = [i * dl for dl in delta]
    y = [v + x for v, x in zip(y, a)]
    if endpoint and num > 1 and i == num - 1:
        y = list(b)
    return y


def piecewise_linear(p0, p1, waypoints, K):
    """generate_piecewise_linear (initial_guess.py:41-58): exactly K samples, (d, K)."""
    pts = [list(p0)] + [list(w) for w in waypoints] + [list(p1)]
    lengths = [_norm([b - a for a, b in zip(pts[i], pts[i + 1])]) for i in range(len(pts) - 1)]
    total = 0.0
    for L in lengths:
        total = total + L
    Ns = [max(2, round(K * L / total)) for L in lengths]
    Ns[-1] = K - sum(Ns[:-1])
    if Ns[-1] < 0:
        raise ValueError(f"Number of samples, {Ns[-1]}, must be non-negative.")
    rows = []
    for idx in range(len(pts) - 1):
        endpoint = idx == len(pts) - 2
        for i in range(Ns[idx]):
            rows.append(_linspace_point(pts[idx], pts[idx + 1], Ns[idx], endpoint, i))
    return np.array(rows).T


def initial_guess_unicycle(p0, p1, obstacles, clearance, K):
    """initial_guess.py:61-107."""
    a = [float(p0[0]), float(p0[1])]; b = [float(p1[0]), float(p1[1])]
    waypoints = []
    for c, r in obstacles:
        c = [float(c[0]), float(c[1])]; rr = float(r) + clearance
        if not segment_hits_ball(a, b, c, rr):
            continue
        T = tangent_points(a, c, rr); G = tangent_points(b, c, rr)
        best = None
        for Ti in T:
            for Gj in G:
                L = (_norm([a[0] - Ti[0], a[1] - Ti[1]]) + _norm([Ti[0] - Gj[0], Ti[1] - Gj[1]])
                     + _norm([Gj[0] - b[0], Gj[1] - b[1]]))
                if best is None or L < best[0]:
                    best = (L, Ti, Gj)
        waypoints.extend([best[1], best[2]])
    path = piecewise_linear(a, b, waypoints, K)
    X0 = np.zeros((3, K)); X0[0:2] = path
    for k in range(K - 1):
        X0[2, k] = math.atan2(path[1, k + 1] - path[1, k], path[0, k + 1] - path[0, k])
    X0[2, -1] = X0[2, -2]
    return X0, np.zeros((2, K))


def initial_guess_si(p0, p1, obstacles, clearance, K):
    """IS_initial_guess.py:87-126."""
    a = [float(x) for x in p0]; b = [float(x) for x in p1]
    waypoints = []
    for c, r in obstacles:
        c = [float(x) for x in c]; rr = float(r) + clearance
        if not segment_hits_ball(a, b, c, rr):
            continue
        waypoints.extend(detour_waypoints(a, b, c, rr))
    X0 = piecewise_linear(a, b, waypoints, K)
    U0 = np.zeros_like(X0)
    dt = 1.0 / (K - 1)
    U0[:, :-1] = (X0[:, 1:] - X0[:, :-1]) / dt
    U0[:, -1] = U0[:, -2]
    return X0, U0


def min_inter_agent_distance(X_list):
    """analysis.py:10-31."""
    N = len(X_list)
    d_mat = np.zeros((N, N))
    for i in range(N):
        for j in range(i + 1, N):
            diff = X_list[i][0:3] - X_list[j][0:3]
            s = diff[0] * diff[0]
            for c in range(1, diff.shape[0]):
                s = s + diff[c] * diff[c]
            d_mat[i, j] = d_mat[j, i] = np.sqrt(s).min()
    return d_mat[d_mat > 0].min(), d_mat


def min_agent_obstacle_distance(X_list, obstacles, robot_radius):
    """analysis.py:34-62."""
    N, M = len(X_list), len(obstacles)
    d_mat = np.full((N, M), np.inf)
    for i in range(N):
        for j, (centre, r_j) in enumerate(obstacles):
            diff = X_list[i][0:3] - np.array(centre, dtype=float)[:, None]
            s = diff[0] * diff[0]
            for c in range(1, diff.shape[0]):
                s = s + diff[c] * diff[c]
            d_mat[i, j] = (np.sqrt(s) - (robot_radius + r_j)).min()
    return d_mat.min(), d_mat
