"""Prototype that decided the stage-3 algorithm (kept as evidence; NOT part of the product or the tests).

OSQP-style ADMM (Ruiz-equilibrated, over-relaxed, adaptive rho, cone/hinge prox blocks) on the reference's own
sub-problems -- the first outer iterations of the shipped unicycle scene at K=50 -- compared with the optimal
value from the exact HiGHS oracle.  Result (profiles/r01_admm_vs_ipm_prototype.md): after 3000 iterations the
relative objective error is 1e-1 ... 1e+1, so the first-order method named in north_star cannot meet the 1e-4
parity gate on these degenerate exact-penalty LPs; the interior-point twin (oracle/ipm_struct.py) reaches 1e-10 in
10-40 iterations.

    python tools/proto_admm.py [max_iter]
"""
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spl, sys, pickle, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import subproblem as spb


def proj_l1cone(v, t):
    """project (v,t) onto {||v||_1 <= t}"""
    a = np.abs(v)
    if a.sum() <= t:
        return v.copy(), t
    if a.max() <= -t:
        return np.zeros_like(v), 0.0
    s = np.sort(a)[::-1]
    cs = 0.0
    n = len(v)
    for m in range(1, n + 1):
        cs += s[m - 1]
        lam = (cs - t) / (m + 1)
        nxt = s[m] if m < n else 0.0
        if lam >= nxt and lam < s[m - 1] + 1e-300:
            break
    lam = max(lam, 0.0)
    return np.sign(v) * np.maximum(a - lam, 0), t + lam


def build(p: spb.Params):
    m, K = p.model, p.K
    n_x, n_u, d = m.n_x, m.n_u, m.d
    ns = n_x + n_u
    nz = ns * K + 4
    iSig, iTnu, iTx, iTu = ns * K, ns * K + 1, ns * K + 2, ns * K + 3
    X = lambda i, k: ns * k + i
    U = lambda i, k: ns * k + n_x + i
    rows, cols, vals = [], [], []
    shift = []      # y = A z - shift   (so sets are origin-based)
    blocks = []     # (type, row indices, params)
    r = 0

    def add(idx, val, sh):
        nonlocal r
        for i, v in zip(idx, val):
            rows.append(r); cols.append(i); vals.append(v)
        shift.append(sh)
        r += 1
        return r - 1

    # nu cones
    for k in range(K - 1):
        A = p.A_bar[:, k].reshape((n_x, n_x), order='F'); B = p.B_bar[:, k].reshape((n_x, n_u), order='F')
        C = p.C_bar[:, k].reshape((n_x, n_u), order='F')
        rr = []
        for i in range(n_x):
            idx = [X(i, k + 1)] + [X(j, k) for j in range(n_x)] + [U(j, k) for j in range(n_u)] + [U(j, k + 1) for j in range(n_u)] + [iSig]
            val = [1.0] + list(-A[i]) + list(-B[i]) + list(-C[i]) + [-p.S_bar[i, k]]
            rr.append(add(idx, val, p.z_bar[i, k]))
        rt = add([iTnu], [1.0], 0.0)
        blocks.append(('l1cone', rr, rt))
    for k in range(K):
        rr = [add([X(i, k)], [1.0], p.X_ref[i, k]) for i in range(n_x)]
        rt = add([iTx], [1.0], 0.0)
        blocks.append(('l1cone', rr, rt))
        rr = [add([U(i, k)], [1.0], p.U_ref[i, k]) for i in range(n_u)]
        rt = add([iTu], [1.0], 0.0)
        blocks.append(('l1cone', rr, rt))
    # trust scalar rows
    r1 = add([iTx, iTu, iSig], [1, 1, 1.0], p.sigma_ref)
    r2 = add([iTx, iTu, iSig], [1, 1, -1.0], -p.sigma_ref)
    blocks.append(('box', [r1, r2], (np.array([-np.inf, -np.inf]), np.array([p.tr_radius, p.tr_radius]))))
    # sigma >= 0
    r3 = add([iSig], [1.0], 0.0)
    blocks.append(('box', [r3], (np.array([0.0]), np.array([np.inf]))))
    # boxes on stage variables
    lo_p, hi_p = m.lower_bound + m.robot_radius, m.upper_bound - m.robot_radius
    for k in range(K):
        rr, lo, hi = [], [], []
        for i in range(n_x):
            l, h = (lo_p, hi_p) if i < d else (-np.inf, np.inf)
            if k == 0: l = max(l, m.x_init[i]); h = min(h, m.x_init[i])
            if k == K - 1: l = max(l, m.x_final[i]); h = min(h, m.x_final[i])
            if np.isfinite(l) or np.isfinite(h):
                rr.append(add([X(i, k)], [1.0], 0.0)); lo.append(l); hi.append(h)
        for i in range(n_u):
            if m.kind == 'unicycle':
                l, h = (0.0, m.v_max) if i == 0 else (-m.w_max, m.w_max)
            else:
                l, h = -m.v_max, m.v_max
            if k == 0 or k == K - 1: l = h = 0.0
            rr.append(add([U(i, k)], [1.0], 0.0)); lo.append(l); hi.append(h)
        blocks.append(('box', rr, (np.array(lo), np.array(hi))))
    # obstacle hinge rows
    for j in range(len(m.obstacles)):
        rr = []
        for k in range(K):
            a = p.obs_a[j][:, k]
            rr.append(add([X(i, k) for i in range(d)], list(a), p.obs_rhs[j] + a.dot(p.obs_c[j])))
        blocks.append(('hinge', rr, p.weight_slack))
    for nb in p.neighbors:
        rr = []
        for k in range(K):
            a = nb['a'][:, k]
            rr.append(add([X(i, k) for i in range(d)], list(a), p.d_min + a.dot(nb['Y'][:, k])))
        blocks.append(('hinge', rr, p.weight_col))
    A = sp.csr_matrix((vals, (rows, cols)), shape=(r, nz))
    q = np.zeros(nz); q[iSig] = p.weight_sigma; q[iTnu] = p.weight_nu
    Pd = np.zeros(nz)
    for nb in p.neighbors:
        for k in range(K):
            for i in range(d):
                q[X(i, k)] += nb['Lam'][i, k] - p.rho * nb['Y'][i, k]; Pd[X(i, k)] += p.rho
    return dict(A=A, shift=np.array(shift), blocks=blocks, q=q, Pd=Pd, nz=nz, ns=ns, K=K, n_x=n_x, n_u=n_u,
                idx=(iSig, iTnu, iTx, iTu))


def prox_all(blocks, v, rho_vec, E=None):
    """v: scaled? here unscaled rows. returns y = prox"""
    y = v.copy()
    for typ, rr, par in blocks:
        if typ == 'box':
            y[rr] = np.clip(v[rr], par[0], par[1])
        elif typ == 'hinge':
            w = par
            vv = v[rr]; rh = rho_vec[rr]
            y[rr] = np.where(vv >= 0, vv, np.where(vv < -w / rh, vv + w / rh, 0.0))
        elif typ == 'l1cone':
            rt = par
            vv, tt = proj_l1cone(v[rr], v[rt])
            y[rr] = vv; y[rt] = tt
    return y


def unpack(pb, z):
    K, ns, n_x, n_u = pb['K'], pb['ns'], pb['n_x'], pb['n_u']
    W = z[:ns * K].reshape((K, ns)).T
    return W[:n_x].copy(), W[n_x:].copy(), z[pb['idx'][0]]


def ruiz(pb, iters=15):
    A = pb['A'].tocsr().copy()
    m, n = A.shape
    D = np.ones(n); E = np.ones(m)
    Pd = pb['Pd'].copy()
    for _ in range(iters):
        As = sp.diags(E) @ A @ sp.diags(D)
        absA = abs(As)
        cn = np.maximum(np.asarray(absA.max(axis=0).todense()).ravel(), D * Pd * D)
        rn = np.asarray(absA.max(axis=1).todense()).ravel()
        # cone-uniform row scaling
        for typ, rr, par in pb['blocks']:
            if typ == 'l1cone':
                allr = list(rr) + [par]
                rn[allr] = rn[allr].max()
        cn[cn < 1e-8] = 1.0; rn[rn < 1e-8] = 1.0
        D /= np.sqrt(cn); E /= np.sqrt(rn)
    return D, E


def admm(pb, rho0=0.1, alpha=1.6, eps_reg=1e-6, max_iter=5000, scale=True, adapt_every=50, verbose=False,
         eval_fn=None, rho_eq_mult=1.0, tol=1e-7, z0=None):
    A0 = pb['A']; m, n = A0.shape
    if scale:
        D, E = ruiz(pb)
    else:
        D, E = np.ones(n), np.ones(m)
    A = (sp.diags(E) @ A0 @ sp.diags(D)).tocsc()
    q0 = pb['q'] * D
    cs = 1.0 / max(np.abs(q0).max(), 1e-8) if scale else 1.0
    q = cs * q0
    Pd = cs * D * pb['Pd'] * D
    shift = E * pb['shift']
    blocks = []
    for typ, rr, par in pb['blocks']:
        if typ == 'box':
            blocks.append((typ, rr, (par[0] * E[rr], par[1] * E[rr])))
        elif typ == 'hinge':
            blocks.append((typ, rr, cs * par / E[rr]))
        else:
            blocks.append((typ, rr, par))
    rho_vec = np.full(m, rho0)
    # equality rows get more rho
    for typ, rr, par in blocks:
        if typ == 'box':
            eq = (par[0] == par[1])
            idx = np.array(rr)[eq]
            rho_vec[idx] = rho0 * rho_eq_mult
    rho_scale = rho_vec / rho0
    rho = rho0

    def factor(rho):
        rv = rho * rho_scale
        H = sp.diags(Pd + eps_reg) + A.T @ sp.diags(rv) @ A
        return spl.splu(H.tocsc()), rv
    lu, rv = factor(rho)
    nfac = 1
    z = np.zeros(n) if z0 is None else z0 / D
    y = prox_all(blocks, A @ z - shift, rv)
    lam = np.zeros(m)
    hist = []
    for it in range(max_iter):
        rhs = eps_reg * z - q + A.T @ (rv * (y + shift) - lam)
        zt = lu.solve(rhs)
        yt = A @ zt - shift
        zn = alpha * zt + (1 - alpha) * z
        yr = alpha * yt + (1 - alpha) * y
        yn = prox_all(blocks, yr + lam / rv, rv)
        lam = lam + rv * (yr - yn)
        z, y = zn, yn
        if (it + 1) % 25 == 0 or it == max_iter - 1:
            Az = A @ z - shift
            rp = np.abs((Az - y) / E).max()
            rd = np.abs((Pd * z + q + A.T @ lam) / D).max() / cs
            hist.append((it + 1, rp, rd))
            if verbose and ((it + 1) % 250 == 0):
                extra = eval_fn(z * D) if eval_fn else ''
                print(it + 1, f'rp={rp:.2e} rd={rd:.2e} rho={rho:.3g}', extra)
            if rp < tol and rd < tol * 100:
                break
            if adapt_every and (it + 1) % adapt_every == 0:
                # OSQP rule in scaled space
                rps = np.abs(Az - y).max() / max(np.abs(Az + shift).max(), np.abs(y + shift).max(), 1e-10)
                rds = np.abs(Pd * z + q + A.T @ lam).max() / max(np.abs(A.T @ lam).max(), np.abs(q).max(), np.abs(Pd * z).max(), 1e-10)
                new = rho * np.sqrt(rps / max(rds, 1e-12))
                new = min(max(new, 1e-6), 1e6)
                if new > 5 * rho or new < rho / 5:
                    rho = new
                    lu, rv = factor(rho); nfac += 1
    return z * D, dict(iters=it + 1, hist=hist, nfac=nfac, lam=lam, y=y)


if __name__ == '__main__':
    from oracle import models as omodels, scvx as oscvx
    from oracle.ipm_struct import StructIPM
    probs = oscvx.scvx_solve(omodels.unicycle(), 50, max_iter=9, return_problems=True)[4]
    for pi in [0, 1, 2, 5, 8]:
        p, r = probs[pi]
        pb = build(p)

        def ev(z):
            X, U, s = unpack(pb, z)
            e = spb.evaluate(p, X, U, s)
            return f"obj={e['obj']:.6f} (rel {abs(e['obj']-r['obj'])/abs(r['obj']):.2e}) viol={e['viol']:.2e}"
        t = time.time()
        z, info = admm(pb, verbose=False, eval_fn=ev, max_iter=int(sys.argv[1]) if len(sys.argv) > 1 else 3000)
        print('ADMM prob', pi, 'opt', r['obj'], 'iters', info['iters'], 'nfac', info['nfac'], ev(z), f'{time.time()-t:.1f}s')
        t = time.time(); s = StructIPM(p).solve(); e = spb.evaluate(p, s['X'], s['U'], s['sigma'])
        print('IPM  prob', pi, 'opt', r['obj'], 'iters', s['iters'], f"obj={e['obj']:.6f} (rel {abs(e['obj']-r['obj'])/abs(r['obj']):.2e}) viol={e['viol']:.2e}", f'{time.time()-t:.1f}s')
