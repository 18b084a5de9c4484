"""Device-resident batched drivers: the performance path of the SCvx inner loop.

`AgentBatch`    -- constraint tables + trajectories of n agents of one model kind, on one GPU.
`BatchedSCvx`   -- SCVXSolver.solve (SCvx/optimization/scvx_solver.py:33-115) for all agents at once:
                   per outer iteration ONE launch each of the FOH kernel, the obstacle-linearisation
                   kernel, the IPM kernel and the bookkeeping kernel; no host round trip inside the loop.
`PipelinedSCvx` -- the same outer loop with the agents split into lanes that advance independently on their own CUDA
                   streams (independent agents need no global barrier between outer iterations).
`BatchedADMM`   -- ADMMCoordinator.solve / SI_ADMMCoordinator.solve (admm_coordinator.py:39-118,
                   si_admm_coordinator.py:42-127) as Jacobi rounds, agents sharded over the ranks of a
                   torch.distributed process group with ONE all-gather of positions per round.
"""
import numpy as np
import torch

from . import _device
from .global_parameters import (CONV_TOL, MAX_ITER, TRUST_RADIUS0, WEIGHT_NU, WEIGHT_SIGMA, WEIGHT_SLACK)

F64 = torch.float64
WEIGHT_COLLISION_SLACK = 1e5     # SCvx/optimization/admm_utils.py:5
_FAR = 1.0e3                     # padding obstacle: centre far outside the workspace, clearance -_FAR (never active)


class AgentBatch:
    def __init__(self, models, K, device=None, min_obstacles=0):
        if not models:
            raise ValueError("empty agent batch")
        mid = models[0].device_model_id
        if mid is None or any(m.device_model_id != mid for m in models):
            raise NotImplementedError("all agents of a batch must share one compiled device model")
        self.models, self.K, self.n = list(models), K, len(models)
        self.model_id = mid
        self.n_x, self.n_u, self.d = _device.MODEL_DIMS[mid]
        self.device = torch.device("cuda") if device is None else torch.device(device)
        tabs = [m.get_constraints() for m in models]
        # obstacle slots per agent: the largest count in the batch (or `min_obstacles`: lanes of one batch pad alike);
        # missing obstacles are padded with discs far outside the workspace that never become active
        self.M = max(max(len(t.obs_clearance) for t in tabs), int(min_obstacles))
        n, M, d = self.n, self.M, self.d
        obs_c = np.full((n, M, d), _FAR)
        obs_clear = np.full((n, M), -_FAR)
        for i, t in enumerate(tabs):
            m_i = len(t.obs_clearance)
            if m_i:
                obs_c[i, :m_i] = t.obs_centres
                obs_clear[i, :m_i] = t.obs_clearance
        self.n_obs = [len(t.obs_clearance) for t in tabs]
        up = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(self.device)   # noqa: E731
        self.x_init = up(np.stack([t.x_init for t in tabs]))
        self.x_final = up(np.stack([t.x_final for t in tabs]))
        self.pos_lo = up(np.array([t.pos_lo for t in tabs]))
        self.pos_hi = up(np.array([t.pos_hi for t in tabs]))
        self.v_max = up(np.array([t.v_max for t in tabs]))
        self.w_max = up(np.array([t.w_max for t in tabs]))
        self.obs_c, self.obs_clear = up(obs_c), up(obs_clear)

    def initial_trajectories(self):
        """model.initialize_trajectory for every agent (straight line, U = 0), built on the device."""
        K = self.K
        a2 = torch.arange(K, dtype=F64, device=self.device) / (K - 1)
        X = self.x_init[:, :, None] * (1.0 - a2) + self.x_final[:, :, None] * a2
        # the reference computes alpha1 = (K-1-k)/(K-1) separately; identical up to 1 ulp
        U = torch.zeros((self.n, self.n_u, K), dtype=F64, device=self.device)
        return X.contiguous(), U


# Barrier start of a warm sub-problem (the reference trajectory is the previous solution): every agent whose previous solve ended
# within MU0_POLICY[0] interior-point iterations starts at mu = MU0_POLICY[1], the others (iteration cap hit) at the cold default.
# Measured on the bench workload (tools/exp_mu0_policy.py, gpurun_out/r02p_mu0_policy.txt): (10, 0.1) 2.66 ms/step, (1000, 1e-2) 2.68,
# (1000, 1e-3) 2.28, (1000, 1e-4) 2.56; no start 3.02.
MU0_POLICY = (1000, 1e-3, 10.0)


class BatchedSCvx:
    """All agents run the reference's outer loop in lock step on the device."""

    def __init__(self, models, K, max_iter=MAX_ITER, tr_radius0=TRUST_RADIUS0, conv_tol=CONV_TOL,
                 weight_nu=WEIGHT_NU, weight_slack=WEIGHT_SLACK, weight_sigma=WEIGHT_SIGMA, n_sub=0,
                 ipm_max_iter=0, device=None, batch=None, adaptive_mu0=False):
        self.batch = batch if batch is not None else AgentBatch(models, K, device)
        # adaptive_mu0: warm sub-problems start at a small barrier parameter (MU0_POLICY) instead of the cold default 10
        self.adaptive_mu0 = bool(adaptive_mu0)
        self.mu0_policy = MU0_POLICY      # (easy_max_iters, mu0_easy, mu0_hard) of scvx_mu0_from_iters
        self.retry_stranded = True        # warm solves that strand from the small start are re-solved cold in the same step
        self._mu0 = None
        b = self.batch
        self.K, self.max_iter, self.tr_radius0, self.conv_tol = K, max_iter, tr_radius0, conv_tol
        self.weight_nu, self.weight_slack, self.weight_sigma = weight_nu, weight_slack, weight_sigma
        self.n_sub, self.ipm_max_iter = n_sub, ipm_max_iter
        n, dev = b.n, b.device
        self.ws = _device.SubproblemWorkspace(b.model_id, n, K, b.M, 0, dev)
        self.mats = tuple(torch.empty((n, r, K - 1), dtype=F64, device=dev)
                          for r in (b.n_x * b.n_x, b.n_x * b.n_u, b.n_x * b.n_u, b.n_x, b.n_x))
        self.obs_a = torch.empty((n, b.M, b.d, K), dtype=F64, device=dev)
        self.obs_b = torch.empty((n, b.M, K), dtype=F64, device=dev)
        self.launches = 0
        self._order = None            # longest-first block order from the previous iteration's IPM iteration counts
        self._order_buf = torch.empty(n, dtype=torch.int32, device=dev)
        self._mu0_buf = torch.empty(n, dtype=F64, device=dev)

    def iterate(self, X, U, sigma, tr, active, metrics_row, solver_events=None):
        """One outer iteration, in place on (X, U, sigma, tr, active): 5 kernel launches (FOH, obstacle linearisation, sub-problem,
        bookkeeping, and the longest-first block order for the NEXT iteration's sub-problem launch).
        solver_events = (start, end) CUDA events recorded around the sub-problem kernel on the launching stream."""
        b = self.batch
        _device.foh(b.model_id, X, U, sigma, self.n_sub, out=self.mats)
        if b.M:
            _device.linearize_obstacles(b.model_id, X, b.obs_c, b.obs_clear, out=(self.obs_a, self.obs_b))
        if solver_events is not None:
            solver_events[0].record(torch.cuda.current_stream())
        _device.solve_subproblem(self.ws, self.mats, X, U, sigma, tr, b.x_init, b.x_final, b.pos_lo, b.pos_hi,
                                 b.v_max, b.w_max, self.obs_a, self.obs_b, self.weight_nu, self.weight_slack,
                                 self.weight_sigma, max_iter=self.ipm_max_iter, block_order=self._order, mu0=self._mu0,
                                 active=active)
        if self._mu0 is not None and self.retry_stranded:
            # retry pass: the few warm solves that stranded from the small barrier start (status != optimal) again from the cold one;
            # every other block returns at once
            _device.solve_subproblem(self.ws, self.mats, X, U, sigma, tr, b.x_init, b.x_final, b.pos_lo, b.pos_hi,
                                     b.v_max, b.w_max, self.obs_a, self.obs_b, self.weight_nu, self.weight_slack,
                                     self.weight_sigma, max_iter=self.ipm_max_iter, active=active, retry_failed=True)
            self.launches += 2      # the list kernel and the retry pass
        if solver_events is not None:
            solver_events[1].record(torch.cuda.current_stream())
        self._order = _device.order_by_iters(self.ws.iters, out=self._order_buf)
        if self.adaptive_mu0:
            self._mu0 = _device.mu0_from_iters(self.ws.iters, self._mu0_buf, *self.mu0_policy)
            self.launches += 1
        _device.outer_update(b.model_id, b.M, self.conv_tol, self.ws.X, self.ws.U, self.ws.nu, self.ws.sigma,
                             self.ws.s_prime, X, U, sigma, tr, active, metrics_row)
        self.launches += 5 if b.M else 4

    # -- host-buffer API: the call a user holding numpy arrays makes ---------------------------------------------
    def make_host_buffers(self):
        """Pinned host staging for `iterate_host`: the iterate (X, U, sigma, tr, active) and the per-step metrics."""
        b = self.batch
        n, K = b.n, self.K
        pin = lambda *shape, dtype=F64: torch.empty(shape, dtype=dtype).pin_memory()   # noqa: E731
        host = {"X": pin(n, b.n_x, K), "U": pin(n, b.n_u, K), "sigma": pin(n), "tr": pin(n),
                "active": torch.ones(n, dtype=torch.int32).pin_memory(), "metrics": pin(n, 6)}
        dev = b.device
        self._dbuf = {"X": torch.empty((n, b.n_x, K), dtype=F64, device=dev), "U": torch.empty((n, b.n_u, K), dtype=F64, device=dev),
                      "sigma": torch.empty(n, dtype=F64, device=dev), "tr": torch.empty(n, dtype=F64, device=dev),
                      "active": torch.empty(n, dtype=torch.int32, device=dev), "metrics": torch.empty((n, 6), dtype=F64, device=dev)}
        self.h2d_bytes = sum(host[k].numel() * host[k].element_size() for k in ("X", "U", "sigma", "tr", "active"))
        self.d2h_bytes = sum(host[k].numel() * host[k].element_size() for k in ("X", "U", "sigma", "tr", "active", "metrics"))
        return host

    def iterate_host(self, host):
        """One outer iteration with HOST buffers: H2D of the iterate, the four kernels, D2H of the new iterate and
        the metrics (nu_norm, slack_norm, dx, du, ds, sigma) -- everything a host-side SCvx loop reads per iteration."""
        d = self._dbuf
        for k in ("X", "U", "sigma", "tr", "active"):
            d[k].copy_(host[k], non_blocking=True)
        self.iterate(d["X"], d["U"], d["sigma"], d["tr"], d["active"], d["metrics"])
        for k in ("X", "U", "sigma", "tr", "active", "metrics"):
            host[k].copy_(d[k], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host

    def solve(self, X0=None, U0=None, initial_sigma=1.0, early_exit=True, check_every=5):
        """Returns dict(X, U, sigma, metrics (iters, n, 6), n_iter (n,), status_hist (iters, n), objective_hist).
        early_exit=False runs exactly max_iter outer iterations (throughput runs)."""
        b = self.batch
        n, dev = b.n, b.device
        if X0 is None:
            X, U = b.initial_trajectories()
        else:
            X, U = _device._dev(X0).clone(), _device._dev(U0).clone()
        sigma = torch.full((n,), float(initial_sigma), dtype=F64, device=dev) if np.isscalar(initial_sigma) \
            else _device._dev(initial_sigma).clone()
        tr = torch.full((n,), float(self.tr_radius0), dtype=F64, device=dev)
        active = torch.ones(n, dtype=torch.int32, device=dev)
        metrics = torch.zeros((self.max_iter, n, 6), dtype=F64, device=dev)
        status = torch.zeros((self.max_iter, n), dtype=torch.int32, device=dev)
        ipm_iters = torch.zeros((self.max_iter, n), dtype=torch.int32, device=dev)
        objective = torch.zeros((self.max_iter, n), dtype=F64, device=dev)
        was_active = torch.zeros((self.max_iter, n), dtype=torch.int32, device=dev)
        done = 0
        for it in range(self.max_iter):
            was_active[it].copy_(active)
            self.iterate(X, U, sigma, tr, active, metrics[it])
            status[it].copy_(self.ws.status); ipm_iters[it].copy_(self.ws.iters); objective[it].copy_(self.ws.objective)
            done = it + 1
            if early_exit and (it + 1) % check_every == 0 and int(active.sum().item()) == 0:
                break
        return {"X": X, "U": U, "sigma": sigma, "tr_radius": tr, "active": active, "metrics": metrics[:done],
                "status": status[:done], "ipm_iters": ipm_iters[:done], "objective": objective[:done],
                "was_active": was_active[:done], "n_outer": done}


def lane_bounds(n, n_lanes):
    """Contiguous lanes of ceil(n / n_lanes) agents (the last one ragged); never more lanes than agents."""
    if n <= 0:
        raise ValueError("empty agent batch")
    n_lanes = max(1, min(int(n_lanes), n))
    per = (n + n_lanes - 1) // n_lanes
    return [(a, min(a + per, n)) for a in range(0, n, per)]


class PipelinedSCvx:
    """`BatchedSCvx` over `n_lanes` sub-batches of the agents, each advancing its own outer loop on its own CUDA stream.

    The agents of SCVXSolver.solve are independent (scvx_solver.py:33-115 holds no cross-agent state), so nothing forces one
    sub-batch to wait for another sub-batch's slowest interior-point solve.  A single launch over all agents ends with most
    SMs idle behind the few agents that need 40 interior-point iterations; with lanes, the tail of one lane's
    sub-problem launch is filled by the other lanes' kernels and throughput approaches the total-work bound.  Results are
    identical to `BatchedSCvx` (same kernels, agents never interact); only the order of execution differs."""

    def __init__(self, models, K, n_lanes=4, max_iter=MAX_ITER, device=None, **kw):
        n = len(models)
        self.bounds = lane_bounds(n, n_lanes)
        self.n, self.K, self.max_iter = n, K, max_iter
        # every lane pads its obstacle table to the batch-wide maximum, so a lane solves exactly the problems the single
        # launch would (padding rows are inactive but do enter the interior-point iterates in the last bits)
        m_all = max(len(m.get_constraints().obs_clearance) for m in models)
        self.engines = [BatchedSCvx(None, K, max_iter=max_iter, device=device,
                                    batch=AgentBatch(models[a:b], K, device, min_obstacles=m_all), **kw) for a, b in self.bounds]
        self.device = self.engines[0].batch.device
        self.streams = [torch.cuda.Stream(device=self.device) for _ in self.engines]
        self.state = None
        self.it = 0

    @property
    def launches(self):
        return getattr(self, "_launch_base", 0) + sum(e.launches for e in self.engines)

    def start(self, X0=None, U0=None, initial_sigma=1.0, tr_radius0=None):
        """Per-lane iterate (X, U, sigma, trust radius, active flags) and the metrics record of every outer iteration."""
        self.state = []
        for eng, (a, b) in zip(self.engines, self.bounds):
            bt = eng.batch
            if X0 is None:
                X, U = bt.initial_trajectories()
            else:
                X, U = _device._dev(X0)[a:b].clone(), _device._dev(U0)[a:b].clone()
            m = b - a
            sig = torch.full((m,), float(initial_sigma), dtype=F64, device=self.device)
            tr = torch.full((m,), float(eng.tr_radius0 if tr_radius0 is None else tr_radius0), dtype=F64, device=self.device)
            act = torch.ones(m, dtype=torch.int32, device=self.device)
            met = torch.zeros((self.max_iter, m, 6), dtype=F64, device=self.device)
            self.state.append([X, U, sig, tr, act, met])
        self.it = 0
        torch.cuda.current_stream(self.device).synchronize()
        return self

    def _fork(self):
        main = torch.cuda.current_stream(self.device)
        go = torch.cuda.Event(); go.record(main)
        for st in self.streams:
            st.wait_event(go)
        return main

    def _join(self, main):
        for st in self.streams:
            done = torch.cuda.Event(); done.record(st)
            main.wait_event(done)

    def run(self, n_iter):
        """Enqueue n_iter outer iterations of every lane (lane-major round robin per iteration, so the host feeds all
        streams evenly) and make the CALLING stream wait for all of them; no host synchronisation."""
        if self.it + n_iter > self.max_iter:
            raise ValueError("PipelinedSCvx.run: more iterations than max_iter")
        main = self._fork()
        for k in range(n_iter):
            it = self.it + k
            for eng, st, (X, U, sig, tr, act, met) in zip(self.engines, self.streams, self.state):
                with torch.cuda.stream(st):
                    eng.iterate(X, U, sig, tr, act, met[it])
        self.it += n_iter
        self._join(main)

    # -- CUDA-graph replay: the host stops pacing the 5 launches x n_lanes of every step ----------------------------------
    def build_graph(self, steps_per_graph=1):
        """Capture `steps_per_graph` outer iterations of ALL lanes (forked onto the lane streams, joined at the end) into one
        CUDA graph.  Every pointer a step uses is fixed (iterate, workspaces, block-order buffers), so a replay is exactly
        the launches `run` would make; the per-step metrics land in a fixed staging table that `run_graph` copies into the
        history.  One eager step runs first if none has yet (the block-order pointer of the first solve is NULL)."""
        if self.state is None:
            raise RuntimeError("PipelinedSCvx.build_graph: call start() first")
        if any(e._order is None for e in self.engines):
            self.run(1)
        self._gsteps = int(steps_per_graph)
        self._gmet = [torch.zeros((self._gsteps, c - a, 6), dtype=F64, device=self.device) for a, c in self.bounds]
        torch.cuda.current_stream(self.device).synchronize()
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        launches0 = self.launches
        with torch.cuda.graph(graph, stream=side):
            main = self._fork()
            for k in range(self._gsteps):
                for eng, st, gm, (X, U, sig, tr, act, _met) in zip(self.engines, self.streams, self._gmet, self.state):
                    with torch.cuda.stream(st):
                        eng.iterate(X, U, sig, tr, act, gm[k])
            self._join(main)
        self._graph = graph
        self._graph_launches = self.launches - launches0      # kernels per replay (capturing launched none of them)
        for e in self.engines:
            e.launches = 0
        self._launch_base = launches0
        return self

    def build_lane_graphs(self):
        """One CUDA graph PER LANE holding one outer iteration of that lane (its 5 launches), replayed on the lane's own stream:
        the lanes stay independent -- `build_graph` joins all lanes at the end of every replay, which was measured to cost the
        whole benefit of the lanes (profiles/r02b_graph_lanes.jsonl) -- and the host issues n_lanes graph launches per step
        instead of 5 x n_lanes kernel launches through Python."""
        if self.state is None:
            raise RuntimeError("PipelinedSCvx.build_lane_graphs: call start() first")
        if any(e._order is None for e in self.engines):
            self.run(1)
        torch.cuda.current_stream(self.device).synchronize()
        self._lmet = [torch.zeros((c - a, 6), dtype=F64, device=self.device) for a, c in self.bounds]
        self._lgraphs = []
        launches0 = self.launches
        for eng, st, lm, (X, U, sig, tr, act, _met) in zip(self.engines, self.streams, self._lmet, self.state):
            g = torch.cuda.CUDAGraph()
            st.synchronize()
            with torch.cuda.graph(g, stream=st):
                eng.iterate(X, U, sig, tr, act, lm)
            self._lgraphs.append(g)
        self._lane_graph_launches = self.launches - launches0
        for e in self.engines:
            e.launches = 0
        self._launch_base = launches0
        return self

    def run_lane_graphs(self, n_iter, keep_history=True):
        """n_iter outer iterations of every lane by replaying the lane graphs on the lane streams; the calling stream waits
        for all lanes at the end (no host synchronisation)."""
        if self.it + n_iter > self.max_iter:
            raise ValueError("PipelinedSCvx.run_lane_graphs: more iterations than max_iter")
        main = self._fork()
        for k in range(n_iter):
            for g, st, lm, s_ in zip(self._lgraphs, self.streams, self._lmet, self.state):
                with torch.cuda.stream(st):
                    g.replay()
                    if keep_history:
                        s_[5][self.it + k].copy_(lm, non_blocking=True)
            self._launch_base += self._lane_graph_launches
        self.it += n_iter
        self._join(main)

    def run_graph(self, n_iter, keep_history=True):
        """n_iter outer iterations by graph replay (n_iter must be a multiple of the captured steps_per_graph)."""
        if n_iter % self._gsteps:
            raise ValueError("PipelinedSCvx.run_graph: n_iter must be a multiple of steps_per_graph")
        if self.it + n_iter > self.max_iter:
            raise ValueError("PipelinedSCvx.run_graph: more iterations than max_iter")
        for _ in range(n_iter // self._gsteps):
            self._graph.replay()
            self._launch_base += self._graph_launches
            if keep_history:
                for gm, s in zip(self._gmet, self.state):
                    s[5][self.it:self.it + self._gsteps].copy_(gm, non_blocking=True)
            self.it += self._gsteps

    # -- host-buffer API: the call a user holding pinned numpy/torch host arrays makes -------------------------------
    def make_host_buffers(self):
        """Pinned host staging of the whole batch (lanes use contiguous slices of it): iterate + per-step metrics."""
        b = self.engines[0].batch
        n, K = self.n, self.K
        pin = lambda *shape, dtype=F64: torch.empty(shape, dtype=dtype).pin_memory()   # noqa: E731
        host = {"X": pin(n, b.n_x, K), "U": pin(n, b.n_u, K), "sigma": pin(n), "tr": pin(n),
                "active": torch.ones(n, dtype=torch.int32).pin_memory(), "metrics": pin(n, 6)}
        dev = self.device
        self._dbuf = []
        for a, c in self.bounds:
            m = c - a
            self._dbuf.append({"X": torch.empty((m, b.n_x, K), dtype=F64, device=dev), "U": torch.empty((m, b.n_u, K), dtype=F64, device=dev),
                               "sigma": torch.empty(m, dtype=F64, device=dev), "tr": torch.empty(m, dtype=F64, device=dev),
                               "active": torch.empty(m, dtype=torch.int32, device=dev),
                               "metrics": torch.empty((m, 6), dtype=F64, device=dev)})
        self.h2d_bytes = sum(host[k].numel() * host[k].element_size() for k in ("X", "U", "sigma", "tr", "active"))
        self.d2h_bytes = sum(host[k].numel() * host[k].element_size() for k in ("X", "U", "sigma", "tr", "active", "metrics"))
        return host

    def run_host(self, host, n_steps=1):
        """n_steps outer iterations with HOST buffers.  Every step of every lane copies its slice of the iterate from the
        pinned host arrays to the device, runs the step's kernels and copies the new iterate + metrics back, all on the
        lane's stream (a lane's next upload is ordered after its previous download, so the host arrays always carry the
        latest iterate); the call returns when everything has landed on the host."""
        main = self._fork()
        for _ in range(n_steps):
            self._enqueue_host_step(host)
        self._join(main)
        main.synchronize()
        return host

    def _host_step_of_lane(self, host, eng, d, a, c):
        for k in ("X", "U", "sigma", "tr", "active"):
            d[k].copy_(host[k][a:c], non_blocking=True)
        eng.iterate(d["X"], d["U"], d["sigma"], d["tr"], d["active"], d["metrics"])
        for k in ("X", "U", "sigma", "tr", "active", "metrics"):
            host[k][a:c].copy_(d[k], non_blocking=True)

    def _enqueue_host_step(self, host):
        for eng, st, d, (a, c) in zip(self.engines, self.streams, self._dbuf, self.bounds):
            with torch.cuda.stream(st):
                self._host_step_of_lane(host, eng, d, a, c)

    def build_lane_host_graphs(self, host):
        """One CUDA graph PER LANE holding one host-buffer step of that lane: the H2D copies of its slice of the pinned host arrays
        (memcpy nodes), the step's kernels and the D2H copies.  Replayed on the lane's own stream by `run_host_lane_graphs`, the
        lanes stay independent (as in `build_lane_graphs`) and the host issues n_lanes graph launches per step instead of
        ~20 x n_lanes copies and kernel launches through Python -- which is what bounds `run_host` beyond a few lanes."""
        if any(e._order is None for e in self.engines) or any(e.adaptive_mu0 and e._mu0 is None for e in self.engines):
            self.run_host(host, 1)
        torch.cuda.current_stream(self.device).synchronize()
        self._hlgraphs, self._hl_host = [], host
        launches0 = self.launches
        for eng, st, d, (a, c) in zip(self.engines, self.streams, self._dbuf, self.bounds):
            g = torch.cuda.CUDAGraph()
            st.synchronize()
            with torch.cuda.graph(g, stream=st):
                self._host_step_of_lane(host, eng, d, a, c)
            self._hlgraphs.append(g)
        self._hl_launches = self.launches - launches0
        for e in self.engines:
            e.launches = 0
        self._launch_base = launches0
        return self

    def run_host_lane_graphs(self, host, n_steps=1):
        """`run_host` by replaying the lane graphs of `build_lane_host_graphs` (same host arrays): returns when every lane's
        result of the last step is on the host."""
        if host is not self._hl_host:
            raise ValueError("PipelinedSCvx.run_host_lane_graphs: the graphs were captured on other host buffers")
        main = self._fork()
        for _ in range(n_steps):
            for g, st in zip(self._hlgraphs, self.streams):
                with torch.cuda.stream(st):
                    g.replay()
            self._launch_base += self._hl_launches
        self._join(main)
        main.synchronize()
        return host

    def build_host_graph(self, host):
        """`run_host(host, 1)` as one CUDA graph: per lane the H2D copies of its slice, the step's kernels and the D2H copies
        (memcpy nodes on pinned memory), forked over the lane streams.  Replay with `run_host_graph`."""
        if any(e._order is None for e in self.engines):
            self.run_host(host, 1)
        torch.cuda.current_stream(self.device).synchronize()
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=self.device)
        launches0 = self.launches
        with torch.cuda.graph(graph, stream=side):
            main = self._fork()
            self._enqueue_host_step(host)
            self._join(main)
        self._hgraph, self._hgraph_launches = graph, self.launches - launches0
        for e in self.engines:
            e.launches = 0
        self._launch_base = launches0
        return self

    def run_host_graph(self, host, n_steps=1):
        for _ in range(n_steps):
            self._hgraph.replay()
            self._launch_base += self._hgraph_launches
        torch.cuda.current_stream(self.device).synchronize()
        return host

    def gather(self):
        """Concatenated iterate of all lanes: dict(X, U, sigma, tr_radius, active, metrics (iterations run, n, 6))."""
        cat = lambda i, dim=0: torch.cat([s[i] for s in self.state], dim=dim)     # noqa: E731
        return {"X": cat(0), "U": cat(1), "sigma": cat(2), "tr_radius": cat(3), "active": cat(4),
                "metrics": cat(5, dim=1)[:self.it], "n_outer": self.it}

    def ipm_iters(self):
        return torch.cat([e.ws.iters for e in self.engines])

    def status(self):
        return torch.cat([e.ws.status for e in self.engines])

    def solve(self, X0=None, U0=None, initial_sigma=1.0, early_exit=True, check_every=5):
        """All agents' outer loops (scvx_solver.py:33-115), lanes pipelined; the host looks at the active flags every
        `check_every` iterations only."""
        self.start(X0, U0, initial_sigma)
        while self.it < self.max_iter:
            self.run(min(check_every, self.max_iter - self.it))
            if early_exit and int(sum(int(s[4].sum().item()) for s in self.state)) == 0:
                break
        return self.gather()


def shard_bounds(N, world, rank):
    """Contiguous block partition of N agents over `world` ranks: (per, i0, i1) with per = ceil(N/world)."""
    per = (N + world - 1) // world
    i0 = min(rank * per, N)
    return per, i0, min(i0 + per, N)


def allgather_shards(X_local, U_local, N, per, dist=None, group=None):
    """The ONE collective of an ADMM round: every rank contributes its shard's (X, U) -- padded to `per`
    agents so that all ranks send the same count -- and receives all N agents' trajectories in agent order.
    Works on any backend (NCCL on the GPUs; gloo in the CPU tests)."""
    if dist is None or dist.get_world_size(group) == 1:
        return X_local, U_local
    world = dist.get_world_size(group)
    nl, n_x, K = X_local.shape
    n_u = U_local.shape[1]
    send = torch.zeros((per, n_x + n_u, K), dtype=X_local.dtype, device=X_local.device)
    if nl:
        send[:nl, :n_x] = X_local
        send[:nl, n_x:] = U_local
    recv = torch.empty((world * per, n_x + n_u, K), dtype=X_local.dtype, device=X_local.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    return recv[:N, :n_x].contiguous(), recv[:N, n_x:].contiguous()


def exchange_states(send, X_all, dist=None, group=None):
    """The ONE collective of an ADMM round.  `send` (per, n_x, K): this rank's new states, zero-padded to `per` agents so that
    every rank contributes the same count; `X_all` (world * per, n_x, K) receives all shards in rank (= agent) order, in place.
    Only the last rank can be ragged, so rows 0..N-1 of X_all are exactly the N agents.  Any backend (NCCL on the GPUs, gloo in
    the CPU tests); a single rank copies."""
    if dist is None or dist.get_world_size(group) == 1:
        X_all[:send.shape[0]].copy_(send)
    else:
        dist.all_gather_into_tensor(X_all, send, group=group)
    return X_all


class BatchedADMM:
    """Jacobi-sweep ADMM over N agents, optionally sharded across a process group.

    Per round, for the local shard: FOH about the current own trajectory, obstacle linearisation, ONE launch that builds the
    inter-agent half-space tables, their right-hand sides and the collapsed augmented-Lagrangian terms in the solver's layout
    (`scvx_admm_round_prep`: multi_agent_model.py:61-79 + agent_solver.py:79-95 collapsed to sum_j Y_j / sum_j Lambda_j,
    SURVEY A.3), the QP/SOCP sub-problem, then ONE all-gather of the new states and the redundant per-j consensus update
    (admm_coordinator.py:80-96).  `si_variant=True` reproduces si_admm_coordinator.py:80-86 (trust region / obstacle
    linearisation about the INITIAL references every round).

    The exchange is zero-copy: the solver writes its states straight into the all-gather's send buffer and every kernel reads
    the gathered states in place (positions are the first d rows of a state block).  Inputs are only the neighbours' STATES:
    the controls of other agents are never needed, so they are gathered once, after the last round.  No PyTorch arithmetic
    runs between the exchange and the sub-problem solve; histories are accumulated into preallocated tables by device copies
    and reduced once at the end.
    """

    def __init__(self, models, d_min, K, rho_admm=1.0, max_iter=10, si_variant=False, group=None, device=None,
                 n_sub=0, ipm_max_iter=0, neighbor_radius=None, neighbor_k=None):
        import torch.distributed as dist
        self.dist = dist if (group is not None or (dist.is_available() and dist.is_initialized())) else None
        self.group = group
        self.rank = self.dist.get_rank(group) if self.dist else 0
        self.world = self.dist.get_world_size(group) if self.dist else 1
        self.N = len(models)
        self.K, self.d_min, self.rho, self.max_iter, self.si_variant = K, float(d_min), float(rho_admm), max_iter, si_variant
        self.n_sub, self.ipm_max_iter = n_sub, ipm_max_iter
        self.neighbor_radius = neighbor_radius
        # neighbor_k: couple every agent with its k nearest neighbours only (compact tables; needed when N^2 tables do not
        # fit: 8192 agents x K=200 would need 26 GB per GPU of half-space tables) -- a documented deviation from the reference
        self.neighbor_k = None if neighbor_k is None else int(min(neighbor_k, max(len(models) - 1, 1)))
        self.per, self.i0, self.i1 = shard_bounds(self.N, self.world, self.rank)
        self.nl = self.i1 - self.i0
        self.all = AgentBatch(models, K, device)                       # tables for shapes / initial Y
        self.local = AgentBatch(models[self.i0:self.i1], K, device) if self.nl else None
        b = self.all
        dev = b.device
        # exchange buffers: `send` holds this rank's (padded) shard of new states, `X_all` every rank's shard in agent order
        self._send = torch.zeros((self.per, b.n_x, K), dtype=F64, device=dev)
        self._X_all = torch.zeros((self.world * self.per, b.n_x, K), dtype=F64, device=dev)
        if self.nl:
            lb = self.local
            n_slots = self.neighbor_k if self.neighbor_k is not None else self.N
            self.ws = _device.SubproblemWorkspace(lb.model_id, self.nl, K, lb.M, n_slots, dev, X_out=self._send[:self.nl])
            self.mats = tuple(torch.empty((self.nl, r, K - 1), dtype=F64, device=dev)
                              for r in (b.n_x * b.n_x, b.n_x * b.n_u, b.n_x * b.n_u, b.n_x, b.n_x))
            self.obs_a = torch.empty((self.nl, lb.M, b.d, K), dtype=F64, device=dev)
            self.obs_b = torch.empty((self.nl, lb.M, K), dtype=F64, device=dev)
            self.tab = _device.AdmmRoundTables(lb.model_id, self.nl, n_slots, K, dev)
            self.col_a, self.col_b = self.tab.col_a, self.tab.col_b
            self._U_cur = torch.empty((self.nl, b.n_u, K), dtype=F64, device=dev)
            need_d2 = self.neighbor_k is not None or self.neighbor_radius is not None
            self._d2 = torch.empty((self.nl, self.N), dtype=F64, device=dev) if need_d2 else None
            self._nbr_idx = torch.empty((self.nl, n_slots), dtype=torch.int32, device=dev) if self.neighbor_k is not None else None
            self._mask_in = torch.empty((self.nl, self.N), dtype=torch.uint8, device=dev) \
                if (self.neighbor_k is None and self.neighbor_radius is not None) else None
            self._order_buf = torch.empty(self.nl, dtype=torch.int32, device=dev)
        self.launches = 0
        self.profile = False             # True: CUDA events around the round's all-gather (see allgather_ms)
        self._gather_events = []

    def _exchange(self):
        """The ONE collective of a round: every rank's send buffer (its new states, padded to `per` agents so that all ranks
        send the same count) lands in X_all in agent order.  One rank: a device-to-device copy."""
        exchange_states(self._send, self._X_all, self.dist, self.group)

    def allgather_ms(self):
        """Device time of the all-gathers of the solves since the last call (profile=True), this rank; synchronises."""
        torch.cuda.synchronize(self.all.device)
        ms = sum(a.elapsed_time(b) for a, b in self._gather_events)
        n = len(self._gather_events)
        self._gather_events = []
        return ms, n

    @property
    def allgather_bytes(self):
        """Bytes every rank RECEIVES per round: the padded state shards of all ranks."""
        b = self.all
        return 0 if self.world == 1 else int(self.world * self.per * b.n_x * self.K * 8)

    def solve(self, X_refs, U_refs, sigma_ref):
        """X_refs (N, n_x, K), U_refs (N, n_u, K) device tensors (every rank holds all of them).
        Returns dict(X, U (all agents, on every rank), primal_hist, dual_hist, objective (rounds, N_local))."""
        b = self.all
        dev, K, N, d, R = b.device, self.K, self.N, b.d, self.max_iter
        X_refs, U_refs = _device._dev(X_refs), _device._dev(U_refs)
        X_all = self._X_all[:N]                                   # gathered states, agent order (contiguous prefix)
        X_all.copy_(X_refs)
        Y = X_refs[:, :d, :].clone()                              # per-j consensus state; a private copy: it is
        #                                                           updated in place and must not alias the caller's references
        Lam = torch.zeros_like(Y)
        pr_hist = torch.zeros((R, N), dtype=F64, device=dev); du_hist = torch.zeros((R, N), dtype=F64, device=dev)
        obj_hist = torch.zeros((R, max(self.nl, 1)), dtype=F64, device=dev); const_hist = torch.zeros_like(obj_hist)
        if self.nl:
            lb = self.local
            X0_loc = X_refs[self.i0:self.i1].clone(); U0_loc = U_refs[self.i0:self.i1].clone()
            self._U_cur.copy_(U0_loc)
            sig = torch.full((self.nl,), float(sigma_ref), dtype=F64, device=dev)
            tr = torch.full((self.nl,), float(TRUST_RADIUS0), dtype=F64, device=dev)
        order = None
        for r in range(R):
            if self.nl:
                X_loc = X_all[self.i0:self.i1]                    # own states of the previous round, read in place
                Xr, Ur = (X0_loc, U0_loc) if self.si_variant else (X_loc, self._U_cur)
                _device.foh(lb.model_id, X_loc, self._U_cur, sig, self.n_sub, out=self.mats)
                if lb.M:
                    _device.linearize_obstacles(lb.model_id, Xr, lb.obs_c, lb.obs_clear, out=(self.obs_a, self.obs_b))
                nbr_idx = mask_in = None
                if self._d2 is not None:
                    # neighbour culling (documented deviations for large N): the k nearest neighbours by minimum distance over
                    # the horizon (compact slot tables) and / or only neighbours that come within neighbor_radius at some node
                    _device.cross_min_dist2(lb.model_id, Xr, X_all, out=self._d2)
                    if self.neighbor_k is not None:
                        nbr_idx = _device.knn_select(self._d2, self.i0, self.neighbor_k, self.neighbor_radius, out=self._nbr_idx)
                        self.last_nbr_idx = nbr_idx
                    else:
                        mask_in = _device.radius_mask(self._d2, self.neighbor_radius, out=self._mask_in)
                    self.launches += 2
                # a_ij about (own reference, neighbour's current trajectory); b = d_min + a.Y_j; sum_j Lambda_j - rho sum_j Y_j
                _device.admm_round_prep(lb.model_id, Xr, X_all, Y, Lam, self.d_min, self.rho, self.i0, self.tab,
                                        nbr_idx=nbr_idx, mask_in=mask_in)
                _device.solve_subproblem(self.ws, self.mats, Xr, Ur, sig, tr, lb.x_init, lb.x_final, lb.pos_lo,
                                         lb.pos_hi, lb.v_max, lb.w_max, self.obs_a, self.obs_b, WEIGHT_NU,
                                         WEIGHT_SLACK, WEIGHT_SIGMA, col_a=self.tab.col_a, col_b=self.tab.col_b,
                                         col_mask=self.tab.mask, quad_rho=self.tab.quad_rho, lin_p=self.tab.lin_p,
                                         weight_col=WEIGHT_COLLISION_SLACK, max_iter=self.ipm_max_iter, block_order=order)
                order = _device.order_by_iters(self.ws.iters, out=self._order_buf)   # longest-first launch order for the next round
                # the objective of agent_solver.py:92-95 = kernel objective + the constant of the collapsed Lagrangian
                obj_hist[r, :self.nl].copy_(self.ws.objective); const_hist[r, :self.nl].copy_(self.tab.aug_const)
                self._U_cur.copy_(self.ws.U)
                self.launches += 5 if lb.M else 4
            if self.profile:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record()
            self._exchange()
            if self.profile:
                ev[1].record(); self._gather_events.append(ev)
            _device.consensus_update_x(X_all, d, Y, Lam, self.rho, out=(pr_hist[r], du_hist[r]))
            self.launches += 1
        # other agents' controls are never an input of a round: one exchange at the end, for the caller
        if self.dist is None or self.world == 1:
            U_all = self._U_cur.clone() if self.nl else torch.empty((0, b.n_u, K), dtype=F64, device=dev)
        else:
            usend = torch.zeros((self.per, b.n_u, K), dtype=F64, device=dev)
            if self.nl:
                usend[:self.nl] = self._U_cur
            urecv = torch.empty((self.world * self.per, b.n_u, K), dtype=F64, device=dev)
            self.dist.all_gather_into_tensor(urecv, usend, group=self.group)
            U_all = urecv[:N]
        return {"X": X_all.clone(), "U": U_all, "Y": Y, "Lambda": Lam,
                "primal_hist": pr_hist.mean(dim=1).cpu().tolist(), "dual_hist": du_hist.mean(dim=1).cpu().tolist(),
                "objective": (obj_hist + const_hist)[:, :self.nl] if self.nl else None,
                "mask": self.tab.mask if self.nl else None}


SLAB_PENALTIES = (1e7, 1e8)      # exact-penalty weights of the hard slab rows, escalated while a slack remains
SLAB_PENALTY = SLAB_PENALTIES[0]  # (see optimization/agent_best_response.py)
SLAB_TOL = 1e-7


class BatchedNash:
    """Iterative best response (NashSolver.solve, SCvx/optimization/nash_solver.py:40-149) as JACOBI sweeps: every agent of
    the local shard computes its best response -- with the ACS inner loop of slab-normal updates -- against the
    trajectories of the previous sweep, all agents in one launch per step; one all-gather per sweep.  (The reference sweeps
    Gauss-Seidel; agent 0's first best response is identical in both orders, which is what the parity test pins.)
    """

    def __init__(self, models, K, max_iter=20, tol=1e-3, max_acs_iters=5, acs_tol=1e-3, group=None, device=None,
                 n_sub=0, ipm_max_iter=0, neighbor_radius=None, n_colors=1):
        import torch.distributed as dist
        from .optimization.agent_best_response import game_tables
        self.dist = dist if (group is not None or (dist.is_available() and dist.is_initialized())) else None
        self.group = group
        self.rank = self.dist.get_rank(group) if self.dist else 0
        self.world = self.dist.get_world_size(group) if self.dist else 1
        self.N, self.K = len(models), K
        self.max_iter, self.tol, self.max_acs_iters, self.acs_tol = max_iter, tol, max_acs_iters, acs_tol
        self.n_sub, self.ipm_max_iter = n_sub, ipm_max_iter
        # neighbor_radius: keep the slab rows of neighbour j only if the two trajectories come within this distance at some
        # knot of the sweep's reference (a documented deviation: the reference keeps all N-1 neighbours; with hundreds of far,
        # inactive penalty rows per stage the interior-point method needs several times the iterations)
        self.neighbor_radius = neighbor_radius
        self.n_colors = max(1, int(n_colors))
        self.per, self.i0, self.i1 = shard_bounds(self.N, self.world, self.rank)
        self.nl = self.i1 - self.i0
        self.all = AgentBatch(models, K, device)
        self.local = AgentBatch(models[self.i0:self.i1], K, device) if self.nl else None
        self._tables = game_tables
        b = self.all
        dev = b.device
        self.radius = torch.as_tensor(np.array([m.collision_radius for m in models[self.i0:self.i1]], dtype=np.float64)).to(dev)
        self.inertia = [getattr(m, "inertia_weight", 0.0) for m in models[self.i0:self.i1]]
        if self.nl:
            lb = self.local
            self.ws = _device.SubproblemWorkspace(lb.model_id, self.nl, K, lb.M, self.N, dev)
            self.mats = tuple(torch.empty((self.nl, r, K - 1), dtype=F64, device=dev)
                              for r in (b.n_x * b.n_x, b.n_x * b.n_u, b.n_x * b.n_u, b.n_x, b.n_x))
            self.obs_a = torch.empty((self.nl, lb.M, b.d, K), dtype=F64, device=dev)
            self.obs_b = torch.empty((self.nl, lb.M, K), dtype=F64, device=dev)
            self.col_a = torch.empty((self.nl, self.N, b.d, K), dtype=F64, device=dev)
            self.col_b = torch.empty((self.nl, self.N, K), dtype=F64, device=dev)
            self.mask = torch.ones((self.nl, self.N), dtype=torch.uint8, device=dev)
            self.mask[torch.arange(self.nl, device=dev), torch.arange(self.i0, self.i1, device=dev)] = 0
            self.self_mask = self.mask.clone()
        self.launches = 0
        self._order = None

    def _cost_tables(self, X_prev_local):
        """quad_diag (nl, ns), lin_w (nl, ns, K), quad_pair (nl, ns), const (nl,) of the local agents."""
        b = self.all
        Xp = X_prev_local.cpu().numpy() if any(w > 0 for w in self.inertia) else None
        qd, lw, qp, cst = [], [], [], []
        for q, m in enumerate(self.local.models):
            t = self._tables(m, Xp[q] if Xp is not None else np.zeros((b.n_x, self.K)), self.K, b.n_x, b.n_u)
            qd.append(t[0]); lw.append(t[1]); qp.append(t[2]); cst.append(t[3])
        up = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(b.device)   # noqa: E731
        return up(np.stack(qd)), up(np.stack(lw)), up(np.stack(qp)), up(np.array(cst))

    def _best_responses(self, X_all, U_all, X_prev_all, sig, tr):
        """Best responses (with the ACS loop) of ALL local agents against the trajectories X_all; the initial slab normals point
        along the sweep-start snapshot X_prev_all (agent_best_response.py:66-72).  Returns X_new, U_new, delta, acs, bad, cap, obj."""
        b, lb = self.all, self.local
        dev, N = b.device, self.N
        X_loc = X_all[self.i0:self.i1].contiguous(); U_loc = U_all[self.i0:self.i1].contiguous()
        _device.foh(lb.model_id, X_loc, U_loc, sig, self.n_sub, out=self.mats)
        if lb.M:
            _device.linearize_obstacles(lb.model_id, X_loc, lb.obs_c, lb.obs_clear, out=(self.obs_a, self.obs_b))
        qd, lw, qp, cst = self._cost_tables(X_loc)
        if self.neighbor_radius is not None:
            d2 = _device.cross_min_dist2(lb.model_id, X_loc, X_all)
            self.mask = self.self_mask & (d2 <= self.neighbor_radius ** 2).to(torch.uint8)
        _, _, deg = _device.slab_normals(lb.model_id, X_loc, X_prev_all, X_all, self.radius, i0=self.i0, out=(self.col_a, self.col_b))
        X_new = X_loc.clone(); U_new = U_loc.clone()
        done = torch.zeros(self.nl, dtype=torch.bool, device=dev)
        acs = torch.zeros(self.nl, dtype=torch.int32, device=dev)
        bad = deg > 0
        not_conv = torch.zeros(self.nl, dtype=torch.bool, device=dev)
        obj = torch.zeros(self.nl, dtype=F64, device=dev)
        for _a in range(self.max_acs_iters):
            live = ~done
            todo = live.clone()                       # agents still waiting for a slack-free solve
            for penalty in ((SLAB_PENALTY,) + tuple(p for p in SLAB_PENALTIES if p > SLAB_PENALTY)):
                _device.solve_subproblem(self.ws, self.mats, X_loc, U_loc, sig, tr, lb.x_init, lb.x_final, lb.pos_lo,
                                         lb.pos_hi, lb.v_max, lb.w_max, self.obs_a, self.obs_b, WEIGHT_NU, WEIGHT_SLACK,
                                         WEIGHT_SIGMA, col_a=self.col_a, col_b=self.col_b, col_mask=self.mask,
                                         weight_col=penalty, max_iter=self.ipm_max_iter, quad_diag=qd, lin_w=lw,
                                         quad_pair=qp, fix_sigma=True, block_order=self._order)
                self._order = _device.order_by_iters(self.ws.iters)
                self.launches += 1
                X_new[todo] = self.ws.X[todo]; U_new[todo] = self.ws.U[todo]
                obj[todo] = self.ws.objective[todo] + cst[todo]
                slack = (self.ws.col_slack * self.mask[:, :, None]).amax(dim=(1, 2)) if N > 1 else torch.zeros(self.nl, device=dev, dtype=F64)
                failed = todo & ((slack > SLAB_TOL) | (self.ws.status == 2))
                not_conv |= todo & (self.ws.status == 1)
                todo = failed
                if not bool(todo.any().item()):
                    break
            bad |= todo                               # still a residual slack at the largest weight: infeasible slabs
            acs += live.to(torch.int32)
            # dual update: normals along (new own position) - (neighbours' current positions)
            _, _, deg = _device.slab_normals(lb.model_id, X_new, X_all, X_all, self.radius, i0=self.i0, out=(self.col_a, self.col_b))
            bad |= deg > 0
            done |= torch.linalg.norm((X_new - X_loc).reshape(self.nl, -1), dim=1) < self.acs_tol
            if bool(done.all().item()):
                break
        delta = torch.linalg.norm((X_new - X_loc).reshape(self.nl, -1), dim=1)
        return X_new, U_new, delta, acs, bad, not_conv, obj

    def solve(self, X_refs, U_refs, sigma_ref=1.0):
        """X_refs (N, n_x, K), U_refs (N, n_u, K) device tensors on every rank.  Returns dict(X, U, change_hist,
        acs_iters (sweeps, nl), infeasible (sweeps, nl): residual slab slack or vanished normal -- the cases in which the
        reference raises; iteration_cap (sweeps, nl): a solve stopped at the interior-point iteration cap).
        With n_colors > 1 a sweep is split into colour phases (agent index mod n_colors): the agents of a phase respond to the
        ALREADY UPDATED trajectories of the earlier phases, one all-gather per phase -- Gauss-Seidel between colours."""
        b = self.all
        dev, K, N = b.device, self.K, self.N
        X_all = _device._dev(X_refs).clone(); U_all = _device._dev(U_refs).clone()
        sig = torch.full((max(self.nl, 1),), float(sigma_ref), dtype=F64, device=dev)[:self.nl]
        tr = torch.full((max(self.nl, 1),), float(TRUST_RADIUS0), dtype=F64, device=dev)[:self.nl]
        change_hist, acs_hist, bad_hist, obj_hist, cap_hist = [], [], [], [], []
        colour = (torch.arange(self.i0, self.i1, device=dev) % self.n_colors) if self.nl else None
        for _ in range(self.max_iter):
            X_prev_all = X_all.clone()
            mx = torch.zeros((), dtype=F64, device=dev)
            acs_s = bad_s = cap_s = obj_s = None
            for c in range(self.n_colors):
                if self.nl:
                    X_new, U_new, delta, acs, bad, cap, obj = self._best_responses(X_all, U_all, X_prev_all, sig, tr)
                    take = colour == c
                    keep = ~take
                    X_new[keep] = X_all[self.i0:self.i1][keep]; U_new[keep] = U_all[self.i0:self.i1][keep]
                    if bool(take.any().item()):
                        mx = torch.maximum(mx, delta[take].max())
                    if acs_s is None:
                        acs_s, bad_s, cap_s, obj_s = torch.zeros_like(acs), torch.zeros_like(bad), torch.zeros_like(cap), torch.zeros_like(obj)
                    acs_s[take] = acs[take]; bad_s[take] = bad[take]; cap_s[take] = cap[take]; obj_s[take] = obj[take]
                else:
                    X_new = torch.empty((0, b.n_x, K), dtype=F64, device=dev); U_new = torch.empty((0, b.n_u, K), dtype=F64, device=dev)
                X_all, U_all = allgather_shards(X_new, U_new, N, self.per, self.dist, self.group)
                X_all, U_all = X_all.clone(), U_all.clone()
            if self.nl:
                acs_hist.append(acs_s); bad_hist.append(bad_s); obj_hist.append(obj_s); cap_hist.append(cap_s)
            if self.dist and self.world > 1:
                mx = mx.clone()
                self.dist.all_reduce(mx, op=self.dist.ReduceOp.MAX, group=self.group)
            change_hist.append(float(mx.item()))
            if change_hist[-1] < self.tol:
                break
        return {"X": X_all, "U": U_all, "change_hist": change_hist,
                "acs_iters": torch.stack(acs_hist) if acs_hist else None,
                "infeasible": torch.stack(bad_hist) if bad_hist else None,
                "objective": torch.stack(obj_hist) if obj_hist else None,
                "iteration_cap": torch.stack(cap_hist) if cap_hist else None}
