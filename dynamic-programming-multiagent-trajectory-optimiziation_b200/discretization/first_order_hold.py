"""FirstOrderHold -- mirrors SCvx/discretization/first_order_hold.py:5-162 on the GPU.

Same constructor, same method names, same array shapes and the same buffer-aliasing behaviour
(`calculate_discretization` returns references to internal host buffers that the next call
overwrites, first_order_hold.py:20-24,87).  The K-1 `odeint` calls are replaced by ONE launch of the
RK4 kernel (`scvx_foh_batched`, csrc/foh.cu); host arrays go through pinned staging buffers.
"""
import numpy as np
import torch

from .. import _device


class FirstOrderHold:
    def __init__(self, model, K, n_sub=0):
        if getattr(model, "device_model_id", None) is None:
            raise NotImplementedError(
                "FirstOrderHold (GPU) needs a model with compiled device dynamics (`device_model_id`); "
                "arbitrary Python f/A/B callables cannot run inside the kernel.")
        self.model = model
        self.K = K
        self.n_x = model.n_x
        self.n_u = model.n_u
        self.n_sub = n_sub           # 0 = sub-step count chosen per interval on the device
        self.model_id = model.device_model_id
        n_x, n_u = self.n_x, self.n_u
        self.dt = 1.0 / (K - 1)
        self.f, self.A, self.B = model.get_equations()

        dev = torch.device("cuda")
        self._dev = dev
        # pinned host staging: [X | U | sigma] in, [A | B | C | S | z] out, one copy each way
        self._n_in = (n_x + n_u) * K + 1
        self._rows_out = (n_x * n_x, n_x * n_u, n_x * n_u, n_x, n_x)
        self._n_out = sum(self._rows_out) * (K - 1)
        self._h_in = torch.empty(self._n_in, dtype=torch.float64).pin_memory()
        self._h_out = torch.empty(self._n_out, dtype=torch.float64).pin_memory()
        self._d_in = torch.empty(self._n_in, dtype=torch.float64, device=dev)
        self._d_out = torch.empty(self._n_out, dtype=torch.float64, device=dev)
        self._np_in = self._h_in.numpy()
        np_out = self._h_out.numpy()
        # public buffers, shapes as in the reference (first_order_hold.py:20-24); views of the pinned block
        offs = np.cumsum((0,) + self._rows_out) * (K - 1)
        self.A_bar, self.B_bar, self.C_bar, self.S_bar, self.z_bar = (
            np_out[offs[i]:offs[i + 1]].reshape(self._rows_out[i], K - 1) for i in range(5))
        self._d_views_out = tuple(
            self._d_out[offs[i]:offs[i + 1]].view(1, self._rows_out[i], K - 1) for i in range(5))
        o1, o2 = n_x * K, (n_x + n_u) * K
        self._dX = self._d_in[:o1].view(1, n_x, K)
        self._dU = self._d_in[o1:o2].view(1, n_u, K)
        self._dS = self._d_in[o2:o2 + 1]

    def _stage_in(self, X, U, sigma):
        n_x, n_u, K = self.n_x, self.n_u, self.K
        X = np.asarray(X, dtype=np.float64)
        U = np.asarray(U, dtype=np.float64)
        if X.shape != (n_x, K) or U.shape != (n_u, K):
            raise ValueError(f"expected X {(n_x, K)} and U {(n_u, K)}, got {X.shape} and {U.shape}")
        self._np_in[:n_x * K] = X.reshape(-1)
        self._np_in[n_x * K:(n_x + n_u) * K] = U.reshape(-1)
        self._np_in[-1] = float(sigma)
        self._d_in.copy_(self._h_in, non_blocking=True)

    def calculate_discretization(self, X, U, sigma):
        """X (n_x, K), U (n_u, K), sigma -> A_bar (n_x^2, K-1), B_bar, C_bar (n_x n_u, K-1), S_bar, z_bar
        (n_x, K-1); column k is the matrix flattened order='F' (first_order_hold.py:52-87)."""
        self._stage_in(X, U, sigma)
        _device.foh(self.model_id, self._dX, self._dU, self._dS, self.n_sub, out=self._d_views_out)
        self._h_out.copy_(self._d_out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.A_bar, self.B_bar, self.C_bar, self.S_bar, self.z_bar

    def integrate_nonlinear_piecewise(self, X_lin, U, sigma):
        """first_order_hold.py:127-140."""
        self._stage_in(X_lin, U, sigma)
        out = _device.integrate_piecewise(self.model_id, self._dX, self._dU, self._dS, self.n_sub)
        return out[0].cpu().numpy()

    def integrate_nonlinear_full(self, x0, U, sigma):
        """first_order_hold.py:142-155."""
        X_dummy = np.zeros((self.n_x, self.K))
        X_dummy[:, 0] = np.asarray(x0, dtype=np.float64).reshape(-1)
        self._stage_in(X_dummy, U, sigma)
        x0d = self._dX[:, :, 0].contiguous()
        out = _device.integrate_full(self.model_id, x0d, self._dU, self._dS, self.n_sub)
        return out[0].cpu().numpy()

    def _dx(self, x, t, u0, u1, sigma):
        """Host twin of first_order_hold.py:157-162 (used by utils-style callers that probe the RHS)."""
        u = u0 + (t / (self.dt * sigma)) * (u1 - u0)
        return np.asarray(self.f(x, u), dtype=float).flatten()
