"""User dynamics on the device: sympy expressions -> CUDA C++ model struct -> NVRTC (row a5, the BaseModel plug-in contract).

The reference accepts any model whose `get_equations()` returns f(x,u), A = df/dx, B = df/du; its own models derive them with
sympy and `lambdify` (SCvx/models/unicycle_model.py:54-63).  Python callables cannot run inside a kernel, so a model that
wants the GPU path hands over the SYMBOLIC right-hand side instead (`BaseModel.symbolic_dynamics()`), and this module

  1. differentiates it (A, B by `jacobian`, exactly what the reference's models do),
  2. prints f, A, B as one common-subexpression-eliminated block of CUDA C++ inside a `struct UserModel` with the interface of
     the shipped device models (csrc/common.cuh: NX, NU, D, eval, f_only),
  3. appends the text of csrc/foh_kernels.cuh -- the very templates the shipped models are compiled from -- and three
     extern "C" kernels instantiating them, and
  4. passes the program to `scvx_user_model_register` (csrc/user_model.cu), which compiles it for sm_100a with NVRTC and returns
     a model id that `scvx_foh_batched`, `scvx_integrate_*_batched` and -- through the model's dimensions -- stages 2-3 accept.

There is no fallback: without libscvx_b200.so, NVRTC or a GPU the registration raises.
"""
import ctypes
import hashlib
import os

HERE = os.path.dirname(os.path.abspath(__file__))
_registered = {}          # sha1 of the program -> model id


def model_struct_source(x_syms, u_syms, f_expr, position_dim):
    """CUDA C++ text of `struct UserModel` for xdot = f_expr(x_syms, u_syms)."""
    import sympy as sp
    from sympy.printing.c import C99CodePrinter

    x_syms, u_syms = list(x_syms), list(u_syms)
    if not (1 <= len(x_syms) <= 8 and 1 <= len(u_syms) <= 8 and 1 <= int(position_dim) <= min(3, len(x_syms))):
        raise ValueError("user dynamics need 1..8 states, 1..8 inputs and a position of 1..3 leading states")
    f = sp.Matrix(f_expr).reshape(len(x_syms), 1)
    nx, nu = len(x_syms), len(u_syms)
    if f.shape[0] != nx:
        raise ValueError(f"f has {f.shape[0]} rows for {nx} states")
    A, B = f.jacobian(sp.Matrix(x_syms)), f.jacobian(sp.Matrix(u_syms))
    subs = {s: sp.Symbol(f"x[{i}]") for i, s in enumerate(x_syms)}
    subs.update({s: sp.Symbol(f"u[{j}]") for j, s in enumerate(u_syms)})
    printer = C99CodePrinter({"contract": False})

    def block(outputs):
        """outputs: list of (lhs text, sympy expr) -> C statements with shared sub-expressions hoisted."""
        exprs = [e.xreplace(subs) for _, e in outputs]
        temps, reduced = sp.cse(exprs, symbols=sp.numbered_symbols("t_"), optimizations="basic")
        lines = [f"    const double {printer.doprint(t)} = {printer.doprint(e)};" for t, e in temps]
        lines += [f"    {lhs} = {printer.doprint(e)};" for (lhs, _), e in zip(outputs, reduced)]
        return "\n".join(lines)

    full = [(f"f[{i}]", f[i]) for i in range(nx)]
    full += [(f"A[{i}][{j}]", A[i, j]) for i in range(nx) for j in range(nx)]
    full += [(f"B[{i}][{j}]", B[i, j]) for i in range(nx) for j in range(nu)]
    return f"""namespace scvx {{
struct UserModel {{
  static constexpr int NX = {nx}, NU = {nu}, D = {int(position_dim)};
  __device__ __forceinline__ static void eval(const double* x, const double* u, double* f, double (*A)[{nx}], double (*B)[{nu}]) {{
    (void)x; (void)u;
{block(full)}
  }}
  __device__ __forceinline__ static void f_only(const double* x, const double* u, double* f) {{
    (void)x; (void)u;
{block(full[:nx])}
  }}
}};
}}  // namespace scvx
""", nx, nu


_KERNELS = r"""
extern "C" __global__ void __launch_bounds__(128)
scvx_user_foh(int n_agents, int K, int n_sub, const double* __restrict__ X, const double* __restrict__ U,
              const double* __restrict__ sigma_arr, double* __restrict__ A_bar, double* __restrict__ B_bar,
              double* __restrict__ C_bar, double* __restrict__ S_bar, double* __restrict__ z_bar) {
  scvx::foh_rk4_body<scvx::UserModel>(n_agents, K, n_sub, X, U, sigma_arr, A_bar, B_bar, C_bar, S_bar, z_bar);
}
extern "C" __global__ void __launch_bounds__(128)
scvx_user_piecewise(int n_agents, int K, int n_sub, const double* __restrict__ X_lin, const double* __restrict__ U,
                    const double* __restrict__ sigma_arr, double* __restrict__ X_nl) {
  scvx::integrate_piecewise_body<scvx::UserModel>(n_agents, K, n_sub, X_lin, U, sigma_arr, X_nl);
}
extern "C" __global__ void __launch_bounds__(128)
scvx_user_full(int n_agents, int K, int n_sub, const double* __restrict__ x0, const double* __restrict__ U,
               const double* __restrict__ sigma_arr, double* __restrict__ X_nl) {
  scvx::integrate_full_body<scvx::UserModel>(n_agents, K, n_sub, x0, U, sigma_arr, X_nl);
}
"""


def program_source(model_src):
    """The whole NVRTC translation unit: stage-1 templates + the user's model + the three kernels."""
    kernels = open(os.path.join(HERE, "csrc", "foh_kernels.cuh")).read().replace("#pragma once", "")
    return "typedef unsigned long long size_t_;\n#define size_t size_t_\n" + kernels + "\n" + model_src + _KERNELS


def _nvrtc_hint():
    """torch ships its own libnvrtc; offer it to the library in case the CUDA toolkit's is not on the loader path."""
    try:
        import nvidia.cuda_nvrtc as pkg                      # noqa: PLC0415
        cand = os.path.join(os.path.dirname(pkg.__file__), "lib", "libnvrtc.so.12")
        if os.path.exists(cand):
            return cand
    except Exception:                                        # noqa: BLE001
        pass
    for cand in ("/usr/local/cuda/lib64/libnvrtc.so.12",):
        if os.path.exists(cand):
            return cand
    return ""


def register(x_syms, u_syms, f_expr, position_dim=2):
    """Compile (once per distinct program) and return the device model id of xdot = f_expr."""
    from . import _lib
    model_src, nx, nu = model_struct_source(x_syms, u_syms, f_expr, position_dim)
    prog = program_source(model_src)
    key = hashlib.sha1(prog.encode()).hexdigest()
    if key in _registered:
        return _registered[key]
    lib = _lib.load()
    mid = ctypes.c_int(-1)
    rc = lib.scvx_user_model_register(prog.encode(), nx, nu, int(position_dim), _nvrtc_hint().encode(), ctypes.byref(mid))
    if rc != 0:
        log = lib.scvx_user_model_log().decode("utf-8", "replace")
        msg = lib.scvx_last_error().decode("utf-8", "replace")
        raise _lib.ScvxError(f"scvx_user_model_register failed with code {rc}: {msg}\n{log}")
    _registered[key] = mid.value
    return mid.value
