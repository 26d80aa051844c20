"""Host-side mirror of the reference's native API (src/interface.jl).

The reference's host layer is Julia (``OSQP.Model``, ``setup!``/``solve!``/
``update!``/``warm_start!``/``update_settings!``); Julia is not present in
this image, so this module replays the same marshalling byte-for-byte over
ctypes: Int64 0-based CSC, upper-triangular P, +-1e30 bound clamp, struct
reads at the offsets of src/types.jl, exit-code -> exception, NULL-safe
cleanup.  Names, argument meaning and error behaviour follow the Julia
functions cited next to each method (index vectors are 0-based here, the
Python convention, where Julia's are 1-based and shifted at src/interface.jl:315-328).

The shared library is the CUDA engine ``osqp.jl_b200/lib/libosqp.so``; it has
no CPU fallback -- if it cannot be loaded or no GPU is present, calls raise.
``OSQP_B200_LIB`` (env) or ``Model(lib=...)`` re-points the handle the way a
JLL override re-points ``OSQP.osqp`` (src/OSQP.jl:7); the tests use that to
drive the CPU oracle through the very same marshalling code.
"""
import ctypes as C
import os
import threading

import numpy as np

from . import types as T
from .constants import (
    MKL_PARDISO_SOLVER,
    OSQP_INFTY,
    QDLDL_SOLVER,
    SOLUTION_PRESENT,
    UPDATABLE_SETTINGS,
    status_map,
)

try:  # scipy is only needed to accept sparse inputs
    import scipy.sparse as sp
except Exception:  # pragma: no cover
    sp = None

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "lib", "libosqp.so")

_libs = {}
_libs_lock = threading.Lock()

_W = C.POINTER(T.Workspace)

# (symbol, restype, argtypes) -- the 30 symbols of SURVEY.md section 8(b)
_SIGNATURES = [
    ("osqp_set_default_settings", None, [C.POINTER(T.Settings)]),
    ("osqp_setup", T.c_int, [C.POINTER(_W), C.POINTER(T.Data), C.POINTER(T.Settings)]),
    ("osqp_solve", T.c_int, [_W]),
    ("osqp_version", C.c_char_p, []),
    ("osqp_cleanup", T.c_int, [_W]),
    ("osqp_update_lin_cost", T.c_int, [_W, T.c_float_p]),
    ("osqp_update_lower_bound", T.c_int, [_W, T.c_float_p]),
    ("osqp_update_upper_bound", T.c_int, [_W, T.c_float_p]),
    ("osqp_update_bounds", T.c_int, [_W, T.c_float_p, T.c_float_p]),
    ("osqp_update_P", T.c_int, [_W, T.c_float_p, T.c_int_p, T.c_int]),
    ("osqp_update_A", T.c_int, [_W, T.c_float_p, T.c_int_p, T.c_int]),
    ("osqp_update_P_A", T.c_int, [_W, T.c_float_p, T.c_int_p, T.c_int, T.c_float_p, T.c_int_p, T.c_int]),
    ("osqp_warm_start", T.c_int, [_W, T.c_float_p, T.c_float_p]),
    ("osqp_warm_start_x", T.c_int, [_W, T.c_float_p]),
    ("osqp_warm_start_y", T.c_int, [_W, T.c_float_p]),
    ("osqp_update_max_iter", T.c_int, [_W, T.c_int]),
    ("osqp_update_eps_abs", T.c_int, [_W, T.c_float]),
    ("osqp_update_eps_rel", T.c_int, [_W, T.c_float]),
    ("osqp_update_eps_prim_inf", T.c_int, [_W, T.c_float]),
    ("osqp_update_eps_dual_inf", T.c_int, [_W, T.c_float]),
    ("osqp_update_rho", T.c_int, [_W, T.c_float]),
    ("osqp_update_alpha", T.c_int, [_W, T.c_float]),
    ("osqp_update_delta", T.c_int, [_W, T.c_float]),
    ("osqp_update_polish", T.c_int, [_W, T.c_int]),
    ("osqp_update_polish_refine_iter", T.c_int, [_W, T.c_int]),
    ("osqp_update_verbose", T.c_int, [_W, T.c_int]),
    ("osqp_update_scaled_termination", T.c_int, [_W, T.c_int]),
    ("osqp_update_check_termination", T.c_int, [_W, T.c_int]),
    ("osqp_update_warm_start", T.c_int, [_W, T.c_int]),
    ("osqp_update_time_limit", T.c_int, [_W, T.c_float]),
]
ABI_SYMBOLS = tuple(s[0] for s in _SIGNATURES)


def load_library(path=None):
    """dlopen the engine (or any library exporting the same ABI) and type its symbols."""
    path = os.path.abspath(path or os.environ.get("OSQP_B200_LIB") or DEFAULT_LIB)
    with _libs_lock:
        lib = _libs.get(path)
        if lib is None:
            if not os.path.exists(path):
                raise OSError(
                    f"{path} not found: build the CUDA engine first "
                    "(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback"
                )
            lib = C.CDLL(path)
            for name, res, args in _SIGNATURES:
                fn = getattr(lib, name)  # AttributeError if the ABI is incomplete
                fn.restype = res
                fn.argtypes = args
            _libs[path] = lib
    return lib


def _fptr(a):
    return a.ctypes.data_as(T.c_float_p)


def _iptr(a):
    return a.ctypes.data_as(T.c_int_p)


class ManagedCcsc:
    """src/types.jl:21-47: Float64 values + 0-based Int64 indices kept alive by Python."""

    def __init__(self, M):
        M = sp.csc_matrix(M)
        M.sort_indices()
        self.m, self.n = M.shape
        self.x = np.ascontiguousarray(M.data, dtype=np.float64)
        self.i = np.ascontiguousarray(M.indices, dtype=np.int64)
        self.p = np.ascontiguousarray(M.indptr, dtype=np.int64)
        self.nzmax = int(self.x.size)
        self.nz = -1

    def ccsc(self):  # src/types.jl:62-72
        return T.Ccsc(self.nzmax, self.m, self.n, _iptr(self.p), _iptr(self.i), _fptr(self.x), self.nz)


def ccsc_to_scipy(c):
    """src/types.jl:49-57 (Base.convert(SparseMatrixCSC, ::Ccsc))."""
    nzmax, n = c.nzmax, c.n
    x = np.array([c.x[k] for k in range(nzmax)], dtype=np.float64)
    i = np.array([c.i[k] for k in range(nzmax)], dtype=np.int64)
    p = np.array([c.p[k] for k in range(n + 1)], dtype=np.int64)
    return sp.csc_matrix((x, i, p), shape=(c.m, c.n))


def linsys_solver_str_to_int(settings):
    """src/interface.jl:749-773."""
    v = settings.get("linsys_solver", None)
    if v is None:
        return
    if isinstance(v, str):
        s = v.lower()
        if s == "qdldl":
            settings["linsys_solver"] = QDLDL_SOLVER
        elif s == "mkl pardiso":
            settings["linsys_solver"] = MKL_PARDISO_SOLVER
        elif s == "":
            settings["linsys_solver"] = QDLDL_SOLVER
        else:
            import warnings

            warnings.warn("Linear system solver not recognized. Using default solver QDLDL.")
            settings["linsys_solver"] = QDLDL_SOLVER
    else:
        import warnings

        warnings.warn("linsys_solver is required to be a string. Using default solver QDLDL.")
        settings["linsys_solver"] = QDLDL_SOLVER


def make_settings(lib, settings):
    """src/types.jl:136-171: defaults come from the C side, then overrides by field name."""
    s = T.Settings()
    lib.osqp_set_default_settings(C.byref(s))
    settings = dict(settings)
    linsys_solver_str_to_int(settings)
    names = {f[0]: f[1] for f in T.Settings._fields_}
    for k, v in settings.items():
        if k not in names:
            raise TypeError(f"type Settings has no field {k}")
        if names[k] is T.c_float:
            setattr(s, k, float(v))
        else:
            setattr(s, k, int(v))
    return s


class Model:
    """src/interface.jl:18-28."""

    def __init__(self, lib=None):
        self._lib = load_library(lib)
        self.workspace = _W()  # C_NULL
        self.lcache = np.zeros(0)
        self.ucache = np.zeros(0)
        self.isempty = True

    def __del__(self):  # finalizer(OSQP.clean!, model), src/interface.jl:25
        try:
            self.clean()
        except Exception:
            pass

    # ------------------------------------------------------------------ setup!
    def setup(self, P=None, q=None, A=None, l=None, u=None, **settings):
        """src/interface.jl:35-162."""
        if P is None:
            if q is not None:
                n = len(q)
            elif A is not None:
                n = A.shape[1]
            else:
                raise RuntimeError("The problem does not have any variables!")
        else:
            n = P.shape[0]
        m = 0 if A is None else A.shape[0]
        if (A is None and (l is not None or u is not None)) or (A is not None and l is None and u is None):
            raise RuntimeError("A must be supplied together with l and u")
        if A is not None and l is None:
            l = -np.inf * np.ones(m)
        if A is not None and u is None:
            u = np.inf * np.ones(m)
        if P is None:
            P = sp.csc_matrix((n, n))
        if q is None:
            q = np.zeros(n)
        if A is None:
            A = sp.csc_matrix((m, n))
            l = np.zeros(m)
            u = np.zeros(m)
        q = np.ascontiguousarray(q, dtype=np.float64)
        l = np.asarray(l, dtype=np.float64)
        u = np.asarray(u, dtype=np.float64)
        if q.size != n:
            raise RuntimeError("Incorrect dimension of q")
        if l.size != m:
            raise RuntimeError("Incorrect dimensions of l")
        if u.size != m:
            raise RuntimeError("Incorrect dimensions of u")
        P = sp.triu(sp.csc_matrix(P), format="csc")  # src/interface.jl:102-104
        u = np.ascontiguousarray(np.minimum(u, OSQP_INFTY))  # :107
        l = np.ascontiguousarray(np.maximum(l, -OSQP_INFTY))  # :108
        self.lcache = np.empty(m)
        self.ucache = np.empty(m)
        managedP = ManagedCcsc(P)
        managedA = ManagedCcsc(sp.csc_matrix(A))
        Pdata = managedP.ccsc()
        Adata = managedA.ccsc()
        stgs = make_settings(self._lib, settings)
        data = T.Data(n, m, C.pointer(Pdata), C.pointer(Adata), _fptr(q), _fptr(l), _fptr(u))
        workspace = _W()
        exitflag = self._lib.osqp_setup(C.byref(workspace), C.byref(data), C.byref(stgs))
        self.workspace = workspace
        if exitflag != 0:
            self.workspace = _W()
            raise RuntimeError("Error in OSQP setup")
        self.isempty = False
        return self

    # ------------------------------------------------------------------ solve!
    def solve(self, results=None):
        """src/interface.jl:164-217."""
        if results is None:
            results = T.Results()
        if self.isempty:
            raise RuntimeError(
                "You are trying to solve an empty model. Please setup the model before calling solve!()."
            )
        rc = self._lib.osqp_solve(self.workspace)  # the reference ignores the return code, :170-175
        if rc is not None and int(rc) >= 100:
            # engine-only failure class (100 + cudaError: launch refused, sticky device error).  libosqp has no such
            # exit; the workspace already reads Unsolved / NaN, and this host mirror additionally refuses to go on.
            raise RuntimeError(f"osqp_solve: CUDA failure in the engine (exit code {int(rc)})")
        workspace = self.workspace.contents
        cinfo = workspace.info.contents
        info = results.info
        info.iter = int(cinfo.iter)
        info.status = status_map[int(cinfo.status_val)]  # KeyError on an unknown code, src/types.jl:240
        info.status_val = int(cinfo.status_val)
        info.status_polish = int(cinfo.status_polish)
        info.obj_val = float(cinfo.obj_val)
        info.pri_res = float(cinfo.pri_res)
        info.dua_res = float(cinfo.dua_res)
        info.setup_time = float(cinfo.setup_time)
        info.solve_time = float(cinfo.solve_time)
        info.update_time = float(cinfo.update_time)
        info.polish_time = float(cinfo.polish_time)
        info.run_time = float(cinfo.run_time)
        info.rho_updates = int(cinfo.rho_updates)
        info.rho_estimate = float(cinfo.rho_estimate)
        solution = workspace.solution.contents
        data = workspace.data.contents
        n, m = int(data.n), int(data.m)
        results.resize(n, m)
        if info.status in SOLUTION_PRESENT:
            C.memmove(results.x.ctypes.data, solution.x, 8 * n)
            if m:
                C.memmove(results.y.ctypes.data, solution.y, 8 * m)
            results.prim_inf_cert.fill(np.nan)
            results.dual_inf_cert.fill(np.nan)
        else:
            results.x.fill(np.nan)
            results.y.fill(np.nan)
            if info.status in ("Primal_infeasible", "Primal_infeasible_inaccurate"):
                C.memmove(results.prim_inf_cert.ctypes.data, workspace.delta_y, 8 * m)
                results.dual_inf_cert.fill(np.nan)
            elif info.status in ("Dual_infeasible", "Dual_infeasible_inaccurate"):
                results.prim_inf_cert.fill(np.nan)
                C.memmove(results.dual_inf_cert.ctypes.data, workspace.delta_x, 8 * n)
            else:
                results.prim_inf_cert.fill(np.nan)
                results.dual_inf_cert.fill(np.nan)
        if info.status == "Non_convex":
            info.obj_val = float("nan")
        return results

    def version(self):
        """src/interface.jl:219-221."""
        return self._lib.osqp_version().decode()

    def clean(self):
        """src/interface.jl:223-233 (also the finalizer; NULL workspace is legal)."""
        lib = getattr(self, "_lib", None)
        if lib is None:
            return
        exitflag = lib.osqp_cleanup(self.workspace)
        self.workspace = _W()
        self.isempty = True
        if exitflag != 0:
            raise RuntimeError("Error in OSQP cleanup")

    def dimensions(self):
        """src/interface.jl:740-747."""
        if not self.workspace:
            raise RuntimeError("Workspace has not been setup yet")
        data = self.workspace.contents.data.contents
        return int(data.n), int(data.m)

    # ------------------------------------------------------------------ update!
    def update_q(self, q):
        """src/interface.jl:235-250."""
        n, m = self.dimensions()
        q = np.ascontiguousarray(q, dtype=np.float64)
        if q.size != n:
            raise RuntimeError(f"q must have length n = {n}")
        if self._lib.osqp_update_lin_cost(self.workspace, _fptr(q)) != 0:
            raise RuntimeError("Error updating q")

    def update_l(self, l):
        """src/interface.jl:252-268."""
        n, m = self.dimensions()
        l = np.asarray(l, dtype=np.float64)
        if l.size != m:
            raise RuntimeError(f"l must have length m = {m}")
        self.lcache[:] = np.maximum(l, -OSQP_INFTY)
        if self._lib.osqp_update_lower_bound(self.workspace, _fptr(self.lcache)) != 0:
            raise RuntimeError("Error updating l")

    def update_u(self, u):
        """src/interface.jl:270-286."""
        n, m = self.dimensions()
        u = np.asarray(u, dtype=np.float64)
        if u.size != m:
            raise RuntimeError(f"u must have length m = {m}")
        self.ucache[:] = np.minimum(u, OSQP_INFTY)
        if self._lib.osqp_update_upper_bound(self.workspace, _fptr(self.ucache)) != 0:
            raise RuntimeError("Error updating u")

    def update_bounds(self, l, u):
        """src/interface.jl:288-313."""
        n, m = self.dimensions()
        l = np.asarray(l, dtype=np.float64)
        u = np.asarray(u, dtype=np.float64)
        if l.size != m:
            raise RuntimeError(f"l must have length m = {m}")
        if u.size != m:
            raise RuntimeError(f"u must have length m = {m}")
        self.lcache[:] = np.maximum(l, -OSQP_INFTY)
        self.ucache[:] = np.minimum(u, OSQP_INFTY)
        if self._lib.osqp_update_bounds(self.workspace, _fptr(self.lcache), _fptr(self.ucache)) != 0:
            raise RuntimeError("Error updating bounds l and u")

    @staticmethod
    def _prep_idx(idx, n, name):
        """src/interface.jl:315-322 (None -> C_NULL = all entries in storage order)."""
        if idx is None:
            return None, None
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        if idx.size != n:
            raise RuntimeError(f"{name} and {name}_idx must have the same length")
        return idx, _iptr(idx)

    def update_P(self, Px, Px_idx=None):
        """src/interface.jl:330-349."""
        Px = np.ascontiguousarray(Px, dtype=np.float64)
        keep, p = self._prep_idx(Px_idx, Px.size, "P")
        if self._lib.osqp_update_P(self.workspace, _fptr(Px), p, Px.size) != 0:
            raise RuntimeError("Error updating P")

    def update_A(self, Ax, Ax_idx=None):
        """src/interface.jl:351-370."""
        Ax = np.ascontiguousarray(Ax, dtype=np.float64)
        keep, p = self._prep_idx(Ax_idx, Ax.size, "A")
        if self._lib.osqp_update_A(self.workspace, _fptr(Ax), p, Ax.size) != 0:
            raise RuntimeError("Error updating A")

    def update_P_A(self, Px, Px_idx, Ax, Ax_idx):
        """src/interface.jl:372-406."""
        Px = np.ascontiguousarray(Px, dtype=np.float64)
        Ax = np.ascontiguousarray(Ax, dtype=np.float64)
        keepP, pp = self._prep_idx(Px_idx, Px.size, "P")
        keepA, pa = self._prep_idx(Ax_idx, Ax.size, "A")
        if self._lib.osqp_update_P_A(self.workspace, _fptr(Px), pp, Px.size, _fptr(Ax), pa, Ax.size) != 0:
            raise RuntimeError("Error updating P and A")

    def update(self, q=None, l=None, u=None, Px=None, Px_idx=None, Ax=None, Ax_idx=None):
        """src/interface.jl:408-440."""
        if q is not None:
            self.update_q(q)
        if l is not None and u is not None:
            self.update_bounds(l, u)
        elif l is not None:
            self.update_l(l)
        elif u is not None:
            self.update_u(u)
        if Px is not None and Ax is not None:
            self.update_P_A(Px, Px_idx, Ax, Ax_idx)
        elif Px is not None:
            self.update_P(Px, Px_idx)
        elif Ax is not None:
            self.update_A(Ax, Ax_idx)

    # ------------------------------------------------------------------ update_settings!
    _INT_SETTINGS = ("max_iter", "polish", "polish_refine_iter", "verbose", "scaled_termination",
                     "check_termination", "warm_start")

    def update_settings(self, **kwargs):
        """src/interface.jl:442-670."""
        if not kwargs:
            return
        data = {}
        for key, value in kwargs.items():
            if key not in UPDATABLE_SETTINGS:
                raise RuntimeError(f"{key} cannot be updated or is not recognized")
            data[key] = value
        # NB the reference looks up :early_terminate for scaled_termination (src/interface.jl:468),
        # which UPDATABLE_SETTINGS rejects above -- unreachable there, unreachable here.
        order = ("max_iter", "eps_abs", "eps_rel", "eps_prim_inf", "eps_dual_inf", "rho", "alpha", "delta",
                 "polish", "polish_refine_iter", "verbose", "check_termination", "warm_start", "time_limit")
        for name in order:
            if name not in data:
                continue
            fn = getattr(self._lib, "osqp_update_" + name)
            v = int(data[name]) if name in self._INT_SETTINGS else float(data[name])
            if fn(self.workspace, v) != 0:
                raise RuntimeError(f"Error updating {name}")

    # ------------------------------------------------------------------ warm_start!
    def warm_start_x(self, x):
        """src/interface.jl:672-684."""
        n, m = self.dimensions()
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.size != n:
            raise RuntimeError("Wrong dimension for variable x")
        if self._lib.osqp_warm_start_x(self.workspace, _fptr(x)) != 0:
            raise RuntimeError("Error in warm starting x")

    def warm_start_y(self, y):
        """src/interface.jl:686-698."""
        n, m = self.dimensions()
        y = np.ascontiguousarray(y, dtype=np.float64)
        if y.size != m:
            raise RuntimeError("Wrong dimension for variable y")
        if self._lib.osqp_warm_start_y(self.workspace, _fptr(y)) != 0:
            raise RuntimeError("Error in warm starting y")

    def warm_start_x_y(self, x, y):
        """src/interface.jl:700-718."""
        n, m = self.dimensions()
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        if x.size != n:
            raise RuntimeError("Wrong dimension for variable x")
        if y.size != m:
            raise RuntimeError("Wrong dimension for variable y")
        if self._lib.osqp_warm_start(self.workspace, _fptr(x), _fptr(y)) != 0:
            raise RuntimeError("Error in warm starting x and y")

    def warm_start(self, x=None, y=None):
        """src/interface.jl:720-732."""
        if x is None and y is None:
            return
        elif x is not None and y is None:
            self.warm_start_x(x)
        elif x is None and y is not None:
            self.warm_start_y(y)
        else:
            self.warm_start_x_y(x, y)
