"""Host side of the batched engine (include/osqp_b200.h ``osqp_batch_*``).

The reference has no batch API (SURVEY.md 8b "Batch extension"); this module is what a Julia helper next to the
unchanged ``src/*.jl`` would look like: the same marshalling conventions as ``interface.py`` (Int64 0-based CSC,
upper-triangular P, +-1e30 bound clamp, NaN-fill by status) applied to ``count`` QPs that share one sparsity
pattern.  A batch lives on one GPU; ``shard_range`` gives the contiguous block of QPs a rank owns when a large batch
is spread over several GPUs (one process per GPU, problem data scattered once at setup, no collective in the loop).
"""
import ctypes as C

import numpy as np

from . import types as T
from .constants import OSQP_INFTY, status_map
from .interface import ManagedCcsc, load_library

try:
    import scipy.sparse as sp
except Exception:  # pragma: no cover
    sp = None


class BatchInfo(C.Structure):
    """include/osqp_b200.h OSQPB200BatchInfo"""

    _fields_ = [("iter", T.c_int), ("status_val", T.c_int), ("obj_val", T.c_float), ("pri_res", T.c_float),
                ("dua_res", T.c_float), ("rho_estimate", T.c_float), ("rho_updates", T.c_int)]


def shard_range(count, world_size, rank):
    """Contiguous block [lo, hi) of a batch of `count` QPs owned by `rank` (blocks of ceil(B/G))."""
    per = -(-count // world_size)
    lo = min(count, rank * per)
    return lo, min(count, lo + per)


_INFO_DTYPE = np.dtype([("iter", "<i8"), ("status_val", "<i8"), ("obj_val", "<f8"), ("pri_res", "<f8"),
                        ("dua_res", "<f8"), ("rho_estimate", "<f8"), ("rho_updates", "<i8")])
assert _INFO_DTYPE.itemsize == C.sizeof(BatchInfo)


class BatchResults:
    def __init__(self, x, y, info):
        self.x, self.y = x, y
        a = np.frombuffer(info, dtype=_INFO_DTYPE).copy()
        self.iter, self.status_val = a["iter"], a["status_val"]
        self.obj_val, self.pri_res, self.dua_res = a["obj_val"], a["pri_res"], a["dua_res"]
        self.rho_estimate, self.rho_updates = a["rho_estimate"], a["rho_updates"]

    @property
    def status(self):
        return [status_map[int(v)] for v in self.status_val]


class BatchModel:
    """`count` QPs  min 1/2 x'P_k x + q_k'x  s.t.  l_k <= A_k x <= u_k  with one pattern for all P_k and one for all A_k."""

    def __init__(self, lib=None):
        self.lib = load_library(lib)
        L = self.lib
        self._h = C.c_void_p()
        fp = T.c_float_p
        L.osqp_batch_setup.restype = T.c_int
        L.osqp_batch_setup.argtypes = [C.POINTER(C.c_void_p), T.c_int, C.POINTER(T.Data), fp, fp, fp, fp, fp,
                                       C.POINTER(T.Settings)]
        L.osqp_batch_solve.restype = T.c_int
        L.osqp_batch_solve.argtypes = [C.c_void_p, fp, fp, C.POINTER(BatchInfo)]
        self._has_view = hasattr(L, "osqp_batch_solve_view")  # the CUDA engine; the CPU oracle has no batch API at all
        if self._has_view:
            L.osqp_batch_solve_view.restype = T.c_int
            L.osqp_batch_solve_view.argtypes = [C.c_void_p, C.POINTER(fp), C.POINTER(fp), C.POINTER(C.POINTER(BatchInfo))]
            L.osqp_batch_input_view.restype = T.c_int
            L.osqp_batch_input_view.argtypes = [C.c_void_p, C.POINTER(fp), C.POINTER(fp), C.POINTER(fp)]
            L.osqp_batch_device_solution.restype = T.c_int
            L.osqp_batch_device_solution.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.osqp_batch_update.restype = T.c_int
        L.osqp_batch_update.argtypes = [C.c_void_p, fp, fp, fp]
        L.osqp_batch_warm_start.restype = T.c_int
        L.osqp_batch_warm_start.argtypes = [C.c_void_p, fp, fp]
        L.osqp_batch_update_setting.restype = T.c_int
        L.osqp_batch_update_setting.argtypes = [C.c_void_p, C.c_char_p, T.c_float]
        L.osqp_batch_last_kernel_ms.restype = T.c_float
        L.osqp_batch_last_kernel_ms.argtypes = [C.c_void_p]
        L.osqp_batch_cleanup.restype = T.c_int
        L.osqp_batch_cleanup.argtypes = [C.c_void_p]
        self.count = self.n = self.m = 0

    @staticmethod
    def _f(a):
        return None if a is None else a.ctypes.data_as(T.c_float_p)

    def setup(self, P_pattern, A_pattern, Px, Ax, q, l, u, **settings):
        """P_pattern / A_pattern: scipy sparse matrices giving the shared patterns (P is reduced to its upper
        triangle, src/interface.jl:102-104); Px [count, nnz(triu P)], Ax [count, nnz(A)] follow the CSC order of
        those patterns; q [count, n]; l, u [count, m]."""
        Pt = sp.triu(sp.csc_matrix(P_pattern), format="csc")
        Pt.sort_indices()
        Ac = sp.csc_matrix(A_pattern)
        Ac.sort_indices()
        n, m = Ac.shape[1], Ac.shape[0]
        Px = np.ascontiguousarray(Px, dtype=np.float64).reshape(-1, Pt.nnz)
        count = Px.shape[0]
        Ax = np.ascontiguousarray(Ax, dtype=np.float64).reshape(count, Ac.nnz)
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(count, n)
        l = np.clip(np.ascontiguousarray(l, dtype=np.float64).reshape(count, m), -OSQP_INFTY, OSQP_INFTY)
        u = np.clip(np.ascontiguousarray(u, dtype=np.float64).reshape(count, m), -OSQP_INFTY, OSQP_INFTY)
        mP, mA = ManagedCcsc(Pt), ManagedCcsc(Ac)
        cP, cA = mP.ccsc(), mA.ccsc()
        data = T.Data(n, m, C.pointer(cP), C.pointer(cA), None, None, None)
        st = T.Settings()
        self.lib.osqp_set_default_settings(C.byref(st))
        for k, v in settings.items():
            if not hasattr(st, k):
                raise ValueError(f"unknown setting {k}")
            setattr(st, k, type(getattr(st, k))(v))
        rc = self.lib.osqp_batch_setup(C.byref(self._h), count, C.byref(data), self._f(Px), self._f(Ax), self._f(q),
                                       self._f(l), self._f(u), C.byref(st))
        if rc != 0:
            self._h = C.c_void_p()
            raise RuntimeError("Error in OSQP batch setup")
        self.count, self.n, self.m = count, n, m
        self.pattern_P, self.pattern_A = Pt, Ac
        return self

    def solve(self, copy=True):
        """Solve every QP of the batch.  copy=True (default) returns private arrays, like `solve!` of the reference
        copies `workspace.solution` out (src/interface.jl:179-191).  copy=False returns zero-copy VIEWS of the engine's
        pinned host mirrors: valid only until the next call on this batch (solve, update, clean) -- for callers that
        consume the result immediately and cannot afford another pass over it."""
        if not self._h:
            raise RuntimeError("You are trying to solve an empty batch. Please setup the batch before calling solve().")
        if self._has_view and not copy:
            px, py, pi = T.c_float_p(), T.c_float_p(), C.POINTER(BatchInfo)()
            rc = self.lib.osqp_batch_solve_view(self._h, C.byref(px), C.byref(py), C.byref(pi))
            if rc != 0:
                raise RuntimeError("Error in OSQP batch solve")
            x = np.ctypeslib.as_array(px, shape=(self.count, self.n))
            y = (np.ctypeslib.as_array(py, shape=(self.count, self.m)) if self.m else np.empty((self.count, 0)))
            info = C.cast(pi, C.POINTER(BatchInfo * self.count)).contents
            self.last_x = x
            return BatchResults(x, y, info)
        x = np.empty((self.count, self.n))
        y = np.empty((self.count, self.m))
        info = (BatchInfo * self.count)()
        rc = self.lib.osqp_batch_solve(self._h, self._f(x), self._f(y), info)
        if rc != 0:
            raise RuntimeError("Error in OSQP batch solve")
        self.last_x = x
        return BatchResults(x, y, info)  # rows without a solution are NaN-filled by the engine (certificates aside)

    def input_views(self):
        """(q [count, n], l [count, m], u [count, m]) numpy views of the engine's PINNED input staging.  Fill them in
        place and pass them to update(): the H2D copy then runs at PCIe speed and no host-side copy is made."""
        if getattr(self, "_in_views", None) is None:
            pq, pl, pu = T.c_float_p(), T.c_float_p(), T.c_float_p()
            if self.lib.osqp_batch_input_view(self._h, C.byref(pq), C.byref(pl), C.byref(pu)) != 0:
                raise RuntimeError("no input staging")
            self._in_views = (np.ctypeslib.as_array(pq, shape=(self.count, self.n)),
                              np.ctypeslib.as_array(pl, shape=(self.count, max(self.m, 1)))[:, : self.m],
                              np.ctypeslib.as_array(pu, shape=(self.count, max(self.m, 1)))[:, : self.m])
        return self._in_views

    def update(self, q=None, l=None, u=None):
        a = []
        views = getattr(self, "_in_views", None) or ()
        for v, w in ((q, self.n), (l, self.m), (u, self.m)):
            if v is None:
                a.append(None)
                continue
            if any(v is t for t in views):  # already in pinned staging; the caller keeps |l|, |u| <= 1e30 (:107-108)
                a.append(v)
                continue
            v = np.ascontiguousarray(v, dtype=np.float64).reshape(self.count, w)
            a.append(v if len(a) == 0 else np.clip(v, -OSQP_INFTY, OSQP_INFTY))  # q is passed as is; l, u are clamped
        if self.lib.osqp_batch_update(self._h, self._f(a[0]), self._f(a[1]), self._f(a[2])) != 0:
            raise RuntimeError("Error updating the batch")

    def warm_start(self, x=None, y=None):
        x = None if x is None else np.ascontiguousarray(x, dtype=np.float64).reshape(self.count, self.n)
        y = None if y is None else np.ascontiguousarray(y, dtype=np.float64).reshape(self.count, self.m)
        if self.lib.osqp_batch_warm_start(self._h, self._f(x), self._f(y)) != 0:
            raise RuntimeError("Error in batch warm start")

    def update_settings(self, **kw):
        for k, v in kw.items():
            if self.lib.osqp_batch_update_setting(self._h, k.encode(), float(v)) != 0:
                raise ValueError(f"cannot update setting {k}")

    @property
    def kernel_ms(self):
        return float(self.lib.osqp_batch_last_kernel_ms(self._h))

    def clean(self):
        self._in_views = None
        if self._h:
            self.lib.osqp_batch_cleanup(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.clean()
        except Exception:
            pass


class _DeviceArray:
    """A raw device pointer as seen by torch.as_tensor (CUDA array interface v3)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False), "version": 3}


def gather_sharded_device(model, count, world_size, rank):
    """All-gather of the last solve's x* straight from the engine's device buffer (osqp_batch_device_solution) into a
    preallocated [count, n] device tensor on every rank: one NCCL all-gather over NVLink, no host round trip.  Shards
    are the contiguous blocks of `shard_range`; a short last shard is padded on the device."""
    import torch
    import torch.distributed as dist

    dx, dy = C.c_void_p(), C.c_void_p()
    if model.lib.osqp_batch_device_solution(model._h, C.byref(dx), C.byref(dy)) != 0:
        raise RuntimeError("no device solution")
    per = -(-count // world_size)
    n = model.n
    local = torch.as_tensor(_DeviceArray(dx.value, (model.count, n)), device="cuda")
    cache = getattr(model, "_gather_cache", None)
    if cache is None or cache[0].shape != (world_size * per, n):
        cache = (torch.empty((world_size * per, n), dtype=torch.float64, device="cuda"),
                 torch.zeros((per, n), dtype=torch.float64, device="cuda"))
        model._gather_cache = cache
    out, pad = cache
    if model.count != per:
        pad[: model.count].copy_(local)
        local = pad
    if world_size > 1:
        dist.all_gather_into_tensor(out, local)
    else:
        out.copy_(local)
    return out[:count]


def gather_sharded(local, count, world_size, rank, device="cpu"):
    """All-gather per-QP rows computed on contiguous shards (shard_range) into the full [count, ...] array on every
    rank.  The only collective of the batched path, and it runs AFTER the solves; backend = whatever
    torch.distributed was initialised with (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist

    if isinstance(local, BatchModel):
        local = local.last_x
    local = np.asarray(local)
    per = -(-count // world_size)
    tail = local.shape[1:]
    buf = torch.zeros((per,) + tail, dtype=torch.float64, device=device)
    if local.shape[0]:
        buf[: local.shape[0]] = torch.as_tensor(local, dtype=torch.float64).to(device)
    out = torch.empty((world_size * per,) + tail, dtype=torch.float64, device=device)
    if world_size > 1:
        dist.all_gather_into_tensor(out, buf)
    else:
        out.copy_(buf)
    return out[:count].cpu().numpy()
