"""ctypes mirrors of the C structs the reference declares in src/types.jl.

Layouts (sizeof / offsets) are asserted in tests/test_abi_layout.py against
SURVEY.md section 8(b).
"""
import ctypes as C

c_int = C.c_longlong  # src/types.jl:5-9 (Cc_int on 64-bit)
c_float = C.c_double
c_int_p = C.POINTER(c_int)
c_float_p = C.POINTER(c_float)


class Ccsc(C.Structure):  # src/types.jl:11-19
    _fields_ = [
        ("nzmax", c_int),
        ("m", c_int),
        ("n", c_int),
        ("p", c_int_p),
        ("i", c_int_p),
        ("x", c_float_p),
        ("nz", c_int),
    ]


class Solution(C.Structure):  # src/types.jl:74-77
    _fields_ = [("x", c_float_p), ("y", c_float_p)]


class CInfo(C.Structure):  # src/types.jl:81-99
    _fields_ = [
        ("iter", c_int),
        ("status", C.c_char * 32),
        ("status_val", c_int),
        ("status_polish", c_int),
        ("obj_val", c_float),
        ("pri_res", c_float),
        ("dua_res", c_float),
        ("setup_time", c_float),
        ("solve_time", c_float),
        ("update_time", c_float),
        ("polish_time", c_float),
        ("run_time", c_float),
        ("rho_updates", c_int),
        ("rho_estimate", c_float),
    ]


class Data(C.Structure):  # src/types.jl:101-109
    _fields_ = [
        ("n", c_int),
        ("m", c_int),
        ("P", C.POINTER(Ccsc)),
        ("A", C.POINTER(Ccsc)),
        ("q", c_float_p),
        ("l", c_float_p),
        ("u", c_float_p),
    ]


class Settings(C.Structure):  # src/types.jl:111-134
    _fields_ = [
        ("rho", c_float),
        ("sigma", c_float),
        ("scaling", c_int),
        ("adaptive_rho", c_int),
        ("adaptive_rho_interval", c_int),
        ("adaptive_rho_tolerance", c_float),
        ("adaptive_rho_fraction", c_float),
        ("max_iter", c_int),
        ("eps_abs", c_float),
        ("eps_rel", c_float),
        ("eps_prim_inf", c_float),
        ("eps_dual_inf", c_float),
        ("alpha", c_float),
        ("linsys_solver", C.c_int),  # enum
        ("delta", c_float),
        ("polish", c_int),
        ("polish_refine_iter", c_int),
        ("verbose", c_int),
        ("scaled_termination", c_int),
        ("check_termination", c_int),
        ("warm_start", c_int),
        ("time_limit", c_float),
    ]


class Workspace(C.Structure):  # src/types.jl:173-217
    _fields_ = [
        ("data", C.POINTER(Data)),
        ("linsys_solver", C.c_void_p),
        ("pol", C.c_void_p),
        ("rho_vec", c_float_p),
        ("rho_inv_vec", c_float_p),
        ("constr_type", c_int_p),
        ("x", c_float_p),
        ("y", c_float_p),
        ("z", c_float_p),
        ("xz_tilde", c_float_p),
        ("x_prev", c_float_p),
        ("z_prev", c_float_p),
        ("Ax", c_float_p),
        ("Px", c_float_p),
        ("Aty", c_float_p),
        ("delta_y", c_float_p),
        ("Atdelta_y", c_float_p),
        ("delta_x", c_float_p),
        ("Pdelta_x", c_float_p),
        ("Adelta_x", c_float_p),
        ("D_temp", c_float_p),
        ("D_temp_A", c_float_p),
        ("E_temp", c_float_p),
        ("settings", C.POINTER(Settings)),
        ("scaling", C.c_void_p),
        ("solution", C.POINTER(Solution)),
        ("info", C.POINTER(CInfo)),
        ("timer", C.c_void_p),
        ("first_run", c_int),
        ("summary_printed", c_int),
    ]


class Info:  # src/types.jl:219-236 (user-facing)
    __slots__ = (
        "iter",
        "status",
        "status_val",
        "status_polish",
        "obj_val",
        "pri_res",
        "dua_res",
        "setup_time",
        "solve_time",
        "update_time",
        "polish_time",
        "run_time",
        "rho_updates",
        "rho_estimate",
    )

    def __init__(self):
        for s in self.__slots__:
            setattr(self, s, None)

    def __repr__(self):
        return "Info(" + ", ".join(f"{s}={getattr(self, s)!r}" for s in self.__slots__) + ")"


class Results:  # src/types.jl:256-272
    def __init__(self):
        import numpy as np

        self.x = np.zeros(0)
        self.y = np.zeros(0)
        self.info = Info()
        self.prim_inf_cert = np.zeros(0)
        self.dual_inf_cert = np.zeros(0)

    def resize(self, n, m):
        import numpy as np

        if self.x.size != n:
            self.x = np.empty(n)
            self.dual_inf_cert = np.empty(n)
        if self.y.size != m:
            self.y = np.empty(m)
            self.prim_inf_cert = np.empty(m)
        return self


class B200Profile(C.Structure):
    """include/osqp_b200.h OSQPB200Profile -- engine extension, not part of the reference ABI."""

    _fields_ = [(k, c_int) for k in ("device", "grid", "block", "lanes_A", "lanes_N", "nnz_A", "nnz_P_full", "launches",
                                     "admm_iters", "pcg_iters", "info_evals", "refreshes")] + \
               [(k, c_float) for k in ("kernel_ms", "polish_ms", "alg_bytes", "spmv_bytes_A", "spmv_bytes_At",
                                       "spmv_bytes_P")] + [("phase_us", c_float * 16)] + \
               [(k, c_int) for k in ("streams", "groups_A", "groups_At", "paired", "fast_kernels")]
