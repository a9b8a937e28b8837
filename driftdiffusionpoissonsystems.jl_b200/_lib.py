"""ctypes binding of libddps_b200.so (the C ABI declared in include/ddps.h).

The product path fails loudly when the CUDA library is missing or no GPU is present: there is no
CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DDPS_B200_LIB", os.path.join(_HERE, "libddps_b200.so"))

c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)
c_int32_p = C.POINTER(C.c_int32)
c_uint8_p = C.POINTER(C.c_uint8)

PRECOND_JACOBI, PRECOND_AMG = 0, 1


class SolverOpts(C.Structure):
    _fields_ = [("lin_rtol", C.c_double), ("newton_eta", C.c_double), ("lin_max_it", C.c_int64),
                ("max_newton", C.c_int32), ("max_gummel", C.c_int32), ("check_every", C.c_int32),
                ("warm_start", C.c_int32), ("verbose", C.c_int32), ("lin_cap_ok", C.c_int32), ("no_alt_traversal", C.c_int32), ("warm_start_v", C.c_int32),
                ("no_sym_scaling", C.c_int32), ("renumber", C.c_int32), ("precond", C.c_int32), ("amg_no_lag", C.c_int32),
                ("newton_damping", C.c_int32), ("reserved0", C.c_int32), ("lin_etol", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("outer_iters", C.c_int32), ("newton_total", C.c_int32), ("pcg_total", C.c_int64),
                ("spmv_total", C.c_int64), ("kernel_launches", C.c_int64), ("control", C.c_double),
                ("t_setup_ms", C.c_double), ("t_solve_ms", C.c_double), ("t_asm_ms", C.c_double),
                ("t_pcg_ms", C.c_double), ("t_wall_ms", C.c_double), ("status", C.c_int32), ("n_negative_area", C.c_int32),
                ("amg_levels", C.c_int32), ("amg_reuses", C.c_int32), ("t_amg_setup_ms", C.c_double),
                ("newton_halvings", C.c_int32), ("reserved", C.c_int32 * 3)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if not k.startswith("reserved")}


# every symbol include/ddps.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "ddps_version": (C.c_int, []),
    "ddps_last_error": (C.c_char_p, [C.c_void_p]),
    "ddps_ctx_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "ddps_ctx_destroy": (C.c_int, [C.c_void_p]),
    "ddps_default_opts": (None, [C.POINTER(SolverOpts)]),
    "ddps_set_opts": (C.c_int, [C.c_void_p, C.POINTER(SolverOpts)]),
    "ddps_comm_unique_id": (C.c_int, [C.c_char_p]),
    "ddps_comm_init": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_int]),
    "ddps_mesh_set": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, c_int64_p, C.c_int64]),
    "ddps_mesh_set_local": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, C.c_int64, c_int64_p, c_int32_p, c_int64_p, C.c_int64, c_int64_p,
                                      C.c_int64, C.c_int64]),
    "ddps_ddpe_fetch_owned": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p, C.c_int64]),
    "ddps_partition_nodes": (C.c_int, [c_double_p, C.c_int64, C.c_int, c_int32_p]),
    "ddps_partition_plan": (C.c_int, [c_double_p, C.c_int64, c_int64_p, C.c_int64, C.c_int, C.c_int, c_int64_p, c_int64_p,
                                      c_int32_p, c_int64_p, c_int64_p, c_int64_p]),
    "ddps_dirichlet_set": (C.c_int, [C.c_void_p, C.c_int, c_int64_p, c_double_p, C.c_int64]),
    "ddps_coeff_set": (C.c_int, [C.c_void_p, C.c_int, c_double_p, C.c_int64]),
    "ddps_functor_set": (C.c_int, [C.c_void_p, C.c_int, c_double_p, C.c_int]),
    "ddps_coeff_functor": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, C.c_int]),
    "ddps_ddpe_begin": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]),
    "ddps_ddpe_step": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, C.POINTER(C.c_int)]),
    "ddps_ddpe_fetch": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p]),
    "ddps_solve_ddpe": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                  c_double_p, c_double_p, c_double_p, C.POINTER(Stats)]),
    "ddps_solve_semlin": (C.c_int, [C.c_void_p, C.c_double, c_double_p, C.POINTER(Stats)]),
    "ddps_get_history": (C.c_int, [C.c_void_p, C.c_int, c_double_p, C.c_int64, c_int64_p]),
    "ddps_get_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "ddps_get_csr_size": (C.c_int, [C.c_void_p, c_int64_p, c_int64_p]),
    "ddps_get_csr": (C.c_int, [C.c_void_p, C.c_int, c_int64_p, c_int64_p, c_double_p]),
    "ddps_get_vector": (C.c_int, [C.c_void_p, C.c_int, c_double_p, C.c_int64]),
    "ddps_get_coeff": (C.c_int, [C.c_void_p, C.c_int, c_double_p, C.c_int64, c_int64_p]),
    "ddps_debug_assemble_base": (C.c_int, [C.c_void_p, C.c_int]),
    "ddps_debug_assemble_newton": (C.c_int, [C.c_void_p, C.c_int, C.c_double, c_double_p, c_double_p, c_double_p,
                                             c_double_p]),
    "ddps_debug_assemble_uv": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, c_double_p,
                                         c_double_p, c_double_p]),
    "ddps_debug_spmv": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p]),
    "ddps_debug_pcg": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_double, C.c_int64, c_int64_p, c_double_p]),
    "ddps_time_kernel": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p]),
    "ddps_amg_host_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64, c_int64_p, c_int64_p, c_double_p, c_uint8_p,
                                       C.c_int, C.c_int]),
    "ddps_amg_host_nlevels": (C.c_int, [C.c_void_p]),
    "ddps_amg_host_level_sizes": (C.c_int, [C.c_void_p, C.c_int, c_int64_p]),
    "ddps_amg_host_level_get": (C.c_int, [C.c_void_p, C.c_int, c_int64_p, c_int64_p, c_double_p, c_int64_p, c_int64_p,
                                          c_double_p, c_int64_p]),
    "ddps_amg_host_free": (None, [C.c_void_p]),
    "ddps_amg_levels": (C.c_int, [C.c_void_p, c_int64_p, c_int64_p]),
    "ddps_amg_numeric": (C.c_int, [C.c_void_p, C.c_int]),
    "ddps_amg_get_level": (C.c_int, [C.c_void_p, C.c_int, c_int64_p, c_int64_p, c_int64_p, c_int64_p, c_int64_p, c_double_p]),
    "ddps_amg_get_values": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p]),
    "ddps_amg_apply": (C.c_int, [C.c_void_p, c_double_p, c_double_p]),
}

_lib = None


def load():
    """Load libddps_b200.so and attach the prototypes.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the DDPS hot path)")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def dptr(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_double_p)


def iptr(a: np.ndarray):
    assert a.dtype == np.int64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_int64_p)
