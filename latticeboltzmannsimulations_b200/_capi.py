"""ctypes binding of the C ABI declared in include/lbm_b200.h (liblbm_b200.so).

The product path has no CPU fallback: if the shared library is missing or no CUDA device is usable the calls
raise (``LBMError``) -- they never route to NumPy or to anything under ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "liblbm_b200.so")

LBM_OK, LBM_EINVAL, LBM_ECUDA, LBM_ENOMEM, LBM_ESTATE = 0, 1, 2, 3, 4
LBM_F32, LBM_F64 = 0, 1
LBM_SRT, LBM_TRT, LBM_MRT = 0, 1, 2
LBM_REGION_ALL, LBM_REGION_EDGE, LBM_REGION_INTERIOR = 0, 1, 2
LBM_ENGINE_AUTO, LBM_ENGINE_LDG, LBM_ENGINE_TMA, LBM_ENGINE_AA = 0, 1, 2, 3

COLLISIONS = {"SRT": LBM_SRT, "TRT": LBM_TRT, "MRT": LBM_MRT}
ENGINES = {"auto": LBM_ENGINE_AUTO, "ldg": LBM_ENGINE_LDG, "tma": LBM_ENGINE_TMA, "aa": LBM_ENGINE_AA}
SEMANTICS = {"C": 0, "A": 1}

# every symbol include/lbm_b200.h declares (tests/test_capi_symbols.py checks header <-> library <-> this list)
SYMBOLS = [
    "lbm_last_error", "lbm_abi_version", "lbm_device_count", "lbm_state_bytes", "lbm_create", "lbm_destroy",
    "lbm_get_layout", "lbm_set_tuning", "lbm_set_reynolds", "lbm_set_rates", "lbm_init_equilibrium", "lbm_upload_f",
    "lbm_download_f", "lbm_step", "lbm_step_region", "lbm_swap", "lbm_step2_region", "lbm_swap2", "lbm_step2_available", "lbm_buffer_ptr", "lbm_halo_pack", "lbm_halo_unpack", "lbm_get_macros",
    "lbm_get_macros_current", "lbm_get_feq", "lbm_equilibrium", "lbm_mean_u", "lbm_set_active", "lbm_converge_check", "lbm_diagnostics", "lbm_sync", "lbm_get_counters", "lbm_engine_name",
]


class LBMError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("lbm_b200 error %d: %s" % (code, msg))
        self.code = code


class Config(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("batch", C.c_int32), ("dtype", C.c_int32),
                ("collision", C.c_int32), ("turb", C.c_int32), ("y0", C.c_int32), ("ny_local", C.c_int32),
                ("device", C.c_int32), ("engine", C.c_int32), ("semantics", C.c_int32), ("reserved", C.c_int32),
                ("ext_f", C.c_void_p * 2)]


class Layout(C.Structure):
    _fields_ = [("elem_size", C.c_int64), ("pitch", C.c_int64), ("rows", C.c_int64), ("plane", C.c_int64),
                ("cavity", C.c_int64), ("state_bytes", C.c_int64), ("ghost2_offset", C.c_int64)]


_lib = None


def load():
    """Load liblbm_b200.so (building it first if the sources are newer and nvcc is present)."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build
    if _build.stale():
        # sources newer than the library (or no library): rebuild; a stale library that cannot be rebuilt is an
        # error, never silently used -- parity tests and benchmarks must run the kernels in the tree
        try:
            _build.build()
        except Exception as exc:  # pragma: no cover - build container always has nvcc
            raise ImportError("liblbm_b200.so is missing or older than its sources and could not be built (%s); run "
                              "`python -m latticeboltzmannsimulations_b200.build`. There is no CPU fallback." % exc)
    lib = C.CDLL(LIB_PATH)
    H = C.c_void_p
    lib.lbm_last_error.restype = C.c_char_p
    lib.lbm_last_error.argtypes = []
    lib.lbm_abi_version.restype = C.c_int
    lib.lbm_device_count.argtypes = [C.POINTER(C.c_int)]
    lib.lbm_state_bytes.argtypes = [C.POINTER(Config), C.POINTER(C.c_size_t)]
    lib.lbm_create.argtypes = [C.POINTER(Config), C.POINTER(H)]
    lib.lbm_destroy.argtypes = [H]
    lib.lbm_get_layout.argtypes = [H, C.POINTER(Layout)]
    lib.lbm_set_tuning.argtypes = [H, C.c_char_p, C.c_int64]
    lib.lbm_set_reynolds.argtypes = [H, C.c_int, C.c_double, C.c_double]
    lib.lbm_set_rates.argtypes = [H, C.c_int] + [C.c_double] * 6
    lib.lbm_init_equilibrium.argtypes = [H]
    lib.lbm_upload_f.argtypes = [H, C.c_void_p, C.c_int, C.c_void_p]
    lib.lbm_download_f.argtypes = [H, C.c_void_p, C.c_int, C.c_void_p]
    lib.lbm_step.argtypes = [H, C.c_int, C.c_int, C.c_void_p]
    lib.lbm_step_region.argtypes = [H, C.c_int, C.c_int, C.c_void_p]
    lib.lbm_swap.argtypes = [H]
    lib.lbm_step2_region.argtypes = [H, C.c_int, C.c_int, C.c_void_p]
    lib.lbm_swap2.argtypes = [H]
    lib.lbm_step2_available.argtypes = [H]
    lib.lbm_buffer_ptr.argtypes = [H, C.c_int, C.POINTER(C.c_void_p)]
    lib.lbm_get_macros.argtypes = [H, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.lbm_get_macros_current.argtypes = [H, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.lbm_get_feq.argtypes = [H, C.c_void_p, C.c_int, C.c_void_p]
    lib.lbm_halo_pack.argtypes = [H, C.c_int, C.c_void_p, C.c_void_p]
    lib.lbm_halo_unpack.argtypes = [H, C.c_int, C.c_void_p, C.c_void_p]
    lib.lbm_equilibrium.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                    C.c_void_p]
    lib.lbm_mean_u.argtypes = [H, C.POINTER(C.c_double), C.c_void_p]
    lib.lbm_set_active.argtypes = [H, C.POINTER(C.c_int32), C.c_void_p]
    lib.lbm_converge_check.argtypes = [H, C.c_double, C.c_int, C.POINTER(C.c_int32), C.c_void_p]
    lib.lbm_diagnostics.argtypes = [H, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p]
    lib.lbm_sync.argtypes = [H]
    lib.lbm_get_counters.argtypes = [H, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.lbm_engine_name.argtypes = [H]
    lib.lbm_engine_name.restype = C.c_char_p
    for name in SYMBOLS:          # fail at load time, not at first use, if the library is stale
        getattr(lib, name)
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != LBM_OK:
        raise LBMError(rc, load().lbm_last_error().decode("utf-8", "replace"))
