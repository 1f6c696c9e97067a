"""Host-side handle over the C ABI: one object = the device state of one (batch of) cavity / y-strip.

Mirrors what the reference scripts keep in module globals (``fin_g``, ``rho_g``, ``u_g`` ... ``MRT_GPU.py:309-328``)
and the calls of their time loop (``:707-757``).  Host arrays use the reference convention: ``f[9, nx, ny]``,
``rho[nx, ny]``, ``u[2, nx, ny]`` with ``y == 0`` the lid (a leading batch axis is added when ``batch > 1``).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _capi

_DTYPES = {"float32": (_capi.LBM_F32, np.float32), "float64": (_capi.LBM_F64, np.float64)}


def _dtype_name(dtype) -> str:
    name = np.dtype(dtype).name if not isinstance(dtype, str) else dtype
    name = {"f32": "float32", "f64": "float64", "fp32": "float32", "fp64": "float64"}.get(name, name)
    if name not in _DTYPES:
        raise ValueError("dtype must be float32 or float64, got %r" % (dtype,))
    return name


def _env_tuning() -> dict:
    import os
    out = {}
    for item in os.environ.get("LBM_B200_TUNING", "").split(","):
        if item.strip():
            k, _, v = item.partition("=")
            out[k.strip()] = int(v)
    return out


def _is_torch_cuda(x) -> bool:
    return type(x).__module__.startswith("torch") and getattr(x, "is_cuda", False)


class CavitySolver:
    def __init__(self, nx: int, ny: int, batch: int = 1, dtype="float64", collision: str = "MRT",
                 turb: bool = False, y0: int = 0, ny_local: Optional[int] = None, device: Optional[int] = None,
                 engine: str = "auto", ext_buffers: Optional[Sequence[int]] = None, semantics: str = "C",
                 tuning: Optional[dict] = None):
        self._lib = _capi.load()
        self._h = C.c_void_p()
        self.nx, self.ny, self.batch = int(nx), int(ny), int(batch)
        self.dtype_name = _dtype_name(dtype)
        code, self.np_dtype = _DTYPES[self.dtype_name]
        if collision not in _capi.COLLISIONS:
            raise ValueError("collision must be one of %s" % sorted(_capi.COLLISIONS))
        if engine not in _capi.ENGINES:
            raise ValueError("engine must be one of %s" % sorted(_capi.ENGINES))
        if semantics not in _capi.SEMANTICS:
            raise ValueError("semantics must be 'C' (MRT_GPU.py) or 'A' (MRT.py)")
        self.collision = collision
        self.semantics = semantics
        self.y0 = int(y0)
        self.ny_local = int(ny) if ny_local is None else int(ny_local)
        cfg = _capi.Config(nx=self.nx, ny=self.ny, batch=self.batch, dtype=code,
                           collision=_capi.COLLISIONS[collision], turb=int(bool(turb)), y0=self.y0,
                           ny_local=0 if ny_local is None else self.ny_local,
                           device=-1 if device is None else int(device), engine=_capi.ENGINES[engine],
                           semantics=_capi.SEMANTICS[semantics], reserved=0)
        if ext_buffers is not None:
            cfg.ext_f[0], cfg.ext_f[1] = int(ext_buffers[0]), int(ext_buffers[1])
        self._cfg = cfg
        _capi.check(self._lib.lbm_create(C.byref(cfg), C.byref(self._h)))
        lay = _capi.Layout()
        _capi.check(self._lib.lbm_get_layout(self._h, C.byref(lay)))
        self.layout = lay
        # kernel-selection knobs (lbm_set_tuning): LBM_B200_TUNING="key=value,..." for the tools, then the argument
        for key, value in list(_env_tuning().items()) + list((tuning or {}).items()):
            self.set_tuning(key, value)

    # -- lifetime ---------------------------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.lbm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @staticmethod
    def state_bytes(nx, ny, batch=1, dtype="float64", ny_local=None) -> int:
        lib = _capi.load()
        code, _ = _DTYPES[_dtype_name(dtype)]
        cfg = _capi.Config(nx=nx, ny=ny, batch=batch, dtype=code, collision=_capi.LBM_MRT, turb=0, y0=0,
                           ny_local=0 if ny_local is None else ny_local, device=-1, engine=0, semantics=0, reserved=0)
        out = C.c_size_t()
        _capi.check(lib.lbm_state_bytes(C.byref(cfg), C.byref(out)))
        return int(out.value)

    def set_tuning(self, key: str, value: int) -> None:
        """Kernel-selection knob (``lbm_set_tuning`` in include/lbm_b200.h); never changes a result."""
        _capi.check(self._lib.lbm_set_tuning(self._h, str(key).encode(), int(value)))

    # -- parameters -------------------------------------------------------------------------------------------
    def set_reynolds(self, Re, uLB: float = 0.08, cavity: int = -1) -> None:
        """``set_omega(uLB, Re, ysize)`` of functions.pyx:38-43; ``Re`` may be a sequence (one per cavity)."""
        if np.ndim(Re) == 0:
            _capi.check(self._lib.lbm_set_reynolds(self._h, int(cavity), float(uLB), float(Re)))
        else:
            if len(Re) != self.batch:
                raise ValueError("need one Re per cavity")
            for b, r in enumerate(Re):
                _capi.check(self._lib.lbm_set_reynolds(self._h, b, float(uLB), float(r)))

    def set_rates(self, uLB, omega_nu, omega_e=1.0, omega_eps=1.2, omega_q=1.2, omega_minus=1.0, cavity=-1):
        _capi.check(self._lib.lbm_set_rates(self._h, int(cavity), float(uLB), float(omega_nu), float(omega_e),
                                            float(omega_eps), float(omega_q), float(omega_minus)))

    # -- state ------------------------------------------------------------------------------------------------
    def init_equilibrium(self) -> None:
        _capi.check(self._lib.lbm_init_equilibrium(self._h))

    def _fshape(self) -> Tuple[int, ...]:
        s = (9, self.nx, self.ny_local)
        return s if self.batch == 1 else (self.batch,) + s

    def _host_arg(self, arr, shape):
        """Validate a caller array (numpy or torch CUDA) -> (pointer, on_device, keepalive)."""
        if _is_torch_cuda(arr):
            import torch
            want = torch.float64 if self.np_dtype is np.float64 else torch.float32
            if arr.dtype != want:
                raise ValueError("Buffer dtype mismatch, expected %r but got %r" % (want, arr.dtype))
            if tuple(arr.shape) != tuple(shape) or not arr.is_contiguous():
                raise ValueError("expected a contiguous tensor of shape %r, got %r" % (shape, tuple(arr.shape)))
            return arr.data_ptr(), 1, arr
        a = np.asarray(arr)
        if a.dtype != self.np_dtype:
            raise ValueError("Buffer dtype mismatch, expected %r but got %r" % (np.dtype(self.np_dtype).name, a.dtype.name))
        if tuple(a.shape) != tuple(shape):
            raise ValueError("expected shape %r, got %r" % (shape, a.shape))
        a = np.ascontiguousarray(a)
        return a.ctypes.data, 0, a

    def upload_f(self, f, stream: int = 0) -> None:
        ptr, on_dev, keep = self._host_arg(f, self._fshape())
        _capi.check(self._lib.lbm_upload_f(self._h, ptr, on_dev, C.c_void_p(stream)))
        del keep

    def download_f(self, out=None, stream: int = 0):
        if out is None:
            out = np.empty(self._fshape(), dtype=self.np_dtype)
        ptr, on_dev, keep = self._host_arg(out, self._fshape())
        if keep is not out and not on_dev:
            raise ValueError("out must be C-contiguous")
        _capi.check(self._lib.lbm_download_f(self._h, ptr, on_dev, C.c_void_p(stream)))
        return out

    # -- stepping ---------------------------------------------------------------------------------------------
    def step(self, nsteps: int = 1, write_macros: bool = True, stream: int = 0) -> None:
        _capi.check(self._lib.lbm_step(self._h, int(nsteps), int(bool(write_macros)), C.c_void_p(stream)))

    def step_region(self, region: int, write_macros: bool = False, stream: int = 0) -> None:
        _capi.check(self._lib.lbm_step_region(self._h, int(region), int(bool(write_macros)), C.c_void_p(stream)))

    def swap(self) -> None:
        _capi.check(self._lib.lbm_swap(self._h))

    def step2_available(self) -> bool:
        """True if the two-step (temporal blocking) kernel can advance this handle right now."""
        return bool(self._lib.lbm_step2_available(self._h))

    def step2_region(self, region: int, write_macros: bool = False, stream: int = 0) -> None:
        _capi.check(self._lib.lbm_step2_region(self._h, int(region), int(bool(write_macros)), C.c_void_p(stream)))

    def swap2(self) -> None:
        _capi.check(self._lib.lbm_swap2(self._h))

    def buffer_ptr(self, which: int) -> int:
        p = C.c_void_p()
        _capi.check(self._lib.lbm_buffer_ptr(self._h, int(which), C.byref(p)))
        return int(p.value)

    def halo_pack(self, direction: int, buf_ptr: int, stream: int = 0) -> None:
        """Nine halo rows for the strip above (0) / below (1) -> one contiguous device buffer ``[9, nx]``."""
        _capi.check(self._lib.lbm_halo_pack(self._h, int(direction), C.c_void_p(buf_ptr), C.c_void_p(stream)))

    def halo_unpack(self, direction: int, buf_ptr: int, stream: int = 0) -> None:
        """One contiguous buffer received from the strip above (0) / below (1) -> ghost rows of that side."""
        _capi.check(self._lib.lbm_halo_unpack(self._h, int(direction), C.c_void_p(buf_ptr), C.c_void_p(stream)))

    def macros(self, current: bool = False, rho_out=None, u_out=None, stream: int = 0):
        """(rho[nx,ny], u[2,nx,ny]) -- by default with the reference's one-step lag (moments of the state that
        entered the last step, MRT_GPU.py:616-631 + :756-757); ``current=True`` evaluates the present state."""
        rs = (self.nx, self.ny_local) if self.batch == 1 else (self.batch, self.nx, self.ny_local)
        us = (2, self.nx, self.ny_local) if self.batch == 1 else (self.batch, 2, self.nx, self.ny_local)
        if rho_out is None:
            rho_out = np.empty(rs, dtype=self.np_dtype)
        if u_out is None:
            u_out = np.empty(us, dtype=self.np_dtype)
        rp, d1, k1 = self._host_arg(rho_out, rs)
        up, d2, k2 = self._host_arg(u_out, us)
        if (k1 is not rho_out and not d1) or (k2 is not u_out and not d2):
            raise ValueError("rho_out and u_out must be C-contiguous arrays of the solver's dtype")
        if d1 != d2:
            raise ValueError("rho_out and u_out must both be host or both be device arrays")
        fn = self._lib.lbm_get_macros_current if current else self._lib.lbm_get_macros
        _capi.check(fn(self._h, rp, up, d1, C.c_void_p(stream)))
        return rho_out, u_out

    def feq(self, out=None, stream: int = 0):
        """Equilibrium of the stored (lagged) ``rho, u``: ``feq[9, nx, ny]`` as ``functions.allfunc`` returns it."""
        if out is None:
            out = np.empty(self._fshape(), dtype=self.np_dtype)
        ptr, on_dev, keep = self._host_arg(out, self._fshape())
        if keep is not out and not on_dev:
            raise ValueError("out must be C-contiguous")
        _capi.check(self._lib.lbm_get_feq(self._h, ptr, on_dev, C.c_void_p(stream)))
        return out

    def mean_u(self, stream: int = 0) -> np.ndarray:
        """Per-cavity ``np.mean(u)`` of the stored velocity field, reduced on the device (MRT_GPU_datagen.py:729)."""
        out = (C.c_double * self.batch)()
        _capi.check(self._lib.lbm_mean_u(self._h, out, C.c_void_p(stream)))
        return np.array(out[:], dtype=np.float64)

    def set_active(self, active, stream: int = 0) -> None:
        """Freeze the cavities whose flag is 0 (per-cavity ``break`` of MRT_GPU_datagen.py:731-733)."""
        flags = np.ascontiguousarray(np.asarray(active, dtype=np.int32))
        if flags.shape != (self.batch,):
            raise ValueError("need one flag per cavity")
        _capi.check(self._lib.lbm_set_active(self._h, flags.ctypes.data_as(C.POINTER(C.c_int32)), C.c_void_p(stream)))

    def converge_check(self, tol: float = 1e-7, hits: int = 6, read_back: bool = True, stream: int = 0):
        """One evaluation of the reference's stopping rule for every cavity, on the device (``lbm_converge_check``);
        returns the per-cavity active flags (``None`` with ``read_back=False``: no host synchronisation at all)."""
        if not read_back:
            _capi.check(self._lib.lbm_converge_check(self._h, float(tol), int(hits), None, C.c_void_p(stream)))
            return None
        out = (C.c_int32 * self.batch)()
        _capi.check(self._lib.lbm_converge_check(self._h, float(tol), int(hits), out, C.c_void_p(stream)))
        return np.array(out[:], dtype=np.int32)

    def diagnostics(self, cavity: int = 0, vortices: bool = True, stream: int = 0):
        """Centre-lines and vortex centres of the stored velocity field, reduced on the device.

        Returns ``(ux_col[ny], uy_row[nx], ((x1, y1), (x2, y2)))`` -- what ``MRT_GPU.py:764-776, 793-800`` computes
        on the host after downloading ``u``: ``u[0, nx//2, :]``, ``u[1, :, ny//2]`` and the two ``nanargmin`` locations.
        """
        ux_col = np.empty(self.ny_local, dtype=self.np_dtype)
        uy_row = np.empty(self.nx, dtype=self.np_dtype)
        loc = (C.c_int32 * 4)()
        _capi.check(self._lib.lbm_diagnostics(self._h, int(cavity), ux_col.ctypes.data, uy_row.ctypes.data,
                                              loc if vortices else None, C.c_void_p(stream)))
        return ux_col, uy_row, (((loc[0], loc[1]), (loc[2], loc[3])) if vortices else None)

    def sync(self) -> None:
        _capi.check(self._lib.lbm_sync(self._h))

    def counters(self) -> Tuple[int, int]:
        a, b = C.c_int64(), C.c_int64()
        _capi.check(self._lib.lbm_get_counters(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    @property
    def engine(self) -> str:
        return self._lib.lbm_engine_name(self._h).decode()
