"""Drop-in for the reference's Cython module ``functions`` (``functions.pyx``), backed by the CUDA library.

``MRT_cython.py:23-24`` does ``from functions import sumf, equ, ucprod, copyfunc, allfunc, set_omega`` and then
calls ``equ`` once (``:210``), ``set_omega`` once (``:232``) and ``allfunc`` every iteration (``:453``).  With
``import latticeboltzmannsimulations_b200.functions as functions`` those call sites run unchanged:

* same names, argument order, array shapes (``[9,nx,ny]``, ``[2,nx,ny]``, ``[nx,ny]`` fp64, C-contiguous) and
  return convention -- ``allfunc`` returns ``(rho_new, u, fin_new, feq)`` with ``u`` and ``feq`` mutated in place
  and fresh ``rho`` / ``fin`` arrays (``functions.pyx:66, 222``);
* same error behaviour for a wrong dtype (``ValueError: Buffer dtype mismatch, expected 'double_t' ...``);
* collision = SRT, like ``functions.pyx:93``.  Wall handling follows the race-free reading of the same scheme,
  i.e. semantics "C" (``MRT_GPU.py`` funBC): ``functions.pyx``'s own boundary fix-ups read populations that
  other OpenMP iterations have not written yet (data race), so they cannot be reproduced
  deterministically; interior nodes, rho, u and feq agree with the compiled reference to 2.2e-16.
* ``set_omega`` keeps module-global state exactly like the reference (not re-entrant).
* The time loop of ``MRT_cython.py:453`` hands back the ``fin`` it just received.  To let that call skip the upload --
  the populations are still on the device -- the returned ``fin`` is marked READ-ONLY: passed back unchanged (same
  object) it is recognised and not copied up again; code that wants to modify it must ``fin = fin.copy()`` first, which
  is then uploaded like any other array (an in-place write raises instead of silently desynchronising the device).
  ``feq`` is evaluated on the device from the step's own ``rho, u`` (no round trip of the moments).

``sumf``, ``ucprod`` and ``copyfunc`` are not on the live path (only in commented-out code of ``MRT_cython.py``);
they are provided as thin NumPy one-liners for import compatibility.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _capi
from .solver import CavitySolver

omega = 0.0      # module globals, as in functions.pyx:11-13
uLatBo = 0.0
_Re = 0.0
_ysize = 0
_solver = None
_solver_key = None
_resident = None      # weakref to the fin array returned last: its populations are still on the device

c = np.array([[0, 0], [1, 0], [0, 1], [-1, 0], [0, -1], [1, 1], [-1, 1], [-1, -1], [1, -1]])   # functions.pyx:9


def _need_f64(name, a, ndim):
    a_np = np.asarray(a)
    if a_np.dtype != np.float64:
        raise ValueError("Buffer dtype mismatch, expected 'double_t' but got %r" % a_np.dtype.name)
    if a_np.ndim != ndim:
        raise ValueError("Buffer has wrong number of dimensions (expected %d, got %d)" % (ndim, a_np.ndim))
    return a_np


def set_omega(uLB, Re, ysize):
    """functions.pyx:38-43 -- ``Re`` and ``ysize`` are integer-typed there; integral floats are accepted."""
    global omega, uLatBo, _Re, _ysize
    if int(Re) != Re or int(ysize) != ysize:
        raise TypeError("an integer is required")
    uLatBo = float(uLB)
    nuLB = uLB * int(ysize) / int(Re)
    omega = 2.0 / (6. * nuLB + 1)
    _Re, _ysize = int(Re), int(ysize)


def equ(rho, ux, uy):
    """functions.pyx:229-267: feq[9,nx,ny] from rho[nx,ny], ux[nx,ny], uy[nx,ny] (computed on the GPU)."""
    rho = np.ascontiguousarray(_need_f64("rho", rho, 2))
    ux = np.ascontiguousarray(_need_f64("ux", ux, 2))
    uy = np.ascontiguousarray(_need_f64("uy", uy, 2))
    feq = np.empty((9,) + rho.shape)
    lib = _capi.load()
    _capi.check(lib.lbm_equilibrium(_capi.LBM_F64, rho.size, rho.ctypes.data, ux.ctypes.data, uy.ctypes.data,
                                    feq.ctypes.data, 0, None))
    return feq


def _get_solver(nx, ny):
    global _solver, _solver_key, _resident
    key = (nx, ny)
    if _solver is None or _solver_key != key:
        if _solver is not None:
            _solver.close()
        _solver = CavitySolver(nx, ny, 1, "float64", "SRT")
        _solver_key = key
        _resident = None
    return _solver


def allfunc(rho, u, fin, feq):
    """functions.pyx:45-222 -- one full step (moments, SRT collision, streaming, walls + lid) on the GPU."""
    _need_f64("rho", rho, 2)
    u = _need_f64("u", u, 3)
    fin = _need_f64("fin", fin, 3)
    feq = _need_f64("feq", feq, 3)
    if omega == 0.0:
        raise RuntimeError("set_omega(uLB, Re, ysize) must be called before allfunc (functions.pyx:38)")
    global _resident
    nx, ny = fin.shape[1], fin.shape[2]
    s = _get_solver(nx, ny)
    s.set_rates(uLatBo, omega, omega_minus=omega)
    # the array this function returned last, passed back untouched (it is read-only): its populations are on the device
    if not (_resident is not None and _resident() is fin and not fin.flags.writeable):
        s.upload_f(fin)
    s.step(1, write_macros=True)
    rho_new = np.empty((nx, ny))
    if u.flags.c_contiguous:
        s.macros(rho_out=rho_new, u_out=u)           # u mutated in place (functions.pyx:73-81)
    else:
        u[...] = s.macros(rho_out=rho_new)[1]
    if feq.flags.c_contiguous:
        s.feq(out=feq)                               # mutated in place (functions.pyx:88)
    else:
        feq[...] = s.feq()
    fin_new = s.download_f()
    fin_new.setflags(write=False)
    _resident = weakref.ref(fin_new)
    return rho_new, u, fin_new, feq


def sumf(fin):                       # functions.pyx:32-33
    return np.sum(fin, axis=0)


def ucprod(c_, fin, rho):            # functions.pyx:269-284
    vel = np.empty((2,) + fin.shape[1:])
    vel[0] = sum(c_[k, 0] * fin[k] for k in range(9)) / rho
    vel[1] = sum(c_[k, 1] * fin[k] for k in range(9)) / rho
    return vel


def copyfunc(fout, fin):             # functions.pyx:286-296
    fout[...] = fin
    return fout
