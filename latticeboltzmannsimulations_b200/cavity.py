"""Script-level entry points of the reference, as callables.

The upstream solvers are flat scripts: edit ``Re``, ``xsize, ysize``, ``uLB``, ``maxIt``, ``RT`` at the top
(``MRT_GPU.py:45-58``), run, and read ``rho[x,y]``, ``u[2,x,y]``, ``fin[9,x,y]`` from the globals
(``:752-760``); ``MRT_GPU_datagen.py`` wraps the same loop in ``for Re in Re_range`` (``:55-57``) and stacks
``f_final[N,9,nx,ny]``, ``u_final[N,2,nx,ny]`` (``:879-902``).  These functions expose exactly those inputs and
outputs; the time loop runs on the GPU through the C ABI.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from .solver import CavitySolver


def run_cavity(nx: int, ny: int, Re: float, uLB: float = 0.08, steps: int = 1000, collision: str = "MRT",
               dtype="float64", turb: bool = False, f0=None, return_f: bool = False, current_macros: bool = False,
               device: Optional[int] = None, engine: str = "auto", semantics: str = "C"):
    """One lid-driven cavity: returns ``(rho[nx,ny], u[2,nx,ny])`` (+ ``f[9,nx,ny]`` with ``return_f``).

    ``rho, u`` carry the reference's one-step lag (they are the moments of the state that entered the last step);
    pass ``current_macros=True`` for the moments of the returned ``f`` instead.  ``f0`` (``[9,nx,ny]``, host array,
    e.g. pinned) replaces the equilibrium start ``rho = 1, u = (uLB,0)`` on the lid row (``MRT_GPU.py:259-267``).
    ``semantics="A"`` (with ``collision="SRT"``) reproduces the NumPy solver ``MRT.py:286-453`` instead of the GPU
    scripts, quirks included, for like-for-like comparisons with that script.
    """
    with CavitySolver(nx, ny, 1, dtype, collision, turb, device=device, engine=engine, semantics=semantics) as s:
        s.set_reynolds(Re, uLB)
        if f0 is None:
            s.init_equilibrium()
        else:
            s.upload_f(f0)
        s.step(int(steps), write_macros=True)
        rho, u = s.macros(current=current_macros)
        if return_f:
            return rho, u, s.download_f()
        return rho, u


def _re_range_array(Re_list) -> np.ndarray:
    """``Re_range`` as the reference stores it: ``np.arange(100, 5100, 10)`` is int64 (``MRT_GPU_datagen.py:55``); a list
    with non-integral entries stays float64."""
    arr = np.asarray(list(Re_list))
    if arr.dtype.kind in "iu":
        return arr.astype(np.int64)
    arr = arr.astype(np.float64)
    return arr.astype(np.int64) if np.all(arr == np.round(arr)) else arr


def datagen(Re_list: Sequence[float], nx: int = 384, ny: int = 384, uLB: float = 0.08, steps: int = 10000,
            collision: str = "MRT", dtype="float32", turb: bool = False, device: Optional[int] = None,
            engine: str = "auto", chunk: Optional[int] = None, converge: bool = False, Pinterval: int = 10000,
            maxIt: int = 3000000, tol: float = 1e-7, hits: int = 6, return_steps: bool = False,
            out_dir: Optional[str] = None):
    """Batched Reynolds sweep of ``MRT_GPU_datagen.py``: one cavity per entry of ``Re_list``.

    Returns ``(f_final[N,9,nx,ny], u_final[N,2,nx,ny], feq_initial[9,nx,ny], Re_range[N])`` in the dtype / ``[x,y]``
    indexing of the files the reference saves (``:899-902``; ``Re_range`` is int64 when the values are integral, like
    the reference's ``np.arange``); with ``out_dir`` the four ``.npy`` files are written there as well
    (``save_dataset``).  Cavities are independent (no communication); the multi-GPU driver
    (``distributed.datagen_sharded``) gives rank r the cavities ``r::world``.

    ``converge=False``: every cavity advances exactly ``steps`` steps (fixed work, used for timing).
    ``converge=True``: the reference's stopping rule (``:716-737``), evaluated per cavity ON THE DEVICE
    (``lbm_converge_check``): at every iteration ``It`` with ``It % Pinterval == 0`` the mean of the stored velocity
    field is compared with the one of the previous check, ``abs(mean(u) - mean(u_past)) / uLB < tol`` increments a
    counter (never reset, as in the reference) and the cavity stops when it exceeds ``hits - 1``; its ``fin`` / ``u`` at
    that moment are what is returned.  A cavity that never converges returns, like the reference, the fields of its
    LAST CHECK (the script only downloads at checks, ``:725-726``, and saves what it downloaded last), i.e. of
    iteration ``Pinterval * ((maxIt - 1) // Pinterval)``.  Stopped cavities are frozen on the device and cost no further
    bandwidth.
    """
    Re_arr = np.asarray(list(Re_list), dtype=np.float64)
    n = len(Re_arr)
    _, npdt = {"float32": (0, np.float32), "float64": (1, np.float64)}[np.dtype(dtype).name]
    f_final = np.empty((n, 9, nx, ny), dtype=npdt)
    u_final = np.empty((n, 2, nx, ny), dtype=npdt)
    steps_done = np.zeros(n, dtype=np.int64)
    feq_initial = None
    chunk = n if chunk is None else max(1, int(chunk))
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        nb = hi - lo
        with CavitySolver(nx, ny, nb, dtype, collision, turb, device=device, engine=engine) as s:
            s.set_reynolds(Re_arr[lo:hi], uLB)
            s.init_equilibrium()
            if feq_initial is None:
                f_init = s.download_f()
                feq_initial = np.array(f_init if nb == 1 else f_init[0])
            if not converge:
                s.step(int(steps), write_macros=True)
                steps_done[lo:hi] = int(steps)
            else:
                active = np.ones(nb, dtype=np.int32)
                it = 0
                while it < maxIt:
                    s.step(1, write_macros=True)          # iteration `it`, followed by the check of :724-737
                    steps_done[lo:hi][active == 1] = it + 1
                    active = s.converge_check(tol, hits)
                    if not active.any() or it + Pinterval >= maxIt:
                        break                             # all stopped, or this was the last check before maxIt
                    s.step(Pinterval - 1, write_macros=False)
                    it += Pinterval
            _, u = s.macros()
            f = s.download_f()
            f_final[lo:hi] = f.reshape(nb, 9, nx, ny)
            u_final[lo:hi] = u.reshape(nb, 2, nx, ny)
    Re_out = _re_range_array(Re_list)
    if out_dir is not None:
        save_dataset(out_dir, f_final, u_final, feq_initial, Re_out)
    if return_steps:
        return f_final, u_final, feq_initial, Re_out, steps_done
    return f_final, u_final, feq_initial, Re_out


def save_dataset(out_dir: str, f_final, u_final, feq_initial, Re_range) -> None:
    """Write the four files of ``MRT_GPU_datagen.py:899-902`` -- ``feq_initial.npy``, ``f_final.npy``, ``u_final.npy``,
    ``Re_range.npy`` -- which the CNN scripts load by exactly these names (``CNN_test.py:18-21``,
    ``CNNTen_384/CNN_Ten.py:22-25``)."""
    import os
    os.makedirs(out_dir, exist_ok=True)
    np.save(os.path.join(out_dir, "feq_initial.npy"), np.asarray(feq_initial))
    np.save(os.path.join(out_dir, "f_final.npy"), np.asarray(f_final))
    np.save(os.path.join(out_dir, "u_final.npy"), np.asarray(u_final))
    np.save(os.path.join(out_dir, "Re_range.npy"), _re_range_array(Re_range))
