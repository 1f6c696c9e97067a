"""Validation helpers on top of ``CavitySolver.diagnostics``: comparison of the centre-line velocities with tabulated
reference stations (e.g. Ghia, Ghia & Shin 1982, the data of the reference's ``GhiaData.csv``).

The reference pairs LBM samples with the Ghia stations through ``int(Y * (ny-1))`` indices measured from the lid and
then reverses the array (``MRT.py:119-120, 559-561``), which pairs a station with the wrong depth; this module maps
stations physically (``y = 1 - j/(ny-1)`` measured from the bottom, ``x = i/(nx-1)``) and interpolates linearly.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np


def centerline_errors(ux_col: np.ndarray, uy_row: np.ndarray, uLB: float, Y: Sequence[float], Ux: Sequence[float],
                      X: Sequence[float], Uy: Sequence[float]) -> Tuple[float, float]:
    """max |u_x(x=0.5, Y)/uLB - Ux| and max |u_y(X, y=0.5)/uLB - Uy| over the given stations."""
    ux_col = np.asarray(ux_col, dtype=np.float64) / uLB
    uy_row = np.asarray(uy_row, dtype=np.float64) / uLB
    ny, nx = len(ux_col), len(uy_row)
    yphys = 1.0 - np.arange(ny) / (ny - 1.0)                 # index 0 is the lid
    ex = np.max(np.abs(np.interp(np.asarray(Y), yphys[::-1], ux_col[::-1]) - np.asarray(Ux)))
    xphys = np.arange(nx) / (nx - 1.0)
    ey = np.max(np.abs(np.interp(np.asarray(X), xphys, uy_row) - np.asarray(Uy)))
    return float(ex), float(ey)


def vortex_positions(loc, nx: int, ny: int):
    """Node indices -> physical (x, y) in [0,1]^2 with y measured from the bottom wall (as tabulated by Ghia)."""
    return [(x / (nx - 1.0), 1.0 - y / (ny - 1.0)) for (x, y) in loc if x >= 0]
