"""Optional VTK export with the call signature of the reference's ``VTKWrapper.saveToVTK`` (``VTKWrapper.py:6-10``;
the call site ``MRT.py:604-610`` is dead upstream: ``SaveVTK = False`` and the import is commented out).

The reference delegates to the vendored third-party ``pyevtk``; this writer has no dependency: it emits a legacy-VTK
``RECTILINEAR_GRID`` file (binary, big-endian as the format demands) with the velocity as a point vector field and the
density as the scalar ``pressure`` (the reference's field names).  Not on the hot path.
"""
from __future__ import annotations

import numpy as np


def saveToVTK(velocity, rho, prefix: str, saveNumber: str, grid) -> str:
    """``velocity``: tuple (ux, uy, uz) of arrays ``[nx, ny, nz]``; ``rho``: ``[nx, ny, nz]``; ``grid``: (X, Y, Z) 1-D
    coordinate arrays.  Writes ``./<prefix>.<saveNumber>.vtk`` and returns the path."""
    X, Y, Z = (np.asarray(g, dtype=np.float64) for g in grid)
    ux, uy, uz = (np.asarray(v, dtype=np.float64) for v in velocity)
    rho = np.asarray(rho, dtype=np.float64)
    shape = (len(X), len(Y), len(Z))
    for name, arr in (("ux", ux), ("uy", uy), ("uz", uz), ("rho", rho)):
        if arr.shape != shape:
            raise ValueError("%s has shape %r, expected %r" % (name, arr.shape, shape))
    path = "./%s.%s.vtk" % (prefix, saveNumber)
    n = shape[0] * shape[1] * shape[2]

    def be(a):                               # VTK legacy order: x fastest, big-endian
        return np.ascontiguousarray(np.transpose(a, (2, 1, 0))).astype(">f8").tobytes()

    with open(path, "wb") as fh:
        fh.write(b"# vtk DataFile Version 3.0\nlid-driven cavity (lbm_b200)\nBINARY\nDATASET RECTILINEAR_GRID\n")
        fh.write(("DIMENSIONS %d %d %d\n" % shape).encode())
        for axis, coord in (("X", X), ("Y", Y), ("Z", Z)):
            fh.write(("%s_COORDINATES %d double\n" % (axis, len(coord))).encode())
            fh.write(coord.astype(">f8").tobytes() + b"\n")
        fh.write(("POINT_DATA %d\nVECTORS velocity double\n" % n).encode())
        vec = np.stack([np.transpose(c, (2, 1, 0)) for c in (ux, uy, uz)], axis=-1)
        fh.write(np.ascontiguousarray(vec).astype(">f8").tobytes() + b"\n")
        fh.write(b"SCALARS pressure double 1\nLOOKUP_TABLE default\n")
        fh.write(be(rho) + b"\n")
    return path


def save_fields(rho, u, prefix: str, index: int) -> str:
    """Convenience wrapper reproducing the conversion of ``MRT.py:604-610`` for 2-D fields ``rho[nx,ny]``, ``u[2,nx,ny]``."""
    nx, ny = rho.shape
    vel = np.reshape(u, (2, nx, ny, 1))
    grid = (np.arange(nx, dtype=np.float64), np.arange(ny, dtype=np.float64), np.arange(1, dtype=np.float64))
    return saveToVTK((vel[0], vel[1], np.zeros((nx, ny, 1))), np.reshape(rho, (nx, ny, 1)), prefix, str(index).zfill(5), grid)
