"""CPU test of the optional VTK writer (reference interface VTKWrapper.saveToVTK, VTKWrapper.py:6-10)."""
import os

import numpy as np

from latticeboltzmannsimulations_b200.vtk_export import save_fields


def test_vtk_roundtrip(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    nx, ny = 7, 5
    rng = np.random.default_rng(0)
    rho, u = rng.uniform(0.9, 1.1, (nx, ny)), rng.uniform(-0.1, 0.1, (2, nx, ny))
    path = save_fields(rho, u, "ldc", 3)
    assert os.path.basename(path) == "ldc.00003.vtk"
    raw = open(path, "rb").read()
    assert raw.startswith(b"# vtk DataFile Version 3.0") and b"DIMENSIONS 7 5 1" in raw
    i = raw.index(b"VECTORS velocity double\n") + len(b"VECTORS velocity double\n")
    vec = np.frombuffer(raw[i:i + nx * ny * 3 * 8], dtype=">f8").reshape(1, ny, nx, 3)
    assert np.array_equal(vec[0, :, :, 0].T, u[0]) and np.array_equal(vec[0, :, :, 1].T, u[1])
    j = raw.index(b"LOOKUP_TABLE default\n") + len(b"LOOKUP_TABLE default\n")
    p = np.frombuffer(raw[j:j + nx * ny * 8], dtype=">f8").reshape(ny, nx)
    assert np.array_equal(p.T, rho)
