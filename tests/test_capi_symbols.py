"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports exactly what include/lbm_b200.h
declares, the ctypes mirror lists the same symbols, and the product fails loudly (no CPU fallback) without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "lbm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lbm_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from latticeboltzmannsimulations_b200 import _capi
    lib = _capi.load()
    names = _header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_capi.SYMBOLS) == names
    assert lib.lbm_abi_version() == 3


def test_config_struct_matches_header_size():
    from latticeboltzmannsimulations_b200 import _capi
    assert ctypes.sizeof(_capi.Config) == 12 * 4 + 2 * 8
    assert ctypes.sizeof(_capi.Layout) == 7 * 8


def test_argument_validation_without_gpu():
    from latticeboltzmannsimulations_b200 import _capi
    lib = _capi.load()
    cfg = _capi.Config(nx=2, ny=64, batch=1, dtype=1, collision=2, turb=0, y0=0, ny_local=0, device=-1, engine=0,
                       semantics=0, reserved=0)
    n = ctypes.c_size_t()
    assert lib.lbm_state_bytes(ctypes.byref(cfg), ctypes.byref(n)) == _capi.LBM_EINVAL
    assert b"nx" in lib.lbm_last_error()
    cfg.nx = 100
    assert lib.lbm_state_bytes(ctypes.byref(cfg), ctypes.byref(n)) == 0
    assert n.value == (9 * (64 + 2) + 6) * 128 * 8    # pitch 128, one ghost row each side, + the 6-row ghost2 tail


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point must raise, never silently compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import latticeboltzmannsimulations_b200 as L
    from latticeboltzmannsimulations_b200 import functions
    with pytest.raises(L.LBMError):
        L.run_cavity(32, 32, 100, steps=1)
    with pytest.raises(L.LBMError):
        functions.equ(np.ones((4, 4)), np.zeros((4, 4)), np.zeros((4, 4)))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "latticeboltzmannsimulations_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), fn
                assert "lbm_oracle" not in txt.replace("oracle/lbm_oracle.py", ""), fn
                assert "host_emu" not in txt, fn          # the CPU emulation of the kernel source is test-only too


def test_functions_shim_signature_errors():
    """Same error behaviour as the Cython module for a wrong dtype."""
    from latticeboltzmannsimulations_b200 import functions
    with pytest.raises(ValueError, match="Buffer dtype mismatch, expected 'double_t'"):
        functions.allfunc(np.ones((4, 4), np.float32), np.zeros((2, 4, 4)), np.zeros((9, 4, 4)), np.zeros((9, 4, 4)))
    with pytest.raises(TypeError):
        functions.set_omega(0.08, 100.5, 64)
    functions.set_omega(0.08, 100.0, 64)
    assert abs(functions.omega - 2.0 / (6 * 0.08 * 64 / 100 + 1)) < 1e-16


def test_dataset_writer_names_and_dtypes(tmp_path):
    """save_dataset writes the four files of MRT_GPU_datagen.py:899-902; Re_range is int64 when integral (np.arange)."""
    from latticeboltzmannsimulations_b200.cavity import save_dataset, _re_range_array
    f = np.zeros((2, 9, 8, 6), np.float32); u = np.zeros((2, 2, 8, 6), np.float32); feq = np.ones((9, 8, 6), np.float32)
    save_dataset(str(tmp_path), f, u, feq, [100.0, 110.0])
    assert sorted(p.name for p in tmp_path.iterdir()) == ["Re_range.npy", "f_final.npy", "feq_initial.npy", "u_final.npy"]
    assert np.load(tmp_path / "Re_range.npy").dtype == np.int64
    assert _re_range_array([100.5, 200.0]).dtype == np.float64 and _re_range_array(np.arange(3)).dtype == np.int64
    assert np.load(tmp_path / "f_final.npy").shape == (2, 9, 8, 6)


def test_batch_limit_is_validated():
    from latticeboltzmannsimulations_b200 import _capi
    lib = _capi.load()
    cfg = _capi.Config(nx=8, ny=8, batch=70000, dtype=1, collision=2, turb=0, y0=0, ny_local=0, device=-1, engine=0,
                       semantics=0, reserved=0)
    out = ctypes.c_size_t()
    assert lib.lbm_state_bytes(ctypes.byref(cfg), ctypes.byref(out)) == _capi.LBM_EINVAL


def test_ext_buffers_come_in_pairs():
    """Caller-owned population buffers (lbm_config_t.ext_f): both or neither, and two different ones -- checked before any
    device is touched."""
    from latticeboltzmannsimulations_b200 import _capi
    lib = _capi.load()
    out = ctypes.c_size_t()
    for a, b in ((4096, 0), (0, 4096), (4096, 4096)):
        cfg = _capi.Config(nx=8, ny=8, batch=1, dtype=1, collision=2, turb=0, y0=0, ny_local=0, device=-1, engine=0,
                           semantics=0, reserved=0)
        cfg.ext_f[0], cfg.ext_f[1] = a, b
        assert lib.lbm_state_bytes(ctypes.byref(cfg), ctypes.byref(out)) == _capi.LBM_EINVAL
        assert b"ext_f" in lib.lbm_last_error()


def test_aa_engine_configuration_is_validated():
    """LBM_ENGINE_AA (one population buffer): whole cavities of semantics C that own their buffer."""
    from latticeboltzmannsimulations_b200 import _capi
    lib = _capi.load()
    out = ctypes.c_size_t()
    base = dict(nx=64, ny=64, batch=2, dtype=1, collision=2, turb=1, y0=0, ny_local=0, device=-1,
                engine=_capi.ENGINES["aa"], semantics=0, reserved=0)
    assert lib.lbm_state_bytes(ctypes.byref(_capi.Config(**base)), ctypes.byref(out)) == 0 and out.value > 0
    for change, word in ((dict(ny_local=32), b"whole"), (dict(semantics=1, collision=0, turb=0), b"semantics"),
                         (dict(engine=4), b"engine")):
        cfg = _capi.Config(**dict(base, **change))
        assert lib.lbm_state_bytes(ctypes.byref(cfg), ctypes.byref(out)) == _capi.LBM_EINVAL
        assert word in lib.lbm_last_error(), lib.lbm_last_error()
    cfg = _capi.Config(**base)
    cfg.ext_f[0], cfg.ext_f[1] = 4096, 8192
    assert lib.lbm_state_bytes(ctypes.byref(cfg), ctypes.byref(out)) == _capi.LBM_EINVAL and b"ext_f" in lib.lbm_last_error()
