"""The C ABI library loads on a CPU-only box and exports every symbol include/imomd_b200.h
declares; compute entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from imomd_b200 import capi, graphs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "imomd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(imomd_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    names = _declared()
    assert len(names) >= 25
    lib = ctypes.CDLL(capi.LIB_CUDA)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_host_library_loads():
    L = capi.host_lib()
    assert L.imomd_host_great_circle_m(37.0, -122.0, 37.0, -122.0) == 0.0
    for name in ("imomd_host_run", "imomd_host_solver_solve", "imomd_host_selection_probabilities"):
        assert hasattr(L, name)


def test_version_and_sizes():
    L = capi.cuda_lib()
    assert b"sm_100a" in L.imomd_version()
    assert capi.sizeof_query_state() > 32 * 32 * 8 * 3


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    g = graphs.jittered_grid(8)
    with pytest.raises(capi.ImomdError):
        capi.DeviceGraph(g)
    with pytest.raises(capi.ImomdError):
        capi.GaContext(0)


def test_argument_validation_needs_no_gpu():
    """graph validation happens before any CUDA call"""
    L = capi.cuda_lib()
    g = graphs.jittered_grid(6)
    col = g.col.copy(); col[0] = 0 if g.col[0] != 0 else 1   # breaks symmetry / creates a self loop
    out = ctypes.c_void_p()
    rc = L.imomd_graph_create(0, g.n, g.row_ptr, col, g.w, g.lat, g.lon, ctypes.byref(out))
    assert rc == 3 and b"arc" in L.imomd_last_error() or b"self loop" in L.imomd_last_error()


def test_facade_constructor_rejects_bad_input_before_touching_the_device():
    """ImomdRRT's constructor (the reference reports nothing here: out-of-range ids are UB upstream)"""
    g = graphs.jittered_grid(12)
    with pytest.raises(capi.ImomdError, match="more than 30 objectives"):
        capi.FacadePlanner(g, list(range(33)))
    with pytest.raises(capi.ImomdError, match="destination id out of range"):
        capi.FacadePlanner(g, [0, g.n + 5, 7])
