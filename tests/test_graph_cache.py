"""GraphCache (host facade): a parsed road graph saved as flat SoA + CSR and loaded back
into containers that iterate in the recorded order -- the order the planner's tie-breaks depend on."""
import os

import numpy as np
import pytest

import goldenlib as gl
from imomd_b200 import capi, graphs


def _cases():
    return [("grid40", lambda: graphs.jittered_grid(40, seed=5)),
            ("road3000", lambda: graphs.road_like(3000, seed=2)),
            ("hubs", lambda: graphs.hub_grid(40, seed=4, hubs=14, spokes=9)),     # degree up to 13
            ("FRB", lambda: gl.graph("FRB")), ("san_pablo", lambda: gl.graph("san_pablo"))]


@pytest.mark.parametrize("name,make", _cases(), ids=[c[0] for c in _cases()])
def test_roundtrip_keeps_container_order(name, make, tmp_path):
    g = make()
    row_ptr, col, w, lat, lon = capi.graph_cache_roundtrip(g, str(tmp_path / "g.cache"))
    # (the C++ wrapper has already checked, element by element, that the loaded containers
    # iterate like the ones that were saved)
    assert np.array_equal(row_ptr, g.row_ptr)
    if g.eu is not None:
        # insertion-ordered edge list: g.col IS the reference containers' iteration order
        # (pinned against the real containers in test_oracle_vs_ref.py)
        assert np.array_equal(col, g.col)
        assert np.array_equal(w.view(np.uint64), g.w.view(np.uint64))
    else:
        # graph exported from the OSM parser without its insertion history: same rows as sets
        for v in range(g.n):
            a, b = slice(row_ptr[v], row_ptr[v + 1]), slice(g.row_ptr[v], g.row_ptr[v + 1])
            assert sorted(zip(col[a], w[a])) == sorted(zip(g.col[b], g.w[b]))
    assert np.array_equal(lat.view(np.uint64), g.lat.view(np.uint64))
    assert np.array_equal(lon.view(np.uint64), g.lon.view(np.uint64))
    assert os.path.getsize(tmp_path / "g.cache") == 32 + 16 * g.n + 8 * (g.n + 1) + 12 * len(g.col)


def test_degree_above_13_is_refused(tmp_path):
    n = 20
    lat = np.linspace(37.0, 37.001, n); lon = np.linspace(-122.0, -121.999, n)
    eu = np.zeros(14, np.uint64); ev = np.arange(1, 15, dtype=np.uint64)      # a 14-spoke star
    ew = np.ones(14)
    g = graphs.RoadGraph(lat=lat, lon=lon, row_ptr=np.zeros(n + 1, np.uint32), col=np.zeros(0, np.uint32),
                         w=np.zeros(0), eu=eu, ev=ev, ew=ew)
    with pytest.raises(capi.ImomdError, match="degree above 13"):
        capi.graph_cache_roundtrip(g, str(tmp_path / "g.cache"))


def test_damaged_file_is_rejected(tmp_path):
    g = graphs.jittered_grid(12, seed=1)
    path = tmp_path / "g.cache"
    capi.graph_cache_roundtrip(g, str(path))
    raw = bytearray(path.read_bytes())
    raw[0] ^= 0xFF                                   # magic
    path.write_bytes(bytes(raw))
    import ctypes as C
    L = capi.host_lib()
    L.imomd_host_cache_load_check.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    err = C.create_string_buffer(256)
    assert L.imomd_host_cache_load_check(str(path).encode(), err, 256) == -1
    assert b"not a graph cache" in err.value
    assert L.imomd_host_cache_load_check(str(tmp_path / "absent").encode(), err, 256) == 1
