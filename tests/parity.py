"""Shared parity checks: an implementation under test (CUDA path or its host simulation)
against the oracle restatement.  `exact_heuristics` is True when both sides use the same
libm (host simulation); on the GPU the min-heuristic *values* come from CUDA's libm and
are compared to 1e-12 relative, everything else bit for bit."""
import numpy as np

NONE32 = 0xFFFFFFFF


def conv_ids(a):
    """uint32 ids with 0xFFFFFFFF -> uint64 with 2^64-1 (the oracle's size_t -1)."""
    a = np.asarray(a).astype(np.uint64)
    a[a == NONE32] = np.uint64(0xFFFFFFFFFFFFFFFF)
    return a


def compare_with_oracle(impl, op, K, exact_heuristics, q=0, check_frontier=True):
    for k in range(K):
        pi, ci = impl.tree(k, q)
        po, co = op.tree(k)
        bad = np.nonzero(pi != po)[0]
        assert bad.size == 0, f"tree {k}: parents differ at {bad[:8]} impl={pi[bad[:8]]} oracle={po[bad[:8]]}"
        badc = np.nonzero(ci.view(np.uint64) != co.view(np.uint64))[0]
        assert badc.size == 0, f"tree {k}: cost bits differ at {badc[:8]} impl={ci[badc[:8]]} oracle={co[badc[:8]]}"
        if check_frontier:
            assert np.array_equal(impl.frontier(k, q), op.frontier(k)), f"tree {k}: frontier differs"
    si, mo, fo = impl.state(q), op.matrices(), op.flags()
    assert np.array_equal(si["D"].view(np.uint64), mo["D"].view(np.uint64)), "distance matrix differs"
    assert np.array_equal(conv_ids(si["conn"]), mo["conn"]), "connection nodes differ"
    pa, pb = si["P"], mo["P"]
    assert np.array_equal(pa.view(np.uint64), pb.view(np.uint64)), "probability matrix differs"
    assert np.array_equal(conv_ids(si["mh_id"]), mo["mh_id"]), "min-heuristic holders differ"
    # min-heuristic VALUES: the device logic evaluates the heuristics' haversine with its own
    # small-argument polynomials (imomd_core.cuh haversine_pc), the oracle with glibc like the
    # reference.  exact_heuristics (host simulation: same libm otherwise): within a few ulp;
    # else (CUDA libm): within the 1e-12 band inside which comparisons are decided on glibc's values.
    a, b = si["mh_val"], mo["mh_val"]
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin)
    assert np.allclose(a[fin], b[fin], rtol=2e-15 if exact_heuristics else 1e-12, atol=0.0)
    assert np.array_equal(si["is_done"], fo["is_done"])
    assert np.array_equal(si["ds_parent"], fo["ds_parent"])
    assert si["connected"] == fo["connected"] and si["dm_updated"] == fo["dm_updated"]
