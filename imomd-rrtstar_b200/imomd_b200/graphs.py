"""Road-graph containers and synthetic generators (host side, numpy only).

The hot path consumes the reference's two containers (README.md:233-236 of the
reference): ``nodes : vector<location_t>`` and ``edges :
vector<unordered_map<size_t,double>>``.  On the device they become a CSR
adjacency *in the containers' native iteration order* plus SoA coordinates
(SURVEY.md section 8 a-1).  This module

* holds that CSR (`RoadGraph`),
* rebuilds the native order from an edge list (`native_csr_from_edges`: what
  libstdc++'s ``unordered_map<size_t,double>`` iterates after the two inserts
  per edge of src/osm_parser.cpp:179-180) so synthetic graphs built here are
  seen identically by the reference and by the CUDA path,
* generates the synthetic workloads named in BASELINE.json (jittered grid,
  road-like graph with degree-2 chains, bug-trap grid).
"""
from __future__ import annotations

import dataclasses
import struct

import numpy as np

EARTH_RADIUS = 6371000  # int in the reference (compute_haversine.hpp:45)
GRAPH_MAGIC = int.from_bytes(b"IMOMDGR1", "little")


def haversine_np(lat1, lon1, lat2, lon2):
    """Vectorised haversine with the reference's formula (compute_haversine.hpp:41-58).

    Used only to *produce input data* (edge weights of synthetic graphs); the
    weights are then handed to every implementation as plain doubles.
    """
    d2r = np.pi / 180
    a = np.sin((lat2 - lat1) * d2r / 2) ** 2 + np.cos(lat1 * d2r) * np.cos(lat2 * d2r) * np.sin(
        (lon2 - lon1) * d2r / 2
    ) ** 2
    return EARTH_RADIUS * (2 * np.arctan2(np.sqrt(a), np.sqrt(1 - a)))


@dataclasses.dataclass
class RoadGraph:
    """CSR road graph + SoA coordinates, adjacency in native iteration order."""

    lat: np.ndarray  # f64[n]
    lon: np.ndarray  # f64[n]
    row_ptr: np.ndarray  # u32[n+1]
    col: np.ndarray  # u32[2E]
    w: np.ndarray  # f64[2E]
    # the edge list the CSR was built from (insertion order), when known
    eu: np.ndarray | None = None
    ev: np.ndarray | None = None
    ew: np.ndarray | None = None
    name: str = "graph"

    @property
    def n(self) -> int:
        return int(self.lat.shape[0])

    @property
    def arcs(self) -> int:
        return int(self.col.shape[0])

    def degree(self) -> np.ndarray:
        return np.diff(self.row_ptr.astype(np.int64))

    def save(self, path: str) -> None:
        extra = {}
        if self.eu is not None:
            extra = dict(eu=self.eu, ev=self.ev, ew=self.ew)
        np.savez_compressed(
            path, lat=self.lat, lon=self.lon, row_ptr=self.row_ptr, col=self.col, w=self.w,
            name=np.array(self.name), **extra
        )

    @staticmethod
    def load(path: str) -> "RoadGraph":
        z = np.load(path)
        return RoadGraph(
            lat=z["lat"], lon=z["lon"], row_ptr=z["row_ptr"], col=z["col"], w=z["w"],
            eu=z["eu"] if "eu" in z else None, ev=z["ev"] if "ev" in z else None,
            ew=z["ew"] if "ew" in z else None, name=str(z["name"]),
        )

    def edge_list(self):
        """(eu, ev, ew) that reproduces this graph; falls back to the CSR arcs u<v."""
        if self.eu is not None:
            return self.eu, self.ev, self.ew
        src = np.repeat(np.arange(self.n, dtype=np.uint64), self.degree())
        dst = self.col.astype(np.uint64)
        keep = src < dst
        return src[keep], dst[keep], self.w[keep]

    def write_probe_file(self, path: str) -> None:
        """Binary edge-list file read by oracle/ref_probe.cpp (--graph)."""
        if self.eu is None:
            raise ValueError("graph has no insertion-ordered edge list")
        with open(path, "wb") as f:
            f.write(struct.pack("<4Q", GRAPH_MAGIC, self.n, len(self.eu), 1))
            f.write(np.ascontiguousarray(self.lat, dtype=np.float64).tobytes())
            f.write(np.ascontiguousarray(self.lon, dtype=np.float64).tobytes())
            f.write(np.ascontiguousarray(self.eu, dtype=np.uint64).tobytes())
            f.write(np.ascontiguousarray(self.ev, dtype=np.uint64).tobytes())
            f.write(np.ascontiguousarray(self.ew, dtype=np.float64).tobytes())

    def components(self) -> np.ndarray:
        """Connected-component label per vertex."""
        from scipy.sparse import csr_matrix
        from scipy.sparse.csgraph import connected_components

        m = csr_matrix(
            (np.ones(self.arcs, dtype=np.int8), self.col.astype(np.int64),
             self.row_ptr.astype(np.int64)), shape=(self.n, self.n),
        )
        return connected_components(m, directed=False)[1]


def native_csr_from_edges(n: int, eu, ev, ew):
    """Adjacency rows in libstdc++ ``unordered_map<size_t,double>`` iteration order.

    The reference fills each row with ``insert({neighbour, weight})`` twice per edge
    (src/osm_parser.cpp:179-180).  With at most 13 keys a row has 13 buckets and an
    identity hash: a new key goes to the head of its bucket's run (bucket = key % 13),
    or to the head of the whole list when that bucket is empty; a duplicate key is
    ignored.  Vectorised over vertices: the k-th insertion of every row at once.
    """
    eu = np.asarray(eu, dtype=np.int64)
    ev = np.asarray(ev, dtype=np.int64)
    ew = np.asarray(ew, dtype=np.float64)
    m = eu.shape[0]
    # insertion stream: (row, key, weight) in program order
    rows = np.empty(2 * m, dtype=np.int64)
    keys = np.empty(2 * m, dtype=np.int64)
    wts = np.empty(2 * m, dtype=np.float64)
    rows[0::2], keys[0::2], wts[0::2] = eu, ev, ew
    rows[1::2], keys[1::2], wts[1::2] = ev, eu, ew
    order = np.argsort(rows, kind="stable")
    rows, keys, wts = rows[order], keys[order], wts[order]
    counts = np.bincount(rows, minlength=n)
    starts = np.concatenate(([0], np.cumsum(counts)))[:-1]
    rank = np.arange(2 * m) - starts[rows]
    maxins = int(counts.max()) if m else 0
    cap = 13
    tab = np.full((n, cap), -1, dtype=np.int64)
    tabw = np.zeros((n, cap), dtype=np.float64)
    size = np.zeros(n, dtype=np.int64)
    for k in range(maxins):
        sel = np.nonzero(rank == k)[0]
        if sel.size == 0:
            break
        r, key, wt = rows[sel], keys[sel], wts[sel]
        cur = tab[r]  # (s, cap)
        valid = np.arange(cap)[None, :] < size[r][:, None]
        dup = ((cur == key[:, None]) & valid).any(axis=1)
        r, key, wt, cur, valid = r[~dup], key[~dup], wt[~dup], cur[~dup], valid[~dup]
        if r.size == 0:
            continue
        if (size[r] >= cap).any():
            raise ValueError("degree above 13: container order not emulated")
        same = ((cur % 13) == (key % 13)[:, None]) & valid
        at = np.where(same.any(axis=1), same.argmax(axis=1), 0)
        curw = tabw[r]
        idx = np.arange(cap)[None, :]
        shifted = np.where(idx > at[:, None], np.roll(cur, 1, axis=1), cur)
        shiftedw = np.where(idx > at[:, None], np.roll(curw, 1, axis=1), curw)
        shifted[np.arange(r.size), at] = key
        shiftedw[np.arange(r.size), at] = wt
        tab[r] = shifted
        tabw[r] = shiftedw
        size[r] += 1
    row_ptr = np.concatenate(([0], np.cumsum(size))).astype(np.uint32)
    mask = np.arange(cap)[None, :] < size[:, None]
    col = tab[mask].astype(np.uint32)
    w = tabw[mask]
    return row_ptr, col, w


def from_edges(lat, lon, eu, ev, ew=None, name="graph") -> RoadGraph:
    lat = np.ascontiguousarray(lat, dtype=np.float64)
    lon = np.ascontiguousarray(lon, dtype=np.float64)
    eu = np.ascontiguousarray(eu, dtype=np.uint64)
    ev = np.ascontiguousarray(ev, dtype=np.uint64)
    if ew is None:
        ew = haversine_np(lat[eu.astype(np.int64)], lon[eu.astype(np.int64)],
                          lat[ev.astype(np.int64)], lon[ev.astype(np.int64)])
    ew = np.ascontiguousarray(ew, dtype=np.float64)
    row_ptr, col, w = native_csr_from_edges(lat.shape[0], eu, ev, ew)
    return RoadGraph(lat=lat, lon=lon, row_ptr=row_ptr, col=col, w=w, eu=eu, ev=ev, ew=ew, name=name)


def jittered_grid(width: int, seed: int = 123, height: int | None = None,
                  lat0: float = 37.0, lon0: float = -122.0, step: float = 1e-4,
                  jitter: float = 2e-5, name: str | None = None) -> RoadGraph:
    """W x H 4-neighbour lattice with jittered coordinates.

    The jitter removes the exact haversine ties a perfect lattice has.  Vertex
    ``r*W + c``; per vertex the edge to ``+1`` is inserted before the edge to ``+W``.
    """
    height = width if height is None else height
    rng = np.random.default_rng(seed)
    r, c = np.divmod(np.arange(width * height, dtype=np.int64), width)
    lat = lat0 + r * step + rng.uniform(-jitter, jitter, size=r.shape)
    lon = lon0 + c * step + rng.uniform(-jitter, jitter, size=r.shape)
    idx = np.arange(width * height, dtype=np.int64)
    right = idx[c + 1 < width]
    down = idx[r + 1 < height]
    # interleave per vertex: (i, i+1) then (i, i+W)
    eu = np.concatenate((right, down))
    ev = np.concatenate((right + 1, down + width))
    key = np.concatenate((right * 2, down * 2 + 1))
    o = np.argsort(key, kind="stable")
    return from_edges(lat, lon, eu[o], ev[o], name=name or f"grid{width}x{height}")


def road_like(n_target: int, seed: int = 5, delete_frac: float = 0.15, mean_shape_points: float = 1.25,
              name: str | None = None) -> RoadGraph:
    """Road-like graph: jittered intersection grid, a fraction of the
    edges deleted, the rest subdivided with Poisson-many degree-2 shape points; only
    the largest component is kept.  About two thirds of the vertices have degree 2."""
    rng = np.random.default_rng(seed)
    # E ~ 2*W^2*(1-delete); N ~ W^2 + E*mean  =>  W from n_target
    per_cell = 1 + 2 * (1 - delete_frac) * mean_shape_points
    width = max(4, int(np.sqrt(n_target / per_cell)))
    base = jittered_grid(width, seed=seed + 1, step=4e-4, jitter=8e-5)
    eu, ev = base.eu.astype(np.int64), base.ev.astype(np.int64)
    keep = rng.random(eu.shape[0]) >= delete_frac
    eu, ev = eu[keep], ev[keep]
    k = rng.poisson(mean_shape_points, size=eu.shape[0])
    n0 = base.n
    total_new = int(k.sum())
    new_ids = n0 + np.arange(total_new, dtype=np.int64)
    first_new = n0 + np.concatenate(([0], np.cumsum(k)))[:-1]
    # coordinates of shape points: evenly spaced on the segment plus a small wiggle
    seg = np.repeat(np.arange(eu.shape[0]), k)
    pos = (np.arange(total_new) - np.repeat(first_new - n0, k) + 1) / (np.repeat(k, k) + 1)
    lat = np.concatenate((base.lat, base.lat[eu[seg]] * (1 - pos) + base.lat[ev[seg]] * pos
                          + rng.uniform(-2e-5, 2e-5, total_new)))
    lon = np.concatenate((base.lon, base.lon[eu[seg]] * (1 - pos) + base.lon[ev[seg]] * pos
                          + rng.uniform(-2e-5, 2e-5, total_new)))
    # chain edges u - s1 - s2 - ... - v
    a_list, b_list = [], []
    nz = k > 0
    a_list.append(eu[~nz]); b_list.append(ev[~nz])
    a_list.append(eu[nz]); b_list.append(first_new[nz])
    inner = np.ones(total_new, dtype=bool)
    last_of_seg = np.cumsum(k)[nz] - 1
    inner[last_of_seg] = False
    a_list.append(new_ids[inner]); b_list.append(new_ids[inner] + 1)
    a_list.append(new_ids[last_of_seg]); b_list.append(ev[nz])
    a = np.concatenate(a_list); b = np.concatenate(b_list)
    g = from_edges(lat, lon, a, b, name="tmp")
    lab = g.components()
    big = np.bincount(lab).argmax()
    keepv = lab == big
    remap = -np.ones(g.n, dtype=np.int64)
    remap[keepv] = np.arange(int(keepv.sum()))
    ke = keepv[a] & keepv[b]
    return from_edges(lat[keepv], lon[keepv], remap[a[ke]], remap[b[ke]],
                      name=name or f"roadlike{int(keepv.sum())}")


def bugtrap_grid(width: int, seed: int = 123, name: str | None = None):
    """Stand-in for the missing sanf.osm bug-trap map: jittered grid
    with a U-shaped band of deleted edges between source and target.  Returns
    (graph, source, target, pseudo_destinations)."""
    g = jittered_grid(width, seed=seed)
    W = width
    eu, ev = g.eu.astype(np.int64), g.ev.astype(np.int64)
    ru, cu = np.divmod(eu, W)
    rv, cv = np.divmod(ev, W)
    c0, c1 = int(0.35 * W), int(0.65 * W)
    r0, r1 = int(0.35 * W), int(0.65 * W)
    # walls: left column c0 (rows r0..r1), right column c1, bottom row r1; open at the top
    def crosses_vertical(col):
        return (ru == rv) & (np.minimum(cu, cv) == col) & (ru >= r0) & (ru <= r1)
    def crosses_horizontal(row):
        return (cu == cv) & (np.minimum(ru, rv) == row) & (cu > c0) & (cu <= c1)
    wall = crosses_vertical(c0) | crosses_vertical(c1) | crosses_horizontal(r1)
    keep = ~wall
    g2 = from_edges(g.lat, g.lon, eu[keep], ev[keep], ew=g.ew[keep], name=name or f"bugtrap{W}")
    source = (r0 + (r1 - r0) // 2) * W + (c0 + c1) // 2           # inside the trap
    target = min(W - 2, r1 + (W - r1) // 2) * W + (c0 + c1) // 2  # below the closed side
    pseudo = [
        (r0 - 3) * W + (c0 + c1) // 2,
        (r0 - 3) * W + max(1, c0 - 3),
        (r0 - 3) * W + min(W - 2, c1 + 4),
        ((r0 + r1) // 2) * W + max(1, c0 - 3),
        ((r0 + r1) // 2) * W + min(W - 2, c1 + 4),
    ]
    return g2, int(source), int(target), [int(p) for p in pseudo]


def hub_grid(width: int, seed: int = 3, hubs: int = 12, spokes: int = 9, name: str | None = None) -> RoadGraph:
    """Jittered grid plus a few high-degree hubs (extra edges to random vertices nearby): max
    degree up to 13, which exercises the widest child-row / adjacency-slot code paths."""
    g = jittered_grid(width, seed=seed)
    rng = np.random.default_rng(seed + 100)
    eu, ev = list(g.eu.astype(np.int64)), list(g.ev.astype(np.int64))
    have = set(zip(eu, ev)) | set(zip(ev, eu))
    deg = g.degree().copy()
    for h in rng.choice(g.n, size=hubs, replace=False):
        r, c = divmod(int(h), width)
        added = 0
        for _ in range(200):
            if added >= spokes or deg[h] >= 13:
                break
            rr = int(np.clip(r + rng.integers(-4, 5), 0, width - 1))
            cc = int(np.clip(c + rng.integers(-4, 5), 0, width - 1))
            v = rr * width + cc
            if v == h or (h, v) in have or deg[v] >= 12:
                continue
            eu.append(int(h)); ev.append(v); have.add((h, v)); have.add((v, h))
            deg[h] += 1; deg[v] += 1; added += 1
    return from_edges(g.lat, g.lon, np.asarray(eu), np.asarray(ev), name=name or f"hubgrid{width}")


def pick_destinations(g: RoadGraph, count: int, seed: int = 7, same_component: bool = True):
    """`count` distinct vertices with non-empty adjacency (largest component when asked)."""
    rng = np.random.default_rng(seed)
    ok = g.degree() > 0
    if same_component:
        lab = g.components()
        big = np.bincount(lab[ok]).argmax()
        ok &= lab == big
    cand = np.nonzero(ok)[0]
    return [int(v) for v in rng.choice(cand, size=count, replace=False)]
