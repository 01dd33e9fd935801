"""ctypes views of DBSCAN on the CPU: the oracle's restatement
(oracle/dbscan_oracle.c in liborc.so) and the REFERENCE's own DBSCAN_simple.h compiled where it lies
(oracle/_ref/libref_dbscan.so, Makefile target ref_dbscan).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref", "libref_dbscan.so")
REFERENCE_ROOT = "/root/reference"


def ref_available():
    return os.path.exists(_REF)


def ref_build():
    if os.path.isdir(REFERENCE_ROOT):
        subprocess.check_call(["make", "-C", _HERE, "ref_dbscan"], stdout=subprocess.DEVNULL)
    return ref_available()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _xyz(points):
    pts = np.asarray(points, dtype=np.float32)
    if pts.shape[1] == 2:
        pts = np.concatenate([pts, np.zeros((len(pts), 1), np.float32)], axis=1)
    return np.ascontiguousarray(pts, dtype=np.float32)


def _split(sizes, members):
    out, o = [], 0
    for s in sizes:
        out.append(members[o:o + s].copy())
        o += s
    return out


def oracle(points, eps, min_pts, min_cluster=1, max_cluster=2**31 - 1):
    """-> (labels, clusters as sorted index arrays in output order, seeds)"""
    from . import orc
    L = orc.lib()
    xyz = _xyz(points)
    n = len(xyz)
    L.orc_dbscan.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p,
                             C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_long, C.c_void_p]
    labels = np.zeros(max(n, 1), np.int32)
    sizes, seeds = np.zeros(n + 1, np.int32), np.zeros(n + 1, np.int32)
    members = np.zeros(2 * n + 2, np.int32)
    nm = C.c_long(0)
    nc = L.orc_dbscan(_p(xyz), n, eps, min_pts, min_cluster, max_cluster, _p(labels), _p(sizes),
                      _p(seeds), n + 1, _p(members), len(members), C.byref(nm))
    return labels[:n].copy(), _split(sizes[:nc], members[:nm.value]), seeds[:nc].copy()


_ref = None


def reference(points, eps, min_pts, min_cluster=1, max_cluster=2**31 - 1):
    """the reference's DBSCANSimpleCluster::extract -> clusters as sorted index arrays, its order"""
    global _ref
    if _ref is None:
        _ref = C.CDLL(_REF)
        _ref.ref_dbscan_simple.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                                           C.c_void_p, C.c_int, C.c_void_p, C.c_long, C.c_void_p]
    xyz = _xyz(points)
    n = len(xyz)
    sizes, members = np.zeros(n + 1, np.int32), np.zeros(2 * n + 2, np.int32)
    nm = C.c_long(0)
    nc = _ref.ref_dbscan_simple(_p(xyz), n, eps, min_pts, min_cluster, max_cluster, _p(sizes), n + 1,
                                _p(members), len(members), C.byref(nm))
    return _split(sizes[:nc], members[:nm.value])


def contract(points, eps, min_pts, min_cluster=1, max_cluster=2**31 - 1):
    """The order-free statement of what the sequential algorithm computes (numpy, O(N^2) memory --
    small inputs only); this is the form the CUDA path implements:
      core(i)      = #{j : |p_j - p_i|^2 <= eps^2} >= min_pts (i itself counted)
      cluster      = connected component of core points under the eps relation; its seed = its
                     lowest-index core point; clusters are discovered in seed order
      border point = non-core with a core neighbour: member of the cluster with the LOWEST seed among
                     its core neighbours' clusters, and ALSO of every other cluster whose SEED POINT
                     is its neighbour (the seed's neighbours are queued whatever their state)
      kept         = clusters with min_cluster <= members <= max_cluster, largest first
                     (ties: lowest seed first -- the reference's std::sort leaves ties unspecified)
    -> (labels, clusters, seeds) like oracle()"""
    xyz = _xyz(points).astype(np.float64)
    n = len(xyz)
    d2 = ((xyz[:, None, :] - xyz[None, :, :]) ** 2).sum(-1)
    adj = d2 <= eps * eps
    core = adj.sum(1) >= min_pts
    root = np.arange(n)
    ci = np.flatnonzero(core)
    # components of core points: propagate the minimum index until stable
    lab = np.where(core, np.arange(n), n)
    while True:
        new = lab.copy()
        for i in ci:
            nb = ci[adj[i, ci]]
            new[i] = lab[nb].min()
        if (new == lab).all():
            break
        lab = new
    members = {int(s): set(np.flatnonzero(lab == s).tolist()) for s in np.unique(lab[core])}
    for b in np.flatnonzero(~core):
        nb = ci[adj[b, ci]]
        if len(nb) == 0:
            continue
        prim = int(lab[nb].min())
        members[prim].add(int(b))
        for j in nb:                       # neighbours that are seed points of other clusters
            if lab[j] == j and j != prim:
                members[int(j)].add(int(b))
    kept = [(len(m), s, np.array(sorted(m), np.int32)) for s, m in members.items()
            if min_cluster <= len(m) <= max_cluster]
    kept.sort(key=lambda t: (-t[0], t[1]))
    labels = np.full(n, -1, np.int32)
    best = np.full(n, 2**31 - 1, np.int64)
    for c, (_, s, m) in enumerate(kept):
        upd = m[best[m] > s]
        labels[upd] = c
        best[upd] = s
    return labels, [m for _, _, m in kept], np.array([s for _, s, _ in kept], np.int32)
