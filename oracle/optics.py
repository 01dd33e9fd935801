"""ctypes view of the OPTICS oracle (oracle/optics_oracle.c in liborc.so) and an independent,
heap-free Python statement of the same ordering for small inputs.  TEST INFRASTRUCTURE ONLY.
The reference's optics.hpp cannot be compiled here (boost.geometry, FunctionalPlus and `geometry`
are not vendored): PARITY UNPINNED by the reference."""
import ctypes as C

import numpy as np


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def oracle(points, min_pts, eps):
    """-> (order uint32[n], reach float64[n] in ordering position; -1 = none)"""
    from . import orc
    L = orc.lib()
    L.orc_optics.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
    L.orc_optics.restype = C.c_int
    pts = np.ascontiguousarray(points, dtype=np.int32)
    n, d = pts.shape
    order = np.zeros(max(n, 1), np.uint32)
    reach = np.zeros(max(n, 1), np.float64)
    assert L.orc_optics(_p(pts), n, d, min_pts, float(eps), _p(order), _p(reach)) == 0
    return order[:n].copy(), reach[:n].copy()


def clusters(reach, threshold):
    from . import orc
    L = orc.lib()
    L.orc_optics_clusters.argtypes = [C.c_void_p, C.c_size_t, C.c_double, C.c_void_p]
    L.orc_optics_clusters.restype = C.c_size_t
    reach = np.ascontiguousarray(reach, dtype=np.float64)
    cl = np.zeros(max(len(reach), 1), np.uint32)
    nc = L.orc_optics_clusters(_p(reach), len(reach), float(threshold), _p(cl))
    return cl[:len(reach)].copy(), int(nc)


def plain(points, min_pts, eps):
    """the walk of optics.hpp:522-560 with a linear scan for the smallest (reachability, index) seed
    instead of a std::set -- O(n^2), small inputs only"""
    pts = np.asarray(points, dtype=np.int64)
    n = len(pts)
    sq = ((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1)
    nb = sq.astype(np.float64) <= eps * eps
    dist = np.sqrt(sq.astype(np.float64))
    reach = np.full(n, -1.0)
    done = np.zeros(n, bool)
    in_seeds = np.zeros(n, bool)
    order = []
    for start in range(n):
        if done[start]:
            continue
        p = start
        in_seeds[:] = False
        while True:
            done[p] = True
            in_seeds[p] = False
            order.append(p)
            idx = np.flatnonzero(nb[p])
            if len(idx) >= min_pts:
                core = np.sqrt(float(np.sort(sq[p, idx])[min_pts - 1]))
                for o in idx:
                    if done[o]:
                        continue
                    nr = max(core, dist[p, o])
                    if reach[o] < 0 or nr < reach[o]:
                        reach[o] = nr
                        in_seeds[o] = True
            cand = np.flatnonzero(in_seeds)
            if len(cand) == 0:
                break
            p = int(cand[np.lexsort((cand, reach[cand]))[0]])
    order = np.array(order, np.uint32)
    return order, reach[order]
