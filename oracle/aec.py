"""ctypes views of the asynchronous-event-clustering consumer on the CPU:
  Oracle     oracle/aec_oracle.c in liborc.so      -- the restatement (the checker)
  Reference  oracle/_ref/libref_aec.so             -- the REFERENCE's own AEClustering.cpp /
             MyCluster.cpp compiled where they lie (oracle/Makefile target ref_aec)
Both expose the same methods so tests can run them side by side.  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref", "libref_aec.so")
REFERENCE_ROOT = "/root/reference"
DEFAULTS = dict(sz_buffer=800, radius=40.0, kappa=0, alpha=0.5, min_n=10)  # AEClustering.cpp:7-18


def ref_available():
    return os.path.exists(_REF)


def ref_build():
    if os.path.isdir(REFERENCE_ROOT):
        subprocess.check_call(["make", "-C", _HERE, "ref_aec"], stdout=subprocess.DEVNULL)
    return ref_available()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class _Base:
    """state accessors shared by both libraries (prefix = 'orc_aec' or 'ref_aec')"""

    def _bind(self, L, prefix):
        vp, i32, dbl = C.c_void_p, C.c_int, C.c_double
        self._L = L
        f = lambda n: getattr(L, prefix + "_" + n)
        f("new").argtypes = [i32, i32, dbl, i32, dbl, i32, C.c_uint]
        f("new").restype = vp
        f("free").argtypes = [vp]
        f("free").restype = None
        f("update").argtypes = [vp, vp, C.c_long]
        f("update").restype = None
        f("n_clusters").argtypes = [vp]
        f("last_updated").argtypes = [vp]
        f("get_clusters").argtypes = [vp, vp, vp, vp, vp]
        f("get_clusters").restype = None
        f("get_points").argtypes = [vp, i32, vp, vp, vp, vp, i32]
        self._f = f

    def __init__(self, init=None, rand_seed=1):
        """init=None: default-constructed, init() never called (what the reference app does,
        store.cpp:42); init=dict(sz_buffer, radius, kappa, alpha, min_n): AEClustering::init"""
        d = dict(DEFAULTS)
        if init is not None:
            d.update(init)
        self.params = d if init is not None else dict(DEFAULTS)
        self._h = self._f("new")(0 if init is None else 1, d["sz_buffer"], d["radius"],
                                 d["kappa"], d["alpha"], d["min_n"], rand_seed)

    def close(self):
        if self._h:
            self._f("free")(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def update(self, e):
        """e: (n, 4) float64 rows {t, x, y, p} -> n calls of AEClustering::update"""
        e = np.ascontiguousarray(e, dtype=np.float64).reshape(-1, 4)
        self._f("update")(self._h, _p(e), len(e))

    def last_updated(self):
        return self._f("last_updated")(self._h)

    def clusters(self):
        """(ids, n, mu[n,2], centroid[n,2]) in cluster order"""
        n = self._f("n_clusters")(self._h)
        ids, ns = np.zeros(n, np.int32), np.zeros(n, np.int32)
        mu, cen = np.zeros((n, 2)), np.zeros((n, 2))
        self._f("get_clusters")(self._h, _p(ids), _p(ns), _p(mu), _p(cen))
        return ids, ns, mu, cen

    def points(self, c, cap=1 << 16):
        """(event ids, xy[n,2], t, pol) of cluster c, front to back"""
        ids, xy = np.zeros(cap, np.int32), np.zeros((cap, 2))
        t, pol = np.zeros(cap), np.zeros(cap, np.int32)
        n = self._f("get_points")(self._h, c, _p(ids), _p(xy), _p(t), _p(pol), cap)
        assert n <= cap
        return ids[:n].copy(), xy[:n].copy(), t[:n].copy(), pol[:n].copy()

    def state(self):
        """everything observable, for state-for-state comparisons"""
        ids, ns, mu, cen = self.clusters()
        pts = [self.points(c) for c in range(len(ids))]
        return dict(ids=ids, n=ns, mu=mu, cen=cen, pts=pts, last=self.last_updated())


class Oracle(_Base):
    def __init__(self, init=None, rand_seed=1):
        from . import orc
        L = orc.lib()
        self._bind(L, "orc_aec")
        L.orc_aec_report.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        super().__init__(init, rand_seed)
        self.centroid_prev = np.zeros((16384, 2))   # store.cpp:188-193

    def coverage(self):
        """(merges, largest merged cluster, rand draws) so far -- tests assert the hard paths ran"""
        out = (C.c_long * 3)()
        self._L.orc_aec_coverage.argtypes = [C.c_void_p, C.c_void_p]
        self._L.orc_aec_coverage.restype = None
        self._L.orc_aec_coverage(self._h, out)
        return tuple(out)

    def report(self):
        """per-slice report (store.cpp:461-521): rows {id, n, cen_x, cen_y, prev_x, prev_y,
        has_arrow, end_x, end_y}"""
        rec = np.zeros((4096, 9))
        n = self._L.orc_aec_report(self._h, _p(self.centroid_prev), len(self.centroid_prev),
                                   _p(rec), len(rec))
        assert 0 <= n <= len(rec)
        return rec[:n].copy()


class Reference(_Base):
    _lib = None

    def __init__(self, init=None, rand_seed=1):
        if Reference._lib is None:
            Reference._lib = C.CDLL(_REF)
        self._bind(Reference._lib, "ref_aec")
        super().__init__(init, rand_seed)


def handoff(unique_coords, unique_count_diff, unique_count):
    """the hand-off loop as written (store.cpp:435-445) -> (n, 4) events {t, x, y, 0}"""
    from . import orc
    L = orc.lib()
    L.orc_aec_handoff.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.orc_aec_handoff.restype = C.c_long
    uc = np.ascontiguousarray(unique_coords, dtype=np.int32)
    assert len(uc) >= unique_count_diff + 2
    out = np.zeros((max(1, (unique_count_diff + 3) // 4), 4))
    n = L.orc_aec_handoff(_p(uc), unique_count_diff, unique_count, _p(out))
    return out[:n].copy()


def glibc_rand_seq(seed, n):
    from . import orc
    L = orc.lib()
    L.orc_glibc_rand_seq.argtypes = [C.c_uint, C.c_void_p, C.c_int]
    L.orc_glibc_rand_seq.restype = None
    out = np.zeros(n, np.int32)
    L.orc_glibc_rand_seq(seed, _p(out), n)
    return out
