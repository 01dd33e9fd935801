"""ctypes views of the corner tracker's pieces on the CPU: the oracle's
restatements (liborc.so: orc_ts_corners, orc_filter_corners) and the REFERENCE's own code compiled
where it lies (oracle/_ref/libref_fct.so, Makefile target ref_fct: the event callback lambda,
CornerFilter::filterCorners and CornerTracker of
event-cam-tracking/event-cam-fast-corner-tracker/metavision_time_surface_periodic_group_track.cpp).
TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref", "libref_fct.so")
REFERENCE_ROOT = "/root/reference"
W, H = 1280, 720   # the reference's callback hard-codes the Gen4 frame in its border test


def ref_available():
    return os.path.exists(_REF)


def ref_build():
    if os.path.isdir(REFERENCE_ROOT):
        subprocess.check_call(["make", "-C", _HERE, "ref_fct"], stdout=subprocess.DEVNULL)
    return ref_available()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


_ref = None


def _lib():
    global _ref
    if _ref is None:
        L = C.CDLL(_REF)
        L.ref_fct_callback.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_long]
        L.ref_fct_callback.restype = C.c_long
        L.ref_fct_filter.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_long]
        L.ref_fct_filter.restype = C.c_long
        L.ref_fct_tracker_new.restype = C.c_void_p
        L.ref_fct_tracker_delete.argtypes = [C.c_void_p]
        L.ref_fct_tracker_update.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_long]
        L.ref_fct_tracker_update.restype = C.c_long
        _ref = L
    return _ref


def reference_callback(ev, surface, flag=1):
    """one callback range through the reference's lambda; surface [720, 1280] int64 is updated in
    place.  -> corner coordinates [n, 2] (x, y) in the order the reference pushes them"""
    ev = np.ascontiguousarray(ev)
    assert surface.shape == (H, W) and surface.dtype == np.int64 and surface.flags.c_contiguous
    out = np.zeros((max(1, len(ev)), 2), np.int32)
    b = ev.ctypes.data
    n = _lib().ref_fct_callback(b, b + ev.nbytes, _p(surface), flag, _p(out), len(out))
    return out[:n].copy()


def reference_filter(xy, width, height, box_size=15):
    """CornerFilter::filterCorners -> [m, 3] (x, y, label)"""
    xy = np.ascontiguousarray(xy, dtype=np.int32).reshape(-1, 2)
    out = np.zeros((max(1, len(xy)), 3), np.int32)
    m = _lib().ref_fct_filter(_p(xy), len(xy), width, height, box_size, _p(out), len(out))
    return out[:m].copy()


def oracle_filter(xy, width, height, box_size=15):
    """the oracle's restatement -> indices of the kept corners (their labels are 0, 1, 2, ...)"""
    from . import orc
    L = orc.lib()
    L.orc_filter_corners.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.orc_filter_corners.restype = C.c_size_t
    xy = np.ascontiguousarray(xy, dtype=np.int32).reshape(-1, 2)
    kept = np.zeros(max(1, len(xy)), np.uint32)
    m = L.orc_filter_corners(_p(xy), len(xy), width, height, box_size, _p(kept))
    return kept[:m].copy()


class ReferenceTracker:
    """the reference's CornerTracker with the app's parameters (FCT:805-813)"""

    def __init__(self):
        self._t = _lib().ref_fct_tracker_new()

    def update(self, xy):
        xy = np.ascontiguousarray(xy, dtype=np.int32).reshape(-1, 2)
        out = np.zeros((4096, 8), np.float32)
        n = _lib().ref_fct_tracker_update(self._t, _p(xy), len(xy), _p(out), len(out))
        return out[:n].copy()

    def __del__(self):
        try:
            _lib().ref_fct_tracker_delete(self._t)
        except Exception:
            pass
