"""ctypes view of the CPU oracle (oracle/liborc.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs — never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liborc.so")

EVENT_DTYPE = np.dtype(
    [("x", "<u2"), ("y", "<u2"), ("p", "<i2"), ("_pad", "<u2"), ("t", "<i8")], align=True
)
assert EVENT_DTYPE.itemsize == 16

KEY_VOXEL, KEY_REF_HASH8192 = 0, 1


class DsParams(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("vx", C.c_int32), ("vy", C.c_int32),
        ("vt_us", C.c_int64), ("t0_us", C.c_int64), ("use_polarity", C.c_int32),
        ("keyfn", C.c_int32), ("algo", C.c_int32), ("count_repeated", C.c_int32),
    ]


class SynthParams(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("first_index", C.c_uint64), ("n_events", C.c_uint64),
        ("rate_eps", C.c_uint64), ("width", C.c_int32), ("height", C.c_int32),
        ("n_blobs", C.c_int32), ("sigma_q8", C.c_int32), ("noise_q16", C.c_int32),
        ("vmax_pps", C.c_int32),
    ]


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("evk_oracle.c", "evk_oracle.h", "aec_oracle.c", "dbscan_oracle.c",
                                             "optics_oracle.c")]
    if (not force and os.path.exists(_LIB)
            and all(os.path.getmtime(_LIB) >= os.path.getmtime(s) for s in src)):
        return _LIB
    subprocess.check_call(["make", "-C", _HERE, "-B", "liborc.so"], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = C.CDLL(_LIB)
        vp, sz = C.c_void_p, C.c_size_t
        L.orc_ref_process_coordinates.argtypes = [vp, C.c_int, vp, vp, vp]
        L.orc_ref_process_coordinates.restype = C.c_int
        L.orc_ref_analyze_coordinates.argtypes = [vp, C.c_int, vp, vp, vp]
        L.orc_ref_analyze_coordinates.restype = C.c_int
        L.orc_ref_kmeans_trip.argtypes = [vp] * 7
        L.orc_ref_kmeans_trip.restype = C.c_float
        L.orc_evt2_decode.argtypes = [vp, sz, vp, sz]
        L.orc_evt2_decode.restype = sz
        L.orc_evt2_encode.argtypes = [vp, sz, vp, sz]
        L.orc_evt2_encode.restype = sz
        L.orc_ts_corners.argtypes = [vp, sz, C.c_int, C.c_int, vp, C.c_int, vp]
        L.orc_ts_corners.restype = sz
        L.orc_evt3_decode.argtypes = [vp, sz, vp, sz]
        L.orc_evt3_decode.restype = sz
        L.orc_evt3_encode.argtypes = [vp, sz, vp, sz]
        L.orc_evt3_encode.restype = sz
        L.orc_downsample.argtypes = [vp, sz, C.POINTER(DsParams), vp, vp, C.POINTER(sz)]
        L.orc_downsample.restype = sz
        L.orc_downsample_mt.argtypes = [vp, sz, C.POINTER(DsParams), C.c_int, C.c_int, vp, vp,
                                        C.POINTER(sz)]
        L.orc_downsample_mt.restype = sz
        L.orc_points.argtypes = [vp, vp, sz, C.c_int, C.c_int64, C.c_float, C.c_float, vp]
        L.orc_points.restype = None
        L.orc_kmeans_assign.argtypes = [vp, sz, C.c_int, vp, C.c_int, C.c_float, C.c_int, vp]
        L.orc_kmeans_assign.restype = None
        L.orc_kmeans_update.argtypes = [vp, sz, C.c_int, vp, C.c_int, vp, vp, vp]
        L.orc_kmeans_update.restype = C.c_float
        L.orc_kmeans.argtypes = [vp, sz, C.c_int, vp, C.c_int, C.c_float, C.c_int, C.c_float,
                                 vp, vp]
        L.orc_kmeans.restype = C.c_int
        L.orc_kmeans_mt.argtypes = [vp, sz, C.c_int, vp, C.c_int, C.c_float, C.c_int, C.c_float,
                                    C.c_int, vp, vp]
        L.orc_kmeans_mt.restype = C.c_int
        L.orc_synth.argtypes = [C.POINTER(SynthParams), vp]
        L.orc_synth.restype = None
        L.orc_synth_mt.argtypes = [C.POINTER(SynthParams), vp, C.c_int]
        L.orc_synth_mt.restype = None
        L.orc_load_csv.argtypes = [C.c_char_p, vp, sz]
        L.orc_load_csv.restype = C.c_long
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def ds_params(width, height, vx=1, vy=1, vt_us=0, t0_us=0, use_polarity=0, keyfn=KEY_VOXEL,
              algo=0, count_repeated=1):
    return DsParams(width, height, vx, vy, vt_us, t0_us, use_polarity, keyfn, algo,
                    count_repeated)


def synth_params(seed, n_events, width, height, rate_eps, n_blobs, first_index=0,
                 sigma_q8=1536, noise_q16=16384, vmax_pps=200):
    return SynthParams(seed, first_index, n_events, rate_eps, width, height, n_blobs, sigma_q8,
                       noise_q16, vmax_pps)


def max_threads():
    return int(lib().orc_max_threads())


def synth(sp, threads=0):
    out = np.zeros(int(sp.n_events), dtype=EVENT_DTYPE)
    if threads and threads > 1:
        lib().orc_synth_mt(C.byref(sp), _p(out), threads)
    else:
        lib().orc_synth(C.byref(sp), _p(out))
    return out


def load_csv(path, cap=1 << 20):
    out = np.zeros(cap, dtype=EVENT_DTYPE)
    n = lib().orc_load_csv(path.encode(), _p(out), cap)
    if n < 0:
        raise IOError(path)
    return out[:n].copy()


def events_from_xy(x, y, t=None, p=None):
    n = len(x)
    ev = np.zeros(n, dtype=EVENT_DTYPE)
    ev["x"], ev["y"] = x, y
    if t is not None:
        ev["t"] = t
    if p is not None:
        ev["p"] = p
    return ev


def downsample(ev, p, threads=0, canonical=True):
    """-> keys[U] u64, first_idx[U] u32, n_repeated (canonical order: ascending first index)"""
    ev = np.ascontiguousarray(ev)
    n = len(ev)
    keys = np.zeros(max(n, 1), dtype=np.uint64)
    first = np.zeros(max(n, 1), dtype=np.uint32)
    rep = C.c_size_t(0)
    if threads and threads > 1:
        U = lib().orc_downsample_mt(_p(ev), n, C.byref(p), threads, int(canonical), _p(keys),
                                    _p(first), C.byref(rep))
    else:
        U = lib().orc_downsample(_p(ev), n, C.byref(p), _p(keys), _p(first), C.byref(rep))
    return keys[:U].copy(), first[:U].copy(), int(rep.value)


def points(ev, first_idx, D=2, t0_us=0, t_scale=1e-3, p_scale=1.0):
    ev = np.ascontiguousarray(ev)
    if first_idx is None:
        U, fp = len(ev), None
    else:
        first_idx = np.ascontiguousarray(first_idx, dtype=np.uint32)
        U, fp = len(first_idx), _p(first_idx)
    pts = np.zeros((U, D), dtype=np.float32)
    lib().orc_points(_p(ev), fp, U, D, t0_us, t_scale, p_scale, _p(pts))
    return pts


def kmeans_assign(pts, cent, max_dist=0.0, use_sqrt=False):
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    cent = np.ascontiguousarray(cent, dtype=np.float32)
    P, D = pts.shape
    labels = np.zeros(P, dtype=np.int32)
    lib().orc_kmeans_assign(_p(pts), P, D, _p(cent), cent.shape[0], max_dist, int(use_sqrt),
                            _p(labels))
    return labels


def kmeans_update(pts, labels, cent):
    """-> new centroids, counts, sums, shift"""
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    labels = np.ascontiguousarray(labels, dtype=np.int32)
    cent = np.array(cent, dtype=np.float32, copy=True)
    P, D = pts.shape
    K = cent.shape[0]
    counts = np.zeros(K, dtype=np.uint64)
    sums = np.zeros((K, D), dtype=np.float64)
    shift = lib().orc_kmeans_update(_p(pts), P, D, _p(labels), K, _p(cent), _p(counts), _p(sums))
    return cent, counts, sums, float(shift)


def kmeans(pts, cent, max_dist=0.0, iters=1, tol=-1.0, threads=0):
    """-> centroids, labels, counts, iterations"""
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    cent = np.array(cent, dtype=np.float32, copy=True)
    P, D = pts.shape
    K = cent.shape[0]
    labels = np.zeros(P, dtype=np.int32)
    counts = np.zeros(K, dtype=np.uint64)
    if threads and threads > 1:
        it = lib().orc_kmeans_mt(_p(pts), P, D, _p(cent), K, max_dist, iters, tol, threads,
                                 _p(labels), _p(counts))
    else:
        it = lib().orc_kmeans(_p(pts), P, D, _p(cent), K, max_dist, iters, tol, _p(labels),
                              _p(counts))
    return cent, labels, counts, int(it)


def ref_process_coordinates(coords, unique_count=0, repeated_count=0):
    """literal kernel: coords int32 [x0,y0,x1,y1,...] -> unique pairs, cumulative counters"""
    coords = np.ascontiguousarray(coords, dtype=np.int32)
    n = len(coords) // 2
    uniq = np.zeros(2 * max(n, 1), dtype=np.int32)
    rc, uc = C.c_int(repeated_count), C.c_int(unique_count)
    u = lib().orc_ref_process_coordinates(_p(coords), n, _p(uniq), C.byref(rc), C.byref(uc))
    return uniq[: 2 * u].reshape(-1, 2).copy(), uc.value, rc.value


def ref_analyze_coordinates(coords):
    coords = np.ascontiguousarray(coords, dtype=np.int32)
    n = len(coords)
    xs = np.zeros(max(n // 2, 1), dtype=np.int32)
    ys = np.zeros_like(xs)
    cs = np.zeros_like(xs)
    u = lib().orc_ref_analyze_coordinates(_p(coords), n, _p(xs), _p(ys), _p(cs))
    return xs[:u].copy(), ys[:u].copy(), cs[:u].copy()


def ref_kmeans_trip(data, centroids, output=None):
    """literal KERNEL_RESTART trip -> dict(centroids, new_centroids, assign, cluster_index,
    scalar_sum, output, error_max)"""
    data = np.ascontiguousarray(data, dtype=np.float32)
    cent = np.array(centroids, dtype=np.float32, copy=True)
    out = np.zeros(32768, dtype=np.float32) if output is None else np.array(
        output, dtype=np.float32, copy=True)
    assign = np.zeros(2048, dtype=np.int32)
    ci = np.zeros(8, dtype=np.int32)
    ss = np.zeros(32, dtype=np.float32)
    nc = np.zeros(16, dtype=np.float32)
    with np.errstate(all="ignore"):
        em = lib().orc_ref_kmeans_trip(_p(data), _p(cent), _p(out), _p(assign), _p(ci), _p(ss),
                                       _p(nc))
    return dict(centroids=cent, new_centroids=nc, assign=assign, cluster_index=ci,
                scalar_sum=ss, output=out, error_max=float(em))


def evt2_encode(ev):
    """events -> RAW EVT 2.0 words (uint32)"""
    ev = np.ascontiguousarray(ev, dtype=EVENT_DTYPE)
    words = np.empty(2 * len(ev) + 1, dtype=np.uint32)
    m = lib().orc_evt2_encode(_p(ev), len(ev), _p(words), len(words))
    if m == C.c_size_t(-1).value:
        raise ValueError("an event does not fit EVT 2.0 (x, y < 2048, 0 <= t < 2^34)")
    return words[:m].copy()


def evt2_decode(words):
    """RAW EVT 2.0 words -> CD events"""
    words = np.ascontiguousarray(words, dtype=np.uint32)
    out = np.zeros(len(words), dtype=EVENT_DTYPE)
    n = lib().orc_evt2_decode(_p(words), len(words), _p(out), len(out))
    return out[:n].copy()


def evt3_encode(ev):
    """events (time-ordered) -> RAW EVT 3.0 words (uint16)"""
    ev = np.ascontiguousarray(ev, dtype=EVENT_DTYPE)
    words = np.empty(5 * len(ev) + 4, dtype=np.uint16)
    m = lib().orc_evt3_encode(_p(ev), len(ev), _p(words), len(words))
    if m == C.c_size_t(-1).value:
        raise ValueError("events do not fit EVT 3.0 (x, y < 2048, time-ordered, gaps < 2^24 us)")
    assert m < len(words)
    return words[:m].copy()


def evt3_decode(words):
    """RAW EVT 3.0 words -> CD events"""
    words = np.ascontiguousarray(words, dtype=np.uint16)
    n = lib().orc_evt3_decode(_p(words), len(words), None, 0)
    out = np.zeros(n, dtype=EVENT_DTYPE)
    lib().orc_evt3_decode(_p(words), len(words), _p(out), n)
    return out


def ts_corners(ev, W, H, surface, literal_break=True):
    """one callback range of the reference's corner tracker: stamps `surface` (int64 [H, W], in
    place) with every event, then tests every event -> indices of the corner events"""
    ev = np.ascontiguousarray(ev, dtype=EVENT_DTYPE)
    assert surface.dtype == np.int64 and surface.shape == (H, W) and surface.flags.c_contiguous
    flags = np.zeros(max(len(ev), 1), np.uint8)
    n = lib().orc_ts_corners(_p(ev), len(ev), W, H, _p(surface), 1 if literal_break else 0, _p(flags))
    idx = np.flatnonzero(flags[:len(ev)]).astype(np.uint32)
    assert len(idx) == n
    return idx
