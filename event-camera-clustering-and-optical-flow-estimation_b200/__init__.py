"""evk_b200 — Python view (ctypes) of the C-ABI library libevk.so (include/evk.h).

The product is the CUDA library; this module only binds it for the tests and bench.py.  The
directory name is not an importable identifier, so load it with `evk_loader.load()` from the
repo root (registers it as module `evk_b200`).  There is no CPU fallback: if libevk.so is missing
or no CUDA device is present every call fails loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EVK_LIB") or os.path.join(_HERE, "libevk.so")  # EVK_LIB: kernel A/B builds

EVENT_DTYPE = np.dtype(
    [("x", "<u2"), ("y", "<u2"), ("p", "<i2"), ("_pad", "<u2"), ("t", "<i8")], align=True
)
assert EVENT_DTYPE.itemsize == 16

KEY_VOXEL, KEY_REF_HASH8192 = 0, 1
ALGO_AUTO, ALGO_TABLE, ALGO_SORT, ALGO_SLAB, ALGO_PARTITION = 0, 1, 2, 3, 4
OWNER_TIME_RANGE, OWNER_MIX64 = 0, 1
STATUS = {0: "EVK_OK", -1: "EVK_ERR_INVALID", -2: "EVK_ERR_CUDA", -3: "EVK_ERR_NOMEM",
          -4: "EVK_ERR_STATE", -5: "EVK_ERR_CAPACITY", -6: "EVK_ERR_IO", -7: "EVK_ERR_COMM"}


class EvkError(RuntimeError):
    def __init__(self, status, msg=""):
        self.status = status
        super().__init__(f"{STATUS.get(status, status)}: {msg}")


class DsParams(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("vx", C.c_int32), ("vy", C.c_int32),
        ("vt_us", C.c_int64), ("t0_us", C.c_int64), ("use_polarity", C.c_int32),
        ("keyfn", C.c_int32), ("algo", C.c_int32), ("count_repeated", C.c_int32),
    ]


class KmParams(C.Structure):
    _fields_ = [
        ("K", C.c_int32), ("D", C.c_int32), ("max_dist", C.c_float), ("iters", C.c_int32),
        ("tol", C.c_float), ("t_scale", C.c_float), ("p_scale", C.c_float),
        ("on_events", C.c_int32),
    ]


class SynthParams(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("first_index", C.c_uint64), ("n_events", C.c_uint64),
        ("rate_eps", C.c_uint64), ("width", C.c_int32), ("height", C.c_int32),
        ("n_blobs", C.c_int32), ("sigma_q8", C.c_int32), ("noise_q16", C.c_int32),
        ("vmax_pps", C.c_int32),
    ]


class AecParams(C.Structure):
    _fields_ = [
        ("use_init", C.c_int32), ("sz_buffer", C.c_int32), ("radius", C.c_double),
        ("kappa", C.c_int32), ("min_n", C.c_int32), ("alpha", C.c_double),
        ("rand_seed", C.c_uint32), ("max_clusters", C.c_int32), ("max_points", C.c_int32),
        ("_pad", C.c_int32),
    ]


CORNER_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("label", "<i4")])
AEC_CLUSTER_DTYPE = np.dtype([("id", "<i4"), ("n", "<i4"), ("mu", "<f8", (2,)),
                              ("centroid", "<f8", (2,))], align=True)
AEC_FLOW_DTYPE = np.dtype([("id", "<i4"), ("n", "<i4"), ("centroid", "<f8", (2,)),
                           ("prev", "<f8", (2,)), ("has_arrow", "<i4"), ("_pad", "<i4"),
                           ("arrow_end", "<f8", (2,))], align=True)
assert AEC_CLUSTER_DTYPE.itemsize == 40 and AEC_FLOW_DTYPE.itemsize == 64


class DbscanParams(C.Structure):
    _fields_ = [("eps", C.c_double), ("min_pts", C.c_int32), ("min_cluster", C.c_int32),
                ("max_cluster", C.c_int32), ("D", C.c_int32), ("t_scale", C.c_double),
                ("t0_us", C.c_int64)]


class StageTimes(C.Structure):
    _fields_ = [
        ("ds_total_ms", C.c_float), ("ds_main_ms", C.c_float), ("ds_compact_ms", C.c_float),
        ("km_total_ms", C.c_float), ("km_assign_ms", C.c_float), ("ds_algo_used", C.c_int32),
        ("km_iters", C.c_int32), ("ds_launches", C.c_int32), ("km_launches", C.c_int32),
    ]


# every symbol include/evk.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "evk_create", "evk_destroy", "evk_last_error", "evk_version", "evk_load_events",
    "evk_append_events", "evk_load_events_soa", "evk_load_coords_i32", "evk_load_csv", "evk_load_evt2", "evk_load_evt3", "evk_load_raw", "evk_synth",
    "evk_num_events", "evk_get_events", "evk_downsample", "evk_get_voxels", "evk_num_voxels", "evk_set_centroids",
    "evk_init_centroids_first_k", "evk_kmeans", "evk_downsample_kmeans", "evk_get_labels", "evk_get_centroids",
    "evk_window_config", "evk_window_config_events", "evk_window_push", "evk_window_flush", "evk_set_profiling",
    "evk_get_stage_times", "evk_timer_start", "evk_timer_stop", "evk_sync", "evk_flush_l2",
    "evk_comm_unique_id", "evk_comm_init", "evk_comm_destroy", "evk_set_shard",
    "evk_downsample_sharded", "evk_kmeans_sharded", "evk_init_centroids_first_k_sharded",
    "evk_downsample_kmeans_sharded", "evk_downsample_kmeans_submit", "evk_downsample_kmeans_wait",
    "evk_downsample_kmeans_sharded_submit", "evk_downsample_kmeans_sharded_wait",
    "evk_aec_create", "evk_aec_destroy", "evk_aec_update", "evk_aec_update_voxels",
    "evk_aec_get_clusters", "evk_aec_get_points", "evk_aec_report",
    "evk_dbscan_points", "evk_dbscan_voxels", "evk_dbscan_get", "evk_dbscan_destroy",
    "evk_ts_create", "evk_ts_destroy", "evk_ts_corners", "evk_ts_get_corners", "evk_ts_get_surface",
    "evk_ts_filter_corners", "evk_filter_corners", "evk_get_filtered_corners",
    "evk_filter_corners_destroy",
    "evk_optics_points", "evk_optics_voxels", "evk_optics_get", "evk_optics_clusters",
    "evk_optics_destroy",
]

_lib = None


def lib():
    """Loads libevk.so.  Raises if it has not been built — there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EvkError(-2, f"{LIB_PATH} is missing: run __graft_entry__.build() "
                           "(nvcc, sm_100a); this package has no CPU or PyTorch fallback")
    L = C.CDLL(LIB_PATH)
    vp, sz, i32 = C.c_void_p, C.c_size_t, C.c_int
    psz = C.POINTER(sz)
    sig = {
        "evk_create": [C.POINTER(vp), i32, sz],
        "evk_destroy": [vp],
        "evk_load_events": [vp, vp, vp],
        "evk_append_events": [vp, vp, vp],
        "evk_load_events_soa": [vp, vp, vp, vp, vp, sz],
        "evk_load_coords_i32": [vp, vp, sz],
        "evk_load_csv": [vp, C.c_char_p],
        "evk_load_evt2": [vp, vp, sz, psz],
        "evk_load_evt3": [vp, vp, sz, psz],
        "evk_load_raw": [vp, C.c_char_p, psz],
        "evk_synth": [vp, C.POINTER(SynthParams)],
        "evk_num_events": [vp, psz],
        "evk_get_events": [vp, vp, sz, sz],
        "evk_downsample": [vp, C.POINTER(DsParams), psz, psz],
        "evk_get_voxels": [vp, vp, vp, vp, sz],
        "evk_num_voxels": [vp, psz, psz],
        "evk_set_centroids": [vp, vp, i32, i32],
        "evk_init_centroids_first_k": [vp, C.POINTER(KmParams)],
        "evk_kmeans": [vp, C.POINTER(KmParams), C.POINTER(i32)],
        "evk_downsample_kmeans": [vp, C.POINTER(DsParams), C.POINTER(KmParams), i32, psz, psz,
                                  C.POINTER(i32)],
        "evk_downsample_kmeans_submit": [vp, C.POINTER(DsParams), C.POINTER(KmParams), i32],
        "evk_downsample_kmeans_wait": [vp, psz, psz, C.POINTER(i32)],
        "evk_get_labels": [vp, vp, sz],
        "evk_downsample_kmeans_sharded_submit": [vp, C.POINTER(DsParams), C.POINTER(KmParams), i32,
                                                 i32],
        "evk_downsample_kmeans_sharded_wait": [vp, psz, psz, C.POINTER(i32)],
        "evk_aec_create": [vp, C.POINTER(AecParams)],
        "evk_aec_destroy": [vp],
        "evk_aec_update": [vp, vp, sz],
        "evk_aec_update_voxels": [vp, C.c_double, sz, sz, sz],
        "evk_aec_get_clusters": [vp, vp, sz, psz, C.POINTER(i32)],
        "evk_aec_get_points": [vp, sz, vp, vp, vp, vp, sz, psz],
        "evk_aec_report": [vp, vp, sz, psz],
        "evk_dbscan_points": [vp, vp, sz, C.POINTER(DbscanParams), psz, psz],
        "evk_dbscan_voxels": [vp, C.POINTER(DbscanParams), psz, psz],
        "evk_dbscan_get": [vp, vp, sz, vp, vp, sz, vp, sz],
        "evk_dbscan_destroy": [vp],
        "evk_ts_create": [vp, i32, i32],
        "evk_ts_destroy": [vp],
        "evk_ts_corners": [vp, i32, psz],
        "evk_ts_get_corners": [vp, vp, sz],
        "evk_ts_get_surface": [vp, vp, sz],
        "evk_ts_filter_corners": [vp, i32, psz],
        "evk_filter_corners": [vp, vp, sz, i32, i32, i32, psz],
        "evk_get_filtered_corners": [vp, vp, sz],
        "evk_filter_corners_destroy": [vp],
        "evk_optics_points": [vp, vp, sz, i32, i32, C.c_double],
        "evk_optics_voxels": [vp, i32, C.c_double],
        "evk_optics_get": [vp, vp, vp, sz, psz],
        "evk_optics_clusters": [vp, C.c_double, vp, sz, psz],
        "evk_optics_destroy": [vp],
        "evk_get_centroids": [vp, vp, vp],
        "evk_window_config": [vp, C.POINTER(DsParams), C.POINTER(KmParams), C.c_int64],
        "evk_window_config_events": [vp, C.POINTER(DsParams), C.POINTER(KmParams), sz],
        "evk_window_push": [vp, vp, vp, C.POINTER(i32)],
        "evk_window_flush": [vp, C.POINTER(i32)],
        "evk_set_profiling": [vp, i32],
        "evk_get_stage_times": [vp, C.POINTER(StageTimes)],
        "evk_timer_start": [vp],
        "evk_timer_stop": [vp, C.POINTER(C.c_float)],
        "evk_sync": [vp],
        "evk_flush_l2": [vp],
        "evk_comm_unique_id": [vp],
        "evk_comm_init": [vp, i32, i32, vp],
        "evk_comm_destroy": [vp],
        "evk_set_shard": [vp, C.c_uint64],
        "evk_downsample_sharded": [vp, C.POINTER(DsParams), i32, psz, psz],
        "evk_kmeans_sharded": [vp, C.POINTER(KmParams), C.POINTER(i32)],
        "evk_init_centroids_first_k_sharded": [vp, C.POINTER(KmParams)],
        "evk_downsample_kmeans_sharded": [vp, C.POINTER(DsParams), C.POINTER(KmParams), i32, i32,
                                          psz, psz, C.POINTER(i32)],
    }
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = i32
    L.evk_last_error.argtypes = [vp]
    L.evk_last_error.restype = C.c_char_p
    L.evk_version.argtypes = []
    L.evk_version.restype = C.c_char_p
    _lib = L
    return L


def ds_params(width, height, vx=1, vy=1, vt_us=0, t0_us=0, use_polarity=0, keyfn=KEY_VOXEL,
              algo=ALGO_AUTO, count_repeated=1):
    return DsParams(width, height, vx, vy, vt_us, t0_us, use_polarity, keyfn, algo,
                    count_repeated)


def km_params(K, D=2, max_dist=0.0, iters=1, tol=-1.0, t_scale=1e-3, p_scale=1.0, on_events=0):
    return KmParams(K, D, max_dist, iters, tol, t_scale, p_scale, on_events)


def synth_params(seed, n_events, width, height, rate_eps, n_blobs, first_index=0,
                 sigma_q8=1536, noise_q16=16384, vmax_pps=200):
    return SynthParams(seed, first_index, n_events, rate_eps, width, height, n_blobs, sigma_q8,
                       noise_q16, vmax_pps)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Evk:
    """One handle = one GPU stream of work, used from one thread (as the reference's single SDK
    thread does pack -> launch -> wait -> consume, ACCEL/store.cpp:370-615).  Method names follow
    the reference's call order: load events -> downsample -> cluster -> labels / centroids."""

    def __init__(self, max_events, device=0):
        self._L = lib()
        self._h = C.c_void_p()
        st = self._L.evk_create(C.byref(self._h), device, max_events)
        if st != 0:
            raise EvkError(st, "evk_create failed (no CUDA device? there is no CPU fallback)")
        self.max_events = max_events

    def _ck(self, st):
        if st != 0:
            raise EvkError(st, self._L.evk_last_error(self._h).decode())

    def close(self):
        if self._h:
            self._L.evk_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- load events
    def load_events(self, ev, append=False):
        ev = np.ascontiguousarray(ev, dtype=EVENT_DTYPE)
        b = ev.ctypes.data
        fn = self._L.evk_append_events if append else self._L.evk_load_events
        self._ck(fn(self._h, b, b + ev.nbytes))
        self._ck(self._L.evk_sync(self._h))  # ev is borrowed only for the duration of the call

    def load_events_ptr(self, addr, n, append=False):
        """pinned host buffer at `addr` holding n records; asynchronous (caller keeps it alive)"""
        fn = self._L.evk_append_events if append else self._L.evk_load_events
        self._ck(fn(self._h, addr, addr + 16 * n))

    def load_events_soa(self, x, y, t=None, p=None):
        x = np.ascontiguousarray(x, dtype=np.uint16)
        y = np.ascontiguousarray(y, dtype=np.uint16)
        t = None if t is None else np.ascontiguousarray(t, dtype=np.int64)
        p = None if p is None else np.ascontiguousarray(p, dtype=np.uint8)
        self._ck(self._L.evk_load_events_soa(self._h, _p(x), _p(y), None if t is None else _p(t),
                                             None if p is None else _p(p), len(x)))
        self._ck(self._L.evk_sync(self._h))

    def load_coords_i32(self, xy):
        xy = np.ascontiguousarray(xy, dtype=np.int32)
        self._ck(self._L.evk_load_coords_i32(self._h, _p(xy), xy.size // 2))
        self._ck(self._L.evk_sync(self._h))

    def load_csv(self, path):
        self._ck(self._L.evk_load_csv(self._h, path.encode()))

    def load_evt2(self, words):
        """RAW EVT 2.0 words (uint32 array). Returns the number of CD events decoded."""
        words = np.ascontiguousarray(words, dtype=np.uint32)
        n = C.c_size_t(0)
        self._ck(self._L.evk_load_evt2(self._h, _p(words), len(words), C.byref(n)))
        return n.value

    def load_evt3(self, words):
        """RAW EVT 3.0 words (uint16) -> events on the device; returns the number of CD events"""
        words = np.ascontiguousarray(words, dtype=np.uint16)
        n = C.c_size_t(0)
        self._ck(self._L.evk_load_evt3(self._h, _p(words), len(words), C.byref(n)))
        return n.value

    def load_evt2_ptr(self, ptr, n_words):
        n = C.c_size_t(0)
        self._ck(self._L.evk_load_evt2(self._h, C.c_void_p(ptr), n_words, C.byref(n)))
        return n.value

    def load_raw(self, path):
        n = C.c_size_t(0)
        self._ck(self._L.evk_load_raw(self._h, path.encode(), C.byref(n)))
        return n.value

    def synth(self, sp):
        self._ck(self._L.evk_synth(self._h, C.byref(sp)))

    @property
    def num_events(self):
        n = C.c_size_t(0)
        self._ck(self._L.evk_num_events(self._h, C.byref(n)))
        return n.value

    def get_events(self, first=0, count=None):
        count = self.num_events - first if count is None else count
        out = np.zeros(count, dtype=EVENT_DTYPE)
        self._ck(self._L.evk_get_events(self._h, _p(out), first, count))
        return out

    # ---- downsample
    def downsample(self, p):
        u, r = C.c_size_t(0), C.c_size_t(0)
        self._ck(self._L.evk_downsample(self._h, C.byref(p), C.byref(u), C.byref(r)))
        self.n_unique = u.value
        return u.value, r.value

    def num_voxels(self):
        """(n_unique, n_repeated) of the current voxel shard"""
        u, r = C.c_size_t(0), C.c_size_t(0)
        self._ck(self._L.evk_num_voxels(self._h, C.byref(u), C.byref(r)))
        return u.value, r.value

    def get_voxels(self, keys=True, reps=True, first=True):
        n = self.n_unique = self.num_voxels()[0]
        k = np.zeros(n, dtype=np.uint64) if keys else None
        r = np.zeros(n, dtype=EVENT_DTYPE) if reps else None
        f = np.zeros(n, dtype=np.uint32) if first else None
        self._ck(self._L.evk_get_voxels(self._h, None if k is None else _p(k),
                                        None if r is None else _p(r),
                                        None if f is None else _p(f), n))
        return k, r, f

    def get_voxels_ptr(self, keys_addr, reps_addr, first_addr, cap):
        """evk_get_voxels into caller-owned (pinned) host buffers given by address (0 = skip)"""
        self._ck(self._L.evk_get_voxels(self._h, C.c_void_p(keys_addr or None),
                                        C.c_void_p(reps_addr or None),
                                        C.c_void_p(first_addr or None), cap))

    def get_labels_ptr(self, addr, cap):
        self._ck(self._L.evk_get_labels(self._h, C.c_void_p(addr), cap))

    # ---- cluster
    def set_centroids(self, c):
        c = np.ascontiguousarray(c, dtype=np.float32)
        self._ck(self._L.evk_set_centroids(self._h, _p(c), c.shape[0], c.shape[1]))

    def init_centroids_first_k(self, km):
        self._ck(self._L.evk_init_centroids_first_k(self._h, C.byref(km)))

    def kmeans(self, km):
        it = C.c_int(0)
        self._ck(self._L.evk_kmeans(self._h, C.byref(km), C.byref(it)))
        self._km = km
        return it.value

    def downsample_kmeans(self, ds, km, init_first_k=True):
        """Fused step (evk_downsample_kmeans). Returns (n_unique, n_repeated, iters_done)."""
        u, r, it = C.c_size_t(0), C.c_size_t(0), C.c_int(0)
        self._ck(self._L.evk_downsample_kmeans(self._h, C.byref(ds), C.byref(km),
                                               1 if init_first_k else 0, C.byref(u), C.byref(r),
                                               C.byref(it)))
        self.n_unique = u.value
        self._km = km
        return u.value, r.value, it.value

    def downsample_kmeans_submit(self, ds, km, init_first_k=True):
        """Queue the fused step on the handle's stream (evk_downsample_kmeans_submit)."""
        self._ck(self._L.evk_downsample_kmeans_submit(self._h, C.byref(ds), C.byref(km),
                                                      1 if init_first_k else 0))
        self._km = km

    def downsample_kmeans_wait(self):
        """The step's one synchronisation. Returns (n_unique, n_repeated, iters_done)."""
        u, r, it = C.c_size_t(0), C.c_size_t(0), C.c_int(0)
        self._ck(self._L.evk_downsample_kmeans_wait(self._h, C.byref(u), C.byref(r), C.byref(it)))
        self.n_unique = u.value
        return u.value, r.value, it.value

    # ---- consumer: asynchronous event clustering (evk_aec_*) ---------------------------------
    def aec_create(self, init=None, rand_seed=1, max_clusters=0, max_points=0):
        """init=None: the default-constructed AEClustering of the reference app; else a dict
        with sz_buffer, radius, kappa, alpha, min_n (AEClustering::init)"""
        p = AecParams()
        p.use_init = 0 if init is None else 1
        if init is not None:
            p.sz_buffer, p.radius, p.kappa = init["sz_buffer"], init["radius"], init["kappa"]
            p.alpha, p.min_n = init["alpha"], init["min_n"]
        p.rand_seed, p.max_clusters, p.max_points = rand_seed, max_clusters, max_points
        self._ck(self._L.evk_aec_create(self._h, C.byref(p)))

    def aec_update(self, e):
        """e: (n, 4) float64 rows {t, x, y, p}"""
        e = np.ascontiguousarray(e, dtype=np.float64).reshape(-1, 4)
        self._ck(self._L.evk_aec_update(self._h, _p(e), len(e)))

    def aec_update_voxels(self, t, start=0, step=1, count=None):
        if count is None:
            nu = getattr(self, "n_unique", 0)
            count = (nu - start + step - 1) // step if nu > start else 0
        self._ck(self._L.evk_aec_update_voxels(self._h, float(t), start, step, count))
        return count

    def aec_clusters(self):
        """(records of AEC_CLUSTER_DTYPE in list order, last updated cluster index)"""
        n, last = C.c_size_t(0), C.c_int(0)
        self._ck(self._L.evk_aec_get_clusters(self._h, None, 0, C.byref(n), C.byref(last)))
        out = np.zeros(n.value, AEC_CLUSTER_DTYPE)
        if n.value:
            self._ck(self._L.evk_aec_get_clusters(self._h, _p(out), len(out), C.byref(n),
                                                  C.byref(last)))
        return out, last.value

    def aec_points(self, c):
        """(event ids, xy[n,2], t, pol) of cluster c, oldest first"""
        n = C.c_size_t(0)
        self._ck(self._L.evk_aec_get_points(self._h, c, None, None, None, None, 0, C.byref(n)))
        k = n.value
        ids, xy, t, pol = np.zeros(k, np.int32), np.zeros((k, 2)), np.zeros(k), np.zeros(k, np.uint8)
        if k:
            self._ck(self._L.evk_aec_get_points(self._h, c, _p(ids), _p(xy), _p(t), _p(pol), k,
                                                C.byref(n)))
        return ids, xy, t, pol.astype(np.int32)

    def aec_state(self):
        """everything observable, shaped like oracle.aec's state()"""
        cl, last = self.aec_clusters()
        pts = [self.aec_points(c) for c in range(len(cl))]
        return dict(ids=cl["id"].copy(), n=cl["n"].copy(), mu=cl["mu"].copy(),
                    cen=cl["centroid"].copy(), pts=pts, last=last)

    def aec_report(self):
        """per-slice report: records of AEC_FLOW_DTYPE (clusters with n >= min_n)"""
        out = np.zeros(1024, AEC_FLOW_DTYPE)
        n = C.c_size_t(0)
        self._ck(self._L.evk_aec_report(self._h, _p(out), len(out), C.byref(n)))
        return out[: n.value].copy()

    # ---- DBSCAN (evk_dbscan_*) ----------------------------------------------------------------
    def _dbscan_results(self, n_points, nc, ne):
        labels = np.zeros(n_points, np.int32)
        sizes, seeds = np.zeros(nc, np.uint32), np.zeros(nc, np.uint32)
        extra = np.zeros((ne, 2), np.uint32)
        self._ck(self._L.evk_dbscan_get(self._h, _p(labels), n_points, _p(sizes), _p(seeds), nc,
                                        _p(extra), ne))
        return labels, sizes, seeds, extra

    def dbscan_points(self, points, eps, min_pts, min_cluster=1, max_cluster=2**31 - 1):
        """points (n, 2) or (n, 3) -> (labels, sizes, seeds, extra pairs (point, cluster))"""
        pts = np.asarray(points, dtype=np.float32)
        if pts.shape[1] == 2:
            pts = np.concatenate([pts, np.zeros((len(pts), 1), np.float32)], axis=1)
        pts = np.ascontiguousarray(pts)
        p = DbscanParams(eps, min_pts, min_cluster, max_cluster, 3, 0.0, 0)
        nc, ne = C.c_size_t(0), C.c_size_t(0)
        self._ck(self._L.evk_dbscan_points(self._h, _p(pts), len(pts), C.byref(p), C.byref(nc),
                                           C.byref(ne)))
        return self._dbscan_results(len(pts), nc.value, ne.value)

    def dbscan_voxels(self, eps, min_pts, min_cluster=1, max_cluster=2**31 - 1, D=2, t_scale=0.0,
                      t0_us=0):
        p = DbscanParams(eps, min_pts, min_cluster, max_cluster, D, t_scale, t0_us)
        nc, ne = C.c_size_t(0), C.c_size_t(0)
        self._ck(self._L.evk_dbscan_voxels(self._h, C.byref(p), C.byref(nc), C.byref(ne)))
        return self._dbscan_results(self.num_voxels()[0], nc.value, ne.value)

    # ---- time surface + corner test (evk_ts_*) ----------------------------------------------
    def ts_create(self, width, height):
        self._ck(self._L.evk_ts_create(self._h, width, height))
        self._ts_shape = (height, width)

    def ts_corners(self, literal_break=True):
        """one callback range = the resident events: stamp the surface, test every event;
        returns the stream indices of the corner events"""
        n = C.c_size_t(0)
        self._ck(self._L.evk_ts_corners(self._h, 1 if literal_break else 0, C.byref(n)))
        idx = np.zeros(n.value, np.uint32)
        if n.value:
            self._ck(self._L.evk_ts_get_corners(self._h, _p(idx), len(idx)))
        return idx

    def _filtered(self, n):
        out = np.zeros(n, CORNER_DTYPE)
        if n:
            self._ck(self._L.evk_get_filtered_corners(self._h, _p(out), n))
        return out

    def ts_filter_corners(self, box_size=15):
        """box non-maximum suppression of the last ts_corners list -> records (x, y, label)"""
        n = C.c_size_t(0)
        self._ck(self._L.evk_ts_filter_corners(self._h, box_size, C.byref(n)))
        return self._filtered(n.value)

    def filter_corners(self, xy, width, height, box_size=15):
        """the same on a host list of (x, y) pairs"""
        xy = np.ascontiguousarray(xy, dtype=np.int32).reshape(-1, 2)
        n = C.c_size_t(0)
        self._ck(self._L.evk_filter_corners(self._h, _p(xy), len(xy), width, height, box_size,
                                            C.byref(n)))
        return self._filtered(n.value)

    # ---- OPTICS (evk_optics_*)
    def _optics_result(self):
        n = C.c_size_t(0)
        self._ck(self._L.evk_optics_get(self._h, None, None, 0, C.byref(n)))
        order, reach = np.zeros(n.value, np.uint32), np.zeros(n.value, np.float64)
        if n.value:
            self._ck(self._L.evk_optics_get(self._h, _p(order), _p(reach), n.value, C.byref(n)))
        return order, reach

    def optics_points(self, pts, min_pts, eps):
        """pts: [n, 2 or 3] integers -> (order, reachability by ordering position; -1 = none)"""
        pts = np.ascontiguousarray(pts, dtype=np.int32)
        self._ck(self._L.evk_optics_points(self._h, _p(pts), len(pts), pts.shape[1], min_pts, eps))
        return self._optics_result()

    def optics_voxels(self, min_pts, eps):
        self._ck(self._L.evk_optics_voxels(self._h, min_pts, eps))
        return self._optics_result()

    def optics_clusters(self, threshold):
        n = C.c_size_t(0)
        self._ck(self._L.evk_optics_get(self._h, None, None, 0, C.byref(n)))
        cl, nc = np.zeros(n.value, np.uint32), C.c_size_t(0)
        self._ck(self._L.evk_optics_clusters(self._h, threshold, _p(cl), n.value, C.byref(nc)))
        return cl, nc.value

    def ts_surface(self):
        out = np.zeros(self._ts_shape, np.int64)
        self._ck(self._L.evk_ts_get_surface(self._h, _p(out), out.size))
        return out

    def get_labels(self, n=None):
        if n is None:
            n = self.num_events if self._km.on_events else self.n_unique
        out = np.zeros(n, dtype=np.int32)
        self._ck(self._L.evk_get_labels(self._h, _p(out), n))
        return out

    def get_centroids(self, K, D):
        c = np.zeros((K, D), dtype=np.float32)
        counts = np.zeros(K, dtype=np.uint64)
        self._ck(self._L.evk_get_centroids(self._h, _p(c), _p(counts)))
        return c, counts

    # ---- streaming windows
    def window_config(self, ds, km, window_us):
        self._ck(self._L.evk_window_config(self._h, C.byref(ds), C.byref(km), window_us))
        self._km = km

    def window_config_events(self, ds, km, n_events):
        """windows of exactly n_events events (the reslicer's make_n_events condition)"""
        self._ck(self._L.evk_window_config_events(self._h, C.byref(ds), C.byref(km), n_events))
        self._km = km

    def window_push(self, ev):
        ev = np.ascontiguousarray(ev, dtype=EVENT_DTYPE)
        done = C.c_int(0)
        b = ev.ctypes.data
        self._ck(self._L.evk_window_push(self._h, b, b + ev.nbytes, C.byref(done)))
        return done.value

    def window_push_ptr(self, addr, n):
        """n records in a (pinned) host buffer at `addr`"""
        done = C.c_int(0)
        self._ck(self._L.evk_window_push(self._h, C.c_void_p(addr), C.c_void_p(addr + 16 * n),
                                         C.byref(done)))
        return done.value

    def window_flush(self):
        done = C.c_int(0)
        self._ck(self._L.evk_window_flush(self._h, C.byref(done)))
        return done.value

    # ---- measurement
    def set_profiling(self, on=True):
        self._ck(self._L.evk_set_profiling(self._h, int(on)))

    def stage_times(self):
        t = StageTimes()
        self._ck(self._L.evk_get_stage_times(self._h, C.byref(t)))
        return t

    def timer_start(self):
        self._ck(self._L.evk_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float(0)
        self._ck(self._L.evk_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def sync(self):
        self._ck(self._L.evk_sync(self._h))

    def flush_l2(self):
        self._ck(self._L.evk_flush_l2(self._h))

    # ---- multi-GPU
    @staticmethod
    def comm_unique_id():
        buf = (C.c_uint8 * 128)()
        st = lib().evk_comm_unique_id(buf)
        if st != 0:
            raise EvkError(st, "ncclGetUniqueId failed")
        return bytes(buf)

    def comm_init(self, rank, world, uid):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        self._ck(self._L.evk_comm_init(self._h, rank, world, buf))

    def set_shard(self, first_global_index):
        self._ck(self._L.evk_set_shard(self._h, first_global_index))

    def downsample_sharded(self, p, owner_mode=OWNER_TIME_RANGE):
        ul, ug = C.c_size_t(0), C.c_size_t(0)
        self._ck(self._L.evk_downsample_sharded(self._h, C.byref(p), owner_mode, C.byref(ul),
                                                C.byref(ug)))
        self.n_unique = ul.value
        return ul.value, ug.value

    def init_centroids_first_k_sharded(self, km):
        self._ck(self._L.evk_init_centroids_first_k_sharded(self._h, C.byref(km)))

    def downsample_kmeans_sharded(self, ds, km, init_first_k=True, owner_mode=0):
        """Fused sharded step. Returns (n_unique_local, n_unique_global, iters_done)."""
        ul, ug, it = C.c_size_t(0), C.c_size_t(0), C.c_int(0)
        self._ck(self._L.evk_downsample_kmeans_sharded(self._h, C.byref(ds), C.byref(km),
                                                       1 if init_first_k else 0, owner_mode,
                                                       C.byref(ul), C.byref(ug), C.byref(it)))
        self.n_unique = ul.value
        self._km = km
        return ul.value, ug.value, it.value

    def downsample_kmeans_sharded_submit(self, ds, km, init_first_k=True, owner_mode=0):
        self._ck(self._L.evk_downsample_kmeans_sharded_submit(
            self._h, C.byref(ds), C.byref(km), 1 if init_first_k else 0, owner_mode))
        self._km = km

    def downsample_kmeans_sharded_wait(self):
        """(n_unique_local, n_unique_global, iters_done) of the last queued sharded step"""
        ul, ug, it = C.c_size_t(0), C.c_size_t(0), C.c_int(0)
        self._ck(self._L.evk_downsample_kmeans_sharded_wait(self._h, C.byref(ul), C.byref(ug),
                                                            C.byref(it)))
        self.n_unique = ul.value
        return ul.value, ug.value, it.value

    def kmeans_sharded(self, km):
        it = C.c_int(0)
        self._ck(self._L.evk_kmeans_sharded(self._h, C.byref(km), C.byref(it)))
        self._km = km
        return it.value
