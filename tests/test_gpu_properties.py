"""Randomised GPU parity (hypothesis, derandomised: the same examples on every run) of the SURVEY 8f
rows against their oracles: the consumer on arbitrary parameters, DBSCAN on arbitrary small lattice
clouds, the corner test on random surfaces, EVT 3.0 on arbitrary ordered streams."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import aec_streams as S
import dbscan_cases as D
import evk_loader
from oracle import aec, dbscan

pytestmark = pytest.mark.gpu
SET = dict(deadline=None, derandomize=True,
           suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])


@pytest.fixture(scope="module")
def evk():
    m = evk_loader.load()
    m.lib()
    return m


@pytest.fixture(scope="module")
def handle(evk):
    with evk.Evk(1 << 16) as h:
        yield h


@settings(max_examples=40, **SET)
@given(seed=st.integers(0, 2**31), n=st.integers(1, 900), sz=st.integers(1, 150),
       radius=st.sampled_from([2.0, 5.0, 12.5, 40.0]), kappa=st.integers(0, 12),
       alpha=st.sampled_from([0.1, 0.3, 0.5, 0.9, 1.0]), min_n=st.integers(0, 8),
       tie=st.integers(1, 40), rand_seed=st.integers(0, 2**31))
def test_consumer_arbitrary_parameters(evk, orc, handle, seed, n, sz, radius, kappa, alpha, min_n,
                                       tie, rand_seed):
    r = np.random.default_rng(seed)
    e = np.zeros((n, 4))
    e[:, 0] = 3.0 + np.floor(np.arange(n) / tie) * 1e-3
    e[:, 1:3] = r.integers(0, 90, size=(n, 2))
    e[:, 3] = r.integers(0, 2, n)
    init = dict(sz_buffer=sz, radius=radius, kappa=kappa, alpha=alpha, min_n=min_n)
    o = aec.Oracle(init, rand_seed=rand_seed)
    handle.aec_create(init, rand_seed=rand_seed)
    half = n // 2
    for part in (e[:half], e[half:]):
        o.update(part)
        handle.aec_update(part)
        S.same_state(handle.aec_state(), o.state())


@settings(max_examples=40, **SET)
@given(seed=st.integers(0, 2**31), n=st.integers(1, 400), span=st.integers(4, 60),
       eps=st.sampled_from([1.0, 1.5, 2.0, 3.0, 5.0]), min_pts=st.integers(1, 7),
       mn=st.integers(1, 6), dims=st.sampled_from([2, 3]))
def test_dbscan_arbitrary_lattice_clouds(evk, orc, handle, seed, n, span, eps, min_pts, mn, dims):
    r = np.random.default_rng(seed)
    pts = r.integers(0, span, size=(n, dims)).astype(np.float32)
    lo, co, so = dbscan.oracle(pts, eps, min_pts, mn, 10**6)
    labels, sizes, seeds, extra = handle.dbscan_points(pts, eps, min_pts, mn, 10**6)
    assert (labels == lo).all() and (seeds == so).all()
    assert D.canon(D.clusters_from(labels, sizes, extra)) == D.canon(co)


@settings(max_examples=25, **SET)
@given(seed=st.integers(0, 2**31), n=st.integers(1, 3000), levels=st.sampled_from([3, 50, 100000]),
       literal=st.booleans())
def test_corners_random_events(evk, orc, handle, seed, n, levels, literal):
    """dense random stamps in a small window (few distinct timestamps -> ties on the circles)"""
    W, H = 64, 48
    r = np.random.default_rng(seed)
    ev = np.zeros(n, orc.EVENT_DTYPE)
    ev["x"], ev["y"] = r.integers(0, W, n), r.integers(0, H, n)
    ev["t"] = np.sort(r.integers(1, levels + 1, n))
    if literal:                                   # keep the as-written break from ending the range
        ev["x"], ev["y"] = np.clip(ev["x"], 4, W - 5), np.clip(ev["y"], 4, H - 5)
    s = np.zeros((H, W), np.int64)
    want = orc.ts_corners(ev, W, H, s, literal)
    handle.ts_create(W, H)
    handle.load_events(ev)
    got = handle.ts_corners(literal)
    assert got.tolist() == want.tolist()
    assert (handle.ts_surface() == s).all()


@settings(max_examples=25, **SET)
@given(seed=st.integers(0, 2**31), n=st.integers(0, 20000), rows=st.integers(1, 40),
       cols=st.integers(1, 300), dt_max=st.integers(0, 5000))
def test_evt3_arbitrary_ordered_streams(evk, orc, handle, seed, n, rows, cols, dt_max):
    r = np.random.default_rng(seed)
    ev = np.zeros(n, orc.EVENT_DTYPE)
    ev["t"] = np.cumsum(r.integers(0, dt_max + 1, n)) if n else 0
    ev["y"], ev["x"], ev["p"] = r.integers(0, rows, n), r.integers(0, cols, n), r.integers(0, 2, n)
    ev = ev[np.lexsort((ev["x"], ev["p"], ev["y"], ev["t"]))]
    w = orc.evt3_encode(ev)
    assert handle.load_evt3(w) == n
    assert handle.get_events().tobytes() == ev.tobytes()


@settings(max_examples=40, **SET)
@given(seed=st.integers(0, 2**31), n=st.integers(1, 60000), W=st.integers(9, 700),
       H=st.integers(9, 500), vx=st.integers(1, 9), vy=st.integers(1, 9),
       vt=st.sampled_from([1, 7, 64, 500, 1000, 12345]), up=st.integers(0, 1),
       dt_max=st.integers(0, 6), hot=st.sampled_from([0.0, 0.3, 0.95]), K=st.integers(1, 40),
       algo=st.sampled_from(["auto", "table", "sort"]))
def test_hot_path_arbitrary_shapes(evk, orc, handle, seed, n, W, H, vx, vy, vt, up, dt_max, hot, K,
                                   algo):
    """the hot path itself on arbitrary sensors, voxel sizes, time bins and duplicate densities:
    every downsample algorithm and the fused step equal the oracle bit for bit"""
    from test_gpu_stress import make_stream
    ev = make_stream(np.random.default_rng(seed), n, W, H, dt_max, hot)
    ok, of, orr = orc.downsample(ev, orc.ds_params(W, H, vx, vy, vt, 0, up))
    a = {"auto": evk.ALGO_AUTO, "table": evk.ALGO_TABLE, "sort": evk.ALGO_SORT}[algo]
    ds = evk.ds_params(W, H, vx, vy, vt, 0, up, algo=a)
    handle.load_events(ev)
    U, R = handle.downsample(ds)
    keys, reps, first = handle.get_voxels()
    assert (U, R) == (len(ok), orr) and (keys == ok).all() and (first == of).all()
    assert reps.tobytes() == ev[of].tobytes()
    K = min(K, len(ok))
    pts = orc.points(ev, of, 2)
    oc, ol, ocnt, _ = orc.kmeans(pts, pts[:K], iters=1)
    U, R, it = handle.downsample_kmeans(evk.ds_params(W, H, vx, vy, vt, 0, up), evk.km_params(K, 2, iters=1),
                                        True)
    assert (U, R, it) == (len(ok), orr, 1)
    cent, counts = handle.get_centroids(K, 2)
    assert (counts == ocnt).all() and (cent == oc).all() and (handle.get_labels() == ol).all()
