"""Round-2 GPU tests: full-size parity against the oracle (BASELINE configs C2 and C3 at the sizes
the metric is quoted on), and regression tests for the advisor's findings on the C-ABI state machine.
Everything goes through the C-ABI of libevk.so; the oracle is only the checker."""
import hashlib

import numpy as np
import pytest

import evk_loader

pytestmark = pytest.mark.gpu
CENT_RTOL = 1e-5   # north_star: centroids agree within 1e-5 relative


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def mix(k):
    """order-independent checksum pair of a key set (xor and sum of mixed keys)"""
    k = k.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        k ^= k >> np.uint64(33)
        k *= np.uint64(0xFF51AFD7ED558CCD)
        k ^= k >> np.uint64(33)
    return int(np.bitwise_xor.reduce(k)), int(k.sum(dtype=np.uint64))


@pytest.fixture(scope="module")
def evk():
    m = evk_loader.load()
    m.lib()
    return m


# --------------------------------------------------------------- full-size parity vs the oracle
def _full_size(evk, orc, seed, n, rate, W, H, blobs, vox, K, iters, fused):
    vx, vy, vt, up = vox
    T = orc.max_threads()
    with evk.Evk(n) as h:
        h.synth(evk.synth_params(seed, n, W, H, rate, blobs))
        ev = h.get_events()                      # the device generator == the oracle's (checked below)
        ref = orc.synth(orc.synth_params(seed, 1 << 20, W, H, rate, blobs), threads=T)
        assert ev[:1 << 20].tobytes() == ref.tobytes()
        ok, of, orr = orc.downsample(ev, orc.ds_params(W, H, vx, vy, vt, 0, up), threads=T)
        pts = orc.points(ev, of, 2)
        oc, ol, ocnt, _ = orc.kmeans(pts, pts[:K], iters=iters, threads=T)
        ds = evk.ds_params(W, H, vx, vy, vt, 0, up)
        km = evk.km_params(K, 2, iters=iters)
        U, R = h.downsample(ds)
        assert h.stage_times().ds_algo_used == evk.ALGO_SLAB   # the fast path, no silent fallback
        assert (U, R) == (len(ok), orr)
        keys, _, first = h.get_voxels(reps=False)
        assert mix(keys) == mix(ok), "voxel key set differs from the oracle's"
        assert (keys == ok).all(), "canonical order of the keys differs"
        assert sha(first) == sha(of), "representatives (first indices) differ"
        h.init_centroids_first_k(km)
        assert h.kmeans(km) == iters
        lab = h.get_labels()
        cent, counts = h.get_centroids(K, 2)
        assert (counts == ocnt).all(), "per-cluster counts differ"
        np.testing.assert_allclose(cent, oc, rtol=CENT_RTOL, atol=0)
        bad = np.nonzero(lab != ol)[0]
        assert len(bad) == 0, f"{len(bad)} labels differ (first at {bad[:5]})"
        if fused:   # the bench's step: downsample + first-K init + one iteration, one graph
            oc1, ol1, ocnt1, _ = orc.kmeans(pts, pts[:K], iters=1, threads=T)
            U1, R1, it = h.downsample_kmeans(ds, evk.km_params(K, 2, iters=1), True)
            assert (U1, R1, it) == (len(ok), orr, 1)
            keys1, _, first1 = h.get_voxels(reps=False)
            assert (keys1 == ok).all() and sha(first1) == sha(of)
            lab1 = h.get_labels()
            cent1, counts1 = h.get_centroids(K, 2)
            assert (counts1 == ocnt1).all() and (lab1 == ol1).all()
            np.testing.assert_allclose(cent1, oc1, rtol=CENT_RTOL, atol=0)


def test_c2_full_size_vs_oracle(evk, orc):
    """BASELINE configs[1]: 10 M DAVIS346 events, 4x4 px x 1 ms voxels, K = 32, 20 iterations"""
    _full_size(evk, orc, 0xE7CA0002, 10_000_000, 10_000_000, 346, 260, 32, (4, 4, 1000, 1), 32, 20,
               fused=False)


def test_c3_full_size_vs_oracle(evk, orc):
    """BASELINE configs[2], the bench workload: 100 M Gen4 events, 2x2 px x 500 us voxels, K = 64;
    separate calls (2 iterations) and the fused step against the oracle at full N"""
    _full_size(evk, orc, 0xE7CA0003, 100_000_000, 100_000_000, 1280, 720, 64, (2, 2, 500, 1), 64, 2,
               fused=True)


def test_c3_unordered_full_size(evk, orc):
    """the general path at the bench size: the C3 stream with its time bins shuffled (every bin's
    events stay together but bins arrive in random order -- out-of-order packets) and a fully
    shuffled 20 M prefix give the oracle's voxel set"""
    n, W, H = 100_000_000, 1280, 720
    T = orc.max_threads()
    ds = evk.ds_params(W, H, 2, 2, 500, 0, 1)
    with evk.Evk(n) as h:
        h.synth(evk.synth_params(0xE7CA0003, n, W, H, 100_000_000, 64))
        ev = h.get_events()
        rng = np.random.default_rng(5)
        # 50 000-event blocks = one 500 us bin each at 100 Mev/s
        blocks = rng.permutation(n // 50_000)
        shuf = ev.reshape(-1, 50_000)[blocks].reshape(-1)
        ok, of, orr = orc.downsample(shuf, orc.ds_params(W, H, 2, 2, 500, 0, 1), threads=T)
        h.load_events(shuf)
        U, R = h.downsample(ds)
        assert (U, R) == (len(ok), orr)
        keys, _, first = h.get_voxels(reps=False)
        assert (keys == ok).all() and sha(first) == sha(of)
        m = 20_000_000
        full = ev[:m][rng.permutation(m)]
        ok, of, orr = orc.downsample(full, orc.ds_params(W, H, 2, 2, 500, 0, 1), threads=T)
        h.load_events(full)
        U, R = h.downsample(ds)
        assert (U, R) == (len(ok), orr)
        keys, _, first = h.get_voxels(reps=False)
        assert (keys == ok).all() and sha(first) == sha(of)


# --------------------------------------------------------------------- advisor regressions ----
def test_ref_hash_kmeans_on_the_frame_edge(evk, orc):
    """EVK_KEY_REF_HASH8192 gates inclusively (x <= width, coordinate_processor.cl:56): a
    representative at x == width or y == height lies outside the W x H pixel images, so k-means
    after that key function must not take the pixel-image path"""
    W, H, K = 64, 48, 4
    rng = np.random.default_rng(3)
    x = rng.integers(0, W + 1, 3000)
    y = rng.integers(0, H + 1, 3000)
    x[:8] = [W, W, 0, W, 5, W, W - 1, W]
    y[:8] = [H, 0, H, H - 1, H, 7, H, H]
    xy = np.stack([x, y], 1).astype(np.int32).ravel()
    ev = orc.events_from_xy(x, y)
    ok, of, orr = orc.downsample(ev, orc.ds_params(W, H, 1, 1, 0, 0, 0, 1))
    pts = orc.points(ev, of, 2)
    assert (pts[:, 0] == W).any() and (pts[:, 1] == H).any()
    with evk.Evk(4096) as h:
        h.load_coords_i32(xy)
        U, R = h.downsample(evk.ds_params(W, H, keyfn=evk.KEY_REF_HASH8192))
        assert (U, R) == (len(ok), orr)
        for iters in (1, 3):
            km = evk.km_params(K, 2, iters=iters)
            h.set_centroids(pts[:K])
            assert h.kmeans(km) == iters
            lab = h.get_labels()
            cent, counts = h.get_centroids(K, 2)
            oc, ol, ocnt, _ = orc.kmeans(pts, pts[:K], iters=iters)
            assert (lab == ol).all() and (counts == ocnt).all()
            np.testing.assert_allclose(cent, oc, rtol=CENT_RTOL, atol=0)


def test_image_reallocation_invalidates_cached_graphs(evk, orc):
    """the pixel images grow with the frame; graphs captured for a smaller frame hold the old
    pointers and must be re-captured (fused step and the Lloyd-loop graph)"""
    n = 300_000
    small = orc.synth(orc.synth_params(11, n, 640, 480, 20_000_000, 8))
    big = orc.synth(orc.synth_params(12, n, 1280, 720, 20_000_000, 8))
    ds_s, ds_b = evk.ds_params(640, 480, 2, 2, 500, 0, 1), evk.ds_params(1280, 720, 2, 2, 500, 0, 1)
    km1, km5 = evk.km_params(16, 2, iters=1), evk.km_params(16, 2, iters=5)

    def fused(h, ev, ds):
        h.load_events(ev)
        U, R, _ = h.downsample_kmeans(ds, km1, True)
        return U, R, h.get_labels().copy(), h.get_centroids(16, 2)[0].copy()

    def loop(h, ev, ds):
        h.load_events(ev)
        h.downsample(ds)
        h.init_centroids_first_k(km5)
        h.kmeans(km5)
        return h.get_labels().copy(), h.get_centroids(16, 2)[0].copy()

    with evk.Evk(n) as h:
        a = fused(h, small, ds_s)
        la = loop(h, small, ds_s)
        b = fused(h, big, ds_b)            # reallocates the images
        lb = loop(h, big, ds_b)
        a2 = fused(h, small, ds_s)         # same key as the first call: must not replay a stale graph
        la2 = loop(h, small, ds_s)
    with evk.Evk(n) as h2:                 # fresh handle: the big frame first
        b_ref = fused(h2, big, ds_b)
        lb_ref = loop(h2, big, ds_b)
    for x, y in ((a, a2), (b, b_ref)):
        assert x[:2] == y[:2] and (x[2] == y[2]).all() and (x[3] == y[3]).all()
    for x, y in ((la, la2), (lb, lb_ref)):
        assert (x[0] == y[0]).all() and (x[1] == y[1]).all()


def test_loading_collects_a_queued_step(evk, orc):
    """a loader called while a fused step is queued collects that step first: its first indices
    never get published over a different event buffer"""
    n, W, H = 200_000, 640, 480
    a = orc.synth(orc.synth_params(21, n, W, H, 20_000_000, 8))
    b = orc.synth(orc.synth_params(22, n, W, H, 20_000_000, 8))
    ds, km = evk.ds_params(W, H, 2, 2, 500, 0, 1), evk.km_params(8, 2, iters=1)
    okb, ofb, orrb = orc.downsample(b, orc.ds_params(W, H, 2, 2, 500, 0, 1))
    with evk.Evk(n) as h:
        h.load_events(a)
        h.downsample_kmeans_submit(ds, km, True)
        h.load_events(b)                         # collects (and discards) the queued step
        with pytest.raises(evk.EvkError) as e:
            h.downsample_kmeans_wait()
        assert e.value.status == -4              # EVK_ERR_STATE: nothing is pending any more
        with pytest.raises(evk.EvkError):
            h.get_voxels()                       # results of the old events are gone
        U, R = h.downsample(ds)
        keys, reps, first = h.get_voxels()
        assert (U, R) == (len(okb), orrb) and (keys == okb).all() and (first == ofb).all()
        assert reps.tobytes() == b[ofb].tobytes()


def test_queued_steps_report_a_rejected_slice(evk, orc):
    """several queued fused steps: if the time-slab path rejects an EARLIER one (unordered slice)
    the wait says so instead of silently skipping that slice"""
    n, W, H = 200_000, 640, 480
    ev = orc.synth(orc.synth_params(23, n, W, H, 20_000_000, 8))
    bad = ev[np.random.default_rng(1).permutation(n)]
    ds, km = evk.ds_params(W, H, 2, 2, 500, 0, 1), evk.km_params(8, 2, iters=1)
    with evk.Evk(n) as h:
        h.load_events(bad)
        h.downsample_kmeans_submit(ds, km, True)     # will be rejected on the device
        # (same buffer, queued behind it: the library cannot rerun step 1 once step 2 is collected)
        h.downsample_kmeans_submit(ds, km, False)
        with pytest.raises(evk.EvkError) as e:
            h.downsample_kmeans_wait()
        assert e.value.status == -4
        # the handle stays usable; one step at a time falls back to the general path
        U, R, _ = h.downsample_kmeans(ds, km, True)
        ok, of, orr = orc.downsample(bad, orc.ds_params(W, H, 2, 2, 500, 0, 1))
        assert (U, R) == (len(ok), orr)
        h.load_events(ev)
        for _ in range(3):
            h.downsample_kmeans_submit(ds, km, True)
        ok, of, orr = orc.downsample(ev, orc.ds_params(W, H, 2, 2, 500, 0, 1))
        assert h.downsample_kmeans_wait()[:2] == (len(ok), orr)


# ------------------------------------------------- D = 3 / 4 k-means with space-time pruning --
@pytest.mark.parametrize("D,K,md", [(3, 32, 0.0), (4, 64, 0.0), (3, 64, 90.0), (4, 12, 0.0)])
def test_kmeans_d3_d4_pruned_vs_oracle(evk, orc, D, K, md):
    """D = 3 / 4 on voxels takes the candidate lists per (time slab, pixel tile) when K > 8: labels,
    counts and centroids must equal the oracle's full scan (ties within 1e-6 aside)"""
    n, W, H = 1_200_000, 640, 480
    ts, ps = 2e-3, 40.0
    ev = orc.synth(orc.synth_params(0xE7CA0007, n, W, H, 20_000_000, 24))
    ok, of, _ = orc.downsample(ev, orc.ds_params(W, H, 2, 2, 500, 0, 1))
    pts = orc.points(ev, of, D, 0, ts, ps)
    iters = 3
    oc_prev, _, _, _ = orc.kmeans(pts, pts[:K], md, iters=iters - 1, threads=orc.max_threads())
    oc, ol, ocnt, _ = orc.kmeans(pts, pts[:K], md, iters=iters, threads=orc.max_threads())
    with evk.Evk(n) as h:
        h.load_events(ev)
        for algo in (evk.ALGO_SLAB, evk.ALGO_PARTITION):
            h.downsample(evk.ds_params(W, H, 2, 2, 500, 0, 1, algo=algo))
            assert h.stage_times().ds_algo_used == algo
            km = evk.km_params(K, D, max_dist=md, iters=iters, t_scale=ts, p_scale=ps)
            h.init_centroids_first_k(km)
            assert h.kmeans(km) == iters
            lab = h.get_labels()
            cent, counts = h.get_centroids(K, D)
            bad = np.nonzero(lab != ol)[0]
            for i in bad:
                assert lab[i] >= 0 and ol[i] >= 0
                dg = ((oc_prev[lab[i]].astype(np.float64) - pts[i]) ** 2).sum()
                dw = ((oc_prev[ol[i]].astype(np.float64) - pts[i]) ** 2).sum()
                assert abs(dg - dw) <= 1e-6 * max(dg, dw)
            if md > 0:
                assert (lab < 0).any()
            assert (counts == ocnt).all()
            np.testing.assert_allclose(cent, oc, rtol=CENT_RTOL, atol=1e-6)
