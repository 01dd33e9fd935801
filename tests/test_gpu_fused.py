"""GPU parity of the fused step (evk_downsample_kmeans): the slab kernel's consumer warps run the
first Lloyd iteration on every voxel as it is emitted.  Results must equal the oracle's and be
bit-identical to the three separate calls (downsample, first-K init, k-means)."""
import numpy as np
import pytest

import evk_loader
from test_gpu_parity import CENT_RTOL, CONFIGS, check_labels

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def evk():
    m = evk_loader.load()
    m.lib()
    return m


def unfused(evk, h, ds, km, init):
    U, R = h.downsample(ds)
    if init is None:
        h.init_centroids_first_k(km)
    else:
        h.set_centroids(init)
    it = h.kmeans(km)
    keys, _, first = h.get_voxels(reps=False)
    return U, R, it, keys, first, h.get_labels(), h.get_centroids(km.K, km.D)


def fused(evk, h, ds, km, init):
    if init is not None:
        h.set_centroids(init)
    U, R, it = h.downsample_kmeans(ds, km, init is None)
    keys, _, first = h.get_voxels(reps=False)
    return U, R, it, keys, first, h.get_labels(), h.get_centroids(km.K, km.D)


def same(a, b):
    assert a[:3] == b[:3]
    for x, y in zip(a[3:6], b[3:6]):
        assert (x == y).all()
    assert (a[6][0] == b[6][0]).all() and (a[6][1] == b[6][1]).all()


@pytest.mark.parametrize("name", list(CONFIGS))
def test_fused_vs_oracle(evk, orc, name):
    seed, n, rate, W, H, blobs, (vx, vy, vt, up), K, iters = CONFIGS[name]
    ev = orc.synth(orc.synth_params(seed, n, W, H, rate, blobs), threads=orc.max_threads())
    ok, of, orr = orc.downsample(ev, orc.ds_params(W, H, vx, vy, vt, 0, up))
    pts = orc.points(ev, of, 2)
    with evk.Evk(n) as h:
        h.load_events(ev)
        for it_n in (1, iters):
            ds = evk.ds_params(W, H, vx, vy, vt, 0, up)
            km = evk.km_params(K, 2, iters=it_n)
            got = fused(evk, h, ds, km, None)
            assert h.stage_times().ds_algo_used == evk.ALGO_SLAB
            assert h.stage_times().km_launches >= 1
            U, R, it, keys, first, lab, (cent, counts) = got
            assert (U, R, it) == (len(ok), orr, it_n)
            assert (keys == ok).all() and (first == of).all()
            oc, ol, ocnt, _ = orc.kmeans(pts, pts[:K], iters=it_n, threads=orc.max_threads())
            oc_prev = orc.kmeans(pts, pts[:K], iters=it_n - 1)[0] if it_n > 1 else pts[:K]
            check_labels(pts, oc_prev, lab, ol)
            assert (counts == ocnt).all()
            np.testing.assert_allclose(cent, oc, rtol=CENT_RTOL, atol=0)
            same(got, unfused(evk, h, ds, km, None))


def test_fused_variants(evk, orc):
    """warm start, distance gate, tolerance stop, K from 1 to 254, no polarity, odd voxel sizes,
    repeated-count off"""
    n, W, H = 600_000, 346, 260
    ev = orc.synth(orc.synth_params(0xE7CA0002, n, W, H, 10_000_000, 16))
    rng = np.random.default_rng(3)
    with evk.Evk(n) as h:
        h.load_events(ev)
        for (vx, vy, vt, up, rep), K, md, iters, tol in [
                ((4, 4, 1000, 1, 1), 8, 0.0, 1, -1.0), ((4, 4, 1000, 1, 1), 1, 0.0, 2, -1.0),
                ((3, 5, 777, 0, 1), 17, 30.0, 3, -1.0), ((2, 2, 500, 1, 0), 200, 0.0, 1, -1.0),
                ((1, 1, 2000, 1, 1), 254, 25.0, 2, -1.0), ((4, 4, 1000, 1, 1), 32, 0.0, 40, 0.25)]:
            ds = evk.ds_params(W, H, vx, vy, vt, 0, up, count_repeated=rep)
            km = evk.km_params(K, 2, max_dist=md, iters=iters, tol=tol)
            for init in (None, np.float32(rng.uniform(0, [W, H], size=(K, 2)))):
                a = fused(evk, h, ds, km, init)
                assert h.stage_times().ds_algo_used == evk.ALGO_SLAB
                same(a, unfused(evk, h, ds, km, init))
                if md > 0:
                    assert (a[5] < 0).any()
        # oracle on one warm-started case
        ok, of, orr = orc.downsample(ev, orc.ds_params(W, H, 4, 4, 1000, 0, 1))
        pts = orc.points(ev, of, 2)
        c0 = np.float32(rng.uniform(0, [W, H], size=(12, 2)))
        U, R, it, keys, first, lab, (cent, counts) = fused(
            evk, h, evk.ds_params(W, H, 4, 4, 1000, 0, 1), evk.km_params(12, 2, iters=1), c0)
        oc, ol, ocnt, _ = orc.kmeans(pts, c0, iters=1)
        assert (U, R) == (len(ok), orr) and (keys == ok).all() and (first == of).all()
        check_labels(pts, c0, lab, ol)
        assert (counts == ocnt).all()
        np.testing.assert_allclose(cent, oc, rtol=CENT_RTOL, atol=0)


def test_fused_falls_back_on_unordered_stream(evk, orc):
    n, W, H, K = 200_000, 346, 260, 8
    ev = orc.synth(orc.synth_params(0xE7CA0001, n, W, H, 10_000_000, 8))
    shuffled = ev[np.random.default_rng(7).permutation(n)]
    c0 = np.float32(np.random.default_rng(8).uniform(0, [W, H], size=(K, 2)))
    ds, km = evk.ds_params(W, H, 4, 4, 1000, 0, 1), evk.km_params(K, 2, iters=2)
    with evk.Evk(n) as h:
        h.load_events(shuffled)
        for init in (None, c0):  # warm start: the caller's centroids survive the abandoned pass
            a = fused(evk, h, ds, km, init)
            assert h.stage_times().ds_algo_used in (evk.ALGO_PARTITION, evk.ALGO_TABLE)  # the general paths
            same(a, unfused(evk, h, ds, km, init))
        # too few voxels to seed K clusters: reported, never fatal
        h.load_events(ev[:3])
        with pytest.raises(evk.EvkError) as e:
            h.downsample_kmeans(ds, km, True)
        assert e.value.status == -1
        # empty stream
        h.load_events(ev[:0])
        h.set_centroids(c0)
        assert h.downsample_kmeans(ds, km, False)[:2] == (0, 0)
        assert (h.get_centroids(K, 2)[0] == c0).all()


def test_fused_full_size(evk):
    """C3 at full size: fused == separate calls (counts, centroids, label histogram, checksums)"""
    n, W, H, K = 100_000_000, 1280, 720, 64
    ds, km = evk.ds_params(W, H, 2, 2, 500, 0, 1), evk.km_params(K, 2, iters=1)
    with evk.Evk(n) as h:
        h.synth(evk.synth_params(0xE7CA0003, n, W, H, 100_000_000, 64))
        U, R, it = h.downsample_kmeans(ds, km, True)
        assert h.stage_times().ds_algo_used == evk.ALGO_SLAB
        cf, nf = h.get_centroids(K, 2)
        lf = h.get_labels()
        U2, R2 = h.downsample(ds)
        h.init_centroids_first_k(km)
        h.kmeans(km)
        cu, nu = h.get_centroids(K, 2)
        lu = h.get_labels()
    assert (U, R) == (U2, R2) and int(nf.sum()) == U
    assert (nf == nu).all() and (cf == cu).all()
    assert (lf == lu).all()
    assert (np.bincount(lf, minlength=K) == nf).all()


def test_submit_wait_pipeline(evk, orc):
    """evk_downsample_kmeans_submit / _wait: queued steps == the synchronous call, bit for bit;
    a warm-started step may be queued behind a cold one; a rejected (unordered) stream is rerun
    on the general path inside wait; wait without a submission is a state error."""
    n, W, H, K = 600_000, 1280, 720, 16
    ev = orc.synth(orc.synth_params(0xE7CA0003, n, W, H, 100_000_000, 16))
    ds, km = evk.ds_params(W, H, 2, 2, 500, 0, 1), evk.km_params(K, 2, iters=1)
    with evk.Evk(n) as h:
        h.load_events(ev)
        ref = fused(evk, h, ds, km, None)
        with pytest.raises(evk.EvkError):
            h.downsample_kmeans_wait()
        for _ in range(3):
            h.downsample_kmeans_submit(ds, km, True)
        U, R, it = h.downsample_kmeans_wait()
        keys, _, first = h.get_voxels(reps=False)
        got = (U, R, it, keys, first, h.get_labels(), h.get_centroids(K, 2))
        assert h.stage_times().ds_algo_used == evk.ALGO_SLAB
        same(ref, got)
        # a getter collects a queued step by itself (no explicit wait): same results
        h.downsample_kmeans_submit(ds, km, True)
        with pytest.raises(evk.EvkError):
            h.num_voxels()                      # (const query: refuses while a step is in flight)
        lab2 = h.get_labels(len(ref[5]))
        assert (lab2 == ref[5]).all() and h.num_voxels() == (ref[0], ref[1])
        with pytest.raises(evk.EvkError):
            h.downsample_kmeans_wait()          # already collected
        # cold step, then a warm-started one queued behind it == two synchronous calls
        h.downsample_kmeans(ds, km, True)
        h.downsample_kmeans(ds, km, False)
        want = h.get_centroids(K, 2)
        want_lab = h.get_labels()
        h.downsample_kmeans_submit(ds, km, True)
        h.downsample_kmeans_submit(ds, km, False)
        h.downsample_kmeans_wait()
        c = h.get_centroids(K, 2)
        assert (c[0] == want[0]).all() and (c[1] == want[1]).all()
        assert (h.get_labels() == want_lab).all()
        # unordered stream: the slab pass gives up, wait reruns the general path
        shuffled = ev[np.random.default_rng(3).permutation(n)]
        h.load_events(shuffled)
        a = unfused(evk, h, ds, km, None)
        h.downsample_kmeans_submit(ds, km, True)
        U, R, it = h.downsample_kmeans_wait()
        assert h.stage_times().ds_algo_used in (evk.ALGO_PARTITION, evk.ALGO_TABLE)
        keys, _, first = h.get_voxels(reps=False)
        same(a, (U, R, it, keys, first, h.get_labels(), h.get_centroids(K, 2)))
