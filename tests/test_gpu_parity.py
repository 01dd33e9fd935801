"""GPU parity tests (run on the B200 box): every call goes through the C-ABI of libevk.so and
is compared with the CPU oracle on the same inputs and with the committed golden fixtures.

Bars (BASELINE.json north_star): voxel key set bit-exact after canonical sort (here: also the
representatives and the canonical order are bit-exact); labels identical except for distance ties
within 1e-6 relative; centroids within 1e-5 relative (integer coordinates make them bit-exact).
"""
import hashlib
import os

import numpy as np
import pytest

import evk_loader

pytestmark = pytest.mark.gpu
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CENT_RTOL = 1e-5   # north_star: centroids agree within 1e-5 relative
TIE_RTOL = 1e-6    # north_star: label mismatches allowed only for distance ties within 1e-6


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def evk():
    m = evk_loader.load()
    m.lib()  # raises if libevk.so is missing: no fallback
    return m


ALGOS = ["AUTO", "TABLE", "SORT", "SLAB"]


def algo_id(evk, name):
    return getattr(evk, "ALGO_" + name)


def check_labels(pts, cent, got, want):
    """identical, or the two candidate distances tie within TIE_RTOL"""
    bad = np.nonzero(got != want)[0]
    for i in bad:
        assert got[i] >= 0 and want[i] >= 0
        dg = ((cent[got[i]].astype(np.float64) - pts[i]) ** 2).sum()
        dw = ((cent[want[i]].astype(np.float64) - pts[i]) ** 2).sum()
        assert abs(dg - dw) <= TIE_RTOL * max(dg, dw), (i, got[i], want[i], dg, dw)


def run_ds(evk, orc, h, ev, W, H, vx, vy, vt, up, algo, keyfn=0, t0=0):
    U, R = h.downsample(evk.ds_params(W, H, vx, vy, vt, t0, up, keyfn, algo_id(evk, algo)))
    keys, reps, first = h.get_voxels()
    ok, of, orr = orc.downsample(ev, orc.ds_params(W, H, vx, vy, vt, t0, up, keyfn))
    assert (U, R) == (len(ok), orr), (algo, U, R, len(ok), orr)
    assert (np.sort(keys) == np.sort(ok)).all(), "voxel key set differs"
    assert (keys == ok).all() and (first == of).all(), "canonical order / representatives differ"
    assert reps.tobytes() == ev[of].tobytes()
    return ok, of, orr


# ------------------------------------------------------------------------------ fixtures ------
def test_f1_kmeans_reference_fixture(evk, orc, golden):
    """KM/assign_to_centers2.c:121-131: 2048 points, 8 centres, threshold 50"""
    g = golden["F1"]
    data = (np.arange(4096) % 100).astype(np.int32)
    c0 = np.array([1, 1, 10, 10, 20, 20, 30, 30, 50, 50, 60, 60, 70, 70, 80, 80],
                  np.float32).reshape(8, 2)
    with evk.Evk(4096) as h:
        h.load_coords_i32(data)
        h.set_centroids(c0)
        km = evk.km_params(8, 2, max_dist=50.0, iters=1, on_events=1)
        assert h.kmeans(km) == 1
        lab = h.get_labels()
        cent, counts = h.get_centroids(8, 2)
    assert lab.tolist() == g["labels"]
    assert counts.tolist() == [123, 205, 205, 287, 328, 205, 205, 490]
    np.testing.assert_allclose(cent.ravel(), g["centroids"], rtol=CENT_RTOL)
    assert cent.ravel().tolist() == np.float32(g["centroids"]).tolist()  # exact sums => bit-exact


def test_f2_warmup_launch(evk, orc, golden):
    """ACCEL/store.cpp:209-215,317-326: 8192 x (0,0) through the reference hash"""
    with evk.Evk(8192) as h:
        h.load_coords_i32(np.zeros(16384, np.int32))
        for algo in ("AUTO", "TABLE", "SORT"):
            U, R = h.downsample(evk.ds_params(1280, 720, keyfn=evk.KEY_REF_HASH8192,
                                              algo=algo_id(evk, algo)))
            keys, reps, first = h.get_voxels()
            assert (U, R) == (1, 1) and keys.tolist() == [0] and first.tolist() == [0]
            assert (int(reps["x"][0]), int(reps["y"][0])) == (0, 0)


@pytest.mark.parametrize("algo", ALGOS)
def test_f3_real_events(evk, orc, golden, f3_events, algo):
    g = golden["F3"]
    ev = f3_events
    with evk.Evk(1024) as h:
        h.load_csv(os.path.join(GOLDEN_DIR, "event_raw_data8.csv"))
        assert h.get_events().tobytes() == ev.tobytes()
        ok, of, orr = run_ds(evk, orc, h, ev, 1280, 720, 1, 1, 0, 0, algo, keyfn=1)
        assert (len(ok), orr) == (315, 5) and ok.tolist() == g["ref_hash"]["keys"]
        for c in g["cases"]:
            ok, of, orr = run_ds(evk, orc, h, ev, 1280, 720, c["vx"], c["vy"], c["vt"],
                                 c["use_p"], algo)
            assert (len(ok), orr) == (c["unique"], c["repeated"])
            assert sha(np.sort(ok)) == c["keys_sorted_sha"] and sha(of) == c["first_sha"]


@pytest.mark.parametrize("D", [2, 3, 4])
def test_f3_kmeans(evk, orc, golden, f3_events, D):
    g = golden["F3"][f"kmeans_D{D}"]
    with evk.Evk(1024) as h:
        h.load_events(f3_events)
        h.downsample(evk.ds_params(1280, 720, 4, 4, 1000, 0, 1))
        km = evk.km_params(4, D, iters=3)
        h.init_centroids_first_k(km)
        assert h.kmeans(km) == 3
        lab = h.get_labels()
        cent, counts = h.get_centroids(4, D)
    assert lab.tolist() == g["labels"] and counts.tolist() == g["counts"]
    np.testing.assert_allclose(cent.ravel(), g["centroids"], rtol=CENT_RTOL)


def test_synth_golden_streams(evk, orc, golden):
    for s in golden["synth"]:
        sp = evk.synth_params(s["seed"], s["n"], s["W"], s["H"], s["rate"], s["blobs"],
                              first_index=s["first_index"])
        vx, vy, vt, up = s["vox"]
        with evk.Evk(s["n"]) as h:
            h.synth(sp)
            ev = h.get_events()
            assert sha(ev) == s["events_sha"], "device generator differs from the golden stream"
            for algo in ALGOS:
                t0 = 0
                U, R = h.downsample(evk.ds_params(s["W"], s["H"], vx, vy, vt, t0, up,
                                                  algo=algo_id(evk, algo)))
                keys, reps, first = h.get_voxels()
                assert (U, R) == (s["unique"], s["repeated"]), algo
                assert sha(np.sort(keys)) == s["keys_sorted_sha"] and sha(first) == s["first_sha"]
            km = evk.km_params(s["K"], 2, iters=3)
            h.init_centroids_first_k(km)
            h.kmeans(km)
            lab = h.get_labels()
            cent, counts = h.get_centroids(s["K"], 2)
        assert sha(lab) == s["labels_sha"] and counts.tolist() == s["counts"]
        np.testing.assert_allclose(cent.ravel(), s["centroids"], rtol=CENT_RTOL)


# ------------------------------------------------------------------------- GPU vs oracle ------
CONFIGS = {
    # name: (seed, n, rate, W, H, blobs, (vx,vy,vt,p), K, iters)   -- reduced-N BASELINE configs
    "C1_davis_1M": (0xE7CA0001, 1_000_000, 10_000_000, 346, 260, 8, (4, 4, 1000, 1), 8, 3),
    "C2_davis_2M": (0xE7CA0002, 2_000_000, 10_000_000, 346, 260, 32, (4, 4, 1000, 1), 32, 20),
    "C3_gen4_3M": (0xE7CA0003, 3_000_000, 100_000_000, 1280, 720, 64, (2, 2, 500, 1), 64, 2),
}


@pytest.mark.parametrize("name", list(CONFIGS))
def test_configs_vs_oracle(evk, orc, name):
    seed, n, rate, W, H, blobs, (vx, vy, vt, up), K, iters = CONFIGS[name]
    ev = orc.synth(orc.synth_params(seed, n, W, H, rate, blobs), threads=orc.max_threads())
    ok, of, orr = orc.downsample(ev, orc.ds_params(W, H, vx, vy, vt, 0, up))
    pts = orc.points(ev, of, 2)
    oc, ol, ocnt, _ = orc.kmeans(pts, pts[:K], iters=iters, threads=orc.max_threads())
    prev = pts[:K].copy()
    with evk.Evk(n) as h:
        h.synth(evk.synth_params(seed, n, W, H, rate, blobs))
        assert h.get_events().tobytes() == ev.tobytes()
        for algo in ALGOS:
            U, R = h.downsample(evk.ds_params(W, H, vx, vy, vt, 0, up, algo=algo_id(evk, algo)))
            assert (U, R) == (len(ok), orr), algo
            keys, reps, first = h.get_voxels()
            assert (keys == ok).all() and (first == of).all(), algo
            assert reps.tobytes() == ev[of].tobytes()
            if algo == "SLAB":
                assert h.stage_times().ds_algo_used == evk.ALGO_SLAB  # no silent fallback
        km = evk.km_params(K, 2, iters=iters)
        h.init_centroids_first_k(km)
        c_init, _ = h.get_centroids(K, 2)
        assert (c_init == prev).all(), "first-K initialisation differs"
        assert h.kmeans(km) == iters
        lab = h.get_labels()
        cent, counts = h.get_centroids(K, 2)
    # labels of the last iteration were assigned against the centroids of iteration iters-1
    oc_prev, _, _, _ = orc.kmeans(pts, pts[:K], iters=iters - 1) if iters > 1 else (prev, 0, 0, 0)
    check_labels(pts, oc_prev, lab, ol)
    assert (counts == ocnt).all()
    np.testing.assert_allclose(cent, oc, rtol=CENT_RTOL, atol=0)


def test_kmeans_extensions(evk, orc):
    """D = 3 / 4, distance gate, tolerance stop, raw-event clustering"""
    n, W, H = 300_000, 346, 260
    ev = orc.synth(orc.synth_params(0xE7CA0001, n, W, H, 10_000_000, 8))
    ok, of, _ = orc.downsample(ev, orc.ds_params(W, H, 4, 4, 1000, 0, 1))
    with evk.Evk(n) as h:
        h.load_events(ev)
        h.downsample(evk.ds_params(W, H, 4, 4, 1000, 0, 1))
        for D, md, ts in [(3, 0.0, 1e-3), (4, 0.0, 1e-2), (2, 40.0, 1e-3), (3, 60.0, 5e-3)]:
            pts = orc.points(ev, of, D, 0, ts, 25.0)
            km = evk.km_params(8, D, max_dist=md, iters=4, t_scale=ts, p_scale=25.0)
            h.init_centroids_first_k(km)
            h.kmeans(km)
            lab = h.get_labels()
            cent, counts = h.get_centroids(8, D)
            oc3, _, _, _ = orc.kmeans(pts, pts[:8], md, iters=3)
            oc, ol, ocnt, _ = orc.kmeans(pts, pts[:8], md, iters=4)
            check_labels(pts, oc3, lab, ol)
            if md > 0:
                assert (lab < 0).any()  # the gate leaves points unassigned (label -1)
            assert (counts == ocnt).all()
            np.testing.assert_allclose(cent, oc, rtol=CENT_RTOL, atol=1e-6)
        # tolerance stop (the reference loops while error_max > 10, assign_to_centers2.c:545)
        pts = orc.points(ev, of, 2)
        km = evk.km_params(8, 2, iters=50, tol=0.5)
        h.init_centroids_first_k(km)
        it = h.kmeans(km)
        oc, ol, ocnt, oit = orc.kmeans(pts, pts[:8], iters=50, tol=0.5)
        assert it == oit and 1 < it < 50
        cent, counts = h.get_centroids(8, 2)
        assert (counts == ocnt).all() and np.allclose(cent, oc, rtol=CENT_RTOL, atol=0)
        # cluster the raw events instead of the voxels
        pts = orc.points(ev, None, 2)
        km = evk.km_params(8, 2, iters=2, on_events=1)
        h.set_centroids(pts[:8])
        h.kmeans(km)
        lab = h.get_labels()
        cent, counts = h.get_centroids(8, 2)
        oc1, _, _, _ = orc.kmeans(pts, pts[:8], iters=1)
        oc, ol, ocnt, _ = orc.kmeans(pts, pts[:8], iters=2)
        check_labels(pts, oc1, lab, ol)
        assert (counts == ocnt).all() and np.allclose(cent, oc, rtol=CENT_RTOL, atol=0)


def test_unordered_stream_falls_back_to_table(evk, orc):
    """the slab kernel requires a stream partitioned by time bin; anything else must be detected
    and handled by the general paths (stable partition by time bin + slab kernel, else the
    table) with identical results"""
    n, W, H = 200_000, 346, 260
    ev = orc.synth(orc.synth_params(0xE7CA0001, n, W, H, 10_000_000, 8))
    rng = np.random.default_rng(7)
    cases = {"shuffled": ev[rng.permutation(n)], "reversed": ev[::-1].copy()}
    swapped = ev.copy()
    swapped[[1000, 150_000]] = swapped[[150_000, 1000]]  # a single pair out of place
    cases["one_swap"] = swapped
    late = ev.copy()
    late["t"][:10] = -5  # before t0: gated out
    cases["before_t0"] = late
    with evk.Evk(n) as h:
        for name, e in cases.items():
            h.load_events(e)
            for algo in ("AUTO", "SLAB", "PARTITION", "TABLE", "SORT"):
                run_ds(evk, orc, h, e, W, H, 4, 4, 1000, 1, algo)
                if algo in ("AUTO", "SLAB", "PARTITION"):
                    assert h.stage_times().ds_algo_used == evk.ALGO_PARTITION, name
            # a key space the partition does not take (no time bins): the table
            run_ds(evk, orc, h, e, W, H, 4, 4, 0, 1, "PARTITION")
            assert h.stage_times().ds_algo_used == evk.ALGO_TABLE, name


def test_edge_cases(evk, orc):
    W, H = 346, 260
    with evk.Evk(4096) as h:
        for algo in ALGOS:
            a = algo_id(evk, algo)
            # empty input
            h.load_events(np.zeros(0, dtype=evk.EVENT_DTYPE))
            assert h.downsample(evk.ds_params(W, H, 4, 4, 1000, 0, 1, algo=a)) == (0, 0)
            assert len(h.get_voxels()[0]) == 0
            # every event gated out
            ev = orc.events_from_xy([346, 400, 10], [10, 10, 300], t=[0, 5, 9])
            h.load_events(ev)
            assert h.downsample(evk.ds_params(W, H, 4, 4, 1000, 0, 1, algo=a)) == (0, 0)
            # single event, maximal duplication, all distinct, sensor corner
            for ev in (orc.events_from_xy([345], [259], t=[17], p=[1]),
                       orc.events_from_xy(np.full(4096, 7), np.full(4096, 9)),
                       orc.events_from_xy(np.arange(4096) % 346, np.arange(4096) // 346,
                                          t=np.arange(4096) * 1000)):
                h.load_events(ev)
                run_ds(evk, orc, h, ev, W, H, 4, 4, 1000, 1, algo)
                run_ds(evk, orc, h, ev, W, H, 1, 1, 0, 0, algo)      # exact pixel dedup (a12)
                run_ds(evk, orc, h, ev, W, H, 3, 5, 777, 1, algo)    # non-power-of-two divisors
                run_ds(evk, orc, h, ev, W, H, 4, 4, 1000, 1, algo, t0=-123)
        # soa loader == aos loader
        ev = orc.synth(orc.synth_params(1, 3000, W, H, 1_000_000, 4))
        h.load_events_soa(ev["x"], ev["y"], ev["t"], (ev["p"] > 0).astype(np.uint8))
        assert h.get_events().tobytes() == ev.tobytes()
        # capacity and state errors are reported, never fatal
        with pytest.raises(evk.EvkError) as e:
            h.load_events(np.zeros(5000, dtype=evk.EVENT_DTYPE))
        assert e.value.status == -5
        with pytest.raises(evk.EvkError) as e:
            h.kmeans(evk.km_params(8))
        assert e.value.status == -4
        with pytest.raises(evk.EvkError) as e:
            h.downsample(evk.ds_params(0, 10))
        assert e.value.status == -1


def test_streaming_windows(evk, orc):
    """50 ms slices as ACCEL/store.cpp:329,349-352: per-window downsample + warm-started k-means"""
    n, W, H, K = 400_000, 346, 260, 8
    ev = orc.synth(orc.synth_params(0xE7CA0005, n, W, H, 2_000_000, K))  # 0.2 s => 4 windows
    win = 50_000
    ds = evk.ds_params(W, H, 4, 4, 1000, 0, 1)
    km = evk.km_params(K, 2, iters=2)
    with evk.Evk(n) as h:
        h.window_config(ds, km, win)
        done = 0
        cents = []
        # feed in uneven chunks, as the SDK callback does (ACCEL/store.cpp:614-615)
        edges = [0, 12345, 100_000, 100_001, 250_000, n]
        for a, b in zip(edges[:-1], edges[1:]):
            d = h.window_push(ev[a:b])
            done += d
            if d:
                cents.append(h.get_centroids(K, 2)[0].copy())
        done += h.window_flush()
        cents.append(h.get_centroids(K, 2)[0].copy())
    assert done == 4
    # oracle: same windows, warm start
    cent = None
    last = None
    for w in range(4):
        m = (ev["t"] >= w * win) & (ev["t"] < (w + 1) * win)
        e = ev[m]
        ok, of, _ = orc.downsample(e, orc.ds_params(W, H, 4, 4, 1000, w * win, 1))
        pts = orc.points(e, of, 2)
        if cent is None:
            cent = pts[:K].copy()
        cent, _, _, _ = orc.kmeans(pts, cent, iters=2)
        last = cent
    np.testing.assert_allclose(cents[-1], last, rtol=CENT_RTOL, atol=0)


def test_streaming_windows_by_event_count(evk, orc):
    """the reslicer's make_n_events condition (SAMP/store.cpp:336): windows of exactly n events,
    time bins starting at the window's first event, centroids warm-started across windows"""
    n, W, H, K, per = 330_000, 346, 260, 8, 100_000
    ev = orc.synth(orc.synth_params(0xE7CA0006, n, W, H, 2_000_000, K))
    ds = evk.ds_params(W, H, 4, 4, 1000, 0, 1)
    km = evk.km_params(K, 2, iters=2)
    got = []
    with evk.Evk(per) as h:
        h.window_config_events(ds, km, per)
        done = 0
        for a, b in zip([0, 7, 99_999, 100_000, 250_001], [7, 99_999, 100_000, 250_001, n]):
            d = h.window_push(ev[a:b])
            done += d
            if d:
                got.append((h.get_voxels(reps=False)[0].copy(), h.get_centroids(K, 2)[0].copy()))
        assert done == 3
        assert h.window_flush() == 1            # the 30 000-event remainder
        got.append((h.get_voxels(reps=False)[0].copy(), h.get_centroids(K, 2)[0].copy()))
        with pytest.raises(evk.EvkError):
            h.window_config_events(ds, km, per + 1)   # beyond the handle capacity
    cent = None
    for w in range(4):
        e = ev[w * per:(w + 1) * per]
        ok, of, _ = orc.downsample(e, orc.ds_params(W, H, 4, 4, 1000, int(e["t"][0]), 1))
        assert (got[w][0] == ok).all()
        pts = orc.points(e, of, 2)
        if cent is None:
            cent = pts[:K].copy()
        cent, _, _, _ = orc.kmeans(pts, cent, iters=2)
        np.testing.assert_allclose(got[w][1], cent, rtol=CENT_RTOL, atol=0)


def test_full_size_properties(evk, orc):
    """BASELINE config C3 at full size (100 M Gen4 events): size-independent properties —
    slab and table agree on counts and on an order-independent checksum of the key set, the
    downsample is idempotent (downsampling the representatives returns the same set), and the
    oracle agrees on a 2 M-event prefix."""
    n, W, H = 100_000_000, 1280, 720
    sp = evk.synth_params(0xE7CA0003, n, W, H, 100_000_000, 64)

    def mix(k):
        k = k.copy()
        with np.errstate(over="ignore"):
            k ^= k >> np.uint64(33); k *= np.uint64(0xFF51AFD7ED558CCD)
            k ^= k >> np.uint64(33)
        return int(np.bitwise_xor.reduce(k)), int(k.sum(dtype=np.uint64))

    with evk.Evk(n) as h:
        h.synth(sp)
        res = {}
        for algo in ("SLAB", "TABLE"):
            U, R = h.downsample(evk.ds_params(W, H, 2, 2, 500, 0, 1, algo=algo_id(evk, algo)))
            assert h.stage_times().ds_algo_used == algo_id(evk, algo)
            keys, reps, first = h.get_voxels(reps=(algo == "SLAB"))
            assert (np.diff(first.astype(np.int64)) > 0).all()  # canonical order, distinct reps
            res[algo] = (U, R, mix(keys), sha(first))
            if algo == "SLAB":
                reps_slab = reps
        assert res["SLAB"] == res["TABLE"]
        U = res["SLAB"][0]
        # idempotence: the representatives are one event per voxel
        h.load_events(reps_slab)
        U2, R2 = h.downsample(evk.ds_params(W, H, 2, 2, 500, 0, 1))
        assert (U2, R2) == (U, 0)
        keys2, _, first2 = h.get_voxels(reps=False)
        assert mix(keys2) == res["SLAB"][2] and (first2 == np.arange(U, dtype=np.uint32)).all()
        # prefix vs oracle
        h.synth(evk.synth_params(0xE7CA0003, 2_000_000, W, H, 100_000_000, 64))
        ev = h.get_events()
        ok, of, orr = orc.downsample(ev, orc.ds_params(W, H, 2, 2, 500, 0, 1))
        Up, Rp = h.downsample(evk.ds_params(W, H, 2, 2, 500, 0, 1))
        keys, _, first = h.get_voxels(reps=False)
        assert (Up, Rp) == (len(ok), orr) and (keys == ok).all() and (first == of).all()


def test_cpp_host_replay(evk, orc):
    """the C++ host layer (host/evk.hpp + store_replay.cpp) replays 50 ms slices through the C-ABI"""
    import subprocess
    exe = os.path.join(evk_loader.PKG_DIR, "store_replay")
    assert os.path.exists(exe), "host demo not built"
    r = subprocess.run([exe, "synth:1500000", "8"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("slice ")]
    assert len(lines) == 3 and r.stdout.strip().endswith("3 slices")  # 0.15 s at 10 Mev/s
    # slice 0 against the oracle: same stream, same 50 ms window
    ev = orc.synth(orc.synth_params(0xE7CA0005, 1_500_000, 1280, 720, 10_000_000, 8))
    e0 = ev[ev["t"] < 50_000]
    ok, of, orr = orc.downsample(e0, orc.ds_params(1280, 720, 4, 4, 1000, 0, 1))
    assert f"events={len(e0)} unique={len(ok)} repeated={orr}" in lines[0]
