"""CPU tests: the C oracle against the committed golden fixtures (tests/golden/golden.json, made
by the independent numpy restatement tests/golden/make_golden.py) and against the known answers
of SURVEY.md 8c that come from the reference's own deterministic inputs."""
import hashlib

import numpy as np
import pytest


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


C0 = np.array([1, 1, 10, 10, 20, 20, 30, 30, 50, 50, 60, 60, 70, 70, 80, 80],
              np.float32).reshape(8, 2)  # KM/assign_to_centers2.c:131


def f1_points():
    return (np.arange(4096) % 100).astype(np.float32).reshape(-1, 2)  # :121-129


def test_f1_contract(orc, golden):
    g = golden["F1"]
    pts = f1_points()
    lab = orc.kmeans_assign(pts, C0, 50.0)
    assert lab.tolist() == g["labels"] and g["unassigned"] == 0
    # sqrt form (what the reference computes, assign_to_centers.cl:17-21) gives the same labels
    assert (orc.kmeans_assign(pts, C0, 50.0, use_sqrt=True) == lab).all()
    cent, counts, sums, _ = orc.kmeans_update(pts, lab, C0)
    assert counts.tolist() == [123, 205, 205, 287, 328, 205, 205, 490] == g["counts"]
    assert sums[:, 0].tolist() == [246, 2050, 4100, 9184, 15416, 12300, 14350, 42610] == g["sum_x"]
    assert sums[:, 1].tolist() == [369, 2255, 4305, 9471, 15744, 12505, 14555, 43100] == g["sum_y"]
    want = [2, 3, 10, 11, 20, 21, 32, 33, 47, 48, 60, 61, 70, 71, 86.95918367, 87.95918367]
    np.testing.assert_allclose(cent.ravel(), want, rtol=1e-7)
    np.testing.assert_allclose(cent.ravel(), g["centroids"], rtol=0, atol=0)


def test_f1_literal_quirks(orc, golden):
    """the reference exactly as written (stride-2 centroid indexing, int abs, selective
    overwrite): SURVEY.md 8c 'quirks' values"""
    data = (np.arange(4096) % 100).astype(np.float32)
    r = orc.ref_kmeans_trip(data, C0.ravel())
    want = [2, 3, 1.8, 10, 10, 11, 7.857142857, 14.28571429, 12.5, 13.125, 21, 44.8, 44.8, 46.2,
            19.32857143, 31.46122449]
    np.testing.assert_allclose(r["new_centroids"], want, rtol=1e-6)
    np.testing.assert_allclose(r["new_centroids"], golden["F1"]["quirk_centroids"], rtol=1e-6)
    assert r["cluster_index"].tolist() == golden["F1"]["counts"]
    # labels are 2k / 255 in the reference (assign_to_centers.cl:12,22,26)
    assert (r["assign"] // 2).tolist() == golden["F1"]["labels"]
    assert r["error_max"] == 60.0  # > 10 => the reference would goto KERNEL_RESTART (:545)
    # second trip re-uploads the stale `output` (D13): sums are polluted, still deterministic
    r2 = orc.ref_kmeans_trip(data, r["centroids"], r["output"])
    assert np.isfinite(r2["scalar_sum"]).all()


def test_f2_warmup(orc, golden):
    uq, uc, rc = orc.ref_process_coordinates(np.zeros(16384, np.int32))
    assert uq.tolist() == [[0, 0]] and (uc, rc) == (1, 1)
    # counters are cumulative in the reference (never reset, ACCEL/store.cpp:267-268)
    _, uc2, rc2 = orc.ref_process_coordinates(np.zeros(16384, np.int32), uc, rc)
    assert (uc2, rc2) == (2, 2)
    ev = orc.events_from_xy(np.zeros(8192), np.zeros(8192))
    k, f, rep = orc.downsample(ev, orc.ds_params(1280, 720, keyfn=orc.KEY_REF_HASH8192))
    g = golden["F2"]
    assert (k.tolist(), f.tolist(), rep) == (g["keys"], g["first"], g["repeated"]) == ([0], [0], 1)


def test_f3_real_events(orc, golden, f3_events):
    g = golden["F3"]
    ev = f3_events
    assert len(ev) == 320 == g["rows"]
    assert g["sha256"] == "207f866e73c7ec85320d6f00ec0b50f73298e39c0956e74363cda3f6657592d8"
    assert (ev["p"] == 0).sum() == 139 and ev["t"].min() == 2458 and ev["t"].max() == 2808
    xy = np.stack([ev["x"], ev["y"]], 1).astype(np.int32).ravel()
    xs, ys, cs = orc.ref_analyze_coordinates(xy)
    assert len(xs) == 320 == g["exact_xy"] and (cs == 1).all()
    uq, uc, rc = orc.ref_process_coordinates(xy)
    assert (uc, rc) == (315, 5)
    k, f, rep = orc.downsample(ev, orc.ds_params(1280, 720, keyfn=orc.KEY_REF_HASH8192))
    assert k.tolist() == g["ref_hash"]["keys"] and f.tolist() == g["ref_hash"]["first"]
    assert rep == 5
    # the contract's canonical order reproduces the literal kernel's output array
    assert (np.stack([ev["x"][f], ev["y"][f]], 1) == uq).all()
    expect = {(4, 4, 1000, 0): 257, (4, 4, 1000, 1): 261, (2, 2, 500, 0): 306, (2, 2, 500, 1): 306,
              (4, 4, 100, 0): 303, (4, 4, 100, 1): 306, (1, 1, 0, 0): 320, (1, 1, 0, 1): 320}
    for c in g["cases"]:
        p = orc.ds_params(1280, 720, c["vx"], c["vy"], c["vt"], 0, c["use_p"])
        k, f, rep = orc.downsample(ev, p)
        assert len(k) == c["unique"] == expect[(c["vx"], c["vy"], c["vt"], c["use_p"])]
        assert rep == c["repeated"]
        assert sha(np.sort(k)) == c["keys_sorted_sha"] and sha(f) == c["first_sha"]
        if c["keys"] is not None:
            assert k.tolist() == c["keys"]
        k2, f2, rep2 = orc.downsample(ev, p, threads=3)
        assert (k2 == k).all() and (f2 == f).all() and rep2 == rep


@pytest.mark.parametrize("D", [2, 3, 4])
def test_f3_kmeans(orc, golden, f3_events, D):
    g = golden["F3"][f"kmeans_D{D}"]
    ev = f3_events
    k, f, _ = orc.downsample(ev, orc.ds_params(1280, 720, 4, 4, 1000, 0, 1))
    pts = orc.points(ev, f, D)
    cent, lab, counts, it = orc.kmeans(pts, pts[:4], iters=3)
    assert it == 3 and lab.tolist() == g["labels"] and counts.tolist() == g["counts"]
    np.testing.assert_allclose(cent.ravel(), g["centroids"], rtol=1e-6)
    cent2, lab2, counts2, _ = orc.kmeans(pts, pts[:4], iters=3, threads=4)
    assert (lab2 == lab).all() and (counts2 == counts).all()
    np.testing.assert_allclose(cent2, cent, rtol=1e-6)


def test_synth_streams(orc, golden):
    for s in golden["synth"]:
        sp = orc.synth_params(s["seed"], s["n"], s["W"], s["H"], s["rate"], s["blobs"],
                              first_index=s["first_index"])
        ev = orc.synth(sp)
        assert sha(ev) == s["events_sha"]
        assert [[int(e["x"]), int(e["y"]), int(e["t"]), int(e["p"])] for e in ev[:8]] == s["head"]
        assert sha(orc.synth(sp, threads=4)) == s["events_sha"]
        assert (np.diff(ev["t"]) >= 0).all()
        vx, vy, vt, up = s["vox"]
        k, f, rep = orc.downsample(ev, orc.ds_params(s["W"], s["H"], vx, vy, vt, 0, up))
        assert (len(k), rep) == (s["unique"], s["repeated"])
        assert sha(np.sort(k)) == s["keys_sorted_sha"] and sha(f) == s["first_sha"]
        pts = orc.points(ev, f, 2)
        cent, lab, counts, _ = orc.kmeans(pts, pts[: s["K"]], iters=3)
        assert sha(lab) == s["labels_sha"] and counts.tolist() == s["counts"]
        np.testing.assert_allclose(cent.ravel(), s["centroids"], rtol=1e-6)


def test_edge_cases(orc):
    p = orc.ds_params(346, 260, 4, 4, 1000, 0, 1)
    ev = orc.events_from_xy([], [])
    k, f, rep = orc.downsample(ev, p)
    assert len(k) == 0 and rep == 0
    # gating: x >= W, y >= H, t < t0 are dropped in VOXEL mode; inclusive bounds in REF mode
    ev = orc.events_from_xy([345, 346, 10, 10], [259, 10, 260, 10], t=[0, 0, 0, -5])
    k, f, rep = orc.downsample(ev, p)
    assert f.tolist() == [0]
    pr = orc.ds_params(346, 260, keyfn=orc.KEY_REF_HASH8192)
    k, f, rep = orc.downsample(ev, pr)
    assert f.tolist() == [0, 1, 2, 3]  # x == W and y == H pass the inclusive gate (D8)
    ev2 = orc.events_from_xy([347, 10], [10, 261])
    assert len(orc.downsample(ev2, pr)[0]) == 0
    # maximal duplication and all-distinct
    ev = orc.events_from_xy(np.full(1000, 7), np.full(1000, 9))
    k, f, rep = orc.downsample(ev, p)
    assert f.tolist() == [0] and rep == 1
    xs = np.arange(1000) % 346
    ev = orc.events_from_xy(xs, np.arange(1000) // 346, t=np.arange(1000) * 1000)
    k, f, rep = orc.downsample(ev, p)
    assert len(k) == 1000 and rep == 0 and (f == np.arange(1000)).all()
    # ties: equidistant centroids -> lowest k; empty cluster keeps its centroid
    pts = np.array([[5, 5], [0, 0]], np.float32)
    cent = np.array([[10, 5], [0, 5], [100, 100]], np.float32)
    lab = orc.kmeans_assign(pts, cent)
    assert lab.tolist() == [0, 1]
    c2, counts, _, _ = orc.kmeans_update(pts, lab, cent)
    assert counts.tolist() == [1, 1, 0] and c2[2].tolist() == [100, 100]
    assert orc.kmeans_assign(pts, cent, 1.0).tolist() == [-1, -1]


def test_evt2_codec(orc):
    """RAW EVT 2.0 (third-party format restated in oracle/evk_oracle.h): known-answer words and
    encode -> decode round trips"""
    kw = np.array([0x80000003, (1 << 28) | (9 << 22) | (5 << 11) | 7,
                   (0 << 28) | (63 << 22) | (2047 << 11) | 2047, 0xA0000000, 0xE1234567,
                   0x80000004, (1 << 28) | (0 << 22) | (1 << 11) | 2], dtype=np.uint32)
    d = orc.evt2_decode(kw)
    assert [(int(e["x"]), int(e["y"]), int(e["p"]), int(e["t"])) for e in d] == \
        [(5, 7, 1, 3 * 64 + 9), (2047, 2047, 0, 3 * 64 + 63), (1, 2, 1, 4 * 64)]
    # CD words before the first EVT_TIME_HIGH decode with time-high 0
    assert int(orc.evt2_decode(np.array([(1 << 28) | (5 << 22)], np.uint32))["t"][0]) == 5
    ev = orc.synth(orc.synth_params(0xE7CA0003, 200_000, 1280, 720, 100_000_000, 64))
    w = orc.evt2_encode(ev)
    assert len(w) == len(ev) + len(np.unique(ev["t"] >> 6))  # one time-high word per 64 us step
    assert orc.evt2_decode(w).tobytes() == ev.tobytes()
    assert len(orc.evt2_decode(np.zeros(0, np.uint32))) == 0
    with pytest.raises(ValueError):
        orc.evt2_encode(orc.events_from_xy([2048], [0]))


def test_host_evt2_writer_matches_oracle_codec(orc):
    """<package>/evt2.py (the vectorised writer bench.py uses to make RAW input) == the oracle's
    encoder word for word, and the oracle decodes its output back to the same events."""
    import importlib.util
    import os
    import evk_loader
    spec = importlib.util.spec_from_file_location("evk_evt2", os.path.join(evk_loader.PKG_DIR, "evt2.py"))
    evt2 = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(evt2)
    for seed, n, W, H, rate in ((1, 100_000, 1280, 720, 100_000_000), (2, 5000, 346, 260, 10_000),
                                (3, 1, 346, 260, 10_000)):
        ev = orc.synth(orc.synth_params(seed, n, W, H, rate, 8))
        w = evt2.encode_evt2(ev)
        assert (w == orc.evt2_encode(ev)).all()
        assert orc.evt2_decode(w).tobytes() == ev.tobytes()
    assert len(evt2.encode_evt2(ev[:0])) == 0
    with pytest.raises(ValueError):
        evt2.encode_evt2(orc.events_from_xy([2048], [0]))


def test_evt3_codec_known_answers_and_round_trips(orc):
    """RAW EVT 3.0 (third-party format, SDK absent: parity unpinned by the reference): hand-derived
    known-answer words for every word kind (row, single event, vector base, VECT_12, VECT_8 with the
    running base, time low / high, a 2^24 us wrap, ignored types) and encode -> decode round trips"""
    kw = np.array([0x8001, 0x6005, 0x0007, 0x2000 | 0x800 | 5,      # th=1 tl=5 y=7, ON event at x=5
                   0x3000 | 10, 0x4000 | 0b101,                     # base 10 OFF: x=10, 12; base -> 22
                   0x5000 | 0x80,                                   # VECT_8: x = 22 + 7; base -> 30
                   0xA123, 0xE000, 0x7ABC,                          # trigger / others / continued
                   0x8000, 0x2003,                                  # th 1 -> 0: one wrap
                   0x4001], np.uint16)                              # vector again: x = 30
    d = orc.evt3_decode(kw)
    assert [(int(e["x"]), int(e["y"]), int(e["p"]), int(e["t"])) for e in d] == [
        (5, 7, 1, 4101), (10, 7, 0, 4101), (12, 7, 0, 4101), (29, 7, 0, 4101),
        (3, 7, 0, (1 << 24) + 5), (30, 7, 0, (1 << 24) + 5)]
    assert len(orc.evt3_decode(np.zeros(0, np.uint16))) == 0
    # events before any time / row word decode with zero state
    assert orc.evt3_decode(np.array([0x2000 | 9], np.uint16)).tolist() == [(9, 0, 0, 0, 0)]
    ev = orc.synth(orc.synth_params(0xE7CA0003, 200_000, 1280, 720, 100_000_000, 64))
    ev["p"] = ev["p"] > 0
    w = orc.evt3_encode(ev)
    assert orc.evt3_decode(w).tobytes() == ev.tobytes()
    # dense rows: sorted columns per (t, row, polarity) -> vector words incl. continued bases
    r = np.random.default_rng(1)
    n = 100_000
    t = np.sort(r.integers(0, 3000, n)).astype(np.int64)
    y, x, p = r.integers(0, 8, n), r.integers(0, 200, n), r.integers(0, 2, n)
    o = np.lexsort((x, p, y, t))
    dense = np.zeros(n, orc.EVENT_DTYPE)
    dense["t"], dense["y"], dense["x"], dense["p"] = t[o], y[o], x[o], p[o]
    w = orc.evt3_encode(dense)
    kinds = np.bincount(w >> 12, minlength=16)
    assert kinds[0x3] > 1000 and kinds[0x4] > 1000 and kinds[0x5] > 1000 and kinds[0x2] > 1000
    assert kinds[0x4] + kinds[0x5] > kinds[0x3]          # some vectors continue a running base
    assert orc.evt3_decode(w).tobytes() == dense.tobytes()
    # a stream longer than 2^24 us: wraps are counted from the EVT_TIME_HIGH words
    long_ev = np.zeros(5000, orc.EVENT_DTYPE)
    long_ev["t"] = np.arange(5000, dtype=np.int64) * 20_000          # 100 s > 5 wraps
    long_ev["x"], long_ev["y"] = np.arange(5000) % 1280, np.arange(5000) % 720
    assert orc.evt3_decode(orc.evt3_encode(long_ev)).tobytes() == long_ev.tobytes()
    with pytest.raises(ValueError):
        orc.evt3_encode(orc.events_from_xy([2048], [0]))
