"""Corner tracker pieces: time surface + corner test and the box-NMS of the corner
list (CornerFilter::filterCorners), pinned against the REFERENCE's own code -- its event callback
lambda and its CornerFilter class are compiled where they lie (oracle/_ref/libref_fct.so), the
committed digests (tests/golden/fct_golden.json) were made from that library.

CPU: oracle == reference (live, where the library exists) and oracle == golden digests.
GPU: the CUDA path (through the C-ABI) == oracle == golden digests."""
import hashlib
import json
import os

import numpy as np
import pytest

import evk_loader
import fct_cases
from oracle import fct

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fct_golden.json")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def golden():
    with open(GOLDEN) as f:
        return json.load(f)


def _xy(ev, idx):
    return np.stack([ev["x"][idx], ev["y"][idx]], 1).astype(np.int32)


def _kept_records(xy, kept):
    return np.concatenate([xy[kept], np.arange(len(kept), dtype=np.int32)[:, None]], 1).astype(np.int32)


# ------------------------------------------------------------------------------------ CPU ------
def test_oracle_equals_reference_callback_and_filter_golden(orc, golden):
    """orc_ts_corners / orc_filter_corners against the digests of the reference's own lambda and
    class: surfaces, corner lists (order included) and kept corners with their labels"""
    for (name, ev, chunk), g in zip(fct_cases.streams(orc), golden["streams"]):
        assert (name, len(ev), chunk) == (g["name"], g["n"], g["chunk"])
        surf = np.zeros((fct.H, fct.W), np.int64)
        for a, gr in zip(range(0, len(ev), chunk), g["ranges"]):
            e = ev[a:a + chunk]
            idx = orc.ts_corners(e, fct.W, fct.H, surf, literal_break=True)
            xy = _xy(e, idx)
            assert (len(xy), sha(xy)) == (gr["corners"], gr["corners_sha"]), (name, a)
            kept = fct.oracle_filter(xy, fct.W, fct.H, 15)
            rec = _kept_records(xy, kept)
            assert (len(rec), sha(rec)) == (gr["kept"], gr["kept_sha"]), (name, a)
        assert sha(surf) == g["surface_sha"]
    # the cropped streams produce corners; in the third one the literal `break` at the first event
    # near the border (FCT:948-955) ends most ranges early
    tot = [sum(r["corners"] for r in g["ranges"]) for g in golden["streams"]]
    assert tot[0] > 1000 and tot[1] > 1000 and tot[2] < 100


def test_oracle_filter_equals_reference_lists(orc, golden):
    for (name, xy, w, h, box), g in zip(fct_cases.filter_lists(), golden["filters"]):
        kept = fct.oracle_filter(xy, w, h, box)
        rec = _kept_records(xy, kept)
        assert (name, len(xy), len(rec), sha(rec)) == (g["name"], g["n"], g["kept"], g["kept_sha"])
        if fct.ref_available():   # live, where the reference library exists
            ref = fct.reference_filter(xy, w, h, box)
            assert ref.shape == rec.shape and (ref == rec).all(), name


def test_filter_known_answers(orc):
    """hand-derived: box 15 -> half 7: two corners conflict iff |dx| <= 14 and |dy| <= 14"""
    xy = np.array([[100, 100], [114, 100], [115, 100], [100, 114], [129, 100], [100, 115]], np.int32)
    kept = fct.oracle_filter(xy, 640, 480, 15)
    # 0 kept; 1 (dx 14) dropped; 2 (dx 15) kept; 3 (dy 14 from 0) dropped; 4: dx 14 from 2 -> dropped;
    # 5: dy 15 from 0, dx 15 from 2 -> kept
    assert kept.tolist() == [0, 2, 5]
    assert fct.oracle_filter(np.zeros((0, 2), np.int32), 640, 480, 15).tolist() == []
    assert fct.oracle_filter(np.array([[5, 5]] * 4, np.int32), 64, 64, 0).tolist() == [0]


@pytest.mark.skipif(not fct.ref_available(), reason="oracle/_ref/libref_fct.so not built")
def test_oracle_equals_reference_callback_live(orc):
    """random ranges with heavy timestamp ties through the reference's lambda and the oracle"""
    r = np.random.default_rng(3)
    surf_a = np.zeros((fct.H, fct.W), np.int64)
    surf_b = np.zeros((fct.H, fct.W), np.int64)
    for rep in range(6):
        n = 30_000
        cx, cy = r.integers(40, 1240), r.integers(40, 680)
        ev = np.zeros(n, orc.EVENT_DTYPE)
        ev["x"] = np.clip(cx + r.normal(0, 6, n).round(), 4, 1275)
        ev["y"] = np.clip(cy + r.normal(0, 6, n).round(), 4, 715)
        ev["t"] = 1000 * rep + r.integers(0, 40, n)        # few distinct stamps: ties on the circles
        c = fct.reference_callback(ev, surf_a, 1)
        idx = orc.ts_corners(ev, fct.W, fct.H, surf_b, literal_break=True)
        assert (surf_a == surf_b).all() and (c == _xy(ev, idx)).all()


# ------------------------------------------------------------------------------------ GPU ------
@pytest.fixture(scope="module")
def evk():
    m = evk_loader.load()
    m.lib()
    return m


@pytest.mark.gpu
def test_cuda_corner_pipeline_equals_reference_golden(evk, orc, golden):
    """stamp -> corner test -> box-NMS on the device, range by range, against the digests of the
    reference's own code"""
    for (name, ev, chunk), g in zip(fct_cases.streams(orc), golden["streams"]):
        with evk.Evk(chunk) as h:
            h.ts_create(fct.W, fct.H)
            for a, gr in zip(range(0, len(ev), chunk), g["ranges"]):
                e = ev[a:a + chunk]
                h.load_events(e)
                idx = h.ts_corners(literal_break=True)
                xy = _xy(e, idx)
                assert (len(xy), sha(xy)) == (gr["corners"], gr["corners_sha"]), (name, a)
                rec = h.ts_filter_corners(15)
                rec = np.stack([rec["x"], rec["y"], rec["label"]], 1).astype(np.int32)
                assert (len(rec), sha(rec)) == (gr["kept"], gr["kept_sha"]), (name, a)
            assert sha(h.ts_surface()) == g["surface_sha"]


@pytest.mark.gpu
def test_cuda_filter_lists_equal_oracle_and_golden(evk, orc, golden):
    with evk.Evk(1024) as h:
        for (name, xy, w, hh, box), g in zip(fct_cases.filter_lists(), golden["filters"]):
            rec = h.filter_corners(xy, w, hh, box)
            rec = np.stack([rec["x"], rec["y"], rec["label"]], 1).astype(np.int32).reshape(-1, 3)
            kept = fct.oracle_filter(xy, w, hh, box)
            assert (rec == _kept_records(xy, kept)).all(), name
            assert (len(rec), sha(rec)) == (g["kept"], g["kept_sha"]), name
        # edge cases: empty list, a single corner, corners outside the frame
        assert len(h.filter_corners(np.zeros((0, 2), np.int32), 64, 64, 15)) == 0
        assert h.filter_corners(np.array([[3, 3]], np.int32), 64, 64, 15)["label"].tolist() == [0]
        xy = np.array([[10, 10], [12, 12], [500, 10], [-40, 3], [11, 11]], np.int32)
        rec = h.filter_corners(xy, 64, 64, 15)
        kept = fct.oracle_filter(xy, 64, 64, 15)
        assert (np.stack([rec["x"], rec["y"]], 1) == xy[kept]).all()
        with pytest.raises(evk.EvkError):
            h.filter_corners(np.zeros((40000, 2), np.int32), 64, 64, 15)
