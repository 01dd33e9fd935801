"""Pins the oracle (and, on the GPU box, the CUDA path) to the REFERENCE's own kernels.

oracle/_ref/libref.so is coordinate_processor.cl + assign_to_centers.cl of the reference compiled
as C where they lie (oracle/Makefile `ref`, oracle/cl_shim.h); tests/golden/ref_kernel_golden.json
holds vectors produced by it (tests/golden/make_ref_golden.py).  The golden tests run everywhere;
the live tests run wherever the library exists (built here, it travels with the snapshot)."""
import hashlib
import json
import os

import numpy as np
import pytest

import evk_loader
from oracle import ref

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CENTERS = np.array([1, 1, 10, 10, 20, 20, 30, 30, 50, 50, 60, 60, 70, 70, 80, 80], np.float32)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def rg():
    with open(os.path.join(GOLDEN_DIR, "ref_kernel_golden.json")) as f:
        return json.load(f)


def random_coords(seed, n, lo=-40, hi_x=1400, hi_y=800):
    rng = np.random.default_rng(seed)
    c = np.empty(2 * n, dtype=np.int32)
    c[0::2] = rng.integers(lo, hi_x, size=n)
    c[1::2] = rng.integers(lo, hi_y, size=n)
    c[:8] = [1280, 720, 1281, 0, 0, 721, -1, 5]
    return c


def f3_coords():
    rows = np.loadtxt(os.path.join(GOLDEN_DIR, "event_raw_data8.csv"), delimiter=",",
                      dtype=np.int64)
    return np.ascontiguousarray(rows[:, :2], dtype=np.int32).ravel()


def to_ref_labels(lab):
    """contract labels (k or -1) in the reference's encoding (2k or 255)"""
    return np.where(lab < 0, 255, 2 * lab).astype(np.int32)


def contract_downsample(orc, coords):
    """the contract's REF_HASH8192 downsample on the kernel's input layout -> (pairs in canonical
    order, unique, repeated); coordinates outside uint16 are gated by the kernel and dropped here"""
    xy = coords.reshape(-1, 2)
    keep = (xy[:, 0] >= 0) & (xy[:, 0] <= 65535) & (xy[:, 1] >= 0) & (xy[:, 1] <= 65535)
    ev = orc.events_from_xy(xy[keep, 0], xy[keep, 1])
    keys, first, rep = orc.downsample(ev, orc.ds_params(1280, 720, keyfn=orc.KEY_REF_HASH8192))
    return np.stack([ev["x"][first], ev["y"][first]], 1).astype(np.int32), len(keys), rep


# ------------------------------------------------------------------ oracle vs golden vectors ---
def test_oracle_matches_reference_kernel_golden(orc, rg):
    cases = [("F2", np.zeros(16384, np.int32))] + [("F3", f3_coords())] + \
            [(c, random_coords(c["seed"], c["n"])) for c in rg["random"]]
    for tag, coords in cases:
        want = rg[tag] if isinstance(tag, str) else tag
        lit, uc, rc = orc.ref_process_coordinates(coords)          # literal restatement
        con, cu, cr = contract_downsample(orc, coords)             # contract (canonical order)
        assert (uc, rc) == (want["unique_count"], want["repeated_count"]) == (cu, cr)
        assert (lit == con).all()
        if "unique_sha" in want:
            assert sha(lit) == want["unique_sha"]
        else:
            assert lit.tolist() == want["unique"]
    # cumulative counters (the kernel never resets them)
    _, uc, rc = orc.ref_process_coordinates(f3_coords(), rg["F3"]["unique_count"],
                                            rg["F3"]["repeated_count"])
    assert (uc, rc) == (rg["F3_second_launch"]["unique_count"],
                        rg["F3_second_launch"]["repeated_count"])
    # k-means assign on the reference's own data: squared-distance contract == length() compare
    data = (np.arange(4096) % 100).astype(np.float32).reshape(-1, 2)
    for use_sqrt in (False, True):
        lab = orc.kmeans_assign(data, CENTERS.reshape(8, 2), 50.0, use_sqrt=use_sqrt)
        assert sha(to_ref_labels(lab)) == rg["F1"]["assign_sha"]
    assert np.bincount(lab, minlength=8).tolist() == rg["F1"]["cluster_index"]
    sums = [(float(data[lab == k, 0].sum()), float(data[lab == k, 1].sum())) for k in range(8)]
    assert [s[0] for s in sums] == rg["F1"]["sum_x"] and [s[1] for s in sums] == rg["F1"]["sum_y"]
    edge = np.array([500, 500, 5.5, 5.5, 15, 15, 0, 0, 40, 40, 65, 65], np.float32).reshape(-1, 2)
    assert to_ref_labels(orc.kmeans_assign(edge, CENTERS.reshape(8, 2), 50.0)).tolist() == \
        rg["assign_edge"]


# ------------------------------------------------------------ oracle vs the kernels, live ------
@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libref.so not built (needs /root/reference)")
def test_oracle_matches_reference_kernel_live(orc):
    rng = np.random.default_rng(99)
    for trial in range(40):
        n = int(rng.integers(1, 30000))
        coords = random_coords(1000 + trial, n, lo=int(rng.integers(-50, 1)),
                               hi_x=int(rng.integers(30, 1500)), hi_y=int(rng.integers(30, 900)))
        ru, ruc, rrc = ref.process_coordinates(coords)
        lit, uc, rc = orc.ref_process_coordinates(coords)
        con, cu, cr = contract_downsample(orc, coords)
        assert (ruc, rrc) == (uc, rc) == (cu, cr)
        assert (ru == lit).all() and (ru == con).all()
    for trial in range(20):
        pts = rng.uniform(-20, 140, size=(2048, 2)).astype(np.float32)
        if trial % 2:
            pts = np.round(pts)  # integer coordinates: exact ties between centres
        cent = rng.uniform(0, 100, size=(8, 2)).astype(np.float32)
        if trial % 4 == 3:
            cent[5] = cent[2]  # duplicate centre: the lower index must win
        ra = ref.assign_to_centers(pts.ravel(), cent.ravel())
        for use_sqrt in (True, False):
            lab = to_ref_labels(orc.kmeans_assign(pts, cent, 50.0, use_sqrt=use_sqrt))
            bad = np.nonzero(lab != ra)[0]
            # the squared-distance contract may differ from length() only where two candidate
            # distances tie within 1e-6 relative (BASELINE.json) or at the threshold itself
            for i in bad:
                d = np.sqrt(((cent.astype(np.float64) - pts[i]) ** 2).sum(1))
                cand = [d[k // 2] if k != 255 else 50.0 for k in (int(lab[i]), int(ra[i]))]
                assert abs(cand[0] - cand[1]) <= 1e-6 * max(cand), (trial, i, lab[i], ra[i], cand)
            if use_sqrt:
                assert len(bad) == 0, "length()-mode oracle must equal the kernel exactly"
        out, ci = ref.assign_data_cluster(pts.ravel(), ra)
        lab = orc.kmeans_assign(pts, cent, 50.0, use_sqrt=True)
        assert np.bincount(lab[lab >= 0], minlength=8).tolist() == ci.tolist()


# ------------------------------------------------------------- CUDA vs the reference vectors ---
@pytest.mark.gpu
def test_cuda_matches_reference_kernel_golden(orc, rg):
    evk = evk_loader.load()
    evk.lib()
    cases = [("F2", np.zeros(16384, np.int32))] + [("F3", f3_coords())] + \
            [(c, random_coords(c["seed"], c["n"])) for c in rg["random"]]
    with evk.Evk(32768) as h:
        for tag, coords in cases:
            want = rg[tag] if isinstance(tag, str) else tag
            h.load_coords_i32(coords)
            for algo in (evk.ALGO_AUTO, evk.ALGO_TABLE, evk.ALGO_SORT):
                U, R = h.downsample(evk.ds_params(1280, 720, keyfn=evk.KEY_REF_HASH8192, algo=algo))
                _, reps, _ = h.get_voxels()
                pairs = np.stack([reps["x"], reps["y"]], 1).astype(np.int32)
                assert (U, R) == (want["unique_count"], want["repeated_count"])
                if "unique_sha" in want:
                    assert sha(pairs) == want["unique_sha"]
                else:
                    assert pairs.tolist() == want["unique"]
        data = (np.arange(4096) % 100).astype(np.int32)
        h.load_coords_i32(data)
        h.set_centroids(CENTERS.reshape(8, 2))
        h.kmeans(evk.km_params(8, 2, max_dist=50.0, iters=1, on_events=1))
        lab = h.get_labels()
        cent, counts = h.get_centroids(8, 2)
        assert sha(to_ref_labels(lab)) == rg["F1"]["assign_sha"]
        assert counts.tolist() == rg["F1"]["cluster_index"]
        want_c = np.array([[sx / n, sy / n] for sx, sy, n in
                           zip(rg["F1"]["sum_x"], rg["F1"]["sum_y"], rg["F1"]["cluster_index"])])
        np.testing.assert_allclose(cent, want_c, rtol=1e-5)
