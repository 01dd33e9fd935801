"""DBSCAN on the downsampled cloud (reference: event-cam-clustering/
point-cloud-clustering/DBSCAN_simple.h, parameters as pcl_cluster.cpp:112-120).

Four statements of the same clustering must agree cluster for cluster (member lists; the order among
equal-size clusters is whatever the reference's std::sort leaves, so lists are canonicalised):
  Reference  DBSCAN_simple.h compiled where it lies against a PCL container shim
             (oracle/_ref/libref_dbscan.so)
  Oracle     oracle/dbscan_oracle.c, the literal seed-queue walk
  Contract   oracle/dbscan.py `contract`, the order-free form (core flags, components of core points,
             border rule, second memberships) -- the form the CUDA path implements
  CUDA       evk_dbscan_* through the C-ABI (csrc/evk_dbscan.cu)"""
import hashlib
import json
import os

import numpy as np
import pytest

import dbscan_cases as D
import evk_loader
from oracle import dbscan

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dbscan_golden.json")


@pytest.fixture(scope="module")
def gold():
    with open(GOLDEN) as f:
        return json.load(f)["cases"]


@pytest.fixture(scope="module")
def evk():
    m = evk_loader.load()
    m.lib()
    return m


def digest(clusters):
    h = hashlib.sha256()
    for c in D.canon(clusters):
        h.update(repr(c).encode())
    return h.hexdigest()


@pytest.mark.parametrize("name", list(D.CASES))
def test_oracle_and_contract_match_reference_golden(orc, gold, name):
    pts, (eps, mp, mn, mx) = D.cloud(name)
    lo, co, so = dbscan.oracle(pts, eps, mp, mn, mx)
    assert digest(co) == gold[name]["digest"]
    assert sorted(len(c) for c in co) == sorted(gold[name]["sizes"])
    assert [len(c) for c in co] == sorted((len(c) for c in co), reverse=True)
    lc, cc, sc = dbscan.contract(pts, eps, mp, mn, mx)
    assert digest(cc) == gold[name]["digest"]
    assert (lc == lo).all() and (sc == so).all()
    members = sum(len(c) for c in co)
    assert members - len(set(int(v) for c in co for v in c)) == gold[name]["second_memberships"]


def test_cases_cover_second_memberships_and_filters(gold):
    assert gold["tight_many_borders"]["second_memberships"] > 0
    assert len(gold["size_filter"]["sizes"]) < len(gold["blobs2d"]["sizes"])
    assert len(gold["app_parameters"]["sizes"]) >= 2


@pytest.mark.skipif(not dbscan.ref_available(), reason="oracle/_ref/libref_dbscan.so not built")
def test_oracle_matches_reference_live(orc):
    r = np.random.default_rng(77)
    for trial in range(6):
        n = 400 + 150 * trial
        eps, mp = [(4.0, 4), (6.0, 5), (3.0, 3), (9.0, 8), (5.0, 4), (2.5, 3)][trial]
        c = r.uniform(20, 180, size=(5, 2))
        pts = np.rint(c[r.integers(0, 5, n)] + r.normal(0, 6, size=(n, 2)))
        isn = r.random(n) < 0.3
        pts[isn] = np.rint(r.uniform(0, 200, size=(int(isn.sum()), 2)))
        ref = dbscan.reference(pts, eps, mp, 2, 300)
        lo, co, so = dbscan.oracle(pts, eps, mp, 2, 300)
        assert D.canon(co) == D.canon(ref)
        assert [len(x) for x in ref] == sorted((len(x) for x in ref), reverse=True)
        assert D.canon(dbscan.contract(pts, eps, mp, 2, 300)[1]) == D.canon(ref)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(D.CASES))
def test_cuda_matches_oracle_and_reference_golden(evk, orc, gold, name):
    pts, (eps, mp, mn, mx) = D.cloud(name)
    lo, co, so = dbscan.oracle(pts, eps, mp, mn, mx)
    with evk.Evk(1024) as h:
        labels, sizes, seeds, extra = h.dbscan_points(pts, eps, mp, mn, mx)
    assert (labels == lo).all()
    assert sizes.tolist() == [len(c) for c in co] and (seeds == so).all()
    got = D.clusters_from(labels, sizes, extra)
    assert D.canon(got) == D.canon(co)
    assert digest(got) == gold[name]["digest"]
    assert len(extra) == gold[name]["second_memberships"]


@pytest.mark.gpu
def test_cuda_voxel_cloud_and_errors(evk, orc):
    """the reference's use: voxel-grid downsample, then DBSCAN on what is left (pcl_cluster.cpp
    :53-57, 112-123) -- here on the voxel shard without leaving the device, D = 2 and D = 3"""
    W, H, n = 346, 260, 60_000
    ev = orc.synth(orc.synth_params(0xE7CA0001, n, W, H, 1_000_000, 8))
    ds = evk.ds_params(W, H, 2, 2, 20_000, 0, 0)
    ok, of, _ = orc.downsample(ev, orc.ds_params(W, H, 2, 2, 20_000, 0, 0))
    reps = ev[of]
    with evk.Evk(n) as h:
        with pytest.raises(evk.EvkError):
            h.dbscan_voxels(5.0, 4)                      # no voxel shard yet
        h.load_events(ev)
        U, _ = h.downsample(ds)
        assert U == len(ok) and U > 3000
        for D_, ts in ((2, 0.0), (3, 1e-3)):
            pts = np.zeros((U, 3), np.float32)
            pts[:, 0], pts[:, 1] = reps["x"], reps["y"]
            if D_ == 3:
                pts[:, 2] = (reps["t"].astype(np.float64) * ts).astype(np.float32)
            lo, co, so = dbscan.oracle(pts, 3.0, 6, 10, 100_000)
            labels, sizes, seeds, extra = h.dbscan_voxels(3.0, 6, 10, 100_000, D=D_, t_scale=ts)
            assert (labels == lo).all() and (seeds == so).all() and len(sizes) > 1
            assert D.canon(D.clusters_from(labels, sizes, extra)) == D.canon(co)
        with pytest.raises(evk.EvkError):
            h.dbscan_points(np.zeros((4, 2)), -1.0, 3)
        labels, sizes, seeds, extra = h.dbscan_points(np.zeros((0, 3)), 2.0, 3)
        assert len(labels) == 0 and len(sizes) == 0
        lab, sz, sd, ex = h.dbscan_points(np.array([[0, 0], [1, 0], [50, 50]], np.float32), 2.0, 2)
        assert lab.tolist() == [0, 0, -1] and sz.tolist() == [2] and sd.tolist() == [0]
