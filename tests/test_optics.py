"""OPTICS reachability ordering (second half).  The reference's optics.hpp needs
boost.geometry, FunctionalPlus and `geometry`, none vendored: it cannot be compiled here, so the
oracle (oracle/optics_oracle.c) is PINNED BY HAND-DERIVED ANSWERS and by an independent restatement
only -- parity unpinned by the reference.  GPU: the CUDA path through the C-ABI == the oracle."""
import os

import numpy as np
import pytest

import evk_loader
from oracle import optics

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def clouds():
    r = np.random.default_rng(41)
    out = []
    for n, d, mp, eps in [(1500, 2, 2, 10.0), (2500, 2, 5, 8.0), (1200, 3, 3, 12.0), (64, 2, 1, 3.0),
                          (400, 2, 4, 1000.0), (900, 2, 2, 0.5)]:
        c = r.integers(0, 600, size=(9, d))
        pts = (c[r.integers(0, 9, n)] + r.normal(0, 7, size=(n, d))).round().astype(np.int32)
        out.append((pts, mp, eps))
    # coincident points and exact ties (lattice)
    g = np.stack(np.meshgrid(np.arange(20), np.arange(20)), -1).reshape(-1, 2).astype(np.int32) * 3
    out.append((np.concatenate([g, g[:50]]), 3, 3.0))
    return out


def test_known_answers():
    # collinear points 0, 1, 2 | 10, 11 with eps 3, min_pts 2 (derived by hand, oracle/optics_oracle.c)
    pts = np.array([[0, 0], [1, 0], [2, 0], [10, 0], [11, 0]], np.int32)
    order, reach = optics.oracle(pts, 2, 3.0)
    assert order.tolist() == [0, 1, 2, 3, 4] and reach.tolist() == [-1.0, 1.0, 1.0, -1.0, 1.0]
    cl, nc = optics.clusters(reach, 3.0)
    assert nc == 2 and cl.tolist() == [0, 0, 0, 1, 1]
    # min_pts larger than any neighbourhood: no core point, everything undefined, index order
    order, reach = optics.oracle(pts, 4, 3.0)
    assert order.tolist() == [0, 1, 2, 3, 4] and (reach == -1).all()
    # the ordering leaves index order when a farther seed is closer in reachability: 0 - 4 - 2 | ...
    pts = np.array([[0, 0], [9, 0], [5, 0], [20, 0], [3, 0]], np.int32)
    order, reach = optics.oracle(pts, 2, 6.0)
    # 0: neighbours {0, 2, 4}, core 3 -> reach[4] = 3, reach[2] = 5; pop 4 (3): neighbours {0, 2, 4, 1},
    # core 2 -> reach[2] = max(2, 2) = 2, reach[1] = 6; pop 2: core 2 -> reach[1] = max(2, 4) = 4; pop 1
    assert order.tolist() == [0, 4, 2, 1, 3] and reach.tolist() == [-1.0, 3.0, 2.0, 4.0, -1.0]


def test_oracle_equals_independent_statement():
    for pts, mp, eps in clouds():
        if len(pts) > 1600:
            continue   # the plain form is O(n^2) in Python
        o1, r1 = optics.oracle(pts, mp, eps)
        o2, r2 = optics.plain(pts, mp, eps)
        assert (o1 == o2).all() and (r1 == r2).all()
        assert sorted(o1.tolist()) == list(range(len(pts)))     # a permutation


@pytest.fixture(scope="module")
def evk():
    m = evk_loader.load()
    m.lib()
    return m


@pytest.mark.gpu
def test_cuda_equals_oracle(evk, orc):
    with evk.Evk(1 << 16) as h:
        for pts, mp, eps in clouds():
            order, reach = h.optics_points(pts, mp, eps)
            o, r = optics.oracle(pts, mp, eps)
            assert (order == o).all() and (reach == r).all(), (len(pts), mp, eps)
            cl, nc = h.optics_clusters(eps)
            ocl, onc = optics.clusters(r, eps)
            assert nc == onc and (cl == ocl).all()
        # the reference app's own shape: integer (x, y) events, min_pts 2, epsilon 10, threshold 10
        ev = orc.load_csv(os.path.join(GOLDEN_DIR, "event_raw_data8.csv"))
        pts = np.stack([ev["x"], ev["y"]], 1).astype(np.int32)
        order, reach = h.optics_points(pts, 2, 10.0)
        o, r = optics.oracle(pts, 2, 10.0)
        assert (order == o).all() and (reach == r).all()
        # on the voxel shard (canonical order), D = 2
        e = orc.synth(orc.synth_params(0xE7CA0021, 60_000, 346, 260, 2_000_000, 6))
        h.load_events(e)
        h.downsample(evk.ds_params(346, 260, 8, 8, 10_000, 0, 0))
        _, reps, _ = h.get_voxels()
        order, reach = h.optics_voxels(3, 12.0)
        o, r = optics.oracle(np.stack([reps["x"], reps["y"]], 1).astype(np.int32), 3, 12.0)
        assert (order == o).all() and (reach == r).all()
        # edge cases
        assert len(h.optics_points(np.zeros((0, 2), np.int32), 2, 5.0)[0]) == 0
        o1, r1 = h.optics_points(np.array([[4, 4]], np.int32), 1, 5.0)
        assert o1.tolist() == [0] and r1.tolist() == [-1.0]
        with pytest.raises(evk.EvkError):
            h.optics_points(np.zeros((70000, 2), np.int32), 2, 5.0)
