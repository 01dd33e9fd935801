"""Randomised stress of the time-slab kernel against the oracle: shapes that maximise races inside a
tile (few cells, thousands of events per cell), bins of one event, bins far larger than a tile,
odd voxel sizes, sensors up to the shared-memory limit.  Every case must take the slab path (no
silent fallback) and match the oracle bit for bit (keys, first indices, representatives, repeated
count) -- then the same through the fused step."""
import numpy as np
import pytest

import evk_loader

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def evk():
    m = evk_loader.load()
    m.lib()
    return m


def make_stream(rng, n, W, H, dt_max, hot):
    """time-ordered stream; `hot` = fraction of events that land on a handful of pixels"""
    ev = np.zeros(n, dtype=[("x", "<u2"), ("y", "<u2"), ("p", "<i2"), ("_pad", "<u2"),
                            ("t", "<i8")])
    ev["t"] = np.cumsum(rng.integers(0, dt_max + 1, size=n))
    ev["x"] = rng.integers(0, W, size=n)
    ev["y"] = rng.integers(0, H, size=n)
    ev["p"] = rng.integers(0, 2, size=n)
    m = rng.random(n) < hot
    k = int(m.sum())
    hx, hy = rng.integers(0, W, size=4), rng.integers(0, H, size=4)
    pick = rng.integers(0, 4, size=k)
    ev["x"][m] = hx[pick]
    ev["y"][m] = hy[pick]
    return ev


CASES = [
    # W, H, vx, vy, vt, use_p, n, dt_max, hot
    (64, 48, 1, 1, 100, 1, 200_000, 1, 0.9),        # ~50 events per us, nearly all on 4 pixels
    (64, 48, 4, 4, 1, 0, 100_000, 3, 0.5),          # bins of about one event
    (346, 260, 4, 4, 1000, 1, 300_000, 0, 0.0),     # every event at t = 0: one huge bin
    (346, 260, 3, 5, 777, 1, 300_000, 2, 0.2),      # non-power-of-two divisors
    (1280, 720, 2, 2, 500, 1, 400_000, 0, 0.3),     # Gen4 key space, one bin of 400 k events
    (1280, 720, 2, 2, 50_000, 0, 400_000, 1, 0.97),  # 97 % of the events on 4 cells
    (1600, 1100, 4, 2, 250, 1, 250_000, 1, 0.1),    # 440 000 cells per bin: near the smem limit
    (17, 3, 1, 1, 10, 1, 50_000, 1, 0.0),           # tiny sensor: 102 cells, 5 k events per bin
]


@pytest.mark.parametrize("case", CASES)
def test_slab_stress(evk, orc, case):
    W, H, vx, vy, vt, up, n, dt_max, hot = case
    rng = np.random.default_rng(hash(case) & 0xFFFFFFFF)
    ev = make_stream(rng, n, W, H, dt_max, hot)
    ok, of, orr = orc.downsample(ev, orc.ds_params(W, H, vx, vy, vt, 0, up))
    pts = orc.points(ev, of, 2)
    K = int(min(16, len(ok)))
    with evk.Evk(n) as h:
        h.load_events(ev)
        ds = evk.ds_params(W, H, vx, vy, vt, 0, up, algo=evk.ALGO_SLAB)
        for rep in range(2):
            U, R = h.downsample(ds)
            assert h.stage_times().ds_algo_used == evk.ALGO_SLAB, "silent fallback"
            keys, reps, first = h.get_voxels()
            assert (U, R) == (len(ok), orr)
            assert (keys == ok).all() and (first == of).all()
            assert reps.tobytes() == ev[of].tobytes()
        km = evk.km_params(K, 2, iters=1)
        U, R, it = h.downsample_kmeans(ds, km, True)
        assert (U, R, it) == (len(ok), orr, 1)
        keys, _, first = h.get_voxels(reps=False)
        assert (keys == ok).all() and (first == of).all()
        oc, ol, ocnt, _ = orc.kmeans(pts, pts[:K], iters=1)
        cent, counts = h.get_centroids(K, 2)
        assert (counts == ocnt).all() and (cent == oc).all()
        assert (h.get_labels() == ol).all()


def test_slab_many_one_tile_bins(evk, orc):
    """Regression (round 2): with bins of one short tile each, every tile is the last tile of its
    bin, and the per-bin constants the output pass reads back through shared memory (first key of
    the bin) must not be replaced before every warp has written the previous bin's records.
    Found by the 4- and 8-GPU parity runs on shuffled shards (280 bins of 2 500 events per rank):
    voxels came out with the NEXT bin's key base.  Two shapes: that shard through the partition
    path, and an ordered stream with bins of about 2 000 events through the slab path."""
    W, H, vx, vy, vt, up = 346, 260, 4, 4, 1000, 1
    n, world = 700_000, 4
    ev_all = orc.synth(orc.synth_params(0xE7CA0002, n * world, W, H, 10_000_000, 32), threads=4)
    shard = ev_all[np.random.default_rng(5).permutation(n * world)][:n]
    ok, of, orr = orc.downsample(shard, orc.ds_params(W, H, vx, vy, vt, 0, up))
    with evk.Evk(n) as h:
        for rep in range(3):
            h.load_events(shard)
            U, R = h.downsample(evk.ds_params(W, H, vx, vy, vt, 0, up, algo=evk.ALGO_PARTITION))
            assert h.stage_times().ds_algo_used == evk.ALGO_PARTITION
            keys, _, first = h.get_voxels(reps=False)
            assert (U, R) == (len(ok), orr)
            assert (keys == ok).all() and (first == of).all()
    n2 = 2_000_000
    ev = orc.synth(orc.synth_params(0xE7CA0005, n2, 1280, 720, 100_000_000, 64), threads=4)
    p = (1280, 720, 2, 2, 20, 0, 1)   # 20 us bins: 1 000 bins of 2 000 events
    ok, of, orr = orc.downsample(ev, orc.ds_params(*p))
    with evk.Evk(n2) as h:
        h.load_events(ev)
        for rep in range(3):
            U, R = h.downsample(evk.ds_params(*p, algo=evk.ALGO_SLAB))
            assert h.stage_times().ds_algo_used == evk.ALGO_SLAB
            keys, _, first = h.get_voxels(reps=False)
            assert (U, R) == (len(ok), orr)
            assert (keys == ok).all() and (first == of).all()
