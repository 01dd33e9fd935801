"""Small shapes for compute-sanitizer (memcheck / racecheck / synccheck): the slab downsample, the
fused step (label map, quads, assign tiles, finalise), the table and sort paths, D=3 k-means.
Results are compared with the oracle so that a sanitizer-clean run is also a correct one."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import evk_loader
from oracle import orc
evk = evk_loader.load()
n, W, H, K = int(os.environ.get("EVK_SAN_EVENTS", 400_000)), 1280, 720, 16
ev = orc.synth(orc.synth_params(0xE7CA0003, n, W, H, 100_000_000, 16))
ok, of, orr = orc.downsample(ev, orc.ds_params(W, H, 2, 2, 500, 0, 1))
with evk.Evk(n) as h:
    h.load_events(ev)
    for algo in (evk.ALGO_SLAB, evk.ALGO_TABLE, evk.ALGO_SORT):
        U, R = h.downsample(evk.ds_params(W, H, 2, 2, 500, 0, 1, algo=algo))
        assert h.stage_times().ds_algo_used == algo
        keys, _, first = h.get_voxels(reps=False)
        assert (U, R) == (len(ok), orr) and (keys == ok).all() and (first == of).all(), algo
    km = evk.km_params(K, 2, iters=1)
    U, R, _ = h.downsample_kmeans(evk.ds_params(W, H, 2, 2, 500, 0, 1), km, True)
    pts = orc.points(ev, of, 2)
    oc, ol, ocnt, _ = orc.kmeans(pts, pts[:K], iters=1)
    assert (h.get_labels() == ol).all() and (h.get_centroids(K, 2)[1] == ocnt).all()
    km3 = evk.km_params(K, 3, iters=2, t_scale=1e-3)
    h.init_centroids_first_k(km3)
    h.kmeans(km3)
    # an unordered slice: detected, rerun on the general path
    h.load_events(ev[::-1].copy())
    U2, R2 = h.downsample(evk.ds_params(W, H, 2, 2, 500, 0, 1))
    assert U2 == U
print("sanitize_small OK", n, U, R, flush=True)
