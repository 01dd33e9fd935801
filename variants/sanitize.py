"""small slab + fused-step runs for compute-sanitizer (racecheck / memcheck)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import evk_loader
from oracle import orc
evk = evk_loader.load()
for (W, H, vox, n, rate, K) in [(1280, 720, (2, 2, 500, 1), 120_000, 100_000_000, 64),
                                (64, 48, (1, 1, 100, 1), 40_000, 50_000_000, 8)]:
    ev = orc.synth(orc.synth_params(7, n, W, H, rate, 8))
    ok, of, orr = orc.downsample(ev, orc.ds_params(W, H, *vox[:3], 0, vox[3]))
    with evk.Evk(n) as h:
        h.load_evt2(orc.evt2_encode(ev))
        ds = evk.ds_params(W, H, *vox[:3], 0, vox[3], algo=evk.ALGO_SLAB)
        U, R, it = h.downsample_kmeans(ds, evk.km_params(K, 2, iters=1), True)
        keys, _, first = h.get_voxels(reps=False)
        assert (U, R) == (len(ok), orr) and (keys == ok).all() and (first == of).all()
        U, R = h.downsample(ds)
        h.init_centroids_first_k(evk.km_params(K, 2, iters=4))
        h.kmeans(evk.km_params(K, 2, iters=4))
        h.get_labels()
print("sanitize run ok")
