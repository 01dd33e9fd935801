import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import evk_loader
evk = evk_loader.load()
n, W, H = 600_000, 346, 260
for rnd in range(3):
    h = evk.Evk(n)
    h.synth(evk.synth_params(0xE7CA0002, n, W, H, 10_000_000, 16))
    for iters in (3, 1, 2, 5):
        ds = evk.ds_params(W, H, 4, 4, 1000, 0, 1)
        km = evk.km_params(8, 2, iters=iters)
        print(rnd, iters, h.downsample_kmeans(ds, km, True), flush=True)
        h.downsample(ds); h.init_centroids_first_k(km); print(h.kmeans(km), flush=True)
    h.close()
    print("closed", rnd, flush=True)
