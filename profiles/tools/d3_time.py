"""C3 with D = 3 and D = 4 k-means on the 55 M voxels (K = 64): time of one assign + accumulate
iteration (CUDA events of the library)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import evk_loader
evk = evk_loader.load()
n, W, H, K = 100_000_000, 1280, 720, 64
h = evk.Evk(n)
h.synth(evk.synth_params(0xE7CA0003, n, W, H, 100_000_000, 64))
u, r = h.downsample(evk.ds_params(W, H, 2, 2, 500, 0, 1))
h.set_profiling(True)
for D in (2, 3, 4):
    for iters in (1, 3):
        km = evk.km_params(K, D, iters=iters, t_scale=1e-3, p_scale=25.0)
        ts = []
        for _ in range(4):
            h.init_centroids_first_k(km)
            h.kmeans(km)
            ts.append(h.stage_times().km_total_ms)
        print(json.dumps({"config": "C3", "voxels": u, "K": K, "D": D, "iters": iters,
                          "km_total_ms": round(min(ts[1:]), 4),
                          "ms_per_iter": round(min(ts[1:]) / iters, 4)}), flush=True)
