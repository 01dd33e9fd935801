"""A/B helper: CUDA-event time of the dominant downsample kernel on the C3 workload, with a
checksum of the result so that every variant is also checked against the others.
EVK_LIB=<path to an alternative libevk.so> selects the build under test (one library per
compile flag: profiles/tools/build_variant.py)."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import evk_loader
evk = evk_loader.load()
n = int(os.environ.get("EVK_AB_EVENTS", 100_000_000))
h = evk.Evk(n)
h.synth(evk.synth_params(0xE7CA0003, n, 1280, 720, 100_000_000, 64))
ds = evk.ds_params(1280, 720, 2, 2, 500, 0, 1, algo=evk.ALGO_SLAB,
                   count_repeated=0 if os.environ.get('EVK_AB_NOREP') else 1)
km = evk.km_params(64, 2, iters=1)
h.set_profiling(True)
ts, used = [], None
for it in range(7):
    try:
        u, r = h.downsample(ds)
        used = h.stage_times().ds_algo_used
    except Exception as e:
        u = r = -1
    ts.append(h.stage_times().ds_main_ms)
keys, _, first = h.get_voxels(reps=False)
digest = hashlib.sha256(keys.tobytes() + first.tobytes()).hexdigest()[:16]
# the fused step, queued (what bench.py times)
h.set_profiling(False)
for _ in range(3):
    h.downsample_kmeans(ds, km, True)
h.timer_start()
for _ in range(20):
    h.downsample_kmeans_submit(ds, km, True)
h.downsample_kmeans_wait()
step_ms = h.timer_stop() / 20
cent, counts = h.get_centroids(64, 2)
digest2 = hashlib.sha256(cent.tobytes() + counts.tobytes()).hexdigest()[:16]
print(os.environ.get("EVK_LIB", "default").split("/")[-1], "algo", used, "U", u, "R", r, "sha", digest, digest2,
      "ds_main_ms", [round(t, 4) for t in ts[2:]], "min", round(min(ts[2:]), 4), "step_ms", round(step_ms, 4), flush=True)
