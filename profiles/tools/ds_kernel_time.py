"""A/B helper: CUDA-event time of the dominant downsample kernel on the C3 workload.
EVK_LIB=<path to an alternative libevk.so> selects the build under test (one library per
compile flag: see DESIGN.md section 7)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import evk_loader
evk = evk_loader.load()
n = 100_000_000
h = evk.Evk(n)
h.synth(evk.synth_params(0xE7CA0003, n, 1280, 720, 100_000_000, 64))
ds = evk.ds_params(1280, 720, 2, 2, 500, 0, 1, algo=evk.ALGO_SLAB)
h.set_profiling(True)
ts = []
for it in range(6):
    try:
        u, r = h.downsample(ds)
    except Exception as e:
        u = r = -1
    ts.append(h.stage_times().ds_main_ms)
print(os.environ.get("EVK_LIB", "default").split("/")[-1], "U", u, "R", r, "ds_main_ms", [round(t, 4) for t in ts[2:]])
