"""ncu target: a few launches of the slab downsample on the C3 workload (EVK_LIB selects the build)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import evk_loader
evk = evk_loader.load()
n = int(os.environ.get("EVK_AB_EVENTS", 100_000_000))
h = evk.Evk(n)
h.synth(evk.synth_params(0xE7CA0003, n, 1280, 720, 100_000_000, 64))
ds = evk.ds_params(1280, 720, 2, 2, 500, 0, 1, algo=evk.ALGO_SLAB)
for _ in range(int(os.environ.get("EVK_AB_REPS", 4))):
    print(h.downsample(ds))
