"""A/B helper: wall time of evk_load_evt2 (H2D of 400 MB of RAW words + device decode)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import evk_loader, torch
from oracle import orc
evk = evk_loader.load()
n = 100_000_000
h = evk.Evk(n)
h.synth(evk.synth_params(0xE7CA0003, n, 1280, 720, 100_000_000, 64))
ev = h.get_events()
w = orc.evt2_encode(ev)
raw = torch.empty(len(w), dtype=torch.int32, pin_memory=True)
raw.numpy().view(np.uint32)[:] = w
for _ in range(2): h.load_evt2_ptr(raw.data_ptr(), len(w))
ts = []
for _ in range(6):
    t0 = time.perf_counter(); h.load_evt2_ptr(raw.data_ptr(), len(w)); ts.append((time.perf_counter() - t0) * 1e3)
print(os.environ.get("EVK_LIB", "new").split("/")[-1], "load_evt2 ms", [round(t, 3) for t in ts], "H2D GB/s if copy only", round(len(w) * 4 / min(ts) / 1e6, 1))
