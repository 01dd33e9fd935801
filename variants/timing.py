import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import evk_loader
evk = evk_loader.load()
L = evk.lib()
n = 100_000_000
h = evk.Evk(n)
h.synth(evk.synth_params(0xE7CA0003, n, 1280, 720, 100_000_000, 64))
ds = evk.ds_params(1280, 720, 2, 2, 500, 0, 1)
out = (C.c_ulonglong * 8)()
for it in range(3):
    h.downsample(ds)
    L.evk_debug_slab_timing(out)
    v = list(out)
    tot = sum(v)
    names = ["bin prologue+top", "tma wait", "classify", "B0 wait+tma issue", "claim", "S1 wait", "resolve", "epilogue"]
    print(it, {nm: f"{100*x/tot:.1f}%" for nm, x in zip(names, v)}, "cycles/warp", tot / (148 * 32))
