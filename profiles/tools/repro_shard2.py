"""repro_shard.py narrowed to one case (davis, world 4), with a diagnosis of the mismatch; EVK_LIB selects the build."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import evk_loader
from oracle import orc
evk = evk_loader.load()
name, seed, n, rate, W, H, blobs, (vx, vy, vt, up) = ("davis", 0xE7CA0002, 700_000, 10_000_000, 346, 260, 32, (4, 4, 1000, 1))
world = 4
total = n * world
ev_all = orc.synth(orc.synth_params(seed, total, W, H, rate, blobs), threads=8)
perm = np.random.default_rng(5).permutation(total)
shard = ev_all[perm][:n]
ok, of, _ = orc.downsample(shard, orc.ds_params(W, H, vx, vy, vt, 0, up), threads=8)
h = evk.Evk(n + (1 << 19))
for cr in (1, 0):
    h.load_events(shard)
    u, rep = h.downsample(evk.ds_params(W, H, vx, vy, vt, 0, up, algo=evk.ALGO_PARTITION, count_repeated=cr))
    keys, _, first = h.get_voxels(reps=False)
    o = np.argsort(keys, kind="stable"); oo = np.argsort(ok, kind="stable")
    same_set = len(keys) == len(ok) and (keys[o] == ok[oo]).all()
    fd = (first[o] != of[oo]) if same_set else None
    print(os.environ.get("EVK_LIB", "lib").split("/")[-1], "count_rep", cr, "used", h.stage_times().ds_algo_used, "U", u, len(ok),
          "same key set", same_set, "first-index mismatches", int(fd.sum()) if fd is not None else -1,
          "gpu first > oracle first", int((first[o][fd] > of[oo][fd]).sum()) if fd is not None else -1, flush=True)
    if fd is not None and fd.any():
        i = np.flatnonzero(fd)[:3]
        for q in i:
            k = keys[o][q]
            allidx = np.flatnonzero((shard["t"] // vt == k // (87 * 65 * 2)))  # events of that bin (approx)
            print("  key", k, "gpu first", first[o][q], "oracle first", of[oo][q])
