"""Single-GPU check of what a rank's local downsample sees in the multi-GPU 'unordered fallback' case:
a fully shuffled stream of n * world events cut into world shards, every shard downsampled on the
device (auto algorithm) and compared with the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import evk_loader
from oracle import orc
evk = evk_loader.load()
cases = [("gen4", 0xE7CA0004, 1_500_000, 100_000_000, 1280, 720, 64, (2, 2, 500, 1)),
         ("davis", 0xE7CA0002, 700_000, 10_000_000, 346, 260, 32, (4, 4, 1000, 1))]
bad = 0
for world in (2, 4, 8):
    for name, seed, n, rate, W, H, blobs, (vx, vy, vt, up) in cases:
        total = n * world
        ev_all = orc.synth(orc.synth_params(seed, total, W, H, rate, blobs), threads=8)
        perm = np.random.default_rng(5).permutation(total)
        ev_shuf = ev_all[perm]
        ds = evk.ds_params(W, H, vx, vy, vt, 0, up)
        h = evk.Evk(n + (1 << 19))
        for r in range(world):
            shard = ev_shuf[r * n:(r + 1) * n]
            ok, of, _ = orc.downsample(shard, orc.ds_params(W, H, vx, vy, vt, 0, up), threads=8)
            for algo in (evk.ALGO_AUTO, evk.ALGO_TABLE):
                h.load_events(shard)
                u, rep = h.downsample(evk.ds_params(W, H, vx, vy, vt, 0, up, algo=algo))
                keys, _, first = h.get_voxels(reps=False)
                used = h.stage_times().ds_algo_used
                good = u == len(ok) and (keys == ok).all() and (first == of).all()
                if not good:
                    bad += 1
                    nk = int((keys[:min(len(keys), len(ok))] != ok[:min(len(keys), len(ok))]).sum()) if len(keys) else -1
                    print(f"MISMATCH {name} world={world} rank={r} algo={algo} used={used} U={u} oracle={len(ok)} keydiff={nk}", flush=True)
        h.close()
        print(f"checked {name} world={world}", flush=True)
print("bad", bad)
