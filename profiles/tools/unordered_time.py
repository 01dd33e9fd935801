"""The general paths on the bench workload (C3, 100 M Gen4 events) when the stream is NOT time-ordered:
fully shuffled, and shuffled in blocks of one time bin (out-of-order packets).  Times every
algorithm through the C-ABI (CUDA events of the library, whole downsample call incl. the slab
attempt that detects the disorder) and checks that all of them return the same voxel shard."""
import hashlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import evk_loader
evk = evk_loader.load()
n, W, H = int(os.environ.get("EVK_AB_EVENTS", 100_000_000)), 1280, 720
h = evk.Evk(n)
h.synth(evk.synth_params(0xE7CA0003, n, W, H, 100_000_000, 64))
ev = h.get_events()
rng = np.random.default_rng(5)
cases = {"ordered": None,
         "bins_shuffled": ev.reshape(-1, 50_000)[rng.permutation(n // 50_000)].reshape(-1),
         "fully_shuffled": ev[rng.permutation(n)]}
h.set_profiling(True)
for name, e in cases.items():
    if e is not None:
        h.load_events(e)
    ref = None
    for algo, aname in ((evk.ALGO_AUTO, "auto"), (evk.ALGO_PARTITION, "partition"),
                        (evk.ALGO_TABLE, "table"), (evk.ALGO_SORT, "sort")):
        ds = evk.ds_params(W, H, 2, 2, 500, 0, 1, algo=algo)
        ts, wall = [], []
        for _ in range(4):
            t0 = time.perf_counter()
            u, r = h.downsample(ds)
            wall.append((time.perf_counter() - t0) * 1e3)
            ts.append(h.stage_times().ds_total_ms)
        keys, _, first = h.get_voxels(reps=False)
        dig = hashlib.sha256(keys.tobytes() + first.tobytes()).hexdigest()[:16]
        ref = ref or dig
        print(json.dumps({"stream": name, "algo": aname, "algo_used": h.stage_times().ds_algo_used,
                          "U": u, "R": r, "same_as_first": dig == ref,
                          "ds_total_ms_device": round(min(ts[1:]), 4),
                          "call_ms_wall": round(min(wall[1:]), 4)}), flush=True)
