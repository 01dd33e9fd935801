"""ncu target: the hash-owned step's kernels with ONE rank (EVK_P2P_SINGLE=1: every record is its own
rank's), C3 workload."""
import os, sys
os.environ["EVK_P2P_SINGLE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # makes libnccl.so.2 resolvable
import evk_loader
evk = evk_loader.load()
n = int(os.environ.get("EVK_AB_EVENTS", 100_000_000))
h = evk.Evk(n + (1 << 19))
h.synth(evk.synth_params(0xE7CA0003, n, 1280, 720, 100_000_000, 64))
h.comm_init(0, 1, evk.Evk.comm_unique_id())
h.set_shard(0)
ds = evk.ds_params(1280, 720, 2, 2, 500, 0, 1)
km = evk.km_params(64, 2, iters=1)
for _ in range(3):
    print(h.downsample_kmeans_sharded(ds, km, True, evk.OWNER_MIX64))
