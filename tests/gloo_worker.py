"""world_size-2 CPU run (gloo) of the sharding protocol of csrc/evk_comm.cu: the ownership rules
of <package>/sharding.py + the oracle as the per-rank downsample / assign, exchanged with
torch.distributed, must reproduce the single-process oracle bit for bit.  Launched by
tests/test_sharding_gloo.py."""
import importlib.util
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402

spec = importlib.util.spec_from_file_location(
    "evk_sharding",
    os.path.join(ROOT, "event-camera-clustering-and-optical-flow-estimation_b200", "sharding.py"))
sh = importlib.util.module_from_spec(spec)
spec.loader.exec_module(sh)


def allgather(obj, world):
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    W, H, K = 346, 260, 16
    vx, vy, vt, up = 4, 4, 1000, 1
    n_total = 700_000  # > 2 * HALO_EVENTS so that each rank holds a full boundary block
    ev_all = orc.synth(orc.synth_params(0xE7CA0002, n_total, W, H, 10_000_000, 16))
    p = orc.ds_params(W, H, vx, vy, vt, 0, up)
    ok_keys, ok_first, ok_rep = orc.downsample(ev_all, p)
    pts_all = orc.points(ev_all, ok_first, 2)
    lo, hi = sh.shard_range(n_total, rank, world)
    assert sh.shard_range(n_total, world - 1, world)[1] == n_total
    own = ev_all[lo:hi]

    # ---- time-range ownership: boundary block to the previous rank, then purely local work
    heads = allgather(own[:sh.HALO_EVENTS].copy(), world)  # stands for the neighbour send/recv
    halo = heads[rank + 1] if rank < world - 1 else own[:0]
    skip, keep, ok = sh.time_range_split(own["t"], halo["t"], rank, world, 0, vt)
    assert ok
    # the peer-memory form of the same exchange: every sender counts its own boundary share and
    # publishes it; the receiver takes exactly that many events of the next shard -- same split
    share, ok2 = sh.boundary_share(own["t"], rank, 0, vt)
    shares = allgather((share, ok2), world)
    assert all(q[1] for q in shares) and share == skip
    assert keep == (shares[rank + 1][0] if rank < world - 1 else 0)
    pulled = heads[rank + 1][:shares[rank + 1][0]] if rank < world - 1 else own[:0]
    assert pulled.tobytes() == halo[:keep].tobytes()
    mine = np.concatenate([own[skip:], halo[:keep]])
    k, f, r = orc.downsample(mine, p)
    f = f.astype(np.int64) + lo + skip  # global first indices
    parts = allgather((k, f, r), world)
    all_k = np.concatenate([q[0] for q in parts])
    all_f = np.concatenate([q[1] for q in parts])
    order = np.argsort(all_f, kind="stable")
    assert len(all_k) == len(ok_keys), "a key is owned twice or lost"
    assert (all_k[order] == ok_keys).all() and (all_f[order] == ok_first).all()
    assert sum(q[2] for q in parts) == ok_rep, "repeated count is not additive over owners"
    # every bin is wholly on one rank
    bins = [np.unique(q[0] // np.uint64(87 * 65 * 2)) for q in parts]
    assert len(np.intersect1d(bins[0], bins[1])) == 0

    # ---- hash ownership: local downsample, bucket by owner, all-to-all, lowest index wins
    k, f, _ = orc.downsample(own, p)
    f = f.astype(np.int64) + lo
    owner = sh.owner_mix64(k, world)
    assert ((owner >= 0) & (owner < world)).all()
    send = [(k[owner == d], f[owner == d]) for d in range(world)]
    recv = [None] * world
    dist.all_to_all_object_list(recv, send) if hasattr(dist, "all_to_all_object_list") else None
    if recv[0] is None:  # older torch: emulate the all-to-all with an allgather
        every = allgather(send, world)
        recv = [every[s][rank] for s in range(world)]
    mk, mf = sh.merge_lowest_index(np.concatenate([q[0] for q in recv]),
                                   np.concatenate([q[1] for q in recv]))
    assert (sh.owner_mix64(mk, world) == rank).all()
    parts = allgather((mk, mf), world)
    all_k = np.concatenate([q[0] for q in parts])
    all_f = np.concatenate([q[1] for q in parts])
    order = np.argsort(all_f, kind="stable")
    assert len(all_k) == len(ok_keys)
    assert (all_k[order] == ok_keys).all() and (all_f[order] == ok_first).all()

    # ---- k-means: exact integer partial sums, one allreduce per iteration
    cent = pts_all[:K].copy()  # first K voxels in canonical order (all on rank 0's shard)
    my_pts = orc.points(ev_all, mf.astype(np.uint32), 2)
    for _ in range(3):
        lab = orc.kmeans_assign(my_pts, cent)
        sums = torch.from_numpy(sh.partial_sums(my_pts, lab, K))
        dist.all_reduce(sums)
        cent = sh.finalise(cent, sums.numpy())
    oc, _, ocnt, _ = orc.kmeans(pts_all, pts_all[:K], iters=3)
    assert (sums.numpy()[:, 0] == ocnt).all()
    assert (cent == oc).all(), "sharded centroids differ from the single-process oracle"

    # ---- bench.py's timing rule: the step time of the job is the max over ranks
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == float(world)
    if rank == 0:
        print("gloo ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
