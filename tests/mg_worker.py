"""Multi-GPU parity worker (launched by tests/test_gpu_multi.py under torchrun, one rank per GPU).

Each rank owns a contiguous index shard of one synthetic stream.  Checks, for both ownership
schemes, that the union of the ranks' voxel shards equals the oracle's voxel set bit-exactly (keys
and global first indices), that no key is owned twice, and that sharded k-means gives the oracle's
centroids and counts on every rank.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import evk_loader  # noqa: E402
from oracle import orc  # noqa: E402


def gather_arrays(a, rank, world):
    """variable-length gather of a 1-D numpy array to every rank"""
    out = [None] * world
    dist.all_gather_object(out, a)
    return out


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    evk = evk_loader.load()
    cases = [
        # name, seed, n_per_rank, rate, W, H, blobs, voxel, K, iters
        ("gen4", 0xE7CA0004, 1_500_000, 100_000_000, 1280, 720, 64, (2, 2, 500, 1), 64, 3),
        ("davis", 0xE7CA0002, 700_000, 10_000_000, 346, 260, 32, (4, 4, 1000, 1), 32, 5),
    ]
    for name, seed, n, rate, W, H, blobs, (vx, vy, vt, up), K, iters in cases:
        total = n * world
        ev_all = orc.synth(orc.synth_params(seed, total, W, H, rate, blobs), threads=4)
        ok, of, orr = orc.downsample(ev_all, orc.ds_params(W, H, vx, vy, vt, 0, up))
        pts = orc.points(ev_all, of, 2)
        oc, ol, ocnt, _ = orc.kmeans(pts, pts[:K], iters=iters, threads=4)
        ds = evk.ds_params(W, H, vx, vy, vt, 0, up)
        km = evk.km_params(K, 2, iters=iters)
        h = evk.Evk(n + (1 << 19), device=local)
        uid = [evk.Evk.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        h.comm_init(rank, world, uid[0])
        h.set_shard(rank * n)
        for mode, mode_name in ((evk.OWNER_TIME_RANGE, "time_range"), (evk.OWNER_MIX64, "mix64")):
            for rep in range(2):  # twice: the table / halo state must be reusable
                h.synth(evk.synth_params(seed, n, W, H, rate, blobs, first_index=rank * n))
                ul, ug = h.downsample_sharded(ds, mode)
                assert ug == len(ok), (name, mode_name, rank, ug, len(ok))
                keys, reps, first = h.get_voxels()
                assert len(keys) == ul
                assert (np.diff(first.astype(np.int64)) > 0).all()
                assert reps.tobytes() == ev_all[first].tobytes(), "representative events differ"
                all_keys = np.concatenate(gather_arrays(keys, rank, world))
                all_first = np.concatenate(gather_arrays(first, rank, world))
                order = np.argsort(all_first, kind="stable")
                assert len(all_keys) == len(ok), "a key is owned by two ranks or lost"
                assert (all_keys[order] == ok).all() and (all_first[order] == of).all(), \
                    (name, mode_name, "voxel set differs from the oracle")
                h.init_centroids_first_k_sharded(km)
                c0, _ = h.get_centroids(K, 2)
                assert (c0 == pts[:K]).all(), "sharded first-K initialisation differs"
                assert h.kmeans_sharded(km) == iters
                cent, counts = h.get_centroids(K, 2)
                assert (counts == ocnt).all(), (name, mode_name, rank)
                assert np.allclose(cent, oc, rtol=1e-5, atol=0)
                lab = h.get_labels()
                # labels follow the local canonical order: compare through the global position
                pos = np.searchsorted(of, first)
                oc_prev = orc.kmeans(pts, pts[:K], iters=iters - 1)[0]
                bad = np.nonzero(lab != ol[pos])[0]
                for i in bad:  # only exact distance ties may differ (1e-6 relative)
                    d = ((oc_prev.astype(np.float64) - pts[pos[i]]) ** 2).sum(1)
                    assert abs(d[lab[i]] - d[ol[pos[i]]]) <= 1e-6 * d.max()
            if rank == 0:
                print(f"mg ok: {name} {mode_name} world={world} U={ug}", flush=True)
        # fused sharded step (one iteration): same voxel set, centroids and counts as the oracle
        km1 = evk.km_params(K, 2, iters=1)
        oc1, ol1, ocnt1, _ = orc.kmeans(pts, pts[:K], iters=1, threads=4)
        for rep in range(2):
            h.synth(evk.synth_params(seed, n, W, H, rate, blobs, first_index=rank * n))
            ul, ug, it = h.downsample_kmeans_sharded(ds, km1, True, evk.OWNER_TIME_RANGE)
            assert (ug, it) == (len(ok), 1), (name, rank, ug, len(ok))
            # (5 kernels with the peer-memory tail, 6 on the NCCL path; the separate calls report others)
            assert h.stage_times().km_launches in (5, 6), "fused sharded pass did not run"
            keys, reps, first = h.get_voxels()
            assert len(keys) == ul and reps.tobytes() == ev_all[first].tobytes()
            all_keys = np.concatenate(gather_arrays(keys, rank, world))
            all_first = np.concatenate(gather_arrays(first, rank, world))
            order = np.argsort(all_first, kind="stable")
            assert len(all_keys) == len(ok)
            assert (all_keys[order] == ok).all() and (all_first[order] == of).all()
            cent, counts = h.get_centroids(K, 2)
            assert (counts == ocnt1).all() and np.allclose(cent, oc1, rtol=1e-5, atol=0)
            lab = h.get_labels()
            assert (lab == ol1[np.searchsorted(of, first)]).all()
        # the same step queued three times behind each other, one wait: same result
        h.synth(evk.synth_params(seed, n, W, H, rate, blobs, first_index=rank * n))
        for _ in range(3):
            h.downsample_kmeans_sharded_submit(ds, km1, True, evk.OWNER_TIME_RANGE)
        ul2, ug2, it2 = h.downsample_kmeans_sharded_wait()
        assert (ul2, ug2, it2) == (ul, ug, 1)
        keys2, _, first2 = h.get_voxels(reps=False)
        assert (np.sort(first2) == np.sort(first)).all()
        cent2, counts2 = h.get_centroids(K, 2)
        assert (counts2 == counts).all() and (cent2 == cent).all()
        assert (h.get_labels() == ol1[np.searchsorted(of, first2)]).all()
        if rank == 0:
            print(f"mg ok: {name} fused sharded step world={world} U={ug} (also queued x3)", flush=True)
        # the hash-owned step (owner = mix64(key) range-reduced to the ranks): local downsample, records
        # written straight into their owners' shards over peer memory, suspects merged, one k-means
        # pass -- the union of the shards is the oracle's voxel set, centroids and counts agree
        for rep in range(2):
            h.synth(evk.synth_params(seed, n, W, H, rate, blobs, first_index=rank * n))
            ul, ug, it = h.downsample_kmeans_sharded(ds, km1, True, evk.OWNER_MIX64)
            assert (ug, it) == (len(ok), 1), (name, rank, ug, len(ok))
            keys, _, first = h.get_voxels(reps=False)
            assert len(keys) == ul and (np.diff(first.astype(np.int64)) > 0).all()
            all_keys = np.concatenate(gather_arrays(keys, rank, world))
            all_first = np.concatenate(gather_arrays(first, rank, world))
            order = np.argsort(all_first, kind="stable")
            assert len(all_keys) == len(ok), "a key is owned by two ranks or lost"
            assert (all_keys[order] == ok).all() and (all_first[order] == of).all()
            cent, counts = h.get_centroids(K, 2)
            assert (counts == ocnt1).all() and np.allclose(cent, oc1, rtol=1e-5, atol=0)
            lab = h.get_labels()
            assert (lab == ol1[np.searchsorted(of, first)]).all()
        if rank == 0:
            print(f"mg ok: {name} hash-owned fused step world={world} U={ug}", flush=True)
        # unordered stream: the time-range scheme must detect it on every rank and fall back
        rng = np.random.default_rng(5)
        perm = rng.permutation(total)
        ev_shuf = ev_all[perm]
        ok2, of2, _ = orc.downsample(ev_shuf, orc.ds_params(W, H, vx, vy, vt, 0, up))
        h.load_events(ev_shuf[rank * n:(rank + 1) * n])
        ul, ug = h.downsample_sharded(ds, evk.OWNER_TIME_RANGE)
        keys, _, first = h.get_voxels(reps=False)
        all_keys = np.concatenate(gather_arrays(keys, rank, world))
        all_first = np.concatenate(gather_arrays(first, rank, world))
        order = np.argsort(all_first, kind="stable")
        assert ug == len(ok2) and (all_keys[order] == ok2).all() and (all_first[order] == of2).all()
        # ... and so must the fused step (general path, same answer)
        h.load_events(ev_shuf[rank * n:(rank + 1) * n])
        ul, ug, it = h.downsample_kmeans_sharded(ds, km1, True, evk.OWNER_TIME_RANGE)
        keys, _, first = h.get_voxels(reps=False)
        all_keys = np.concatenate(gather_arrays(keys, rank, world))
        all_first = np.concatenate(gather_arrays(first, rank, world))
        order = np.argsort(all_first, kind="stable")
        assert ug == len(ok2) and (all_keys[order] == ok2).all() and (all_first[order] == of2).all()
        pts2 = orc.points(ev_shuf, of2, 2)
        oc2, _, ocnt2, _ = orc.kmeans(pts2, pts2[:K], iters=1, threads=4)
        cent, counts = h.get_centroids(K, 2)
        assert (counts == ocnt2).all() and np.allclose(cent, oc2, rtol=1e-5, atol=0)
        h.load_events(ev_shuf[rank * n:(rank + 1) * n])
        ul, ug, it = h.downsample_kmeans_sharded(ds, km1, True, evk.OWNER_MIX64)
        keys, _, first = h.get_voxels(reps=False)
        all_keys = np.concatenate(gather_arrays(keys, rank, world))
        all_first = np.concatenate(gather_arrays(first, rank, world))
        order = np.argsort(all_first, kind="stable")
        assert ug == len(ok2) and (all_keys[order] == ok2).all() and (all_first[order] == of2).all()
        cent, counts = h.get_centroids(K, 2)
        assert (counts == ocnt2).all() and np.allclose(cent, oc2, rtol=1e-5, atol=0)
        if rank == 0:
            print(f"mg ok: {name} unordered fallback world={world} U={ug}", flush=True)
        h.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
