"""Host-side statement of the multi-GPU sharding rules of csrc/evk_comm.cu.

Pure arithmetic on numpy arrays — no device work, no oracle: which rank owns what.  The CUDA
library implements the same rules on the device (k_halo_range / k_p2p_tick_range, owner_of); bench.py uses
`shard_range`, and tests/test_sharding_gloo.py runs the whole exchange protocol over these
functions on CPU with world_size 2 (gloo) against the single-process answer.
"""
import numpy as np

HALO_EVENTS = 1 << 18  # boundary block a rank sends to its predecessor (csrc/evk_comm.cu: halo)


def shard_range(n_total, rank, world):
    """contiguous index shard [lo, hi) of `rank` (events are sharded by index)"""
    per = n_total // world
    lo = rank * per
    hi = n_total if rank == world - 1 else lo + per
    return lo, hi


def time_bin(t, t0_us, vt_us):
    """tbin = (t - t0) / vt, integer division (include/evk.h: evk_ds_params)"""
    t = np.asarray(t, dtype=np.int64)
    return (t - t0_us) // vt_us if vt_us > 0 else np.zeros_like(t)


def time_range_split(t_own, t_halo, rank, world, t0_us, vt_us):
    """EVK_OWNER_TIME_RANGE: a time bin is owned by the rank whose shard holds its first event.

    t_own  : timestamps of this rank's shard (time-ordered)
    t_halo : timestamps of the head of the next rank's shard (its boundary block), or empty
    returns (skip, keep, ok): this rank works on own[skip:] + halo[:keep]; ok = False when the
    first bin does not end inside the block (k_halo_range's give-up conditions).
    """
    skip = keep = 0
    ok = True
    if rank > 0:
        if len(t_own) == 0 or t_own[0] < t0_us:
            return 0, 0, False
        b0 = time_bin(t_own[0], t0_us, vt_us)
        head = time_bin(t_own[:HALO_EVENTS], t0_us, vt_us)
        skip = int(np.searchsorted(head, b0, side="right"))  # events sharing the first bin
        if skip >= min(HALO_EVENTS, len(t_own)):
            ok = False
    if rank < world - 1:
        if len(t_halo) == 0 or t_halo[0] < t0_us:
            return skip, 0, False
        b0 = time_bin(t_halo[0], t0_us, vt_us)
        keep = int(np.searchsorted(time_bin(t_halo, t0_us, vt_us), b0, side="right"))
        if keep >= len(t_halo):
            ok = False
    return skip, keep, ok


def boundary_share(t_own, rank, t0_us, vt_us):
    """The sender's half of the peer-memory exchange (csrc/evk_comm.cu k_p2p_tick_range): how many of
    this rank's leading events belong to the previous rank's last time bin.  The count travels with
    the "my events are loaded" flag and the previous rank pulls exactly that many events
    (k_p2p_pull_exact), so no rank searches a received block.  Returns (share, ok); rank 0 sends
    nothing.  Equal to time_range_split's (skip, ok) of this rank and to its predecessor's keep."""
    if rank == 0:
        return 0, True
    skip, _, ok = time_range_split(t_own, t_own[:0], rank, rank + 1, t0_us, vt_us)
    return skip, ok


def mix64(z):
    """evk_mix64 of csrc/evk_internal.cuh on a uint64 array"""
    z = np.asarray(z, dtype=np.uint64).copy()
    with np.errstate(over="ignore"):
        z ^= z >> np.uint64(33)
        z *= np.uint64(0xFF51AFD7ED558CCD)
        z ^= z >> np.uint64(33)
        z *= np.uint64(0xC4CEB9FE1A85EC53)
        z ^= z >> np.uint64(33)
    return z


def owner_mix64(keys, world):
    """EVK_OWNER_MIX64: the top 32 bits of mix64(key ^ golden), range-reduced to [0, world) by a
    multiply: (hi32 * world) >> 32 (csrc/evk_comm.cu: owner_of)"""
    k = np.asarray(keys, dtype=np.uint64) ^ np.uint64(0x9E3779B97F4A7C15)
    hi = mix64(k) >> np.uint64(32)
    return ((hi * np.uint64(world)) >> np.uint64(32)).astype(np.int64)


def merge_lowest_index(keys, first):
    """owner-side merge: one record per key, the lowest global first index wins"""
    order = np.lexsort((first, keys))
    k, f = np.asarray(keys)[order], np.asarray(first)[order]
    head = np.ones(len(k), dtype=bool)
    head[1:] = k[1:] != k[:-1]
    return k[head], f[head]


def partial_sums(points_xy, labels, K):
    """exact integer partial sums [K, 3] = (count, sum x, sum y): what the ranks allreduce"""
    out = np.zeros((K, 3), dtype=np.int64)
    m = labels >= 0
    np.add.at(out[:, 0], labels[m], 1)
    np.add.at(out[:, 1], labels[m], points_xy[m, 0].astype(np.int64))
    np.add.at(out[:, 2], labels[m], points_xy[m, 1].astype(np.int64))
    return out


def finalise(cent, sums):
    """centroid = exact sum / count rounded once to fp32; empty clusters keep their centroid"""
    cent = np.array(cent, dtype=np.float32, copy=True)
    nz = sums[:, 0] > 0
    cent[nz, 0] = (sums[nz, 1].astype(np.float64) / sums[nz, 0]).astype(np.float32)
    cent[nz, 1] = (sums[nz, 2].astype(np.float64) / sums[nz, 0]).astype(np.float32)
    return cent
