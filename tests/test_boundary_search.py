"""The boundary search of the fused sharded step (k_p2p_tick_range, csrc/evk_comm.cu): one CTA of NT
threads finds the first offset of a boundary block whose time bin differs from the first event's --
a monotone predicate on a time-ordered stream -- in rounds of NT independent probes.  This is the
kernel's arithmetic restated on the host (same probe positions, same updates of lo / hi) and checked
against the plain scan, including the shapes the GPU runs never see (blocks shorter than the CTA,
runs that end at the block's first or last event, no end inside the block)."""
import numpy as np
import pytest


def search(pred, nt=1024):
    """pred: bool array, false...false true...true; returns the first true offset, len(pred) if none;
    also the number of rounds (each round = one set of independent loads on the device)."""
    lo, hi, rounds = 0, len(pred), 0
    while lo < hi:
        rounds += 1
        span = hi - lo
        tid = np.arange(nt)
        if span <= nt:  # dense: every offset of the range has its own thread
            p = lo + tid
            live = tid < span
            hit = live & pred[np.minimum(p, len(pred) - 1)]
            f = int(np.flatnonzero(hit)[0]) if hit.any() else nt
            lo = hi = lo + f if f < nt else hi
        else:
            p = lo + span * (tid + 1) // (nt + 1)
            hit = pred[p]
            f = int(np.flatnonzero(hit)[0]) if hit.any() else nt
            p_f = lo + span * (f + 1) // (nt + 1)
            p_b = lo + span * f // (nt + 1)
            if f < nt:
                hi = p_f
            if f > 0:
                lo = p_b + 1
    return lo, rounds


@pytest.mark.parametrize("n", [1, 2, 31, 1023, 1024, 1025, 1026, 2049, 65536, 262144, 1 << 20])
def test_boundary_search_matches_scan(n):
    rng = np.random.default_rng(n)
    ends = {0, 1, n - 1, n, n // 2, n // 3, max(0, n - 2)} | set(int(v) for v in rng.integers(0, n + 1, size=40))
    for end in sorted(e for e in ends if 0 <= e <= n):
        pred = np.arange(n) >= end
        got, rounds = search(pred)
        assert got == end, (n, end, got)
        assert rounds <= 3, (n, end, rounds)   # 1024-way: two rounds for the 256 Ki-event block


def test_boundary_search_small_cta():
    """the same loop with a 32-thread 'CTA' (the round-1 form of the search: a warp, six rounds)"""
    for n in (1, 5, 33, 34, 1000, 262144):
        for end in (0, 1, n // 2, n - 1, n):
            pred = np.arange(n) >= end
            assert search(pred, nt=32)[0] == end
