"""Deterministic event streams {t, x, y, p} for the AEC consumer tests (numpy PCG64: the same
bytes on every box for a given seed)."""
import hashlib

import numpy as np

INITS = {
    "default": None,   # the reference app: init() never called
    "paper": dict(sz_buffer=200, radius=10.0, kappa=10, alpha=0.5, min_n=5),
    "alpha03": dict(sz_buffer=300, radius=25.0, kappa=4, alpha=0.3, min_n=3),
    "exhaustive": dict(sz_buffer=150, radius=12.0, kappa=200, alpha=0.7, min_n=4),  # kappa > n
}


def stream(seed, n, W=1280, H=720, blobs=6, noise=0.25, dt=1e-4, tie=1, sigma=8.0):
    """blobs moving linearly + uniform noise; `tie` consecutive events share a timestamp (the
    reference hands a whole slice over with ONE pseudo-time, store.cpp:440)"""
    r = np.random.default_rng(seed)
    c0 = r.uniform([100, 100], [W - 100, H - 100], size=(blobs, 2))
    v = r.uniform(-300, 300, size=(blobs, 2))
    e = np.zeros((n, 4))
    t = np.floor(np.arange(n) / tie) * dt * tie + 5.0
    k = r.integers(0, blobs, n)
    isn = r.random(n) < noise
    xy = c0[k] + v[k] * (t - 5.0)[:, None] + r.normal(0, sigma, size=(n, 2))
    xy[isn] = r.uniform([0, 0], [W, H], size=(int(isn.sum()), 2))
    e[:, 0] = t
    e[:, 1:3] = np.clip(np.rint(xy), 0, [W - 1, H - 1])
    e[:, 3] = r.integers(0, 2, n)
    return e


def unsorted_stream(seed, n):
    """times jitter backwards: the stored lists are no longer time-sorted (the merge must still
    take 'the list whose head is oldest')"""
    e = stream(seed, n, tie=1)
    r = np.random.default_rng(seed + 1000)
    e[:, 0] += r.uniform(-0.004, 0.004, n)
    e[0, 0] = 5.0
    e[:, 0] = np.maximum(e[:, 0], 5.0)
    return e


def digest(state):
    """sha256 over everything observable (cluster order, ids, n, mu, stored events)"""
    h = hashlib.sha256()
    for k in ("ids", "n", "mu"):
        h.update(np.ascontiguousarray(state[k]).tobytes())
    h.update(np.int64(state["last"]).tobytes())
    for p in state["pts"]:
        for a in p:
            h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def same_state(a, b):
    assert (a["ids"] == b["ids"]).all(), (a["ids"], b["ids"])
    assert (a["n"] == b["n"]).all()
    assert (a["mu"] == b["mu"]).all()          # bit-exact doubles
    assert a["last"] == b["last"]
    nn = a["n"] > 0
    assert (a["cen"][nn] == b["cen"][nn]).all()
    for p, q in zip(a["pts"], b["pts"]):
        for x, y in zip(p, q):
            assert (np.asarray(x) == np.asarray(y)).all()


CASES = [  # (name, init key, stream kind, seed, n events, chunk)
    ("default_distinct_t", "default", "sorted", 1, 6000, 500),
    ("default_slices", "default", "tie50", 2, 6000, 750),
    ("paper_sampling", "paper", "tie7", 3, 5000, 333),
    ("alpha03", "alpha03", "sorted", 4, 5000, 1000),
    ("exhaustive_min", "exhaustive", "tie7", 5, 4000, 400),
    ("unsorted_times", "alpha03", "unsorted", 6, 4000, 500),
    ("dense_merges", "paper", "dense", 7, 4000, 500),
]


def make(kind, seed, n):
    if kind == "sorted":
        return stream(seed, n)
    if kind == "tie50":
        return stream(seed, n, tie=50)
    if kind == "tie7":
        return stream(seed, n, tie=7)
    if kind == "unsorted":
        return unsorted_stream(seed, n)
    if kind == "dense":   # few wide blobs, no noise: neighbouring clusters keep merging
        return stream(seed, n, blobs=3, noise=0.02, sigma=30.0, tie=3)
    raise ValueError(kind)
