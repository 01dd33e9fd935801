"""Deterministic point clouds for the DBSCAN tests (numpy PCG64)."""
import numpy as np

# name -> (seed, n, eps, min_pts, min_cluster, max_cluster, dims, noise fraction, blob sigma)
CASES = {
    "blobs2d": (0, 900, 6.0, 4, 1, 2**31 - 1, 2, 0.25, 6.0),
    "size_filter": (1, 1200, 5.0, 6, 5, 200, 2, 0.25, 6.0),
    "loose": (2, 800, 8.0, 3, 1, 2**31 - 1, 2, 0.3, 6.0),
    "tight_many_borders": (3, 1500, 3.0, 4, 1, 2**31 - 1, 2, 0.25, 6.0),
    "app_parameters": (4, 2500, 20.0, 20, 100, 25000, 2, 0.2, 25.0),   # pcl_cluster.cpp:112-120
    "blobs3d": (5, 1200, 7.0, 5, 3, 2**31 - 1, 3, 0.25, 6.0),
    "duplicates": (6, 1500, 2.0, 5, 1, 2**31 - 1, 2, 0.1, 3.0),        # many coincident points
}


def cloud(name):
    seed, n, eps, mp, mn, mx, dims, noise, sigma = CASES[name]
    r = np.random.default_rng(seed)
    span = 600.0 if name == "app_parameters" else 200.0
    c = r.uniform(0.1 * span, 0.9 * span, size=(6, dims))
    k = r.integers(0, 6, n)
    pts = np.rint(c[k] + r.normal(0, sigma, size=(n, dims)))
    isn = r.random(n) < noise
    pts[isn] = np.rint(r.uniform(0, span, size=(int(isn.sum()), dims)))
    return pts.astype(np.float32), (eps, mp, mn, mx)


def canon(clusters):
    """clusters as a sorted list of member tuples (the reference's order among equal sizes is
    whatever std::sort leaves)"""
    return sorted((tuple(int(v) for v in c) for c in clusters), key=lambda t: (-len(t), t))


def clusters_from(labels, sizes, extra):
    """member lists from the C-ABI's results (labels + second memberships)"""
    out = [[] for _ in sizes]
    for i, l in enumerate(labels):
        if l >= 0:
            out[l].append(i)
    for p, c in extra:
        out[int(c)].append(int(p))
    return [np.array(sorted(m), np.int32) for m in out]
