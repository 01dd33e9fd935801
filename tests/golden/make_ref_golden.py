"""Golden vectors produced by the REFERENCE's own kernels (oracle/_ref/libref.so: the .cl files of
/root/reference compiled as C, oracle/Makefile target `ref`).  Run in the development container:

    python tests/golden/make_ref_golden.py      # writes tests/golden/ref_kernel_golden.json

The GPU box has no /root/reference: tests there use this file (and the prebuilt _ref library when it
travelled with the snapshot)."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def random_coords(seed, n, lo=-40, hi_x=1400, hi_y=800):
    rng = np.random.default_rng(seed)
    c = np.empty(2 * n, dtype=np.int32)
    c[0::2] = rng.integers(lo, hi_x, size=n)
    c[1::2] = rng.integers(lo, hi_y, size=n)
    return c


def main():
    assert ref.build(), "reference sources not present"
    out = {"_made_by": "tests/golden/make_ref_golden.py from coordinate_processor.cl and "
                       "assign_to_centers.cl of the reference, compiled as C (oracle/cl_shim.h)"}
    # F2: the warm-up launch, 8192 x (0, 0)   (ACCEL/store.cpp:209-215,317-326)
    u, uc, rc = ref.process_coordinates(np.zeros(16384, np.int32))
    out["F2"] = dict(unique=u.tolist(), unique_count=uc, repeated_count=rc)
    # F3: the reference's 320-event CSV
    rows = np.loadtxt(os.path.join(HERE, "event_raw_data8.csv"), delimiter=",", dtype=np.int64)
    coords = np.ascontiguousarray(rows[:, :2], dtype=np.int32).ravel()
    u, uc, rc = ref.process_coordinates(coords)
    out["F3"] = dict(unique_count=uc, repeated_count=rc, unique_sha=sha(u), unique_head=u[:8].tolist())
    # cumulative counters over two launches (they are never reset by the kernel)
    u2, uc2, rc2 = ref.process_coordinates(coords, uc, rc)
    out["F3_second_launch"] = dict(unique_count=uc2, repeated_count=rc2)
    # random launches, including gated coordinates (negative, beyond 1280 / 720, the inclusive edge)
    cases = []
    for seed, n in ((1, 8192), (2, 8192), (3, 1000), (4, 20000)):
        c = random_coords(seed, n)
        c[:8] = [1280, 720, 1281, 0, 0, 721, -1, 5]
        u, uc, rc = ref.process_coordinates(c)
        cases.append(dict(seed=seed, n=n, unique_count=uc, repeated_count=rc, unique_sha=sha(u)))
    out["random"] = cases
    # F1: the k-means fixture of KM/assign_to_centers2.c:121-131
    data = (np.arange(4096) % 100).astype(np.float32)
    centers = np.array([1, 1, 10, 10, 20, 20, 30, 30, 50, 50, 60, 60, 70, 70, 80, 80], np.float32)
    a = ref.assign_to_centers(data, centers)
    o, ci = ref.assign_data_cluster(data, a)
    out["F1"] = dict(assign_sha=sha(a), assign_head=a[:16].tolist(), cluster_index=ci.tolist(),
                     sum_x=[float(o[k * 4096:k * 4096 + 2048].sum(dtype=np.float64)) for k in range(8)],
                     sum_y=[float(o[k * 4096 + 2048:(k + 1) * 4096].sum(dtype=np.float64))
                            for k in range(8)])
    # far-away points stay unassigned (255), ties go to the lower centre
    data2 = np.array([500, 500, 5.5, 5.5, 15, 15, 0, 0, 40, 40, 65, 65], np.float32)
    out["assign_edge"] = ref.assign_to_centers(data2, centers).tolist()
    with open(os.path.join(HERE, "ref_kernel_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out)[:600])


if __name__ == "__main__":
    main()
