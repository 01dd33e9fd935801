"""Generates the committed golden fixtures under tests/golden/.

The reference is C/OpenCL and cannot be built or imported here, so the goldens
are produced by an INDEPENDENT numpy restatement of the reference kernels' arithmetic (this
file shares no code with oracle/evk_oracle.c) applied to the reference's own deterministic
inputs:
  F1  KM/assign_to_centers2.c:121-131  (data[i] = i % 100, 8 initial centres, threshold 50)
  F2  ACCEL/store.cpp:209-215,317-326  (warm-up launch on 8192 x (0,0))
  F3  optics-clustering/test/event_raw_data8.csv (320 real events; copied as a data fixture)
plus two small synthetic streams (SURVEY.md 8d generator, restated here with numpy uint64).
tests/test_oracle.py checks the C oracle against these files; tests/test_gpu_parity.py checks
the CUDA path against them and against the oracle.

Run from the repo root (only needed to regenerate):  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import shutil

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CSV = "/root/reference/event-cam-clustering/optics-clustering/test/event_raw_data8.csv"
M64 = (1 << 64) - 1


# ---------------------------------------------------------------- downsample restatement ----
def voxel_keys(x, y, t, p, W, H, vx, vy, vt, t0, use_p):
    x = x.astype(np.int64); y = y.astype(np.int64); t = t.astype(np.int64)
    valid = (x < W) & (y < H) & (t >= t0)
    NX = -(-W // vx); NY = -(-H // vy)
    tb = (t - t0) // vt if vt > 0 else np.zeros_like(t)
    k = (tb * NY + y // vy) * NX + x // vx
    if use_p:
        k = k * 2 + (p > 0)
    return k.astype(np.uint64), valid


def ref_keys(x, y, W=1280, H=720):
    x = x.astype(np.int64); y = y.astype(np.int64)
    valid = (x >= 0) & (x <= W) & (y >= 0) & (y <= H)      # coordinate_processor.cl:56
    return ((x * 1619 + y * 31) % 8192).astype(np.uint64), valid   # :12


def first_occurrence(keys, valid):
    """unique keys with lowest stream index, canonical order = ascending first index;
    repeated = number of keys hit at least twice (coordinate_processor.cl:73-75)"""
    idx = np.nonzero(valid)[0]
    uk, first, cnt = np.unique(keys[idx], return_index=True, return_counts=True)
    first = idx[first]
    order = np.argsort(first, kind="stable")
    return uk[order], first[order].astype(np.uint32), int((cnt >= 2).sum())


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---------------------------------------------------------------- k-means restatement -------
def f32(a):
    return np.asarray(a, dtype=np.float32)


def assign(pts, cent, max_dist):
    """fp32 squared distance: t0 = dx*dx (rounded), d2 = fma(dy,dy,t0) (+ one fma per extra dim);
    strict '<', lowest k wins ties; gate d2 < max_dist^2 (assign_to_centers.cl:11-25)."""
    P, D = pts.shape
    best = np.full(P, np.inf if not (max_dist > 0) else np.float32(max_dist) * np.float32(max_dist),
                   dtype=np.float32)
    lab = np.full(P, -1, dtype=np.int32)
    for k in range(cent.shape[0]):
        dx = f32(cent[k, 0] - pts[:, 0])
        dy = f32(cent[k, 1] - pts[:, 1])
        d = f32(dy.astype(np.float64) * dy.astype(np.float64) + f32(dx * dx).astype(np.float64))
        for j in range(2, D):
            dj = f32(cent[k, j] - pts[:, j]).astype(np.float64)
            d = f32(dj * dj + d.astype(np.float64))
        m = d < best
        lab[m] = k
        best[m] = d[m]
    return lab


def update(pts, lab, cent):
    K, D = cent.shape
    new = cent.copy()
    counts = np.zeros(K, dtype=np.uint64)
    sums = np.zeros((K, D), dtype=np.float64)
    for k in range(K):
        m = lab == k
        counts[k] = m.sum()
        if counts[k]:
            sums[k] = pts[m].astype(np.float64).sum(axis=0)
            new[k] = f32(sums[k] / float(counts[k]))
    return new, counts, sums


# ---------------------------------------------------------------- synthetic stream ----------
def sm64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def ih8(u):
    s = np.zeros(u.shape, dtype=np.int64)
    for b in range(8):
        s += ((u >> np.uint64(8 * b)) & np.uint64(0xFF)).astype(np.int64)
    return s - 1020


def tdiv(a, b):
    """C integer division (truncation toward zero)"""
    q = np.abs(a) // b
    return np.where(a < 0, -q, q)


def reflect(pos, ext):
    r = np.mod(pos, 2 * ext)          # numpy mod is already non-negative for positive modulus
    return np.where(r >= ext, 2 * ext - 1 - r, r)


def synth(seed, first, n, rate, W, H, n_blobs, sigma_q8=1536, noise_q16=16384, vmax=200):
    with np.errstate(over="ignore"):
        i = np.arange(first, first + n, dtype=np.uint64)
        u0 = sm64(np.uint64(seed) ^ i); u1 = sm64(u0); u2 = sm64(u1)
        t = ((i * np.uint64(1000000)) // np.uint64(rate)).astype(np.int64)
        noise = (u0 & np.uint64(0xFFFF)).astype(np.int64) < noise_q16
        lo, hi = (u1 & np.uint64(0xFFFFFFFF)), (u1 >> np.uint64(32))
        xn = ((lo * np.uint64(W)) >> np.uint64(32)).astype(np.int64)
        yn = ((hi * np.uint64(H)) >> np.uint64(32)).astype(np.int64)
        m = ((((u0 >> np.uint64(16)) & np.uint64(0xFFFF)) * np.uint64(n_blobs)) >> np.uint64(16))
        b0 = sm64(np.uint64(seed) ^ np.uint64(0xB10B00000000) ^ (m << np.uint64(8))); b1 = sm64(b0)
        wq, hq = W * 256, H * 256
        c0x = (((b0 & np.uint64(0xFFFFFFFF)) * np.uint64(wq)) >> np.uint64(32)).astype(np.int64)
        c0y = (((b0 >> np.uint64(32)) * np.uint64(hq)) >> np.uint64(32)).astype(np.int64)
        vspan = 2 * vmax * 256 + 1
        vx = (((b1 & np.uint64(0xFFFFFFFF)) * np.uint64(vspan)) >> np.uint64(32)).astype(np.int64) - vmax * 256
        vy = (((b1 >> np.uint64(32)) * np.uint64(vspan)) >> np.uint64(32)).astype(np.int64) - vmax * 256
        cx = reflect(c0x + tdiv(vx * t, 1000000), wq)
        cy = reflect(c0y + tdiv(vy * t, 1000000), hq)
        px = cx + tdiv(ih8(u1) * sigma_q8, 209)
        py = cy + tdiv(ih8(u2) * sigma_q8, 209)
        x = np.where(noise, xn, px >> 8); y = np.where(noise, yn, py >> 8)
        x = np.clip(x, 0, W - 1); y = np.clip(y, 0, H - 1)
        p = ((u0 >> np.uint64(32)) & np.uint64(1)).astype(np.int64)
    return x, y, t, p


def kmeans_golden(x, y, t, p, first, K, D, iters, max_dist=0.0, t0=0, t_scale=1e-3, p_scale=1.0):
    cols = [f32(x[first]), f32(y[first])]
    if D > 2:
        cols.append(f32(f32(t[first] - t0) * np.float32(t_scale)))
    if D > 3:
        cols.append(f32(f32(p[first] > 0) * np.float32(p_scale)))
    pts = np.stack(cols, axis=1)
    cent = pts[:K].copy()          # first K voxel representatives in canonical order
    for _ in range(iters):
        lab = assign(pts, cent, max_dist)
        cent, counts, _ = update(pts, lab, cent)
    return lab, cent, counts


def main():
    out = {}
    # ---- F1 ---------------------------------------------------------------------------------
    data = f32(np.arange(4096) % 100)
    pts = data.reshape(-1, 2)
    c0 = f32([1, 1, 10, 10, 20, 20, 30, 30, 50, 50, 60, 60, 70, 70, 80, 80]).reshape(8, 2)
    lab = assign(pts, c0, 50.0)
    c1, counts, sums = update(pts, lab, c0)
    # literal host formula, assign_to_centers2.c:507-512 on 32 group sums of a fresh output
    ss = np.zeros(32, dtype=np.float64)
    for k in range(8):
        ss[4 * k] = sums[k, 0]; ss[4 * k + 2] = sums[k, 1]
    quirk = np.zeros(16)
    for j in range(0, 16, 2):
        quirk[j] = (ss[j] + ss[j + 1]) / counts[j // 2]
        quirk[j + 1] = (ss[j + 2] + ss[j + 3]) / counts[j // 2]
    out["F1"] = dict(labels_sha=sha(lab), labels=lab.tolist(), counts=counts.tolist(),
                     sum_x=sums[:, 0].tolist(), sum_y=sums[:, 1].tolist(),
                     centroids=c1.astype(float).ravel().tolist(),
                     quirk_centroids=quirk.tolist(), unassigned=int((lab < 0).sum()))
    # ---- F2 ---------------------------------------------------------------------------------
    k, valid = ref_keys(np.zeros(8192, np.int64), np.zeros(8192, np.int64))
    uk, first, rep = first_occurrence(k, valid)
    out["F2"] = dict(unique=len(uk), repeated=rep, keys=uk.tolist(), first=first.tolist())
    # ---- F3 ---------------------------------------------------------------------------------
    dst = os.path.join(HERE, "event_raw_data8.csv")
    if os.path.exists(REF_CSV):
        shutil.copyfile(REF_CSV, dst)
    raw = np.loadtxt(dst, delimiter=",", dtype=np.int64)
    x, y, t, p = raw[:, 0], raw[:, 1], raw[:, 2], raw[:, 3]
    f3 = dict(sha256=hashlib.sha256(open(dst, "rb").read()).hexdigest(), rows=int(len(raw)),
              exact_xy=int(len(np.unique(raw[:, :2], axis=0))), cases=[])
    k, valid = ref_keys(x, y)
    uk, first, rep = first_occurrence(k, valid)
    f3["ref_hash"] = dict(unique=len(uk), repeated=rep, keys=uk.tolist(), first=first.tolist())
    for (vx, vy, vt) in [(4, 4, 1000), (2, 2, 500), (4, 4, 100), (1, 1, 0)]:
        for up in (0, 1):
            k, valid = voxel_keys(x, y, t, p, 1280, 720, vx, vy, vt, 0, up)
            uk, first, rep = first_occurrence(k, valid)
            f3["cases"].append(dict(vx=vx, vy=vy, vt=vt, use_p=up, unique=len(uk), repeated=rep,
                                    keys_sorted_sha=sha(np.sort(uk)), first_sha=sha(first),
                                    keys=uk.tolist() if (vx, vt, up) == (4, 1000, 1) else None))
    # k-means on the F3 voxels (4,4,1000,p): K=4, D=2 and D=3, 3 iterations
    k, valid = voxel_keys(x, y, t, p, 1280, 720, 4, 4, 1000, 0, 1)
    uk, first, rep = first_occurrence(k, valid)
    for D in (2, 3, 4):
        lab, cent, counts = kmeans_golden(x, y, t, p, first, 4, D, 3)
        f3[f"kmeans_D{D}"] = dict(labels=lab.tolist(), centroids=cent.astype(float).ravel().tolist(),
                                  counts=counts.tolist())
    out["F3"] = f3
    # ---- synthetic ---------------------------------------------------------------------------
    syn = []
    for name, seed, n, rate, W, H, blobs, vox, K in [
        ("davis_20k", 0xE7CA0001, 20000, 10_000_000, 346, 260, 8, (4, 4, 1000, 1), 8),
        ("gen4_50k", 0xE7CA0003, 50000, 100_000_000, 1280, 720, 64, (2, 2, 500, 1), 64),
        ("gen4_shard", 0xE7CA0003, 4096, 100_000_000, 1280, 720, 64, (2, 2, 500, 1), 16),
    ]:
        first_index = 99_000_000 if name == "gen4_shard" else 0
        x, y, t, p = synth(seed, first_index, n, rate, W, H, blobs)
        ev = np.zeros(n, dtype=[("x", "<u2"), ("y", "<u2"), ("p", "<i2"), ("_pad", "<u2"), ("t", "<i8")])
        ev["x"], ev["y"], ev["p"], ev["t"] = x, y, p, t
        k, valid = voxel_keys(x, y, t, p, W, H, vox[0], vox[1], vox[2], 0, vox[3])
        uk, first, rep = first_occurrence(k, valid)
        lab, cent, counts = kmeans_golden(x, y, t, p, first, K, 2, 3)
        syn.append(dict(name=name, seed=seed, first_index=first_index, n=n, rate=rate, W=W, H=H,
                        blobs=blobs, vox=list(vox), K=K, events_sha=sha(ev), head=[
                            [int(x[i]), int(y[i]), int(t[i]), int(p[i])] for i in range(8)],
                        unique=len(uk), repeated=rep, keys_sorted_sha=sha(np.sort(uk)),
                        first_sha=sha(first), labels_sha=sha(lab),
                        centroids=cent.astype(float).ravel().tolist(), counts=counts.tolist()))
    out["synth"] = syn
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", os.path.join(HERE, "golden.json"))
    for s in syn:
        print(s["name"], "U =", s["unique"], "rep =", s["repeated"])


if __name__ == "__main__":
    main()
