"""Inputs of the corner tracker tests (time surface + corner test + box-NMS), shared by the golden
generator (tests/golden/make_fct_golden.py, run against the reference's own code) and the tests."""
import numpy as np


def streams(orc):
    """(name, events, events per callback range): Gen4 frames (the reference's callback hard-codes
    1280 x 720), events kept 4 px away from the border so that its literal `break` does not end every
    range at the first border event -- one stream keeps them to exercise exactly that"""
    out = []
    for name, seed, n, rate, blobs, chunk, crop in (("blobs12", 0xE7CA0011, 400_000, 20_000_000, 12, 50_000, True),
                                                    ("dense64", 0xE7CA0012, 600_000, 50_000_000, 64, 100_000, True),
                                                    ("with_border", 0xE7CA0013, 200_000, 20_000_000, 8, 40_000, False)):
        ev = orc.synth(orc.synth_params(seed, n, 1280, 720, rate, blobs))
        if crop:
            m = (ev["x"] >= 4) & (ev["x"] < 1276) & (ev["y"] >= 4) & (ev["y"] < 716)
            ev = ev[m]
        out.append((name, ev, chunk))
    return out


def filter_lists():
    """(name, xy [n, 2], width, height, box size) for CornerFilter::filterCorners"""
    r = np.random.default_rng(17)
    out = []
    out.append(("uniform_3000", np.stack([r.integers(0, 1280, 3000), r.integers(0, 720, 3000)], 1), 1280, 720, 15))
    c = r.integers(40, 600, size=(12, 2))
    pts = (c[r.integers(0, 12, 5000)] + r.normal(0, 9, size=(5000, 2))).round().astype(int)
    out.append(("clusters_5000", np.clip(pts, 0, [639, 479]), 640, 480, 15))
    out.append(("duplicates", np.repeat(r.integers(0, 200, size=(50, 2)), 20, axis=0), 200, 200, 9))
    out.append(("on_the_border", np.array([[0, 0], [7, 0], [14, 0], [15, 0], [199, 199], [192, 199],
                                           [199, 184], [0, 199], [8, 192]]), 200, 200, 15))
    out.append(("box_1", np.stack([r.integers(0, 64, 2000), r.integers(0, 64, 2000)], 1), 64, 64, 1))
    out.append(("box_even_16", np.stack([r.integers(0, 300, 4000), r.integers(0, 300, 4000)], 1), 300, 300, 16))
    out.append(("dense_20000", np.stack([r.integers(0, 1280, 20000), r.integers(0, 720, 20000)], 1), 1280, 720, 15))
    return [(n, np.ascontiguousarray(xy, dtype=np.int32), w, h, b) for n, xy, w, h, b in out]
