"""One pass over the SURVEY 8f rows (RAW EVT 3.0 ingest, consumer, corner test, DBSCAN) for the ncu
launch list:  ncu --metrics gpu__time_duration.sum --clock-control none --csv
              --log-file gpurun_out/launches_next_rows.csv python profiles/next_rows.py
Prints wall-clock figures too when run without ncu."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import aec_streams as S  # noqa: E402
import evk_loader  # noqa: E402
from oracle import orc  # noqa: E402  (input generation only: the EVT 3.0 writer)

evk = evk_loader.load()
W, H = 1280, 720
out = {}


def timed(fn, reps=5):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    return (time.perf_counter() - t0) / reps, r


n = 5_000_000
with evk.Evk(n) as h:
    h.synth(evk.synth_params(0xE7CA0003, n, W, H, 100_000_000, 64))
    ev = h.get_events()
    ev["p"] = ev["p"] > 0
    w3 = orc.evt3_encode(ev)
    dt, m = timed(lambda: h.load_evt3(w3))
    out["evt3_ingest"] = {"events": n, "words": int(len(w3)), "ms": dt * 1e3,
                          "Mev_per_s_from_pageable_host": n / dt / 1e6}
    h.ts_create(W, H)
    dt, c = timed(lambda: h.ts_corners(False))
    out["corners"] = {"events_per_range": n, "ms": dt * 1e3, "Mev_per_s": n / dt / 1e6,
                      "corners": int(len(c))}
    U, _ = h.downsample(evk.ds_params(W, H, 4, 4, 10_000, 0, 0))
    dt, r = timed(lambda: h.dbscan_voxels(6.0, 8, 20, 1_000_000), reps=3)
    out["dbscan_voxels"] = {"points": int(U), "eps": 6.0, "min_pts": 8, "ms": dt * 1e3,
                            "Mpoints_per_s": U / dt / 1e6, "clusters": int(len(r[1])),
                            "second_memberships": int(len(r[3]))}
    h.aec_create(None, max_points=1 << 15)
    e = S.stream(31, 20_000, tie=1250)
    t0 = time.perf_counter()
    for i in range(0, len(e), 1250):
        h.aec_update(e[i:i + 1250])
    out["consumer"] = {"events": len(e), "us_per_event": (time.perf_counter() - t0) / len(e) * 1e6}
print(json.dumps(out))
