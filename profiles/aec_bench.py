"""Latency of the asynchronous event clustering consumer: the CUDA path
(evk_aec_update through the C-ABI, host buffers, one call per slice) beside the reference's own
code (oracle/_ref/libref_aec.so, one host thread -- the algorithm is sequential) and the C oracle.
The algorithm is sequential by definition, so the figure is microseconds per event, not GB/s.

    python profiles/aec_bench.py > profiles/r01/aec_bench.jsonl     (on the B200 box)
"""
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
from oracle import aec  # noqa: E402


def run(evk, name, init, e, slice_len):
    n = len(e)
    rec = {"case": name, "events": n, "slice": slice_len, "params": init or "reference app defaults"}
    with evk.Evk(1024) as h:
        h.aec_create(init)
        h.aec_update(e[:slice_len])           # warm-up: module load, staging allocation
        h.aec_create(init)
        t0 = time.perf_counter()
        for i in range(0, n, slice_len):
            h.aec_update(e[i:i + slice_len])
        dt = time.perf_counter() - t0
        st = h.aec_state()
    rec["gpu_us_per_event"] = dt / n * 1e6
    rec["gpu_ms_per_slice"] = dt / (n / slice_len) * 1e3
    rec["clusters"] = int(len(st["ids"]))
    for label, cls in (("oracle", aec.Oracle), ("reference", aec.Reference)):
        if label == "reference" and not aec.ref_available():
            continue
        o = cls(init)
        t0 = time.perf_counter()
        for i in range(0, n, slice_len):
            o.update(e[i:i + slice_len])
        dt = time.perf_counter() - t0
        rec[label + "_cpu_us_per_event"] = dt / n * 1e6
        S.same_state(st, o.state())
    rec["parity"] = "state-for-state equal (cluster order, ids, n, mu bits, stored events)"
    return rec


def main():
    evk = evk_loader.load()
    evk.lib()
    cases = [
        ("reference app: defaults, 1250-event slices sharing one pseudo-time", None,
         S.stream(31, 50_000, tie=1250), 1250),
        ("defaults, distinct timestamps", None, S.stream(32, 50_000, tie=1), 1250),
        ("init(200, 10, kappa 10, 0.5, 5): random sampling", S.INITS["paper"],
         S.stream(33, 50_000, tie=7), 1250),
        ("quiet scene: 3 blobs, 2 % noise", None, S.stream(34, 50_000, blobs=3, noise=0.02, tie=50),
         1250),
    ]
    for name, init, e, sl in cases:
        print(json.dumps(run(evk, name, init, e, sl)), flush=True)


if __name__ == "__main__":
    main()
