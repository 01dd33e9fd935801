"""Generates tests/golden/fct_golden.json from the REFERENCE's own corner tracker code compiled in
place (oracle/_ref/libref_fct.so: the event callback lambda and CornerFilter::filterCorners of
event-cam-tracking/event-cam-fast-corner-tracker/metavision_time_surface_periodic_group_track.cpp).
Run in the container that has /root/reference:  python tests/golden/make_fct_golden.py"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import fct, orc  # noqa: E402
import fct_cases  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    assert fct.ref_build(), "needs /root/reference"
    out = {"streams": [], "filters": []}
    for name, ev, chunk in fct_cases.streams(orc):
        surf = np.zeros((fct.H, fct.W), np.int64)
        ranges = []
        for a in range(0, len(ev), chunk):
            c = fct.reference_callback(ev[a:a + chunk], surf, 1)
            f = fct.reference_filter(c, fct.W, fct.H, 15)
            ranges.append({"corners": int(len(c)), "corners_sha": sha(c), "kept": int(len(f)),
                           "kept_sha": sha(f)})
        out["streams"].append({"name": name, "n": int(len(ev)), "chunk": chunk, "ranges": ranges,
                               "surface_sha": sha(surf)})
    for name, xy, w, h, box in fct_cases.filter_lists():
        f = fct.reference_filter(xy, w, h, box)
        out["filters"].append({"name": name, "n": int(len(xy)), "kept": int(len(f)), "kept_sha": sha(f)})
    with open(os.path.join(ROOT, "tests", "golden", "fct_golden.json"), "w") as fo:
        json.dump(out, fo, indent=1)
    print(json.dumps({k: len(v) for k, v in out.items()}))


if __name__ == "__main__":
    main()
