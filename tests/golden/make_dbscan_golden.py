"""Golden vectors of DBSCAN, produced by the REFERENCE's own DBSCAN_simple.h
(oracle/_ref/libref_dbscan.so: compiled where it lies, oracle/Makefile target ref_dbscan).
Run in the development container:  python tests/golden/make_dbscan_golden.py
Writes tests/golden/dbscan_golden.json: per case the cluster sizes in the reference's order and a
digest of the canonicalised member lists."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import dbscan_cases as D  # noqa: E402
from oracle import dbscan  # noqa: E402


def digest(clusters):
    h = hashlib.sha256()
    for c in D.canon(clusters):
        h.update(repr(c).encode())
    return h.hexdigest()


def main():
    assert dbscan.ref_build(), "reference sources not present"
    out = {"generator": "oracle/_ref/libref_dbscan.so (reference DBSCAN_simple.h)", "cases": {}}
    for name in D.CASES:
        pts, (eps, mp, mn, mx) = D.cloud(name)
        cl = dbscan.reference(pts, eps, mp, mn, mx)
        members = sum(len(c) for c in cl)
        distinct = len(set(int(v) for c in cl for v in c))
        out["cases"][name] = {"sizes": [len(c) for c in cl], "digest": digest(cl),
                              "second_memberships": members - distinct}
        print(name, len(cl), "clusters", [len(c) for c in cl][:8], "second memberships",
              members - distinct)
    with open(os.path.join(HERE, "dbscan_golden.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
