"""Golden vectors of the AEC consumer, produced by the REFERENCE's own AEClustering.cpp /
MyCluster.cpp (oracle/_ref/libref_aec.so: compiled where they lie, oracle/Makefile target ref_aec).
Run in the development container (needs /root/reference):  python tests/golden/make_aec_golden.py
Writes tests/golden/aec_golden.json: for every case, the state digest after every chunk and the
final cluster table in clear."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import aec_streams as S  # noqa: E402
from oracle import aec  # noqa: E402


def main():
    assert aec.ref_build(), "reference sources not present"
    out = {"generator": "oracle/_ref/libref_aec.so (reference AEClustering.cpp + MyCluster.cpp)",
           "cases": {}}
    for name, init, kind, seed, n, chunk in S.CASES:
        e = S.make(kind, seed, n)
        r = aec.Reference(S.INITS[init], rand_seed=1)
        digests = []
        for i in range(0, n, chunk):
            r.update(e[i:i + chunk])
            digests.append(S.digest(r.state()))
        st = r.state()
        out["cases"][name] = {
            "digests": digests, "n_clusters": int(len(st["ids"])),
            "ids": st["ids"].tolist(), "n": st["n"].tolist(),
            "mu_hex": [[float(v).hex() for v in row] for row in st["mu"]],
            "last": int(st["last"]),
        }
        print(name, len(st["ids"]), "clusters, max n", int(st["n"].max()))
    with open(os.path.join(HERE, "aec_golden.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
