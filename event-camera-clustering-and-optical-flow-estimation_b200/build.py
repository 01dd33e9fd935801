"""Builds libevk.so (the C-ABI library) in-tree with nvcc for sm_100a.

    python event-camera-clustering-and-optical-flow-estimation_b200/build.py [--force]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libevk.so")
SOURCES = ["evk_api.cu", "evk_downsample.cu", "evk_slab.cu", "evk_partition.cu", "evk_kmeans.cu", "evk_synth.cu",
           "evk_comm.cu", "evk_evt2.cu", "evk_evt3.cu", "evk_aec.cu", "evk_corner.cu", "evk_dbscan.cu", "evk_optics.cu"]
NVCC = os.environ.get("EVK_NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = os.environ.get("EVK_NVCC_EXTRA", "").split() + ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall",
         "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def _deps():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.endswith(".o")]
    out += [os.path.join(HERE, "host", f) for f in os.listdir(os.path.join(HERE, "host"))]
    inc = os.path.join(os.path.dirname(HERE), "include")
    out += [os.path.join(inc, f) for f in os.listdir(inc)]
    return out


def build(force=False, verbose=False):
    if (not force and os.path.exists(LIB)
            and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in _deps())):
        return LIB
    objs = []
    log = []
    for s in SOURCES:
        o = os.path.join(CSRC, s[:-3] + ".o")
        cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, s), "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append(r.stderr)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed on " + s)
        objs.append(o)
    cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++",
           "-o", LIB] + objs + ["-lcudart", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    # C++ host demo over the C-ABI (the reference's store sample without the SDK)
    demo = os.path.join(HERE, "store_replay")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", os.path.join(HERE, "host", "store_replay.cpp"),
           "-L" + HERE, "-levk", "-Wl,-rpath,$ORIGIN", "-o", demo]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("host demo build failed")
    with open(os.path.join(HERE, "build_ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
