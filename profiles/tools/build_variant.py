"""A/B builds: one libevk_<name>.so per set of compile flags.

    python profiles/tools/build_variant.py <name> [file.cu ...] -- -DFLAG=1 ...

Recompiles only the listed sources (default: evk_slab.cu) with the extra flags and links them with
the objects of the regular build (run the package's build.py first).  The variants live under
gpurun_out/../variants/ (git-ignored *.so) and are selected with EVK_LIB=<path>."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = os.path.join(ROOT, "event-camera-clustering-and-optical-flow-estimation_b200")
CSRC = os.path.join(PKG, "csrc")
OUT = os.path.join(ROOT, "variants")
NVCC = "/usr/local/cuda/bin/nvcc"
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-ccbin",
         "/usr/bin/g++", "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall", "--expt-relaxed-constexpr"]


def main():
    argv = sys.argv[1:]
    extra = []
    if "--" in argv:
        i = argv.index("--")
        argv, extra = argv[:i], argv[i + 1:]
    name, files = argv[0], argv[1:] or ["evk_slab.cu"]
    os.makedirs(OUT, exist_ok=True)
    objs = []
    for f in sorted(os.listdir(CSRC)):
        if not f.endswith(".cu"):
            continue
        if f in files:
            o = os.path.join(OUT, f"{name}_{f[:-3]}.o")
            subprocess.run([NVCC] + FLAGS + extra + ["-c", os.path.join(CSRC, f), "-o", o], check=True)
        else:
            o = os.path.join(CSRC, f[:-3] + ".o")
        objs.append(o)
    lib = os.path.join(OUT, f"libevk_{name}.so")
    subprocess.run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin",
                    "/usr/bin/g++", "-o", lib] + objs + ["-lcudart", "-ldl"], check=True)
    print(lib)


if __name__ == "__main__":
    main()
