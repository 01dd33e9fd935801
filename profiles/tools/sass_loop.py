"""Static size of k_slab_main's tile loop in a variant object (no GPU needed).

    python profiles/tools/sass_loop.py variants/trim4l_evk_slab.o [--dump]

Prints, for the <COUNT_REP=true, POW2=true> instance, the instruction count between the tile loop's
mbarrier try-wait and its back edge, and a mnemonic histogram.  A proxy for the per-tile-warp dynamic
count the ncu source page gives (rare paths are in the static count too)."""
import collections
import re
import subprocess
import sys


def main():
    obj = sys.argv[1]
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", out)
    body = next(f for f in funcs if f.startswith("_ZN") and "k_slab_main" in f.split("\n")[0]
                and "Lb1ELb1E" in f.split("\n")[0])
    ins = []
    for line in body.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    first = next(i for i, (_, s) in enumerate(ins) if "SYNCS.PHASECHK" in s)
    # the loop's back edge: last branch that targets an address at or before the try-wait block
    head = ins[first][0]
    last = first
    for i, (a, s) in enumerate(ins):
        m = re.search(r"BRA\s+(?:P\d,\s*)?`?\(?\.?L?_?x?_?\d*\)?|BRA.*0x([0-9a-f]+)", s)
        t = re.search(r"0x([0-9a-f]+)", s) if "BRA" in s else None
        if t and i > first and head - 0x200 <= int(t.group(1), 16) <= head:
            last = i
    loop = ins[first:last + 1]
    hist = collections.Counter(re.sub(r"^@!?U?P\d\s+", "", s).split()[0].split(".")[0] for _, s in loop)
    print(f"{obj}: function {len(ins)} instr, tile loop ~{len(loop)} instr "
          f"({ins[first][0]:#x}..{ins[last][0]:#x})")
    print("  " + " ".join(f"{k}:{v}" for k, v in hist.most_common(24)))
    if "--dump" in sys.argv:
        for a, s in loop:
            print(f"{a:06x} {s}")


main()
