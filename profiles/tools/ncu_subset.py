"""Committed summary of an ncu report: the metrics the design notes quote, one (metric, unit, value)
row per metric and captured kernel.

    python profiles/tools/ncu_subset.py gpurun_out/r02/ncu_slab_final.ncu-rep profiles/r02/ncu_slab_final_raw_subset.csv

Reads `ncu -i <rep> --page raw --csv` (the .ncu-rep itself stays in gpurun_out/, scratch)."""
import csv
import io
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(Kernel Name|Block Size|Grid Size|gpu__time_duration|dram__bytes|dram__throughput|gpu__dram_throughput|"
    r"smsp__inst_executed\.sum|smsp__inst_executed_op_shared|smsp__issue_active|sm__issue_active|"
    r"sm__inst_executed_pipe_|sm__pipe_|sm__throughput|sm__warps_active|launch__registers|launch__occupancy|"
    r"launch__shared_mem|l1tex__data_pipe_lsu_wavefronts|l1tex__data_bank_conflicts_pipe_lsu_mem_shared|"
    r"l1tex__throughput|lts__throughput|l1tex__t_sectors_pipe_lsu_mem_global_op_(ld|st)\.sum|"
    r"l1tex__t_requests_pipe_lsu_mem_global_op_(ld|st)\.sum|smsp__average_warps_issue_stalled|"
    r"smsp__pcsamp_warps_issue_stalled|smsp__thread_inst_executed_per_inst)")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True,
                         check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "metric", "unit", "value"])
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            short = re.sub(r"\(.*", "", name.replace("void ", "").replace("<unnamed>::", ""))
            for h, u, v in zip(hdr, units, r):
                if KEEP.match(h) and ".min." not in h and ".max." not in h:
                    w.writerow([short, h, u, v])
    print(out, len(rows) - 2, "kernel(s)")


main()
