#!/usr/bin/env python
"""bench.py — Mevents/s of the fused hot path: voxel-hash downsample + one k-means iteration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl evk|reference]

Workload at N=1 (BASELINE.json configs[2], the config the metric is quoted on): 100 M synthetic
Prophesee Gen4 (1280x720) events at 100 Mev/s, 2x2 px x 500 us voxels (+ polarity), k-means K=64,
D=2.  A step = ONE call of the fused entry point evk_downsample_kmeans (downsample -> first-K
centroid initialisation -> one assign + accumulate iteration -> finalise; a replayed CUDA graph,
one host synchronisation).  --unfused runs the three separate calls instead (same results).
N>1: weak scaling, each rank owns a contiguous 100 M-event index shard of an N x 100 M-event stream
(configs[3]); evk_downsample_kmeans_sharded: boundary blocks move between neighbours (time-range
ownership; --owner mix64 = hash ownership with an NCCL all-to-all of voxels) and the K x (D+1)
partial sums are allreduced.

`value`  : events resident in HBM, timed with CUDA events on the library's stream.
`e2e`    : the same step through the C-ABI with HOST buffers, everything the reference reads back
           inside the timed region: H2D of the events (pinned), the step, D2H of the unique voxels'
           keys (ACCEL/store.cpp:412-422 reads the unique coordinates every slice), of the labels
           (KM/assign_to_centers2.c:259-265 reads the assignments) and of centroids + counts.  Two
           handles alternate so that one slice's read-back overlaps the next slice's upload (PCIe is
           full duplex); `e2e_serial` is the same on ONE handle, `e2e_centroids` reads back
           centroids + counts only (round 1's e2e), `h2d_only` is the bare copy (the host limit).
`c5`, `c1`: BASELINE configs[4] (50 ms windows pushed from host memory: per-window latency) and
           configs[0] (the reference's CPU-runnable case) beside their CPU timings.
`roofline`: dominant kernel, algorithmic bytes / its CUDA-event duration, against
           MEASURED_PEAKS.json hbm_gbs (fallback 6650 GB/s).
`cpu_baseline`: the CPU oracle (a port of the reference's semantics; the reference's own OpenCL
           code cannot be built here) on all host threads over a bounded prefix of the workload.
--impl reference times that same oracle as the reference arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 1280, 720
RATE = 100_000_000
VOX = (2, 2, 500, 1)
K, D = 64, 2
SEED = 0xE7CA0003
N_BLOBS = 64


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="evk", choices=["evk", "reference"])
    ap.add_argument("--events", type=int, default=100_000_000, help="events per GPU")
    ap.add_argument("--cpu-sample", type=int, default=20_000_000)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--owner", default="time", choices=["time", "mix64"])
    ap.add_argument("--no-repeated", action="store_true", help="do not count repeated keys")
    ap.add_argument("--unfused", action="store_true",
                    help="three separate calls (downsample, init, k-means) instead of the fused step")
    ap.add_argument("--algo", default="auto", choices=["auto", "table", "sort", "slab", "partition"])
    ap.add_argument("--sync-steps", action="store_true",
                    help="one host synchronisation per timed step instead of a queued pipeline")
    ap.add_argument("--config", default="c3", choices=["c3", "c4"],
                    help="c3: weak scaling, --events per GPU (default); c4: ONE 1 B-event stream "
                         "strong-scaled over the GPUs (BASELINE configs[3])")
    ap.add_argument("--c4-events", type=int, default=1_000_000_000)
    ap.add_argument("--no-extras", action="store_true", help="skip the c5 / c1 / c4 / consumer legs")
    ap.add_argument("--c5-windows", type=int, default=200)
    return ap.parse_args()


def host_threads():
    """threads this process may use on the box -- never OMP_NUM_THREADS (torchrun exports 1)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def pin_to_gpu_numa(local_rank):
    """Bind this rank (and the pinned buffers it allocates afterwards) to the NUMA node of its GPU:
    8 ranks pinning 1.6 GB each on node 0 halve the host->device rate of the far GPUs."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        node = int(open(base + "/numa_node").read())
        cpus = open(base + "/local_cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            if "-" in part:
                a, b = part.split("-")
                ids.update(range(int(a), int(b) + 1))
            elif part:
                ids.add(int(part))
        allowed = set(os.sched_getaffinity(0))
        ids &= allowed
        if ids:
            os.sched_setaffinity(0, ids)
        return {"gpu": bdf, "numa_node": node, "cpus": len(ids) or len(allowed)}
    except Exception as e:   # containers without sysfs: leave the affinity alone
        return {"error": str(e)[:80]}


_POLL_SRC = r"""
import sys, time
import pynvml as n
n.nvmlInit()
dev = n.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
mx = n.nvmlDeviceGetMaxClockInfo(dev, n.NVML_CLOCK_SM)
def pick(*names):
    for k in names:
        if hasattr(n, k):
            return getattr(n, k)
    return 0
bits = [pick("nvmlClocksEventReasonHwSlowdown", "nvmlClocksThrottleReasonHwSlowdown"),
        pick("nvmlClocksEventReasonHwThermalSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
        pick("nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
        pick("nvmlClocksEventReasonSwPowerCap", "nvmlClocksThrottleReasonSwPowerCap")]
reasons = pick("nvmlDeviceGetCurrentClocksEventReasons", "nvmlDeviceGetCurrentClocksThrottleReasons")
n.nvmlDeviceGetClockInfo(dev, n.NVML_CLOCK_SM)
print("ready", flush=True)
while True:
    t = time.time()
    sm = n.nvmlDeviceGetClockInfo(dev, n.NVML_CLOCK_SM)
    try:
        r = reasons(dev) if reasons else 0
    except Exception:
        r = 0
    try:
        pw = n.nvmlDeviceGetPowerUsage(dev) / 1000.0
    except Exception:
        pw = 0.0
    print(t, sm, mx, pw, *[1 if (r & b) else 0 for b in bits], flush=True)
    time.sleep(0.001)
"""


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region.  The region is tens of
    milliseconds long, so NVML is polled every ~1-2 ms (nvidia-smi -lms cannot sample that fast) --
    from a separate PROCESS, so that this process's interpreter lock (the timed loop is a tight loop
    of C-ABI calls) cannot starve the sampler.  Rows carry wall-clock stamps; stop(t0, t1) keeps the
    rows taken inside [t0, t1].  nvidia-smi is the fallback when pynvml is missing."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc, self.source = index, [], None, "none"

    def start(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        idx = self.index  # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES indices
        if vis and all(v.strip().isdigit() for v in vis.split(",")):
            idx = int(vis.split(",")[self.index])
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", _POLL_SRC, str(idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            if self.proc.stdout.readline().strip() != "ready":   # first NVML queries are slow
                raise RuntimeError("sampler did not start")
            self.source = "nvml"
        except Exception:
            if self.proc:
                self.proc.kill()
            try:
                self.proc = subprocess.Popen(
                    ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                     "--format=csv,noheader,nounits", "-lms", "20"],
                    stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.source = "nvidia-smi"
            except Exception:
                self.proc = None
                return
        self.t = threading.Thread(target=self._read, daemon=True)
        self.t.start()

    def _read(self):
        for line in self.proc.stdout:
            if self.source == "nvml":
                f = line.split()
                self.rows.append([float(f[0])] + f[1:4] + ["Active" if v == "1" else "Not Active"
                                                           for v in f[4:8]])
            else:
                self.rows.append([time.time()] + [c.strip() for c in line.split(",")])

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock source"]}
        time.sleep(0.02)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        self.t.join(timeout=1)
        rows = self.rows
        if t0 is not None:
            inside = [r for r in rows if t0 <= r[0] <= t1]
            rows = inside or rows   # (a region shorter than one poll: keep the surrounding rows)
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for n, v in zip(names, r[4:8]):
                    if str(v).lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm),
                "window": "timed steps + the synchronous stage-time pass of the same steps",
                "reasons": sorted(reasons), "source": self.source}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu summary, if any"""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def cpu_port(n_sample, steps, threads):
    """CPU oracle (port) on a bounded prefix of the same stream. Returns Mev/s, description."""
    from oracle import orc
    orc.build()
    threads = threads or host_threads()
    ev = orc.synth(orc.synth_params(SEED, n_sample, W, H, RATE, N_BLOBS), threads=threads)
    p = orc.ds_params(W, H, VOX[0], VOX[1], VOX[2], 0, VOX[3])
    best = None
    U = 0
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        keys, first, rep = orc.downsample(ev, p, threads=threads, canonical=False)
        pts = orc.points(ev, first, D)
        # first-K initialisation needs the K lowest first indices only
        import numpy as np
        init = pts[np.argsort(first, kind="stable")[:K]]
        orc.kmeans(pts, init, iters=1, threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        U = len(keys)
    return n_sample / best / 1e6, threads, U, best


def consumer_leg(evk):
    """SURVEY 8f rank 1, the reference's consumer of the downsampled coordinates (asynchronous event
    clustering, ACCEL/AEClustering.cpp): sequential by definition, so the figure is microseconds per
    event.  The CUDA path (evk_aec_update through the C-ABI, host buffers, one call per 1250-event
    slice) beside the reference's OWN code (oracle/_ref/libref_aec.so: its sources compiled where
    they lie; one host thread) on the reference app's configuration; states must be equal."""
    import numpy as np
    from oracle import aec
    r = np.random.default_rng(31)
    n, sl = 20_000, 1250
    c0 = r.uniform([100, 100], [1180, 620], size=(6, 2))
    v = r.uniform(-300, 300, size=(6, 2))
    t = np.floor(np.arange(n) / sl) * 0.05 + 5.0          # one pseudo-time per 50 ms slice
    k = r.integers(0, 6, n)
    xy = c0[k] + v[k] * (t - 5.0)[:, None] + r.normal(0, 8, size=(n, 2))
    noise = r.random(n) < 0.25
    xy[noise] = r.uniform([0, 0], [1280, 720], size=(int(noise.sum()), 2))
    e = np.zeros((n, 4))
    e[:, 0], e[:, 1:3] = t, np.clip(np.rint(xy), 0, [1279, 719])
    out = {"workload": f"{n} events in {sl}-event slices (one pseudo-time per slice), default-"
                       "constructed AEClustering (szBuffer 800, radius 40, alpha 0.5, minN 10)",
           "unit": "us/event"}
    with evk.Evk(1024) as h:
        h.aec_create(None)
        h.aec_update(e[:sl])
        h.aec_create(None)
        t0 = time.perf_counter()
        for i in range(0, n, sl):
            h.aec_update(e[i:i + sl])
        out["value"] = (time.perf_counter() - t0) / n * 1e6
        st = h.aec_state()
    kind, cls = ("reference", aec.Reference) if aec.ref_available() else ("port", aec.Oracle)
    o = cls(None)
    t0 = time.perf_counter()
    for i in range(0, n, sl):
        o.update(e[i:i + sl])
    out["cpu_baseline"] = {"value": (time.perf_counter() - t0) / n * 1e6, "unit": "us/event",
                           "cores": 1, "kind": kind}
    so = o.state()
    out["clusters"] = int(len(st["ids"]))
    out["state_equal"] = bool((st["ids"] == so["ids"]).all() and (st["n"] == so["n"]).all()
                              and (st["mu"] == so["mu"]).all())
    return out


def cpu_c1(threads_list):
    """BASELINE configs[0], the reference's CPU-runnable case: 1 M synthetic DAVIS346 events,
    4x4 px x 1 ms voxels, k-means K = 8 (one iteration), on the CPU oracle."""
    import numpy as np
    from oracle import orc
    orc.build()
    n, w, h, k = 1_000_000, 346, 260, 8
    ev = orc.synth(orc.synth_params(0xE7CA0001, n, w, h, 10_000_000, 8))
    p = orc.ds_params(w, h, 4, 4, 1000, 0, 1)
    out = {}
    for t in threads_list:
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            keys, first, rep = orc.downsample(ev, p, threads=t, canonical=False)
            pts = orc.points(ev, first, 2)
            init = pts[np.argsort(first, kind="stable")[:k]]
            orc.kmeans(pts, init, iters=1, threads=t)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        out[t] = (n / best / 1e6, best * 1e3, len(keys))
    return out


def run_reference(args, rank, world, out):
    """The reference arm: the reference's CPU semantics (oracle port; its OpenCL / Metavision host
    code cannot be built here) on ALL host threads of the box, rank 0 only.  The CPU rate does not
    depend on how many GPUs the other arm uses: the sample is a prefix of the same stream."""
    if rank != 0:
        return
    threads = host_threads()
    n_sample = min(args.cpu_sample, args.events)
    mev, threads, U, best = cpu_port(n_sample, args.steps + args.warmup, threads)
    sample = (f"first {n_sample} events of the workload stream (U={U}), best of "
              f"{args.steps + args.warmup} passes, generation excluded")
    line = {
        "impl": "reference", "metric": "Mevents/s downsample+k-means iteration", "value": mev,
        "unit": "Mevents/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": best * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.config == "c4" else "weak",
        "vs_baseline": None, "dtype": "u64 keys / f32 distances", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": mev, "unit": "Mevents/s", "cores": threads, "kind": "port",
                         "sample": sample,
                         "threads_source": "sched_getaffinity (OMP_NUM_THREADS is ignored: torchrun "
                                           "exports 1)"},
        "e2e": {"value": mev, "unit": "Mevents/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference's OpenCL/Metavision code cannot be built here; this is the CPU oracle "
                "port of its semantics on all host threads of the box (one host whatever N is)",
    }
    print(json.dumps(line), file=out, flush=True)


def workload_config(args, world):
    if args.config == "c4":
        per = args.c4_events // world
        return {"workload": f"C4 synthetic Prophesee Gen4 {W}x{H}, ONE {args.c4_events}-event stream "
                            f"at {RATE // 1_000_000} Mev/s index-sharded over {world} GPUs ({per} "
                            f"events/GPU, strong scaling), voxels {VOX[0]}x{VOX[1]} px x {VOX[2]} us"
                            f"{' x polarity' if VOX[3] else ''}, k-means K={K} D={D}, 1 iteration",
                "events_per_gpu": per, "events_total": per * world, "K": K, "D": D,
                "voxel": list(VOX), "seed": hex(SEED), "parallelism": f"index-sharded x{world}",
                "l2_policy": "inputs larger than L2; no flush needed"}
    return {"workload": f"C3 synthetic Prophesee Gen4 {W}x{H}, {args.events} events/GPU at "
                        f"{RATE // 1_000_000} Mev/s, voxels {VOX[0]}x{VOX[1]} px x {VOX[2]} us"
                        f"{' x polarity' if VOX[3] else ''}, k-means K={K} D={D}, 1 iteration",
            "events_per_gpu": args.events, "events_total": args.events * world, "K": K, "D": D,
            "voxel": list(VOX), "seed": hex(SEED), "parallelism": f"index-sharded x{world}",
            "l2_policy": "inputs (1.6 GB/GPU) larger than L2; no flush needed"}


def main():
    args = parse()
    # the JSON line is the only thing on stdout: libraries (NCCL banner, warnings) go to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    try:
        _main(args, real_stdout)
    finally:
        real_stdout.flush()


def c5_leg(evk, torch, device, n_windows):
    """BASELINE configs[4]: 50 ms windows at 100 Mev/s (5 M events each) pushed from pinned host
    memory through evk_window_push -- per-window latency = H2D + downsample + two warm-started
    k-means iterations + the host synchronisation.  Windows are tumbling 50 ms slices, which is
    what the reference's reslicer produces (ACCEL/store.cpp:329,349-352: make_n_us(50000))."""
    import numpy as np
    per, win_us = 5_000_000, 50_000
    ds = evk.ds_params(W, H, VOX[0], VOX[1], VOX[2], 0, VOX[3])
    km = evk.km_params(K, D, iters=2)
    gen = evk.Evk(per, device=device)
    h = evk.Evk(per + 1, device=device)
    bufs = [torch.empty(per * 16, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    h.window_config(ds, km, win_us)
    lat = []
    for w in range(n_windows + 3):
        buf = bufs[w & 1]
        gen.synth(evk.synth_params(SEED, per, W, H, RATE, N_BLOBS, first_index=w * per))
        buf.numpy().view(evk.EVENT_DTYPE)[:] = gen.get_events()
        t0 = time.perf_counter()
        # window w completes when the first event of window w+1 arrives: push the window, then
        # flush it (the reslicer's callback fires on the slice boundary)
        done = h.window_push_ptr(buf.data_ptr(), per)
        done += h.window_flush()
        dt = (time.perf_counter() - t0) * 1e3
        assert done == 1, done
        if w >= 3:
            lat.append(dt)
    u, r = h.num_voxels()
    h.close()
    gen.close()
    lat.sort()
    return {"windows": len(lat), "events_per_window": per, "window_ms": 50.0,
            "p50_ms": lat[len(lat) // 2], "p99_ms": lat[min(len(lat) - 1, int(len(lat) * 0.99))],
            "max_ms": lat[-1], "mean_ms": sum(lat) / len(lat),
            "real_time_factor": 50.0 / lat[len(lat) // 2],
            "includes": "H2D of the window's 80 MB from pinned host memory, downsample, 2 warm-started "
                        "k-means iterations, one host synchronisation (evk_window_push + _flush)",
            "windows_are": "tumbling 50 ms slices (the reference's make_n_us(50000) reslicer)",
            "last_window_voxels": u}


def c1_leg(evk, device):
    """BASELINE configs[0] on the GPU (fused step) beside the CPU oracle, one and all threads"""
    n, w, hh, k = 1_000_000, 346, 260, 8
    with evk.Evk(n, device=device) as h:
        h.synth(evk.synth_params(0xE7CA0001, n, w, hh, 10_000_000, 8))
        ds = evk.ds_params(w, hh, 4, 4, 1000, 0, 1)
        km = evk.km_params(k, 2, iters=1)
        for _ in range(3):
            h.downsample_kmeans(ds, km, True)
        h.timer_start()
        for _ in range(20):
            h.downsample_kmeans_submit(ds, km, True)
        u, r, _ = h.downsample_kmeans_wait()
        ms = h.timer_stop() / 20
    T = host_threads()
    cpu = cpu_c1([1, T] if T > 1 else [1])
    return {"workload": "C1: 1 M synthetic DAVIS346 (346x260) events, 4x4 px x 1 ms voxels, K=8, "
                        "1 iteration (BASELINE configs[0], the reference's CPU-runnable case)",
            "gpu_ms_per_step": ms, "gpu_Mevents_per_s": n / ms / 1e3, "unique_voxels": u,
            "cpu_baseline_c1": {"unit": "Mevents/s", "kind": "port",
                                "single_thread": {"value": cpu[1][0], "ms": cpu[1][1], "cores": 1},
                                "all_threads": {"value": cpu[T][0], "ms": cpu[T][1], "cores": T}}}


def _main(args, real_stdout):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world, real_stdout)
        return
    import numpy as np
    import torch
    import torch.distributed as dist
    import evk_loader
    evk = evk_loader.load()
    evk.lib()  # fails loudly when libevk.so is missing: there is no fallback
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    numa = pin_to_gpu_numa(local_rank) if world > 1 else {"note": "single rank: not pinned"}
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    strong = args.config == "c4"
    n = args.c4_events // world if strong else args.events
    algo = {"auto": evk.ALGO_AUTO, "table": evk.ALGO_TABLE, "sort": evk.ALGO_SORT,
            "slab": evk.ALGO_SLAB, "partition": evk.ALGO_PARTITION}[args.algo]
    ds = evk.ds_params(W, H, VOX[0], VOX[1], VOX[2], 0, VOX[3], algo=algo,
                       count_repeated=0 if args.no_repeated else 1)
    km = evk.km_params(K, D, iters=1)
    # sharded runs receive the next rank's boundary block behind their own events
    h = evk.Evk(n + (1 << 19) if world > 1 else n, device=local_rank)
    # index sharding of the N x n-event stream (the rule is stated in <package>/sharding.py)
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "evk_sharding", os.path.join(evk_loader.PKG_DIR, "sharding.py"))
    sharding = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sharding)
    shard_lo, shard_hi = sharding.shard_range(n * world, rank, world)
    assert shard_hi - shard_lo == n
    h.synth(evk.synth_params(SEED, n, W, H, RATE, N_BLOBS, first_index=shard_lo))
    h.sync()
    if world > 1:
        uid = [evk.Evk.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        h.comm_init(rank, world, uid[0])
        h.set_shard(shard_lo)
    owner = evk.OWNER_TIME_RANGE if args.owner == "time" else evk.OWNER_MIX64

    state = {}

    def step(hh=None):
        hh = hh or h
        if world > 1 and args.unfused:
            ul, ug = hh.downsample_sharded(ds, owner)
            hh.init_centroids_first_k_sharded(km)
            hh.kmeans_sharded(km)
            state["U_local"], state["U"] = ul, ug
        elif world > 1:
            ul, ug, _ = hh.downsample_kmeans_sharded(ds, km, True, owner)
            state["U_local"], state["U"] = ul, ug
        elif args.unfused:
            u, r = hh.downsample(ds)
            hh.init_centroids_first_k(km)
            hh.kmeans(km)
            state["U_local"] = state["U"] = u
            state["R"] = r
        else:
            u, r, _ = hh.downsample_kmeans(ds, km, True)
            state["U_local"] = state["U"] = u
            state["R"] = r

    # the timed region runs WITHOUT the stage-time instrumentation (CUDA event records between the
    # kernels of the step's graph cost about 20 us per step); the stage times come from the second,
    # synchronous pass below.
    # the timed region lasts tens of milliseconds: the sampler also covers the warm-up steps of the
    # same workload right before it
    h.set_profiling(False)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    # ---- timed region: K steps, inputs resident in HBM -----------------------------------------
    # One GPU, fused step: the K steps are QUEUED (evk_downsample_kmeans_submit) and collected by
    # one evk_downsample_kmeans_wait -- a slice pipeline never idles the device between slices
    # (the synchronous call costs one host wake-up + relaunch per step); sharded runs queue
    # evk_downsample_kmeans_sharded_submit the same way.  Unfused runs synchronise inside every step.
    pipelined = not args.unfused and not args.sync_steps
    launches = 0
    clk_t0 = time.time()
    h.timer_start()
    if pipelined and world > 1:
        for _ in range(args.steps):
            h.downsample_kmeans_sharded_submit(ds, km, True, owner)
        state["U_local"], state["U"], _ = h.downsample_kmeans_sharded_wait()
    elif pipelined:
        for _ in range(args.steps):
            h.downsample_kmeans_submit(ds, km, True)
        u, r, _ = h.downsample_kmeans_wait()
        state["U_local"] = state["U"] = u
        state["R"] = r
    else:
        for _ in range(args.steps):
            step()
    total_ms = h.timer_stop()
    barrier()
    # ---- stage times: the same K steps again, synchronous, CUDA events around every stage ------
    # (the events are nodes of the step's graph: they can only be read once a replay has finished)
    ds_main = ds_total = km_total = 0.0
    h.set_profiling(True)
    step()   # (untimed: the instrumented graph is captured here)
    sync_t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
        t = h.stage_times()
        ds_main += t.ds_main_ms; ds_total += t.ds_total_ms; km_total += t.km_total_ms
        launches += t.ds_launches + t.km_launches + (1 if args.unfused else 0)
    sync_ms_per_step = (time.perf_counter() - sync_t0) * 1e3 / args.steps
    h.set_profiling(False)
    barrier()
    clocks = sampler.stop(clk_t0, time.time()) if rank == 0 else None
    algo_used = h.stage_times().ds_algo_used
    ms_per_step = max_over_ranks(total_ms) / args.steps
    value = n * world / (ms_per_step * 1e-3) / 1e6

    if strong:   # C4 is a device-resident scaling measurement: no host legs
        if rank == 0:
            peak, _ = peaks()
            U = state["U_local"]
            step_bytes = 16.0 * n + 36.0 * U
            line = {"metric": "Mevents/s downsample+k-means iteration", "value": value,
                    "unit": "Mevents/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                    "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                    "vs_baseline": None, "dtype": "u64 keys / f32 distances / exact u64 sums",
                    "data": "synthetic", "config": workload_config(args, world),
                    "owner": args.owner, "unique_voxels_rank0": U, "unique_voxels_total": state["U"],
                    "ds_algo": {1: "table", 2: "sort", 3: "slab", 4: "partition"}.get(algo_used),
                    "stage_ms": {"downsample_dominant_kernel": ds_main / args.steps,
                                 "downsample_total": ds_total / args.steps,
                                 "kmeans_iteration": km_total / args.steps},
                    "step_roofline_frac_per_gpu": step_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                    "gpu_launches": launches, "clocks": clocks, "e2e": None}
            print(json.dumps(line), file=real_stdout, flush=True)
        h.close()
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- end to end through the C-ABI with host buffers ----------------------------------------
    # (a) everything the reference reads back (e2e), (b) centroids + counts only, (c) the bare H2D
    U_cap = state["U_local"] + (1 << 16)
    host = torch.empty(n * 16, dtype=torch.uint8, pin_memory=True)
    host_np = host.numpy().view(evk.EVENT_DTYPE)
    host_np[:] = h.get_events()
    e2e_steps = max(1, args.e2e_steps)
    d2h_full = lambda u: 8 * u + 4 * u + K * D * 4 + K * 8 + 64

    def timed(fn, reps):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return max_over_ranks((time.perf_counter() - t0) / reps)

    def e2e_centroids_step():
        h.load_events_ptr(host.data_ptr(), n)
        step()
        return h.get_centroids(K, D)[0]   # D2H + sync

    cent = e2e_centroids_step()
    e2e_cent_s = timed(e2e_centroids_step, e2e_steps)

    def h2d_only():
        h.load_events_ptr(host.data_ptr(), n)
        h.sync()

    h2d_s = timed(h2d_only, e2e_steps)

    out_keys = [torch.empty(U_cap, dtype=torch.int64, pin_memory=True) for _ in range(2)]
    out_lab = [torch.empty(U_cap, dtype=torch.int32, pin_memory=True) for _ in range(2)]

    def read_back(hh, slot):
        u = hh.num_voxels()[0]
        hh.get_voxels_ptr(out_keys[slot].data_ptr(), 0, 0, U_cap)   # canonical order, D2H + sync
        hh.get_labels_ptr(out_lab[slot].data_ptr(), U_cap)
        return u, hh.get_centroids(K, D)[0]

    def e2e_serial_step():
        h.load_events_ptr(host.data_ptr(), n)
        step()
        return read_back(h, 0)

    u_s, cent_s = e2e_serial_step()
    assert u_s == state["U_local"] and (cent_s == cent).all()
    e2e_serial_s = timed(e2e_serial_step, e2e_steps)

    # two handles alternate: slice i's read-back (D2H) overlaps slice i+1's upload (H2D)
    e2e_pipe_s = None
    if world == 1 and not args.unfused:
        h2 = evk.Evk(n, device=local_rank)
        hs = [h, h2]

        def e2e_pipe(reps):
            for i in range(reps + 1):
                if i < reps:
                    a = hs[i & 1]
                    a.load_events_ptr(host.data_ptr(), n)         # asynchronous (pinned source)
                    a.downsample_kmeans_submit(ds, km, True)      # asynchronous
                if i > 0:
                    b = hs[(i - 1) & 1]
                    b.downsample_kmeans_wait()
                    u, c = read_back(b, (i - 1) & 1)
                    assert u == state["U_local"]
            return c

        cp = e2e_pipe(2)
        assert (cp == cent).all(), "pipelined e2e changed the result"
        torch.cuda.synchronize()
        reps = max(4, 2 * e2e_steps)
        t0 = time.perf_counter()
        e2e_pipe(reps)
        torch.cuda.synchronize()
        e2e_pipe_s = (time.perf_counter() - t0) / reps
        h2.close()
    e2e_s = e2e_pipe_s if e2e_pipe_s is not None else e2e_serial_s
    e2e_val = n * world / e2e_s / 1e6

    # ---- end to end from the sensor's RAW EVT 2.0 words (4 B per event over PCIe) --------------
    e2e_raw = None
    if world == 1 and not args.no_extras and args.config == "c3":
        spec2 = importlib.util.spec_from_file_location(
            "evk_evt2", os.path.join(evk_loader.PKG_DIR, "evt2.py"))
        evt2 = importlib.util.module_from_spec(spec2)
        spec2.loader.exec_module(evt2)
        words_np = evt2.encode_evt2(host_np)   # host-side writer of the recording format
        raw = torch.empty(len(words_np), dtype=torch.int32, pin_memory=True)
        raw.numpy().view(np.uint32)[:] = words_np
        n_words = len(words_np)

        def raw_step():
            assert h.load_evt2_ptr(raw.data_ptr(), n_words) == n
            step()
            return h.get_centroids(K, D)[0]

        cent_raw = raw_step()
        assert (cent_raw == cent).all(), "RAW ingest changed the result"
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            raw_step()
        torch.cuda.synchronize()
        raw_s = (time.perf_counter() - t0) / e2e_steps
        e2e_raw = {"value": n / raw_s / 1e6, "unit": "Mevents/s", "ms_per_step": raw_s * 1e3,
                   "h2d_bytes_per_step": 4 * n_words, "d2h_bytes_per_step": K * D * 4 + K * 8 + 64,
                   "input": "RAW EVT 2.0 words (evk_load_evt2), decoded on the device; centroids "
                            "+ counts read back"}

    if rank == 0:
        peak, peak_src = peaks()
        U = state["U_local"]
        ds_ms, km_ms = ds_main / args.steps, km_total / args.steps
        ds_bytes, km_bytes = 16.0 * n + 16.0 * U, 20.0 * U
        names = {evk.ALGO_SLAB: "k_slab_main", evk.ALGO_TABLE: "k_table_insert",
                 evk.ALGO_SORT: "sort+unique", evk.ALGO_PARTITION: "k_part_scatter + k_slab_main"}
        traffic = ncu_traffic()
        fused = not args.unfused and algo_used == evk.ALGO_SLAB
        if ds_ms >= km_ms:
            kern, a_bytes, a_ms = names.get(algo_used, "?"), ds_bytes, ds_ms
        else:
            kern, a_bytes, a_ms = "k_km_assign_tiles", km_bytes, km_ms
        achieved = a_bytes / (a_ms * 1e-3) / 1e9 if a_ms > 0 else 0.0
        step_bytes = 16.0 * n + 36.0 * U
        tr = traffic.get(kern) if isinstance(traffic.get(kern), dict) else {"bytes": traffic.get(kern)}
        line = {
            "metric": "Mevents/s downsample+k-means iteration", "value": value,
            "unit": "Mevents/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "u64 keys / f32 distances / exact u64 sums",
            "data": "synthetic", "config": workload_config(args, world),
            "unique_voxels_per_gpu": U, "repeated": state.get("R"),
            "ds_algo": {1: "table", 2: "sort", 3: "slab", 4: "partition"}.get(algo_used, str(algo_used)),
            "owner": args.owner if world > 1 else None,
            "roofline": {"bound": "hbm", "kernel": kern, "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                         "frac_of_8000_nominal": achieved / 8000.0,
                         "algorithmic_bytes_per_launch": a_bytes, "kernel_ms": a_ms,
                         "traffic": tr.get("bytes"), "traffic_source": tr.get("source")},
            "stage_ms": {"downsample_dominant_kernel": ds_ms, "downsample_total": ds_total / args.steps,
                         "kmeans_iteration": km_ms, "fused": fused,
                         "kmeans_byte_model": "20 B/voxel algorithmic (16-B record + label); the "
                                              "kernel itself reads 4 B (xy) + writes 4 B per voxel, "
                                              "so its 20U/t figure may exceed the HBM peak",
                         "measured": "CUDA events around every stage, second pass of the same "
                                     "K steps run synchronously (one host sync per step)",
                         "ms_per_step_synchronous": sync_ms_per_step},
            "submission": ("pipelined: K x evk_downsample_kmeans%s_submit + one "
                           "evk_downsample_kmeans%s_wait" % (("_sharded",) * 2 if world > 1 else ("", ""))
                           if pipelined
                           else "synchronous: one host sync per step"),
            "step_roofline": {"algorithmic_bytes": step_bytes,
                              "achieved": step_bytes / (ms_per_step * 1e-3) / 1e9,
                              "frac": step_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                              "frac_of_8000_nominal": step_bytes / (ms_per_step * 1e-3) / 1e9 / 8000.0,
                              "note": "16N+36U bytes over the whole step (ms_per_step)"},
            "e2e": {"value": e2e_val, "unit": "Mevents/s", "h2d_bytes_per_step": 16 * n,
                    "d2h_bytes_per_step": d2h_full(U), "steps": e2e_steps,
                    "ms_per_step": e2e_s * 1e3,
                    "result_read": "unique voxel keys (canonical order) + labels + centroids + counts",
                    "mode": ("two handles alternating (read-back of slice i overlaps upload of "
                             "slice i+1)" if e2e_pipe_s is not None else "one handle, serial")},
            "e2e_serial": {"value": n * world / e2e_serial_s / 1e6, "unit": "Mevents/s",
                           "ms_per_step": e2e_serial_s * 1e3, "h2d_bytes_per_step": 16 * n,
                           "d2h_bytes_per_step": d2h_full(U),
                           "result_read": "keys + labels + centroids + counts, one handle"},
            "e2e_centroids": {"value": n * world / e2e_cent_s / 1e6, "unit": "Mevents/s",
                              "ms_per_step": e2e_cent_s * 1e3, "h2d_bytes_per_step": 16 * n,
                              "d2h_bytes_per_step": K * D * 4 + K * 8 + 64,
                              "result_read": "centroids + counts (+ voxel counters)"},
            "h2d_only": {"ms_per_step": h2d_s * 1e3, "GBps_per_gpu": 16 * n / h2d_s / 1e9,
                         "GBps_all_gpus": 16 * n * world / h2d_s / 1e9,
                         "note": "bare evk_load_events from pinned host memory on every rank at once: "
                                 "the host-side limit of any e2e number", "numa": numa},
            "gpu_launches": launches, "clocks": clocks,
        }
        if e2e_raw:
            line["e2e_raw_evt2"] = e2e_raw
    h.close()
    if rank == 0 and world == 1 and not args.no_extras and args.config == "c3":
        line["extra_keys"] = {"c5": c5_leg(evk, torch, local_rank, args.c5_windows),
                              "c1": c1_leg(evk, local_rank)}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            n_sample = min(args.cpu_sample, n)
            mev, threads, Us, best = cpu_port(n_sample, 2, host_threads())
            line["cpu_baseline"] = {
                "value": mev, "unit": "Mevents/s", "cores": threads, "kind": "port",
                "sample": f"first {n_sample} events of the workload stream (U={Us}), best of 2 "
                          f"passes ({best:.2f} s each), generation excluded"}
            if not args.no_extras:
                line["consumer"] = consumer_leg(evk)
        print(json.dumps(line), file=real_stdout, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
