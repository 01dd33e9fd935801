"""Measures the BASELINE.json configs other than the bench workload on one B200 (not bench lines:
parity of these shapes is in tests/).  Prints one JSON object per config.

    python profiles/extra_configs.py > profiles/r01/extra_configs.jsonl
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import evk_loader  # noqa: E402

evk = evk_loader.load()


def timed(h, fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    h.timer_start()
    for _ in range(reps):
        fn()
    return h.timer_stop() / reps


def batch(name, seed, n, rate, W, H, blobs, vox, K, iters, D=2):
    h = evk.Evk(n)
    h.synth(evk.synth_params(seed, n, W, H, rate, blobs))
    ds = evk.ds_params(W, H, vox[0], vox[1], vox[2], 0, vox[3])
    km = evk.km_params(K, D, iters=iters, t_scale=1e-3)
    res = {}

    def step():
        res["u"], res["r"], res["it"] = h.downsample_kmeans(ds, km, True)

    ms = timed(h, step)
    h.set_profiling(True)
    step()
    t = h.stage_times()
    U = res["u"]
    out = {"config": name, "events": n, "unique": U, "repeated": res["r"], "K": K, "D": D,
           "iters": res["it"], "ms_total": ms, "Mevents_per_s": n / ms / 1e3,
           "ms_downsample": t.ds_total_ms, "ms_kmeans_all_iters": t.km_total_ms,
           "ms_per_kmeans_iter": t.km_total_ms / max(1, res["it"]),
           "algorithmic_GBps": (16 * n + 16 * U + iters * 20 * U) / ms / 1e6}
    print(json.dumps(out), flush=True)
    h.close()


def windows(n_win=40, per_win=5_000_000):
    W, H, K = 1280, 720, 64
    n = n_win * per_win
    g = evk.Evk(n)
    g.synth(evk.synth_params(0xE7CA0005, n, W, H, 100_000_000, 64))
    import torch
    host = torch.empty(n * 16, dtype=torch.uint8, pin_memory=True)
    ev = host.numpy().view(evk.EVENT_DTYPE)
    ev[:] = g.get_events()
    g.close()
    h = evk.Evk(per_win + 1024)
    h.window_config(evk.ds_params(W, H, 2, 2, 500, 0, 1), evk.km_params(K, 2, iters=2), 50_000)
    lat = []
    # one push per 50 ms of event time (the SDK callback granularity is finer; this is the
    # worst case for latency: the whole window arrives at once, H2D included)
    bounds = np.searchsorted(ev["t"], np.arange(0, n_win + 1) * 50_000)
    for w in range(n_win):
        a, b = bounds[w], bounds[w + 1]
        t0 = time.perf_counter()
        # the first event of the next window closes this one
        d = h.window_push(ev[a:min(b + 1, n)])
        if w == n_win - 1:
            d += h.window_flush()
        lat.append((time.perf_counter() - t0) * 1e3)
    lat = np.array(lat[2:])
    print(json.dumps({"config": "C5 streaming 50 ms windows at 100 Mev/s, K=64, 2 warm-started "
                                "iterations per window, events pushed from pinned host memory",
                      "windows": int(len(lat)), "events_per_window": per_win,
                      "latency_ms_p50": float(np.percentile(lat, 50)),
                      "latency_ms_p99": float(np.percentile(lat, 99)),
                      "latency_ms_max": float(lat.max()),
                      "real_time_factor": 50.0 / float(np.percentile(lat, 50))}), flush=True)
    h.close()


if __name__ == "__main__":
    batch("C1 DAVIS346 1M events, 4x4 px x 1 ms, K=8, 1 iter", 0xE7CA0001, 1_000_000, 10_000_000,
          346, 260, 8, (4, 4, 1000, 1), 8, 1)
    batch("C2 DAVIS346 10M events, 4x4 px x 1 ms, K=32, 20 iters", 0xE7CA0002, 10_000_000,
          10_000_000, 346, 260, 32, (4, 4, 1000, 1), 32, 20)
    batch("C3 Gen4 100M events, 2x2 px x 500 us, K=64, 1 iter", 0xE7CA0003, 100_000_000,
          100_000_000, 1280, 720, 64, (2, 2, 500, 1), 64, 1)
    batch("C3 variant D=3 (x, y, t)", 0xE7CA0003, 100_000_000, 100_000_000, 1280, 720, 64,
          (2, 2, 500, 1), 64, 1, D=3)
    windows()
