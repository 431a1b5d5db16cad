#!/usr/bin/env python
"""Where a C3 batch (16 x 14-qubit density matrices) spends its time: the engine calls one by one, CUDA events + host clock."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import bench  # noqa: E402
import qcdenoise_b200 as qcd  # noqa: E402

eng = qcd.get_engine(0)
progs, _, sites = bench._noisy_programs(qcd, 14, range(16), bench.SEED + 14)
eng.upload(progs)


def timed(name, fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:34s} gpu {e0.elapsed_time(e1) / reps:8.2f} ms   host {1e3 * (time.perf_counter() - t0) / reps:8.2f} ms", flush=True)
    return out


for ts in (0, 1):
    eng.set_option("time_sweeps", ts)
    p = timed(f"run_density_matrix time_sweeps={ts}", lambda: eng.run_density_matrix(32))
    if ts:
        st = eng.stats()
        print("   sweep_ms", round(st["sweep_ms"], 2), "launches", st["sweep_launches"], "kinds", [round(x, 2) for x in st["kind_ms"]])
timed("sample_from_probs", lambda: eng.sample_from_probs(p, 1024, seed=1))
timed("stats()", lambda: eng.stats())
