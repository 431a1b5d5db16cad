#!/usr/bin/env python
"""GPU-only: run_trajectories on resident programs (reference-shaped stochastic CX circuits, 1024 shots) at n = 10, 11, 12
with the register-only kernel enabled from 10 index bits (direct_min_bits 10) and from 13 (tile kernel below)."""
import json
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import qcdenoise_b200 as qcd  # noqa: E402
from qcdenoise_b200 import batch  # noqa: E402

eng = qcd.Engine(0)
for n, B in ((10, 2000), (11, 2000), (12, 2000)):
    random.seed(1)
    gs = qcd.GraphState(qcd.GraphDB(), n_qubits=n)
    edges, n_edges = gs.sample_batch(B, seed=3)
    rng = np.random.default_rng(2)
    builder = qcd.CXGateCircuit(n_qubits=n, stochastic=True)
    coins = rng.integers(0, 2, size=(int(n_edges.sum()), 1))
    gammas = rng.uniform(0.0, 0.1, size=(int(coins.sum()), 2))
    enc = batch.build_gate_lists(builder, edges, n_edges, coins=coins, gammas=gammas)
    eng.upload(enc)
    out = {"n": n, "circuits": B}
    ref = None
    for mb in (13, 10):
        eng.set_option("direct_min_bits", mb)
        h, _ = eng.run_trajectories(1024, seed=7, precision=32)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            h, _ = eng.run_trajectories(1024, seed=7, precision=32)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        st = eng.stats()
        if ref is None:
            ref = h.clone()
        out[f"min_bits_{mb}"] = {"ms": round(ms, 2), "circuits_per_s": round(B / ms * 1e3), "node_sweeps": st["node_sweeps"],
                                 "same_histograms": bool(torch.equal(h, ref))}
    print(json.dumps(out), flush=True)
