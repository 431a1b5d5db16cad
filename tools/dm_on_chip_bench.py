#!/usr/bin/env python
"""GPU-only timing of the density-matrix routes on C2-shaped batches (device noise, CX graph circuits): the on-chip
interpreter (dm_cluster.cu) against the lowered sweep route, programs resident, CUDA events around run_density_matrix.
    python tools/dm_on_chip_bench.py [n_qubits ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import qcdenoise_b200 as qcd  # noqa: E402
from qcdenoise_b200 import batch  # noqa: E402


def main():
    ns = [int(a) for a in sys.argv[1:]] or [6, 7, 8, 9]
    eng = qcd.Engine(0)
    spec = qcd.NoiseSpec(specs={"qubit": {"T1": 100.0, "T2": 100.0, "readout_error": 0.25, "prob_meas0_prep1": 0.1,
                                          "prob_meas1_prep0": 0.1},
                                "cx": {"gate_error": 0.25, "gate_length": 500.0}, "h": {"gate_error": 0.05, "gate_length": 50.0}})
    for n in ns:
        B = {9: 1200}.get(n, 4000)
        gs = qcd.GraphState(qcd.GraphDB(), n_qubits=n)
        edges, n_edges = gs.sample_batch(B, seed=3)
        enc = batch.build_gate_lists(qcd.CXGateCircuit(n_qubits=n, stochastic=False), edges, n_edges)
        props = batch.draw_device_props(spec, enc, np.random.default_rng(1))
        t1t2, gp, _ = batch.device_noise_inputs(props)
        out = {"n_qubits": n, "circuits": B, "cx_per_circuit": float(n_edges.mean())}
        for on_chip in (2, 0):
            if n == 9 and not on_chip:
                continue
            eng.set_option("dm_on_chip", on_chip)
            eng.upload_device_noise(enc, t1t2, gp)
            p = eng.run_density_matrix(32)          # first run: host preparation / lowering
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                p = eng.run_density_matrix(32, out=p)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            out["on_chip" if on_chip else "lowered"] = {"ms": ms, "us_per_circuit": 1e3 * ms / B, "circuits_per_s": B / ms * 1e3}
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
