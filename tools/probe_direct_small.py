#!/usr/bin/env python
"""Register-only sweep kernel on states below 13 index bits (engine option direct_min_bits): complex128 outcomes against the
compiled oracle shot for shot at n = 10, 11, 12, and sweep statistics."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

import qcdenoise_b200 as qb  # noqa: E402
from oracle import cpu_build  # noqa: E402
from tests import circuits  # noqa: E402

eng = qb.Engine(0)
for n in (10, 11, 12):
    rng = np.random.default_rng(n)
    edges = circuits.ref_graph(10, 1) + [(9, 10), (10, 11)][: n - 10]
    batch = [circuits.graph_circuit_amp_damp(n, edges, rng, max_prob=0.25, all_sites=(i == 0)) for i in range(3)]
    eng.upload([qb.Circuit(n, ops) for ops in batch])
    shots, seed = 700, 5
    for mb in (13, 10):
        eng.set_option("direct_min_bits", mb)
        res = {}
        for prec in (64, 32):
            h, o = eng.run_trajectories(shots, seed=seed, precision=prec, want_outcomes=True)
            res[prec] = o.cpu().numpy().reshape(len(batch), shots).astype(np.int64)
        st = eng.stats()
        ok = True
        for c, ops in enumerate(batch):
            want = cpu_build.run_trajectories(qb.encode_programs([qb.Circuit(n, ops)]), shots, seed, circuit=c).astype(np.int64)
            ok = ok and np.array_equal(res[64][c], want)
        print(f"n={n} direct_min_bits={mb}: complex128 == oracle: {ok}; complex64 differs in {(res[32] != res[64]).sum()} shots; "
              f"sweep launches {st['sweep_launches']}, node sweeps {st['node_sweeps']}", flush=True)
