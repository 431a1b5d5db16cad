#!/usr/bin/env python
"""BASELINE.json configs[4] ("C5"): 28-qubit statevector noisy trajectories, circuit batch sharded across the GPUs of
one box.  64 reference-shaped circuits x 8 shots; rank r simulates a contiguous slice of whole circuits (no state is
sharded, no data-path communication), then ONE all_gather assembles the sparse outcome lists and the 29 Toth
generator parity sums per circuit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_c5_multi.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import qcdenoise_b200 as qcd  # noqa: E402
from qcdenoise_b200 import distributed as qd  # noqa: E402
from tools.run_configs import graphs, nxg  # noqa: E402

N, CIRCUITS, SHOTS = 28, 64, 8


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gs = graphs(N)
    np.random.seed(1234)                       # every rank builds the same 64 circuits, then keeps its slice
    sampler = qcd.UnitaryNoiseSampler(None, n_shots=SHOTS, method="statevector")
    progs, masks = [], []
    for i in range(CIRCUITS):
        g = nxg(N, gs[i % len(gs)])
        cd = qcd.CXGateCircuit(n_qubits=N, stochastic=True).build(g)
        sampler.build_noise_model(cd["ops"])
        progs.append(sampler.noise_model.instantiate(cd["circuit"])[0])
        masks.append([qcd.pauli_mask(nm.replace("X", "Z")) for nm in qcd.TothStabilizer(g, N).build()])
    lo, hi = qd.circuit_slice(CIRCUITS, rank, world)
    eng = qcd.get_engine(local)
    eng.set_option("time_sweeps", 1)
    eng.upload(progs[lo:hi])
    my_masks = np.array(masks[lo:hi], dtype=np.uint32)

    def step():
        _, out = eng.run_trajectories(SHOTS, seed=1234, precision=32, want_outcomes=True, first_circuit=lo)
        st = eng.stats()
        out = out.reshape(hi - lo, SHOTS)
        ss = eng.parity_outcomes(out, my_masks)
        rows = torch.cat([out.to(torch.int64), ss], dim=1)
        return st, (qd.gather_rows(rows.contiguous()) if world > 1 else rows)

    step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    st, rows = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    assert rows.shape == (CIRCUITS, SHOTS + N + 1)
    assert int(rows[:, -1].min()) == SHOTS and int(rows[:, -1].max()) == SHOTS     # identity setting
    if rank == 0:
        gbs = st["sweep_bytes"] / (st["sweep_ms"] * 1e-3) / 1e9
        print(json.dumps({"config": "C5: 28-qubit trajectories, 64 circuits x 8 shots, sharded by circuit index",
                          "n_gpus": world, "ms": ms, "shots_per_s": CIRCUITS * SHOTS / (ms * 1e-3),
                          "circuits_per_s": CIRCUITS / (ms * 1e-3), "rank0_sweep_GBps_algorithmic": gbs,
                          "rank0_states_swept": st["node_sweeps"], "rank0_max_live_states": st["max_live_states"]}),
              flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
