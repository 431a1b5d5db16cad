"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shot-space partition + one collective reproduces the
single-process result exactly.  The per-rank simulation is stood in for by the oracle (the CUDA engine cannot run
here); on the GPU box the same functions run over NCCL (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sim
from tests import circuits

N, SHOTS, SEED = 4, 101, 7


def _batch():
    rng = np.random.default_rng(3)
    return [circuits.graph_circuit_amp_damp(N, circuits.ref_graph(N, i), rng, max_prob=0.3, all_sites=True)
            for i in range(3)]


def _oracle_slice(batch, begin, end):
    h = np.zeros((len(batch), 2 ** N), dtype=np.int64)
    for c, ops in enumerate(batch):
        lo, hi = max(begin, c * SHOTS), min(end, (c + 1) * SHOTS)
        if lo < hi:
            out = sim.run_trajectories(ops, N, hi - lo, SEED, circuit=c, traj_begin=lo - c * SHOTS)
            h[c] = np.bincount(out.astype(np.int64), minlength=2 ** N)
    return torch.from_numpy(h)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from qcdenoise_b200 import distributed as qd
    batch = _batch()
    hist = qd.sharded_histograms(lambda b, e: _oracle_slice(batch, b, e), len(batch), SHOTS)
    cb, ce = qd.circuit_slice(len(batch))
    rows = qd.gather_rows(torch.arange(cb, ce, dtype=torch.float64).reshape(-1, 1) * torch.ones(1, 5, dtype=torch.float64))
    # self-scheduled partition: units pulled from one shared counter, results assembled by ONE all_reduce
    import time
    wq = qd.WorkQueue(7, order=[6, 5, 4, 3, 2, 1, 0])
    passes = []
    for _ in range(2):
        full = torch.zeros((7, 2 ** N), dtype=torch.int64)
        mine = []
        for unit in wq.epoch():
            time.sleep(0.02 * (1 + rank))           # unequal speeds: the faster rank wins more units
            full[unit] = _oracle_slice(batch[:1], unit * 10, unit * 10 + 10)[0]
            mine.append(unit)
        own = torch.zeros(7, dtype=torch.int64)
        own[mine] = 1
        passes.append((qd.assemble_rows(full).numpy(), qd.assemble_rows(own).numpy(), mine))
        dist.barrier()
    if rank == 0:
        q.put((hist.numpy(), rows.numpy(), qd.shot_slice(len(batch), SHOTS, 0, world), qd.shot_slice(len(batch), SHOTS, 1, world),
               passes))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(300)
def test_two_rank_partition_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    hist, rows, s0, s1, passes = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    batch = _batch()
    want = _oracle_slice(batch, 0, len(batch) * SHOTS).numpy()
    assert np.array_equal(hist, want)                      # the slice boundary (151) falls inside circuit 1
    assert s0 == (0, 151) and s1 == (151, 303)
    assert np.array_equal(rows[:, 0], np.arange(3.0)) and rows.shape == (3, 5)
    # work queue: every unit ran on exactly one rank, in both passes; the assembled rows equal the single-process ones
    want_units = np.stack([_oracle_slice(batch[:1], u * 10, u * 10 + 10)[0].numpy() for u in range(7)])
    for full, owners, mine in passes:
        assert np.array_equal(owners, np.ones(7, dtype=np.int64)) and np.array_equal(full, want_units)
        assert 0 < len(mine) < 7 and mine == sorted(mine, reverse=True)         # heaviest-first order respected


def test_work_queue_without_a_process_group_is_a_local_counter():
    from qcdenoise_b200 import distributed as qd
    wq = qd.WorkQueue(5, order=[4, 0, 3, 1, 2])
    assert list(wq.epoch()) == [4, 0, 3, 1, 2] and list(wq.epoch()) == [4, 0, 3, 1, 2]
    assert list(qd.WorkQueue(0).epoch()) == []
    with pytest.raises(ValueError):
        qd.WorkQueue(3, order=[0, 0, 1])
    t = torch.arange(6).reshape(2, 3)
    assert qd.assemble_rows(t) is t


def test_slices_cover_everything():
    from qcdenoise_b200 import distributed as qd
    for world in (1, 2, 3, 8):
        edges = [qd.shot_slice(22, 1_000_000, r, world) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == 22_000_000
        assert all(a[1] == b[0] for a, b in zip(edges[:-1], edges[1:]))
        ce = [qd.circuit_slice(10, r, world) for r in range(world)]
        assert ce[0][0] == 0 and ce[-1][1] == 10 and all(a[1] == b[0] for a, b in zip(ce[:-1], ce[1:]))
