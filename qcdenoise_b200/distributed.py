"""Multi-GPU plumbing: one process per GPU, circuits / shots partitioned by index, ONE collective at the end.

The path shards naturally: the unit of work is (circuit, shot); every random draw is a pure function
of (seed, circuit, shot, stream), so the union of the per-rank results does not depend on the number of ranks.  This
mirrors the reference's only parallelism, `mp.Pool(n_procs).map(...)` followed by `np.concatenate`
(older_version/circuit_sampling_utils.py:81-89), with torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU
tests) instead of fork + pickle.  No state is ever sharded and there is no per-gate communication.
"""
import torch
import torch.distributed as dist

__all__ = ["shot_slice", "circuit_slice", "sharded_histograms", "gather_rows", "WorkQueue", "assemble_rows"]


def shot_slice(n_circuits, shots_per_circuit, rank=None, world=None):
    """Contiguous slice [begin, end) of the flattened [circuit][shot] index space owned by `rank`."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    total = n_circuits * shots_per_circuit
    return (total * rank) // world, (total * (rank + 1)) // world


def circuit_slice(n_circuits, rank=None, world=None):
    """Contiguous slice of whole circuits owned by `rank` (density-matrix mode, dataset generation)."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    return (n_circuits * rank) // world, (n_circuits * (rank + 1)) // world


def sharded_histograms(run_slice, n_circuits, shots_per_circuit, group=None):
    """Every rank simulates its slice of the shot space and the partial histograms are summed over ranks.

    run_slice(begin, end) -> integer tensor [n_circuits, 2^n] holding the counts of the global shots [begin, end)
    (e.g. `lambda b, e: engine.run_trajectories(shots, seed=s, shot_range=(b, e))[0]`).  Returns the complete
    histograms on every rank.  A slice boundary may fall inside a circuit: counts are additive, so the assembly is one
    all_reduce(SUM) (= the all-gather + local sum of SURVEY.md 8e); integer arithmetic, hence exact."""
    if not (dist.is_available() and dist.is_initialized()):
        return run_slice(0, n_circuits * shots_per_circuit)
    b, e = shot_slice(n_circuits, shots_per_circuit)
    hist = run_slice(b, e)
    dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return hist


def gather_rows(rows, group=None):
    """all_gather of per-rank row blocks [count_r, width] (whole circuits per rank: probability vectors, dataset
    rows, parity sums) into one [sum count_r, width] tensor in rank order.  Ranks may own different counts."""
    if not (dist.is_available() and dist.is_initialized()):
        return rows
    world = dist.get_world_size(group)
    counts = [torch.zeros(1, dtype=torch.int64, device=rows.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device), group=group)
    counts = [int(c.item()) for c in counts]
    width = rows.shape[1]
    top = max(counts)
    padded = torch.zeros((top, width), dtype=rows.dtype, device=rows.device)
    padded[: rows.shape[0]] = rows
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


class WorkQueue:
    """Self-scheduled partition of independent work units (graphs, blocks of circuits) over the ranks.

    A static slice per rank balances COUNTS, not COST: the trajectory tree of a graph circuit with 17 noise sites has
    ~10x the states of one with 7, so equal counts leave most GPUs idle behind the unlucky one.  Here every rank pulls
    the index of its next unit from one shared counter (an atomic `add` on the process group's key-value store -- host
    side, no GPU collective, no data on the wire) whenever it becomes free; `order` lists the units heaviest first so
    that the cheap ones fill the gaps at the end.  Because every draw is keyed by the GLOBAL circuit index, the results
    do not depend on which rank ran which unit.  Without an initialised process group it is a plain local counter.

        q = WorkQueue(n_graphs, order=np.argsort(-cost))
        for unit in q.epoch():          # every rank calls epoch() once per pass, the same number of times
            ...
    """

    def __init__(self, n_units, order=None, name="qcd_wq"):
        self.n_units = int(n_units)
        self.order = list(range(self.n_units)) if order is None else [int(u) for u in order]
        if sorted(self.order) != list(range(self.n_units)):
            raise ValueError("order must be a permutation of range(n_units)")
        self.name = name
        self._epoch = 0
        self._store = None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            from torch.distributed import distributed_c10d
            self._store = distributed_c10d._get_default_store()

    def epoch(self):
        """Iterator over the units this rank wins in one pass over the queue."""
        self._epoch += 1
        key = f"{self.name}/{self._epoch}"
        local = iter(range(self.n_units))
        while True:
            k = (self._store.add(key, 1) - 1) if self._store is not None else next(local, self.n_units)
            if k >= self.n_units:
                return
            yield self.order[k]


def assemble_rows(full, group=None):
    """Assembly after a self-scheduled pass: `full` is an integer tensor [n_units, width] in which this rank filled the
    rows of the units it ran and left the others ZERO; one all_reduce(SUM) completes it on every rank (exactly one rank
    contributes to a row, so the sum is exact for any integer dtype)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)
    return full
