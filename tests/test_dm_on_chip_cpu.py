"""Host side of the on-chip density-matrix interpreter (csrc/dm_cluster.cu): the pass scheduler.  A schedule is valid when
every non-identity op runs exactly once, every op of a pass acts inside the pass's qubit pair, and two ops that share a
qubit keep their order (ops on disjoint qubits commute) -- checked structurally and, through the numpy oracle, numerically:
the density matrix of the re-ordered gate list equals the one of the list as written.  No GPU."""
import numpy as np
import pytest

import qcdenoise_b200 as qb
from oracle import sim
from qcdenoise_b200.engine import debug_dm_schedule
from tests import circuits


def _cases():
    out = []
    for seed in range(6):
        rng = np.random.default_rng(seed)
        n = 2 + seed % 4
        out.append((n, circuits.random_circuit(n, 60, rng)))
    rng = np.random.default_rng(11)
    out.append((5, circuits.device_noise_circuit(5, circuits.ref_graph(5, 0), rng)[0]))
    out.append((8, circuits.device_noise_circuit(8, circuits.ref_graph(8, 3), rng)[0]))
    out.append((4, circuits.graph_circuit_amp_damp(4, circuits.ref_graph(4, 1), rng, max_prob=0.3, all_sites=True)))
    out.append((1, [("h", (0,)), ("t", (0,)), ("id", (0,)), ("rx", (0,), 0.3)]))
    out.append((3, [("h", (0,)), ("h", (1,)), ("h", (2,))] * 20))                 # no two-qubit gate at all, long runs
    return out


@pytest.mark.parametrize("case", range(len(_cases())))
def test_schedule_is_a_valid_reordering(case):
    n, ops = _cases()[case]
    sched = debug_dm_schedule(qb.Circuit(n, ops))
    order = [i for _, idx in sched for i in idx]
    live = [i for i, op in enumerate(ops) if op[0] != "id"]
    assert sorted(order) == live                                                   # every op once, identities dropped
    pos = {i: k for k, i in enumerate(order)}
    for (q0, q1), idx in sched:
        allowed = {q0} | ({q1} if q1 is not None else set())
        assert 0 < len(idx) <= 24
        for i in idx:
            assert set(ops[i][1]) <= allowed, (ops[i], allowed)
    for a in live:
        for b in live:
            if a < b and set(ops[a][1]) & set(ops[b][1]):
                assert pos[a] < pos[b], (ops[a], ops[b])
    if n <= 5:
        want = sim.density_matrix([ops[i] for i in live], n)
        got = sim.density_matrix([ops[i] for i in order], n)
        np.testing.assert_allclose(got, want, atol=1e-12)


def test_noisy_graph_circuit_needs_one_pass_per_entangling_gate():
    """The reference's device-noise circuit (H layer, CX per edge, H, every gate followed by its depolarizing mixture and
    relaxation channels: ~130 ops at n = 8) collapses to about one pass per CX."""
    rng = np.random.default_rng(2)
    edges = circuits.ref_graph(8, 5)
    ops = circuits.device_noise_circuit(8, edges, rng)[0]
    sched = debug_dm_schedule(qb.Circuit(8, ops))
    assert len(sched) <= len(edges) + 4 and len(ops) > 100
    assert all(q1 is not None for (_, q1), _ in sched)


def test_schedule_rejects_malformed_programs():
    from qcdenoise_b200 import _lib
    bad = qb.Circuit(3, [("h", (0,)), ("cx", (0, 1))])
    bad.ops.append(("cx", (1, 1)))
    with pytest.raises((ValueError, _lib.QcdError)):
        debug_dm_schedule(bad)
