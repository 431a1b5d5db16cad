"""Sibling groups of the leaf level (option "leaf_cross", csrc/sweep_direct.cu KIND 4 + the compose step of scan_l1_kernel).

The n + 2 circuits of a graph -- the dataset circuit and its n + 1 Toth settings (stabilizers.py:115-121, :136-156, :194-199) --
end in the same final sweep up to ONE Hadamard on the setting's qubit k.  Unstored leaves with the same parent and branch word
therefore share passes over the parent that reduce C0 = sum |amp|^2 and X_k = sum Re(conj(amp[x]) amp[x ^ 2^k]) per
32-amplitude chunk; member k's chunk sums are (C0[c] + C0[c ^ k]) / 2 +- X_k[c].  Checked here against one sweep per leaf
(option off: identical complex128 outcomes, complex64 within rounding of a CDF boundary) and against the oracle shot for shot."""
import numpy as np
import pytest

from oracle import cpu_build, sim
from tests import circuits

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import qcdenoise_b200 as qb
    e = qb.Engine(0)
    yield e
    for k, v in (("leaf_cross", 1), ("nostore_max", 1024), ("max_slots", 0), ("max_chunk", 1 << 25), ("table_cap_log2", 22)):
        e.set_option(k, v)
    e.close()


def _circ(n, ops):
    import qcdenoise_b200 as qb
    return qb.Circuit(n, ops)


def _graph_and_settings(n, edges, rng, extra=()):
    """[dataset circuit, X_v Z_nbrs settings for every vertex v (basis change = h on v), identity setting] + extras."""
    ops = circuits.graph_circuit_amp_damp(n, edges, rng, max_prob=0.3, all_sites=True)
    batch = [ops]
    for v in range(n):
        batch.append(ops + [("h", (v,))])
    batch.append(list(ops))
    for sub in extra:
        batch.append(ops + list(sub))
    return batch


def _edges(n, gi):
    e = circuits.ref_graph(n if n <= 10 else 14, gi)
    if n not in (10, 14):
        e = [x for x in e if max(x) < n] + [(n - 2, n - 1)]
    return e


@pytest.mark.parametrize("n,gi", [(10, 0), (10, 1), (12, 2), (13, 0)])
def test_sibling_groups_equal_one_sweep_per_leaf(eng, n, gi):
    from qcdenoise_b200.ir import encode_programs
    rng = np.random.default_rng(900 + 10 * n + gi)
    # extras: a Y-type basis change (complex 2x2: not eligible, keeps its own sweep), a second member on qubit n - 1
    # (duplicate bit: keeps its own sweep), a rotation on a high qubit (real 2x2 that is not a Hadamard)
    batch = _graph_and_settings(n, _edges(n, gi), rng,
                                extra=([("sdg", (n - 1,)), ("h", (n - 1,))], [("h", (n - 1,))], [("ry", (n - 2,), 0.7)]))
    B = len(batch)
    eng.upload([_circ(n, ops) for ops in batch])
    eng.set_option("nostore_max", 1 << 20)          # every leaf unstored: all of them are candidates
    shots, seed = 300, 5
    res = {}
    for prec in (64, 32):
        for on in (1, 0):
            eng.set_option("leaf_cross", on)
            h, o = eng.run_trajectories(shots, seed=seed, precision=prec, want_outcomes=True)
            res[prec, on] = (h.cpu().numpy(), o.cpu().numpy().reshape(B, shots).astype(np.int64), eng.stats())
        st_on, st_off = res[prec, 1][2], res[prec, 0][2]
        assert st_off["cross_leaves"] == 0 and st_off["cross_passes"] == 0
        assert st_on["cross_leaves"] > 0.25 * st_on["leaves"], (st_on["cross_leaves"], st_on["leaves"])
        assert st_on["cross_passes"] < 0.7 * st_on["cross_leaves"]
        assert st_on["leaves"] == st_off["leaves"] and st_on["node_sweeps"] < st_off["node_sweeps"]
    assert np.array_equal(res[64, 1][1], res[64, 0][1]), np.count_nonzero(res[64, 1][1] != res[64, 0][1])
    assert np.array_equal(res[64, 1][0], res[64, 0][0])
    assert np.count_nonzero(res[32, 1][1] != res[32, 0][1]) <= max(2, B * shots // 1000)
    # ... and against the oracle, shot for shot
    run = cpu_build.run_trajectories if n >= 12 else None
    for c in (0, 1, 3, n - 1, n, n + 1, B - 3, B - 2, B - 1):
        if run is not None:
            want = run(encode_programs([_circ(n, batch[c])]), shots, seed, circuit=c).astype(np.int64)
        else:
            want = sim.run_trajectories(batch[c], n, shots, seed, circuit=c).astype(np.int64)
        assert np.array_equal(res[64, 1][1][c], want), f"circuit {c}"
        assert np.count_nonzero(res[32, 1][1][c] != want) <= 3


def test_sibling_groups_with_stored_leaves_and_cut_launches(eng):
    """Default nostore_max (big leaves are stored and keep their own sweep), few state slots (leaf launches cut into small
    chunks: groups split across launches), small grouping table, trees cut at circuit boundaries: same outcomes."""
    n = 12
    rng = np.random.default_rng(77)
    batch = _graph_and_settings(n, _edges(n, 1), rng)
    B = len(batch)
    eng.upload([_circ(n, ops) for ops in batch])
    shots, seed = 3000, 11
    eng.set_option("nostore_max", 1024)
    eng.set_option("leaf_cross", 0)
    h0, o0 = eng.run_trajectories(shots, seed=seed, precision=64, want_outcomes=True)
    h0, o0 = h0.cpu().numpy(), o0.cpu().numpy()
    st0 = eng.stats()
    eng.set_option("leaf_cross", 1)
    for opts in ({}, {"max_slots": 30}, {"table_cap_log2": 6}, {"max_chunk": 4 * shots + 100}):
        for k, v in opts.items():
            eng.set_option(k, v)
        h, o = eng.run_trajectories(shots, seed=seed, precision=64, want_outcomes=True)
        st = eng.stats()
        assert np.array_equal(o.cpu().numpy(), o0) and np.array_equal(h.cpu().numpy(), h0), opts
        assert st["cross_leaves"] > 0 and st["leaves"] == st0["leaves"], opts
        for k, v in (("max_slots", 0), ("table_cap_log2", 22), ("max_chunk", 1 << 25)):
            eng.set_option(k, v)


def test_sibling_group_expectation_values(eng):
    """Sampled <S_k> of every Toth setting with sibling groups on: within 5 sigma of the exact value (oracle/pauli_prop.py),
    complex64, 100 000 shots per setting."""
    from oracle import pauli_prop
    n = 12
    rng = np.random.default_rng(5)
    edges = _edges(n, 0)
    batch = _graph_and_settings(n, edges, rng)
    eng.upload([_circ(n, ops) for ops in batch])
    eng.set_option("leaf_cross", 1)
    eng.set_option("nostore_max", 1024)
    shots = 100_000
    hist, _ = eng.run_trajectories(shots, seed=3, precision=32)
    assert eng.stats()["cross_leaves"] > 0
    nbrs = {v: [b if a == v else a for a, b in edges if v in (a, b)] for v in range(n)}
    masks = np.zeros((len(batch), 1), dtype=np.uint32)
    for v in range(n):
        masks[1 + v, 0] = (1 << v) | sum(1 << q for q in nbrs[v])
    ss, tot = eng.parity_counts(hist, masks)
    mean = ss.cpu().numpy()[:, 0] / tot.cpu().numpy()
    for v in range(n):
        exact = pauli_prop.expectation(batch[1 + v], n, int(masks[1 + v, 0]))      # <prod Z> after the setting's basis change
        sig = np.sqrt(max(1 - mean[1 + v] ** 2, 0) / shots)
        assert abs(mean[1 + v] - exact) <= 5 * sig + 1e-12, (v, mean[1 + v], exact, sig)
