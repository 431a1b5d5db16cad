"""GPU parity tests: the CUDA engine, called through the C ABI (ctypes), against the CPU oracle on the same seeded
inputs.  Tolerances (BASELINE.json north_star): exact probabilities 1e-5 in complex64 / 1e-10 in complex128;
integer work (outcomes with identical Philox draws, histograms, parity sums) bit-exact."""
import numpy as np
import pytest

from oracle import ref_glue, sim
from tests import circuits

pytestmark = pytest.mark.gpu

TOL = {32: 1e-5, 64: 1e-10}


@pytest.fixture(scope="module")
def eng():
    import qcdenoise_b200 as qb
    e = qb.Engine(0)
    yield e
    e.close()


def _circ(n, ops):
    import qcdenoise_b200 as qb
    return qb.Circuit(n, ops)


def _reset_opts(eng):
    for k, v in (("tile_bits", 0), ("max_slots", 0), ("table_cap_log2", 22), ("max_chunk", 1 << 25), ("use_direct", 1),
                 ("share_prefix", 1), ("nostore_max", 1024), ("direct_min_bits", 10), ("leaf_cross", 1)):
        eng.set_option(k, v)


# ------------------------------------------------------------------------------------------- density matrix
@pytest.mark.parametrize("precision", [32, 64])
@pytest.mark.parametrize("n,tile_bits", [(4, 0), (5, 9), (6, 0), (7, 10), (8, 0)])
def test_dm_probs_graph_amp_damp(eng, precision, n, tile_bits):
    _reset_opts(eng)
    eng.set_option("tile_bits", tile_bits)
    rng = np.random.default_rng(n)
    src = {6: 5}.get(n, n)
    batch = []
    for i in range(4):
        g = circuits.ref_graph(src, i)
        if n == 6:
            g = g + [(4, 5)]
        batch.append(circuits.graph_circuit_amp_damp(n, g, rng, max_prob=0.3))
    eng.upload([_circ(n, ops) for ops in batch])
    probs = eng.run_density_matrix(precision).cpu().numpy()
    for i, ops in enumerate(batch):
        np.testing.assert_allclose(probs[i], sim.density_matrix_probs(ops, n), atol=TOL[precision], rtol=0)
    _reset_opts(eng)


@pytest.mark.parametrize("precision", [32, 64])
def test_dm_probs_random_and_device(eng, precision):
    _reset_opts(eng)
    for seed in range(4):
        rng = np.random.default_rng(seed)
        n = 5
        ops = circuits.random_circuit(n, 40, rng)
        dev, conf = circuits.device_noise_circuit(n, circuits.ref_graph(n, seed), rng)
        eng.upload([_circ(n, ops), _circ(n, dev)])
        probs = eng.run_density_matrix(precision)
        want = [sim.density_matrix_probs(ops, n), sim.density_matrix_probs(dev, n)]
        got = probs.cpu().numpy()
        for a, b in zip(got, want):
            np.testing.assert_allclose(a, b, atol=TOL[precision], rtol=0)
        # classical readout error on the exact distribution
        eng.apply_readout(probs, np.stack([conf, conf]))
        got = probs.cpu().numpy()
        np.testing.assert_allclose(got[1], sim.apply_readout_to_probs(want[1], n, conf), atol=TOL[precision], rtol=0)


def test_dm_C1_depolarizing_witness(eng):
    """BASELINE config C1: 4-qubit linear graph, depolarizing noise; exact probs + Toth stabilizer expectations."""
    _reset_opts(eng)
    n, edges = 4, [(0, 1), (1, 2), (2, 3)]
    base = circuits.depolarizing_circuit(n, edges)
    nbrs = {v: [b if a == v else a for a, b in edges if v in (a, b)] for v in range(n)}
    gens = ref_glue.get_generators(n, range(n), nbrs)
    settings = gens + ["I" * n]
    circs = [base + ref_glue.stabilizer_basis_change(name) for name in settings]
    eng.upload([_circ(n, ops) for ops in circs])
    masks = np.array([[sum(1 << i for i, ch in enumerate(name[::-1]) if ch != "I")] for name in settings], dtype=np.uint32)
    for precision in (32, 64):
        probs = eng.run_density_matrix(precision)
        ev = eng.parity_probs(probs, masks).cpu().numpy()[:, 0]
        for i, name in enumerate(settings):
            p = sim.density_matrix_probs(circs[i], n)
            want = float(np.dot(p, ref_glue.construct_diagonal(name)))
            assert abs(ev[i] - want) < TOL[precision]


# ------------------------------------------------------------------------------------------- trajectories
def _oracle_outcomes(ops, n, shots, seed, circuit, readout=None):
    return sim.run_trajectories(ops, n, shots, seed, circuit=circuit, readout=readout).astype(np.int64)


@pytest.mark.parametrize("n,tile_bits", [(4, 0), (6, 0), (8, 0), (8, 7), (10, 8)])
def test_trajectories_match_oracle_fp64(eng, n, tile_bits):
    """complex128 + identical Philox draws: outcomes agree shot for shot; histograms bit-exact."""
    _reset_opts(eng)
    eng.set_option("tile_bits", tile_bits)
    rng = np.random.default_rng(40 + n)
    src = {6: 5}.get(n, n)
    batch = []
    for i in range(3):
        g = circuits.ref_graph(src, i)
        if n == 6:
            g = g + [(4, 5)]
        batch.append(circuits.graph_circuit_amp_damp(n, g, rng, max_prob=0.35, all_sites=(i == 0)))
    eng.upload([_circ(n, ops) for ops in batch])
    shots, seed = 500, 77
    hist, out = eng.run_trajectories(shots, seed=seed, precision=64, want_outcomes=True)
    out = out.cpu().numpy().reshape(3, shots)
    hist = hist.cpu().numpy()
    for c, ops in enumerate(batch):
        want = _oracle_outcomes(ops, n, shots, seed, c)
        assert np.array_equal(out[c], want), f"circuit {c}: {np.count_nonzero(out[c] != want)} outcomes differ"
        assert np.array_equal(hist[c], sim.histogram(want, n))
    st = eng.stats()
    assert st["trajectories"] == 3 * shots and st["leaves"] <= 3 * shots
    _reset_opts(eng)


def test_trajectories_all_noise_kinds_fp64(eng):
    _reset_opts(eng)
    n = 5
    rng = np.random.default_rng(8)
    dev, conf = circuits.device_noise_circuit(n, circuits.ref_graph(n, 0), rng, gate_error=0.2, gate_len_ns=(500, 5000))
    dep = circuits.depolarizing_circuit(n, circuits.ref_graph(n, 1), p1=0.05, p2=0.2)
    rnd = circuits.random_circuit(n, 30, rng)
    batch = [dev, dep, rnd]
    eng.upload([_circ(n, ops) for ops in batch])
    shots, seed = 400, 3
    readout = np.stack([conf] * 3)
    _, out = eng.run_trajectories(shots, seed=seed, precision=64, want_outcomes=True, readout=readout)
    out = out.cpu().numpy().reshape(3, shots)
    for c, ops in enumerate(batch):
        want = _oracle_outcomes(ops, n, shots, seed, c, readout=conf)
        assert np.array_equal(out[c], want), f"circuit {c}: {np.count_nonzero(out[c] != want)} outcomes differ"


def test_dedup_equals_literal_per_shot(eng):
    _reset_opts(eng)
    n = 8
    rng = np.random.default_rng(5)
    ops = circuits.graph_circuit_amp_damp(n, circuits.ref_graph(n, 3), rng, max_prob=0.3, all_sites=True)
    eng.upload([_circ(n, ops)] * 2)
    for precision in (32, 64):
        h1, o1 = eng.run_trajectories(300, seed=9, precision=precision, want_outcomes=True)
        s1 = eng.stats()
        h2, o2 = eng.run_trajectories(300, seed=9, precision=precision, want_outcomes=True, dedup=False)
        s2 = eng.stats()
        assert bool((o1 == o2).all()) and bool((h1 == h2).all())
        assert s2["leaves"] == 600 and s1["node_sweeps"] < s2["node_sweeps"]


def test_memory_limited_recursion_and_table_split(eng):
    """Few state slots (depth-first chunks) and a tiny grouping table give the same outcomes."""
    _reset_opts(eng)
    n = 8
    rng = np.random.default_rng(6)
    ops = circuits.graph_circuit_amp_damp(n, circuits.ref_graph(n, 4), rng, max_prob=0.35, all_sites=True)
    eng.upload([_circ(n, ops)])
    _, ref = eng.run_trajectories(400, seed=1, precision=64, want_outcomes=True)
    n_sweeps = 1 + sum(1 for op in ops if op[0] == "unitary")
    eng.set_option("max_slots", n_sweeps + 3)
    eng.set_option("table_cap_log2", 4)
    eng.set_option("max_chunk", 150)
    _, got = eng.run_trajectories(400, seed=1, precision=64, want_outcomes=True)
    assert bool((ref == got).all())
    assert eng.stats()["max_live_states"] <= n_sweeps + 3
    eng.set_option("max_slots", 2)
    from qcdenoise_b200._lib import QcdError
    with pytest.raises(QcdError):
        eng.run_trajectories(10, seed=1, precision=64)
    _reset_opts(eng)


def test_shot_range_partition_is_rank_invariant(eng):
    """The multi-GPU partition: any split of the flattened shot range gives the same union."""
    _reset_opts(eng)
    n = 6
    rng = np.random.default_rng(12)
    batch = [circuits.graph_circuit_amp_damp(n, circuits.ref_graph(5, i) + [(4, 5)], rng, max_prob=0.3) for i in range(3)]
    eng.upload([_circ(n, ops) for ops in batch])
    shots = 200
    whole, _ = eng.run_trajectories(shots, seed=4, precision=32)
    acc = None
    for lo, hi in ((0, 130), (130, 470), (470, 600)):
        acc, _ = eng.run_trajectories(shots, seed=4, precision=32, shot_range=(lo, hi), hist=acc)
    assert bool((whole == acc).all())


def test_trajectories_fp32_statistics(eng):
    """complex64 trajectories: histogram vs exact density-matrix probabilities (chi-square, TVD)."""
    _reset_opts(eng)
    n = 6
    rng = np.random.default_rng(21)
    ops = circuits.graph_circuit_amp_damp(n, circuits.ref_graph(5, 2) + [(2, 5)], rng, max_prob=0.3, all_sites=True)
    eng.upload([_circ(n, ops)])
    shots = 200000
    hist, _ = eng.run_trajectories(shots, seed=31, precision=32)
    h = hist.cpu().numpy()[0].astype(np.float64)
    p = sim.density_matrix_probs(ops, n)
    assert h.sum() == shots
    tvd = 0.5 * np.abs(h / shots - p).sum()
    assert tvd < 0.02
    sel = p * shots > 5
    chi2 = (((h - p * shots) ** 2) / (p * shots))[sel].sum()
    dof = sel.sum() - 1
    assert chi2 < dof + 6 * np.sqrt(2 * dof)


def test_empty_and_edge_inputs(eng):
    _reset_opts(eng)
    from qcdenoise_b200._lib import QcdError
    # empty circuit: |0...0> with certainty
    eng.upload([_circ(3, [])])
    hist, _ = eng.run_trajectories(50, seed=1, precision=32)
    assert hist.cpu().numpy()[0].tolist() == [50] + [0] * 7
    assert eng.run_density_matrix(64).cpu().numpy()[0].tolist() == [1.0] + [0.0] * 7
    # empty shot range is a no-op
    h, _ = eng.run_trajectories(50, seed=1, precision=32, shot_range=(10, 10))
    assert int(h.sum()) == 0
    with pytest.raises(QcdError):
        eng.run_trajectories(50, seed=1, precision=16)
    with pytest.raises(QcdError):
        eng.run_trajectories(50, seed=1, shot_range=(0, 51))
    # 1-qubit: amplitude damping on |1> gives P(0) = gamma
    from oracle import noise
    g = 0.3
    eng.upload([_circ(1, [("x", (0,)), ("kraus", (0,), noise.amplitude_damping_kraus(g))])])
    p = eng.run_density_matrix(64).cpu().numpy()[0]
    np.testing.assert_allclose(p, [g, 1 - g], atol=1e-12)


# ------------------------------------------------------------------------------------------- sampling / parity
@pytest.mark.parametrize("n", [3, 8, 13])
def test_sample_from_probs_matches_oracle(eng, n):
    import torch
    rng = np.random.default_rng(n)
    p = rng.random((3, 2 ** n)) ** 4
    p /= p.sum(axis=1, keepdims=True)
    conf = np.array([[[0.97, 0.03], [0.08, 0.92]]] * n)
    probs = torch.tensor(p, dtype=torch.float64, device="cuda")
    shots, seed = 2000, 17
    hist = eng.sample_from_probs(probs, shots, seed=seed, first_circuit=5, readout=np.stack([conf] * 3)).cpu().numpy()
    for v in range(3):
        want = sim.sample_from_probs(p[v], shots, seed, circuit=5 + v, readout=conf)
        assert np.array_equal(hist[v], sim.histogram(want, n))


@pytest.mark.parametrize("n", [4, 10, 16])
def test_parity_counts_bit_exact(eng, n, goldens):
    import torch
    rng = np.random.default_rng(n)
    h = rng.integers(0, 1000, size=(3, 2 ** n)).astype(np.int32)
    h[rng.random(h.shape) < 0.5] = 0
    masks = rng.integers(0, 2 ** n, size=(3, n + 1)).astype(np.uint32)
    masks[:, -1] = 0
    ss, tot = eng.parity_counts(torch.tensor(h, device="cuda"), masks)
    ss, tot = ss.cpu().numpy(), tot.cpu().numpy()
    x = np.arange(2 ** n, dtype=np.uint64)
    for v in range(3):
        assert tot[v] == h[v].sum()
        for m in range(n + 1):
            par = np.array([bin(int(t)).count("1") & 1 for t in (x & np.uint64(masks[v, m]))]) if n <= 10 else \
                (np.unpackbits((x & np.uint64(masks[v, m])).view(np.uint8)).reshape(-1, 64).sum(axis=1) & 1)
            assert ss[v, m] == int((h[v].astype(np.int64) * (1 - 2 * par.astype(np.int64))).sum())


def test_parity_counts_golden_G1(eng, goldens):
    """Reference notebook counts -> <S> exactly as recorded (graph_witness_refractored.ipynb#c45-47)."""
    import torch
    g = goldens["G1"]
    names = g["stabilizer_order"]
    n = 4
    h = np.zeros((len(names), 16), dtype=np.int32)
    for i, counts in enumerate(g["counts"]):
        for key, c in counts.items():
            h[i, int(key, 2)] = c
    masks = np.array([[sum(1 << i for i, ch in enumerate(name[::-1]) if ch != "I")] for name in names], dtype=np.uint32)
    ss, tot = eng.parity_counts(torch.tensor(h, device="cuda"), masks)
    ss, tot = ss.cpu().numpy()[:, 0], tot.cpu().numpy()
    got = dict(zip(names, ss / tot))
    for name, (mean, std) in g["stabilizer_measurements"]:
        assert got[name] == mean
        shots = int(tot[names.index(name)])
        assert np.sqrt(max(1 - mean * mean, 0) / shots) == pytest.approx(std, rel=1e-12, abs=1e-18)


# ------------------------------------------------------------------------------------------- full-size properties
def test_n20_ideal_graph_state_stabilizers(eng):
    """BASELINE-size state (n = 20): every Toth generator of the ideal graph state has <S> = +1 on sampled counts
    (size-independent property: all sampled outcomes must have even parity under the generator mask)."""
    _reset_opts(eng)
    n = 20
    edges = circuits.ref_graph(n, 0)
    base, _ = ref_glue.cx_gate_circuit(n, edges)
    nbrs = {v: [b if a == v else a for a, b in edges if v in (a, b)] for v in range(n)}
    settings = []
    for v in range(0, n, 5):
        name = ["I"] * n
        name[v] = "X"
        for w in nbrs[v]:
            name[w] = "Z"
        settings.append("".join(name)[::-1])
    circs = [base + [("h", (i,)) for i, ch in enumerate(s[::-1]) if ch == "X"] for s in settings]
    eng.upload([_circ(n, ops) for ops in circs])
    masks = np.array([[sum(1 << i for i, ch in enumerate(s[::-1]) if ch != "I")] for s in settings], dtype=np.uint32)
    hist, _ = eng.run_trajectories(4096, seed=2, precision=32)
    ss, tot = eng.parity_counts(hist, masks)
    assert tot.cpu().tolist() == [4096] * len(settings)
    assert ss.cpu().numpy()[:, 0].tolist() == [4096] * len(settings)


def test_large_batch_parallel_lowering_matches_single_uploads(eng):
    """>= 64 circuits are lowered on several host threads and merged; results must equal one-circuit uploads."""
    _reset_opts(eng)
    n = 5
    rng = np.random.default_rng(77)
    batch = []
    for i in range(96):
        kind = i % 3
        if kind == 0:
            batch.append(circuits.graph_circuit_amp_damp(n, circuits.ref_graph(n, i), rng, max_prob=0.3))
        elif kind == 1:
            batch.append(circuits.device_noise_circuit(n, circuits.ref_graph(n, i), rng)[0])
        else:
            batch.append(circuits.random_circuit(n, 25, rng))
    eng.upload([_circ(n, ops) for ops in batch])
    probs = eng.run_density_matrix(64).cpu().numpy()
    _, out = eng.run_trajectories(64, seed=5, precision=64, want_outcomes=True)
    out = out.cpu().numpy().reshape(96, 64)
    for i in (0, 1, 2, 31, 32, 33, 64, 65, 95):
        np.testing.assert_allclose(probs[i], sim.density_matrix_probs(batch[i], n), atol=1e-10, rtol=0)
        want = sim.run_trajectories(batch[i], n, 64, 5, circuit=i)
        assert np.array_equal(out[i], want.astype(np.int64))


@pytest.mark.parametrize("n", [13, 14, 16])
def test_midsize_trajectories_match_compiled_oracle(eng, n):
    """n >= 13: single-pass sweeps run on the register-only kernel, multi-pass ones on the tile kernel.  complex128
    outcomes must equal the compiled CPU oracle (oracle/cpu_sim.c, itself pinned to the numpy oracle) shot for shot;
    complex64 may differ only where a draw lands within float rounding of a CDF boundary."""
    from oracle import cpu_build
    from qcdenoise_b200.ir import encode_programs
    _reset_opts(eng)
    rng = np.random.default_rng(1000 + n)
    base = circuits.ref_graph(14 if n >= 14 else 10, 1)
    extra = [(i, i + 1) for i in range(max(a for e in base for a in e), n - 1)]
    g = base + extra
    batch = [circuits.graph_circuit_amp_damp(n, g, rng, max_prob=0.3, all_sites=True),
             circuits.graph_circuit_amp_damp(n, g, rng, max_prob=0.2, builder="cz"),
             circuits.device_noise_circuit(n, g[:8], rng, gate_error=0.05)[0],
             circuits.random_circuit(n, 40, rng)]
    eng.upload([_circ(n, ops) for ops in batch])
    shots, seed = 160, 21
    _, o64 = eng.run_trajectories(shots, seed=seed, precision=64, want_outcomes=True)
    _, o32 = eng.run_trajectories(shots, seed=seed, precision=32, want_outcomes=True)
    o64 = o64.cpu().numpy().reshape(len(batch), shots)
    o32 = o32.cpu().numpy().reshape(len(batch), shots)
    for c, ops in enumerate(batch):
        want = cpu_build.run_trajectories(encode_programs([_circ(n, ops)]), shots, seed, circuit=c).astype(np.int64)
        assert np.array_equal(o64[c], want), f"circuit {c}: {np.count_nonzero(o64[c] != want)} outcomes differ"
        assert np.count_nonzero(o32[c] != want) <= 3
    # the tile kernel alone (register-only kernel switched off) gives the same complex128 outcomes
    eng.set_option("use_direct", 0)
    _, t64 = eng.run_trajectories(shots, seed=seed, precision=64, want_outcomes=True)
    eng.set_option("use_direct", 1)
    assert np.array_equal(t64.cpu().numpy().reshape(len(batch), shots), o64)


def test_dm_direct_and_tile_kernels_agree(eng):
    _reset_opts(eng)
    n = 8
    rng = np.random.default_rng(5)
    batch = [circuits.graph_circuit_amp_damp(n, circuits.ref_graph(n, i), rng, max_prob=0.3) for i in range(3)]
    eng.upload([_circ(n, ops) for ops in batch])
    a = eng.run_density_matrix(64).cpu().numpy()
    eng.set_option("use_direct", 0)
    b = eng.run_density_matrix(64).cpu().numpy()
    eng.set_option("use_direct", 1)
    np.testing.assert_allclose(a, b, atol=1e-13, rtol=0)
    np.testing.assert_allclose(a[0], sim.density_matrix_probs(batch[0], n), atol=1e-10, rtol=0)


def test_parity_outcomes_matches_parity_counts(eng):
    import torch
    n = 12
    rng = np.random.default_rng(3)
    out = rng.integers(0, 2 ** n, size=(3, 5000)).astype(np.int32)
    masks = rng.integers(0, 2 ** n, size=(3, 13)).astype(np.uint32)
    ss = eng.parity_outcomes(torch.tensor(out, device="cuda"), masks).cpu().numpy()
    hist = np.stack([np.bincount(o, minlength=2 ** n) for o in out]).astype(np.int32)
    ref, tot = eng.parity_counts(torch.tensor(hist, device="cuda"), masks)
    assert np.array_equal(ss, ref.cpu().numpy()) and tot.cpu().tolist() == [5000] * 3


def test_n28_ghz_outcomes_are_all_zero_or_all_one(eng):
    """Maximum size of the BASELINE configs (n = 28, 2 GiB complex64 state): a GHZ ladder -- the legacy reference
    circuit (circuit_constructors.py:1169-1192) -- may only produce 00...0 and 11...1; exercises 2^28 index
    arithmetic, scattered high tile bits and multi-pass sweeps without a CPU oracle of that size."""
    _reset_opts(eng)
    n = 28
    ops = [("h", (0,))] + [("cx", (q, q + 1)) for q in range(n - 1)]
    eng.upload([_circ(n, ops)])
    _, out = eng.run_trajectories(256, seed=3, precision=32, want_outcomes=True)
    vals, counts = np.unique(out.cpu().numpy().astype(np.int64) & 0xFFFFFFFF, return_counts=True)
    assert set(vals.tolist()) == {0, 2 ** n - 1}
    assert abs(int(counts[0]) - 128) < 60          # a fair coin over 256 shots
    ss = eng.parity_outcomes(out.reshape(1, -1), np.array([3, 2 ** n - 1, 1 | (1 << 27)], dtype=np.uint32)).cpu().numpy()[0]
    assert ss[0] == 256 and ss[1] == 256 and ss[2] == 256      # Z_i Z_j = +1 for every pair, even-weight parity too


def test_two_qubit_kraus_channels(eng):
    """4x4 Kraus sets (tensor products, as Aer presents `e1.tensor(e2)`): 16x16 superoperators in density-matrix mode,
    two-qubit per-branch matrices in trajectory mode; n = 7 / 13 also take the register-only kernel."""
    from tests.test_schedule_cpu import _two_qubit_kraus_circuit
    from oracle import cpu_build
    from qcdenoise_b200.ir import encode_programs
    _reset_opts(eng)
    rng = np.random.default_rng(23)
    for n in (4, 7):
        ops = _two_qubit_kraus_circuit(n, rng)
        eng.upload([_circ(n, ops)])
        want = sim.density_matrix_probs(ops, n)
        np.testing.assert_allclose(eng.run_density_matrix(64).cpu().numpy()[0], want, atol=1e-10, rtol=0)
        np.testing.assert_allclose(eng.run_density_matrix(32).cpu().numpy()[0], want, atol=1e-5, rtol=0)
        _, out = eng.run_trajectories(300, seed=8, precision=64, want_outcomes=True)
        assert np.array_equal(out.cpu().numpy().astype(np.uint64), sim.run_trajectories(ops, n, 300, 8))
    n = 13
    ops = _two_qubit_kraus_circuit(n, rng)
    eng.upload([_circ(n, ops)])
    _, out = eng.run_trajectories(200, seed=8, precision=64, want_outcomes=True)
    want = cpu_build.run_trajectories(encode_programs([_circ(n, ops)]), 200, 8)
    assert np.array_equal(out.cpu().numpy().astype(np.uint64), want)


# ------------------------------------------------------------------------------------------- shared trajectory trees
def _settings_batch(n, rng, graphs):
    """[graph circuit, its n + 1 Toth settings] per graph, plus one unrelated circuit at the end: runs of circuits that
    share everything up to their final sweep (stabilizers.py:194-199)."""
    batch = []
    for gi, all_sites in graphs:
        edges = circuits.ref_graph(n if n <= 10 else 14, gi)
        if n == 13:
            edges = [e for e in edges if max(e) < 13] + [(11, 12)]
        ops = circuits.graph_circuit_amp_damp(n, edges, rng, max_prob=0.3, all_sites=all_sites)
        batch.append(ops)
        for v in range(n):
            batch.append(ops + ([("sdg", (v,))] if v % 3 == 2 else []) + [("h", (v,))])
        batch.append(list(ops))
    batch.append(circuits.random_circuit(n, 30, rng))
    return batch


@pytest.mark.parametrize("n", [7, 13])
def test_shared_prefix_tree_is_bit_identical_to_per_circuit_trees(eng, n):
    """One trajectory tree per run of circuits with a common prefix == one tree per circuit: same histograms, same
    outcomes (every draw is keyed by the trajectory's own circuit and shot), fewer state sweeps; also when a run is cut
    by chunk boundaries, by the grouping-table cap, by a partial shot range and under the depth-first schedule."""
    from oracle import cpu_build
    from qcdenoise_b200.ir import encode_programs
    _reset_opts(eng)
    eng.set_option("leaf_cross", 0)        # sibling groups (tests/test_leaf_cross_gpu.py) round complex64 sums differently
    rng = np.random.default_rng(70 + n)
    batch = _settings_batch(n, rng, [(0, True), (1, False)])
    B = len(batch)
    eng.upload([_circ(n, ops) for ops in batch])
    shots, seed = 200, 31
    res = {}
    for prec in (32, 64):
        for share in (1, 0):
            eng.set_option("share_prefix", share)
            h, o = eng.run_trajectories(shots, seed=seed, precision=prec, want_outcomes=True)
            res[prec, share] = (h.cpu().numpy(), o.cpu().numpy(), eng.stats())
        assert np.array_equal(res[prec, 1][0], res[prec, 0][0]) and np.array_equal(res[prec, 1][1], res[prec, 0][1])
        assert res[prec, 1][2]["node_sweeps"] < 0.75 * res[prec, 0][2]["node_sweeps"]
        assert res[prec, 1][2]["leaves"] == res[prec, 0][2]["leaves"]
    eng.set_option("share_prefix", 1)
    out = res[64, 1][1].reshape(B, shots).astype(np.int64)
    for c in (0, 1, n, n + 1, n + 2, n + 3, B - 2, B - 1):
        run = cpu_build.run_trajectories if n >= 12 else None
        if run is not None:
            want = run(encode_programs([_circ(n, batch[c])]), shots, seed, circuit=c).astype(np.int64)
        else:
            want = sim.run_trajectories(batch[c], n, shots, seed, circuit=c).astype(np.int64)
        assert np.array_equal(out[c], want), f"circuit {c}"
    # runs cut by chunk boundaries / the table cap / few slots / a partial shot range
    ref_h, ref_o = res[64, 1][0], res[64, 1][1]
    for opts in ({"max_chunk": 3 * shots + 17}, {"table_cap_log2": 5}, {"max_slots": 40}, {"max_chunk": shots // 3}):
        for k, v in opts.items():
            eng.set_option(k, v)
        h, o = eng.run_trajectories(shots, seed=seed, precision=64, want_outcomes=True)
        assert np.array_equal(h.cpu().numpy(), ref_h) and np.array_equal(o.cpu().numpy(), ref_o), opts
        _reset_opts(eng)
    lo, hi = 2 * shots + 50, (B - 3) * shots + 120
    _, o = eng.run_trajectories(shots, seed=seed, precision=64, want_outcomes=True, shot_range=(lo, hi))
    assert np.array_equal(o.cpu().numpy(), ref_o[lo:hi])
    # literal mode (one state per shot) is unaffected by the option
    _, o = eng.run_trajectories(16, seed=seed, precision=64, want_outcomes=True, dedup=False)
    assert np.array_equal(o.cpu().numpy().reshape(B, 16), ref_o.reshape(B, shots)[:, :16])


def test_first_circuit_makes_a_circuit_split_reproduce_the_whole_batch(eng):
    """qcd_run_sv_trajectories(first_circuit): a rank that uploads only its slice of the circuits draws what a single
    GPU draws for the whole batch -- the 2-way split reproduces the unsplit histograms bit for bit."""
    _reset_opts(eng)
    n = 7
    rng = np.random.default_rng(5)
    batch = _settings_batch(n, rng, [(2, True)]) + [circuits.device_noise_circuit(n, circuits.ref_graph(7, 3), rng)[0]]
    B = len(batch)
    conf = np.stack([np.array([[[1 - a, a], [b, 1 - b]] for a, b in rng.uniform(0, 0.1, (n, 2))]) for _ in range(B)])
    eng.upload([_circ(n, ops) for ops in batch])
    for prec in (32, 64):
        whole, _ = eng.run_trajectories(300, seed=9, precision=prec, readout=conf)
        whole = whole.cpu().numpy()
        cut = 4                                   # inside the run of settings
        parts = []
        for lo, hi in ((0, cut), (cut, B)):
            eng.upload([_circ(n, ops) for ops in batch[lo:hi]])
            h, _ = eng.run_trajectories(300, seed=9, precision=prec, readout=conf[lo:hi], first_circuit=lo)
            parts.append(h.cpu().numpy())
        assert np.array_equal(np.concatenate(parts), whole)
        eng.upload([_circ(n, ops) for ops in batch])


@pytest.mark.parametrize("n", [13, 16])
def test_stored_and_unstored_leaves_sample_the_same_outcomes(eng, n):
    """A leaf sampled by few trajectories is never written: its chunk is recomputed from the parent per shot
    (sample_nostore_kernel); a leaf above the `nostore_max` threshold is stored and sampled from memory (sample_kernel).
    Both must give the oracle's outcomes shot for shot, whatever the threshold (0 = every leaf stored)."""
    from oracle import cpu_build
    from qcdenoise_b200.ir import encode_programs
    _reset_opts(eng)
    rng = np.random.default_rng(900 + n)
    batch = [circuits.graph_circuit_amp_damp(n, circuits.ref_graph(14 if n == 13 else n, i)[: n - 1] if n == 13 else
                                             circuits.ref_graph(14, i) + [(13, 14), (14, 15)], rng, max_prob=0.15)
             for i in range(2)]
    batch = [[op for op in ops if max(op[1]) < n] for ops in batch]
    eng.upload([_circ(n, ops) for ops in batch])
    shots, seed = 3000, 77
    ref = None
    for thr in (0, 8, 256, 1 << 20):
        eng.set_option("nostore_max", thr)
        res = {}
        for prec in (64, 32):
            h, o = eng.run_trajectories(shots, seed=seed, precision=prec, want_outcomes=True)
            res[prec] = (h.cpu().numpy(), o.cpu().numpy())
            assert res[prec][0].sum(axis=1).tolist() == [shots] * len(batch)
        if ref is None:
            ref = res
            out = res[64][1].reshape(len(batch), shots).astype(np.int64)
            for c, ops in enumerate(batch):
                want = cpu_build.run_trajectories(encode_programs([_circ(n, ops)]), shots, seed, circuit=c).astype(np.int64)
                assert np.array_equal(out[c], want), f"circuit {c}"
            assert (res[32][1] != res[64][1]).sum() <= 6          # complex64: a few draws on a CDF boundary
        else:
            for prec in (64, 32):
                assert np.array_equal(res[prec][0], ref[prec][0]) and np.array_equal(res[prec][1], ref[prec][1]), (thr, prec)
    _reset_opts(eng)


def test_release_memory_returns_the_arena_and_keeps_results(eng):
    import torch
    _reset_opts(eng)
    n = 14
    rng = np.random.default_rng(5)
    ops = circuits.graph_circuit_amp_damp(n, circuits.ref_graph(n, 0), rng, max_prob=0.2)
    eng.upload([_circ(n, ops)])
    h1, o1 = eng.run_trajectories(500, seed=3, precision=32, want_outcomes=True)
    torch.cuda.synchronize()
    free_before, _ = torch.cuda.mem_get_info()
    eng.set_option("release_memory", 1)
    free_after, _ = torch.cuda.mem_get_info()
    assert free_after > free_before + (1 << 20)                 # the arena went back to the device
    h2, o2 = eng.run_trajectories(500, seed=3, precision=32, want_outcomes=True)
    assert torch.equal(h1, h2) and torch.equal(o1, o2)


@pytest.mark.parametrize("n", [10, 11, 12])
def test_small_states_on_both_sweep_kernels(eng, n):
    """States below 13 index bits: the register-only kernel (default from 10 bits) and the shared-memory tile kernel
    (`direct_min_bits` 13) both reproduce the compiled oracle shot for shot in complex128; complex64 agrees with it up to a
    few draws on a CDF boundary."""
    from oracle import cpu_build
    from qcdenoise_b200.ir import encode_programs
    _reset_opts(eng)
    rng = np.random.default_rng(n)
    edges = circuits.ref_graph(10, 1) + [(9, 10), (10, 11)][: n - 10]
    batch = [circuits.graph_circuit_amp_damp(n, edges, rng, max_prob=0.25, all_sites=(i == 0)) for i in range(3)]
    eng.upload([_circ(n, ops) for ops in batch])
    shots, seed = 700, 5
    want = np.stack([cpu_build.run_trajectories(encode_programs([_circ(n, ops)]), shots, seed, circuit=c).astype(np.int64)
                     for c, ops in enumerate(batch)])
    for min_bits in (13, 10):
        eng.set_option("direct_min_bits", min_bits)
        _, o64 = eng.run_trajectories(shots, seed=seed, precision=64, want_outcomes=True)
        _, o32 = eng.run_trajectories(shots, seed=seed, precision=32, want_outcomes=True)
        got = o64.cpu().numpy().reshape(len(batch), shots).astype(np.int64)
        assert np.array_equal(got, want), min_bits
        assert (o32.cpu().numpy().reshape(len(batch), shots) != got).sum() <= 4, min_bits
    _reset_opts(eng)
