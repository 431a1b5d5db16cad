"""CPU checks of the host scheduler (csrc/lower.cpp, through the C ABI's qcd_debug_schedule_json): the emitted
sweep program, interpreted with numpy exactly the way the kernels execute it, must reproduce the oracle --
density-matrix probabilities to 1e-12 and trajectory outcomes shot for shot."""
import numpy as np
import pytest

import qcdenoise_b200 as qb
from oracle import sim
from tests import circuits, schedule_interp as si


def _sched(n, ops, mode, tile_bits=0, precision=64):
    return qb.debug_schedule(qb.Circuit(n, ops), mode=mode, precision=precision, tile_bits=tile_bits)


@pytest.mark.parametrize("n,tile_bits", [(4, 0), (5, 9), (6, 9), (7, 0), (7, 10)])
def test_dm_graph_amp_damp(n, tile_bits):
    rng = np.random.default_rng(100 + n)
    for i in range(3):
        ops = circuits.graph_circuit_amp_damp(n, circuits.ref_graph(n if n in (4, 5, 7) else 5, i)[: n - 1] if n == 6
                                              else circuits.ref_graph(n, i), rng, max_prob=0.3)
        prog = _sched(n, ops, 1, tile_bits)
        np.testing.assert_allclose(si.run_dm(prog), sim.density_matrix_probs(ops, n), atol=1e-12)


@pytest.mark.parametrize("builder", ["cx", "cz", "cphase"])
def test_dm_builders(builder):
    rng = np.random.default_rng(7)
    n = 5
    ops = circuits.graph_circuit_amp_damp(n, circuits.ref_graph(n, 2), rng, builder=builder, all_sites=True)
    np.testing.assert_allclose(si.run_dm(_sched(n, ops, 1)), sim.density_matrix_probs(ops, n), atol=1e-12)


@pytest.mark.parametrize("seed", range(6))
def test_dm_random_circuits(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(2, 6))
    ops = circuits.random_circuit(n, 40, rng)
    tb = [0, 9, 9][seed % 3]
    np.testing.assert_allclose(si.run_dm(_sched(n, ops, 1, tb)), sim.density_matrix_probs(ops, n), atol=1e-12)


def test_dm_device_noise():
    rng = np.random.default_rng(3)
    n = 5
    ops, _ = circuits.device_noise_circuit(n, circuits.ref_graph(n, 1), rng)
    np.testing.assert_allclose(si.run_dm(_sched(n, ops, 1)), sim.density_matrix_probs(ops, n), atol=1e-12)


@pytest.mark.parametrize("seed", range(4))
def test_ideal_statevector(seed):
    rng = np.random.default_rng(50 + seed)
    n = int(rng.integers(3, 10))
    ops = circuits.random_circuit(n, 60, rng, noise_sites=False)
    prog = _sched(n, ops, 0, [0, 7, 7, 8][seed])
    state, _ = si.run_trajectory(prog, 1, 0, 0)
    np.testing.assert_allclose(state, sim.ideal_statevector(ops, n), atol=1e-12)


def _check_trajectories(n, ops, shots, seed, tile_bits=0, circuit=0):
    prog = _sched(n, ops, 0, tile_bits)
    got = si.run_shots(prog, seed, circuit, shots)
    want = sim.run_trajectories(ops, n, shots, seed, circuit=circuit)
    assert np.array_equal(got, want), f"{np.count_nonzero(got != want)} of {shots} outcomes differ"
    return prog


@pytest.mark.parametrize("n,tile_bits", [(4, 0), (6, 0), (8, 0), (8, 7)])
def test_trajectories_graph_amp_damp(n, tile_bits):
    rng = np.random.default_rng(200 + n)
    g = circuits.ref_graph(8 if n == 8 else 4 if n == 4 else 5, 1)
    if n == 6:
        g = g + [(4, 5)]
    ops = circuits.graph_circuit_amp_damp(n, g, rng, max_prob=0.35, all_sites=True)
    prog = _check_trajectories(n, ops, 300, seed=99, tile_bits=tile_bits, circuit=3)
    # one sweep per state-dependent site group + the prep sweep, not one per gate
    n_sites = sum(1 for op in ops if op[0] == "unitary")
    assert len(prog["sweeps"]) <= n_sites + 1 + (2 if tile_bits else 0) * n_sites


def test_trajectories_depolarizing_C1():
    ops = circuits.depolarizing_circuit(4, circuits.ref_graph(4, 2), p1=0.05, p2=0.2)
    _check_trajectories(4, ops, 400, seed=5)


def test_trajectories_device_noise():
    rng = np.random.default_rng(11)
    n = 5
    ops, _ = circuits.device_noise_circuit(n, circuits.ref_graph(n, 0), rng, gate_error=0.2, gate_len_ns=(500, 5000))
    _check_trajectories(n, ops, 300, seed=2)


@pytest.mark.parametrize("seed", range(6))
def test_trajectories_random_circuits(seed):
    rng = np.random.default_rng(300 + seed)
    n = int(rng.integers(2, 7))
    ops = circuits.random_circuit(n, 30, rng)
    _check_trajectories(n, ops, 200, seed=seed, tile_bits=[0, 7][seed % 2])


def test_schedule_fuses_reference_circuit():
    """n=20 reference circuit: sweeps = state-dependent site groups + 1 (+ tile-fit splits), every non-diagonal
    target inside the tile (asserted by the interpreter), <= 4 register bits per pass."""
    rng = np.random.default_rng(1)
    n = 20
    ops = circuits.graph_circuit_amp_damp(n, circuits.ref_graph(n, 0), rng)
    K = sum(1 for op in ops if op[0] == "unitary")
    prog = _sched(n, ops, 0, precision=32)
    assert K + 1 <= len(prog["sweeps"]) <= 2 * K + 2
    for sw in prog["sweeps"]:
        assert sw["k"] == 13 and sw["L"] >= 5 and len(sw["high"]) == sw["k"] - sw["L"]
        for ps in sw["passes"]:
            assert len(ps["rbits"]) == 4


def test_invalid_programs_rejected():
    from qcdenoise_b200 import _lib
    with pytest.raises(_lib.QcdError):
        qb.debug_schedule(qb.Circuit(3, [("pauli", (0,), np.array([0.5, 0.2, 0.2, 0.2]))]))
    with pytest.raises(_lib.QcdError):
        qb.debug_schedule(qb.Circuit(3, [("cx", (0, 5))]))


def _two_qubit_kraus_circuit(n, rng):
    """Two-qubit Kraus channels (tensor products of single-qubit damping / relaxation operators, given as ONE
    4x4 Kraus set the way Aer's `error1.tensor(error2)` presents them) on several pairs, between entangling gates."""
    from oracle import noise
    ops = [("h", (q,)) for q in range(n)]
    for (a, b) in [(0, 1), (2, 1), (n - 1, 0), (1, n - 2)]:
        ops.append(("cx", (a, b)))
        k0 = noise.amplitude_damping_kraus(float(rng.uniform(0.05, 0.4)))
        k1 = noise.thermal_relaxation_kraus(50.0, 70.0, float(rng.uniform(5, 40)))
        ops.append(("kraus", (a, b), noise.tensor_kraus(k1, k0)))      # k0 acts on a (LSB), k1 on b
        ops.append(("h", (b,)))
    return ops


def test_two_qubit_kraus_dm_and_trajectories():
    rng = np.random.default_rng(17)
    for n in (4, 5):
        ops = _two_qubit_kraus_circuit(n, rng)
        np.testing.assert_allclose(si.run_dm(_sched(n, ops, 1)), sim.density_matrix_probs(ops, n), atol=1e-12)
        _check_trajectories(n, ops, 300, seed=4)


def test_dm_asymmetric_pauli_mixtures():
    """Density-matrix lowering builds Pauli-mixture superoperators straight from the Paulis' monomial form: check it
    with probabilities that distinguish X, Y and Z (depolarizing mixtures would not) on 1 and 2 qubits."""
    rng = np.random.default_rng(17)
    n = 4
    for _ in range(4):
        ops = [("h", (q,)) for q in range(n)] + [("ry", (1,), 0.7), ("rx", (2,), 1.1), ("cx", (0, 1)), ("cz", (2, 3))]
        p1 = rng.dirichlet(np.ones(4))
        p2 = rng.dirichlet(np.ones(16))
        p2[int(rng.integers(1, 16))] = 0.0
        p2 /= p2.sum()
        ops += [("pauli", (1,), p1), ("t", (1,)), ("pauli", (3, 0), p2), ("ry", (3,), 0.3), ("s", (0,)), ("h", (0,)),
                ("pauli", (0, 2), rng.dirichlet(np.ones(16))), ("h", (2,)), ("rx", (3,), 0.9)]
        np.testing.assert_allclose(si.run_dm(_sched(n, ops, 1)), sim.density_matrix_probs(ops, n), atol=1e-12)


def test_prefix_signature_groups_a_circuit_with_its_stabilizer_settings():
    """Circuits that differ only after their last state-dependent noise site (a graph circuit and its Toth settings,
    stabilizers.py:194-199) get the same prefix signature -- the engine runs ONE trajectory tree for such a run down to
    the parents of the leaves; different channel parameters or an extra gate before the last site give another one."""
    import qcdenoise_b200 as qb
    n = 6
    edges = circuits.ref_graph(5, 1) + [(3, 5)]
    rng = np.random.default_rng(11)
    ops = circuits.graph_circuit_amp_damp(n, edges, rng, max_prob=0.3, all_sites=True)
    sig = lambda o: tuple(qb.debug_schedule(qb.Circuit(n, o))["prefix_sig"])      # noqa: E731
    base = sig(ops)
    assert base[2] >= 2                                                             # shared sweeps before the final one
    for q in range(n):
        assert sig(ops + [("h", (q,))]) == base
        assert sig(ops + [("sdg", (q,)), ("h", (q,))]) == base
    other = circuits.graph_circuit_amp_damp(n, edges, np.random.default_rng(12), max_prob=0.3, all_sites=True)
    assert sig(other) != base
    first_kraus = next(i for i, op in enumerate(ops) if op[0] == "kraus")
    assert sig(ops[:first_kraus] + [("x", (0,))] + ops[first_kraus:]) != base
