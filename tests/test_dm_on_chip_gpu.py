"""GPU parity of the on-chip density-matrix interpreter (csrc/dm_cluster.cu, option "dm_on_chip"; forced with the value 2 here:
the default policy only takes clusters of up to 4 CTAs): the whole density matrix
in the shared memory of a thread-block cluster (1 / 4 / 16 CTAs for n <= 7 / 8 / 9 qubits in complex64; 1 / 2 / 8 for
n <= 6 / 7 / 8 in complex128), the gate list interpreted there.  Against the numpy oracle's density matrix (1e-5 complex64,
1e-10 complex128, the north_star bars) and against the lowered sweep route on the same inputs.
Reference seam: AerSimulator density_matrix method behind samplers.py:308-326."""
import numpy as np
import pytest

from oracle import noise, sim
from tests import circuits

pytestmark = pytest.mark.gpu
TOL = {32: 1e-5, 64: 1e-10}


@pytest.fixture(scope="module")
def eng():
    import qcdenoise_b200 as qb
    e = qb.Engine(0)
    yield e
    e.set_option("dm_on_chip", 1)
    e.set_option("dm_max_clusters", 0)
    e.close()


def _circ(n, ops):
    import qcdenoise_b200 as qb
    return qb.Circuit(n, ops)


def _kraus2(rng):
    """A random two-qubit channel with 3 Kraus operators (isometry slices)."""
    a = rng.normal(size=(12, 4)) + 1j * rng.normal(size=(12, 4))
    q, _ = np.linalg.qr(a)
    return [q[4 * k:4 * k + 4, :] for k in range(3)]


@pytest.mark.parametrize("precision", [32, 64])
@pytest.mark.parametrize("n", [1, 2, 3, 5, 6, 7, 8])
def test_every_opcode_against_the_oracle(eng, n, precision):
    """Random circuits over every opcode of the IR (rotations, custom 1q / 2q unitaries, CX / CZ / SWAP in both qubit orders,
    Kraus lists, 1q / 2q Pauli mixtures, reset, thermal relaxation, a 2q Kraus channel), on the on-chip route and on the
    lowered route."""
    batch = []
    for seed in range(3):
        rng = np.random.default_rng(1000 * n + seed)
        ops = circuits.random_circuit(n, 50, rng) if n > 1 else \
            [("h", (0,)), ("kraus", (0,), noise.amplitude_damping_kraus(0.2)), ("t", (0,)), ("rx", (0,), 0.7),
             ("pauli", (0,), noise.depolarizing_pauli_probs(0.1, 1)), ("ry", (0,), -1.1), ("reset", (0,)), ("h", (0,))]
        if n >= 2:
            ops.insert(len(ops) // 2, ("kraus", (n - 1, 0), _kraus2(rng)))
            ops.append(("cx", (n - 1, 0)))
        batch.append(ops)
    eng.upload([_circ(n, ops) for ops in batch])
    want = [sim.density_matrix_probs(ops, n) for ops in batch]
    for on_chip in (2, 0):                                  # 2: on chip whatever the cluster size (complex128 n = 8: 8 CTAs)
        eng.set_option("dm_on_chip", on_chip)
        got = eng.run_density_matrix(precision).cpu().numpy()
        assert (eng.stats()["sweep_launches"] == 0) == (on_chip == 2)
        for i in range(len(batch)):
            np.testing.assert_allclose(got[i], want[i], atol=TOL[precision], rtol=0, err_msg=f"on_chip={on_chip} circuit {i}")
    eng.set_option("dm_on_chip", 2)


@pytest.mark.parametrize("precision", [32, 64])
@pytest.mark.parametrize("builder", ["cx", "cz", "cphase"])
def test_device_noise_circuits_of_every_builder(eng, builder, precision):
    """The pass shapes of the reference's three builders under device noise (CX and CZ entangling gates run the fused
    straight-line pass, CPhase the generic loop), local and cluster-spanning passes, against the oracle."""
    for n in (6, 8):
        rng = np.random.default_rng(40 + n)
        batch = [circuits.device_noise_circuit(n, circuits.ref_graph(8 if n == 8 else 5, i) + ([(4, 5)] if n == 6 else []),
                                               rng, builder=builder)[0] for i in range(3)]
        eng.upload([_circ(n, ops) for ops in batch])
        eng.set_option("dm_on_chip", 2)
        got = eng.run_density_matrix(precision).cpu().numpy()
        eng.set_option("dm_on_chip", 2)
        for i, ops in enumerate(batch):
            np.testing.assert_allclose(got[i], sim.density_matrix_probs(ops, n), atol=TOL[precision], rtol=0,
                                       err_msg=f"{builder} n={n} circuit {i}")


def test_nine_qubits_sixteen_cta_cluster(eng):
    n = 9
    rng = np.random.default_rng(9)
    batch = [circuits.device_noise_circuit(n, circuits.ref_graph(n, i), rng)[0] for i in range(2)]
    batch.append(circuits.random_circuit(n, 40, rng))
    eng.upload([_circ(n, ops) for ops in batch])
    eng.set_option("dm_on_chip", 2)                               # 16-CTA clusters are not the default route (slower end to end)
    got = eng.run_density_matrix(32).cpu().numpy()
    for i, ops in enumerate(batch):
        np.testing.assert_allclose(got[i], sim.density_matrix_probs(ops, n), atol=1e-5, rtol=0)
    assert eng.stats()["sweep_launches"] == 0                     # nothing was lowered to sweeps
    eng.set_option("dm_on_chip", 1)
    auto = eng.run_density_matrix(32).cpu().numpy()               # default policy: n = 9 goes to the lowered sweeps
    assert eng.stats()["sweep_launches"] > 0
    np.testing.assert_allclose(auto, got, atol=2e-6, rtol=0)


@pytest.mark.parametrize("precision", [32, 64])
def test_many_circuits_walk_over_few_clusters(eng, precision):
    """More circuits than resident clusters (and a cap of 3 clusters): each cluster walks over several circuits and has to
    re-initialise its shared memory between them; a circuit slice selects the right programs."""
    n = 8
    rng = np.random.default_rng(5)
    batch = [circuits.device_noise_circuit(n, circuits.ref_graph(n, i), rng)[0] for i in range(11)]
    eng.upload([_circ(n, ops) for ops in batch])
    eng.set_option("dm_on_chip", 2)
    eng.set_option("dm_max_clusters", 3)
    got = eng.run_density_matrix(precision).cpu().numpy()
    part = eng.run_density_matrix(precision, circuit_range=(4, 9)).cpu().numpy()
    eng.set_option("dm_max_clusters", 0)
    for i, ops in enumerate(batch):
        np.testing.assert_allclose(got[i], sim.density_matrix_probs(ops, n), atol=TOL[precision], rtol=0)
    assert np.array_equal(part, got[4:9])


def test_native_device_noise_upload_runs_on_chip(eng):
    """The batch route of simulate(): noise-free gate lists + calibration numbers, channels written natively -- on chip
    the result equals the explicit programs and the lowered route."""
    import qcdenoise_b200 as qb
    from qcdenoise_b200 import batch
    n, B = 8, 300
    gs = qb.GraphState(qb.GraphDB(), n_qubits=n)
    edges, n_edges = gs.sample_batch(B, seed=3)
    enc = batch.build_gate_lists(qb.CXGateCircuit(n_qubits=n, stochastic=False), edges, n_edges)
    spec = qb.NoiseSpec(specs={"qubit": {"T1": 100.0, "T2": 100.0, "readout_error": 0.25, "prob_meas0_prep1": 0.1,
                                         "prob_meas1_prep0": 0.1},
                               "cx": {"gate_error": 0.25, "gate_length": 500.0}, "h": {"gate_error": 0.05, "gate_length": 50.0}})
    props = batch.draw_device_props(spec, enc, np.random.default_rng(1))
    t1t2, gp, _ = batch.device_noise_inputs(props)
    eng.upload_device_noise(enc, t1t2, gp)
    eng.set_option("dm_on_chip", 2)
    a = eng.run_density_matrix(64).cpu().numpy()
    assert eng.stats()["sweep_launches"] == 0
    eng.set_option("dm_on_chip", 0)
    b = eng.run_density_matrix(64).cpu().numpy()
    assert eng.stats()["sweep_launches"] > 0
    eng.set_option("dm_on_chip", 2)
    np.testing.assert_allclose(a, b, atol=1e-12, rtol=0)
    np.testing.assert_allclose(a.sum(axis=1), 1.0, atol=1e-12)
    from qcdenoise_b200.engine import expand_device_noise
    x_ops, x_off, x_par, _ = expand_device_noise(enc, t1t2, gp)
    eng.upload((x_ops, x_off, x_par, n))
    np.testing.assert_allclose(eng.run_density_matrix(64).cpu().numpy(), a, atol=1e-13, rtol=0)


@pytest.mark.parametrize("n", [5, 8])
def test_read_outs_that_need_the_whole_matrix(eng, n):
    """Toth settings and the full stabilizer group are read off the density matrix the interpreter wrote to the state
    arena: identical to the lowered route."""
    import networkx as nx
    import qcdenoise_b200 as qb
    rng = np.random.default_rng(n)
    edges = circuits.ref_graph(n, 2)
    batch = [circuits.graph_circuit_amp_damp(n, edges, rng, max_prob=0.3, all_sites=True) for _ in range(5)]
    eng.upload([_circ(n, ops) for ops in batch])
    g = nx.Graph()
    g.add_nodes_from(range(n))
    g.add_edges_from(edges)
    H = np.array([[1, 1], [1, -1]], dtype=np.complex128) / np.sqrt(2)
    sq = np.broadcast_to(np.array([k if k < n else 0xFF for k in range(n + 1)], dtype=np.uint8), (5, n + 1)).copy()
    sop = np.broadcast_to(np.kron(H.conj(), H), (5, n + 1, 4, 4)).copy()
    res = {}
    for on_chip in (2, 0):
        eng.set_option("dm_on_chip", on_chip)
        eng.set_option("max_slots", 2)                     # several arena chunks
        res[on_chip] = (eng.run_density_matrix_settings(sq, sop, precision=64).cpu().numpy(),
                        qb.group_expectations(eng, g, precision=64).cpu().numpy())
    eng.set_option("max_slots", 0)
    eng.set_option("dm_on_chip", 2)
    np.testing.assert_allclose(res[2][0], res[0][0], atol=1e-12, rtol=0)
    np.testing.assert_allclose(res[2][1], res[0][1], atol=1e-12, rtol=0)
    for i, ops in enumerate(batch):                         # the unrotated setting is the circuit's own distribution
        np.testing.assert_allclose(res[2][0][i, n], sim.density_matrix_probs(ops, n), atol=1e-10, rtol=0)


def test_refused_cluster_launch_falls_back_to_lowered_sweeps(eng):
    """A device that cannot co-schedule the cluster (16 CTAs is a non-portable size) must not fail the call: the context
    switches to the lowered sweep route for good and gives the same distribution."""
    n = 8
    rng = np.random.default_rng(21)
    batch = [circuits.device_noise_circuit(n, circuits.ref_graph(n, i), rng)[0] for i in range(3)]
    eng.upload([_circ(n, ops) for ops in batch])
    eng.set_option("dm_on_chip", 2)
    want = eng.run_density_matrix(64).cpu().numpy()
    assert eng.stats()["sweep_launches"] == 0
    eng.set_option("dm_test_refuse", 1)
    try:
        got = eng.run_density_matrix(64).cpu().numpy()
        assert eng.stats()["sweep_launches"] > 0
        np.testing.assert_allclose(got, want, atol=1e-12, rtol=0)
        again = eng.run_density_matrix(64).cpu().numpy()             # no second attempt at the cluster kernel
        np.testing.assert_allclose(again, want, atol=1e-12, rtol=0)
    finally:
        eng.set_option("dm_test_refuse", 0)
    eng.run_density_matrix(64)
    assert eng.stats()["sweep_launches"] == 0                        # back on chip


@pytest.mark.parametrize("precision", [32, 64])
def test_fused_pass_shapes_and_near_misses(eng, precision):
    """Every variant of the fused pass [S0] [S1] CX|CZ [DEPOL] [S0'] [S1'] (pieces present or absent, control on either slot,
    X-shaped and dense real superoperators) and the near-misses that must take the generic per-record loop instead (a complex
    superoperator, a non-uniform Pauli mixture, a Pauli mixture after a superoperator, two entangling gates in one pass),
    local (n = 6) and cluster-spanning (n = 8, qubits 6 and 7), against the oracle."""
    dep = noise.depolarizing_pauli_probs(0.07, 2)
    skew = np.array(dep, dtype=float)
    skew[5] += 0.01
    skew[0] -= 0.01
    damp = lambda g: ("kraus", None, noise.amplitude_damping_kraus(g))        # noqa: E731
    relax = ("kraus", None, noise.thermal_relaxation_kraus(60.0, 80.0, 0.4))

    def on(op, *qs):
        return (op[0], tuple(qs)) + tuple(op[2:])

    def variants(a, b):
        h = ("h", None)
        yield [on(("cx", None), a, b)]                                                          # bare CX
        yield [on(("cx", None), a, b), ("pauli", (a, b), dep)]                                  # CX DEPOL
        yield [on(h, a), on(("cx", None), b, a), ("pauli", (b, a), dep), on(damp(0.2), a)]      # control on slot 1
        yield [on(h, a), on(h, b), on(("cz", None), a, b), ("pauli", (a, b), dep), on(relax, a), on(h, b), on(damp(0.1), b)]
        yield [on(h, b), on(damp(0.3), b), on(("cx", None), a, b), on(relax, b)]                # no DEPOL, one slot only
        yield [on(("t", None), a), on(("cx", None), a, b), ("pauli", (a, b), dep)]              # complex S0 -> generic
        yield [on(h, a), on(("cx", None), a, b), ("pauli", (a, b), tuple(skew))]                # non-uniform mixture -> generic
        yield [on(("cx", None), a, b), on(damp(0.2), a), ("pauli", (a, b), dep)]                # mixture after S0' -> generic
        yield [on(("cx", None), a, b), ("pauli", (a, b), dep), on(("cx", None), b, a), on(h, a)]  # two gates -> generic

    for n, pairs in ((6, [(0, 1), (4, 2)]), (8, [(7, 0), (3, 6), (6, 7)])):
        batch = []
        for a, b in pairs:
            for body in variants(a, b):
                # a prologue that makes the state generic (entangled, mixed, complex) before the pass under test
                pro = [("h", (q,)) for q in range(n)] + [("cz", (q, (q + 1) % n)) for q in range(0, n - 1, 2)] + \
                      [("s", (a,)), ("kraus", (b,), noise.amplitude_damping_kraus(0.15)), ("ry", ((a + 2) % n,), 0.4)]
                batch.append(pro + body + [("h", ((b + 1) % n,))])
        eng.upload([_circ(n, ops) for ops in batch])
        eng.set_option("dm_on_chip", 2)
        got = eng.run_density_matrix(precision).cpu().numpy()
        assert eng.stats()["sweep_launches"] == 0
        for i, ops in enumerate(batch):
            np.testing.assert_allclose(got[i], sim.density_matrix_probs(ops, n), atol=TOL[precision], rtol=0,
                                       err_msg=f"n={n} circuit {i}")
