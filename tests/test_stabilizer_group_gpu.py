"""SURVEY 8 f3 on the GPU: device enumeration of the signed stabilizer group (golden G5, the host enumeration), exact
<S> of ALL group elements read off one density matrix (dense oracle at n <= 7; exact Pauli propagation at n = 10, 14)
and the noise-robust witnesses through the public API.  Reference: circuit_constructors.py:537-590,
stabilizers.py:39-79,171-172, witnesses.py:112-144."""
import json
import os

import networkx as nx
import numpy as np
import pytest

from oracle import pauli_prop, ref_glue, sim
from tests import circuits

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
TOL = {32: 2e-5, 64: 1e-10}


@pytest.fixture(scope="module")
def eng():
    import qcdenoise_b200 as qb
    e = qb.Engine(0)
    yield e
    e.close()


def _graph(n, edges):
    g = nx.Graph()
    g.add_nodes_from(range(n))
    g.add_edges_from(edges)
    return g


def test_device_enumeration_golden_G5_and_host(eng):
    import qcdenoise_b200 as qb
    with open(os.path.join(HERE, "golden", "witness_goldens.json")) as f:
        g5 = json.load(f)["G5"]
    for case in g5:
        n = len(case["generators"][0])
        js = qb.JungStabilizer(nx.empty_graph(n), n)
        js.generators = case["generators"]
        assert js.find_stabilizers(engine=eng) == case["stabilizers"]
    for n, idx in ((7, 0), (10, 1), (14, 2)):
        g = _graph(n, circuits.ref_graph(n, idx))
        host = qb.JungStabilizer(g, n).find_stabilizers(max_elements=3000)
        assert qb.JungStabilizer(g, n).find_stabilizers(max_elements=3000, engine=eng) == host
    # a slice of a 20-generator group: elements [begin, end) only, (x, z, sign) bit for bit against the host arithmetic
    n = 20
    g = _graph(n, circuits.ref_graph(n, 0))
    js = qb.JungStabilizer(g, n)
    masks = js.generator_masks()
    begin, end = 777_000, 777_000 + 4096
    x, z, s = (t.cpu().numpy() for t in eng.stabilizer_group([m[0] for m in masks], [m[1] for m in masks], begin, end))
    for t in range(0, 4096, 97):
        sel, hx, hz, power = begin + t, 0, 0, 0
        for j, (gx, gz) in enumerate(masks):
            if (sel >> (n - 1 - j)) & 1:
                power = (power + 2 * bin(hz & gx).count("1")) % 4
                hx, hz = hx ^ gx, hz ^ gz
        power = (power - bin(hx & hz).count("1")) % 4
        assert (int(x[t]), int(z[t]), int(s[t])) == (hx, hz, power // 2) and power in (0, 2)
    with pytest.raises(ArithmeticError):          # X and Z on one qubit do not commute
        eng.stabilizer_group([1, 0], [0, 1])


@pytest.mark.parametrize("precision", [32, 64])
@pytest.mark.parametrize("n", [3, 4, 5, 6, 7])
def test_group_expectations_against_dense_oracle(eng, n, precision):
    import qcdenoise_b200 as qb
    rng = np.random.default_rng(100 + n)
    src = {3: 4, 6: 5}.get(n, n)
    edges = [e for e in circuits.ref_graph(src, 1) if max(e) < n] + ([(4, 5)] if n == 6 else [])
    g = _graph(n, edges)
    gens = ref_glue.get_generators(n, list(g.nodes()), {v: list(g.neighbors(v)) for v in g.nodes()})
    group = ref_glue.full_stabilizer_group(gens, n)
    batch = [circuits.graph_circuit_amp_damp(n, edges, rng, max_prob=0.3, all_sites=True) for _ in range(3)]
    batch.append(ref_glue.cx_gate_circuit(n, edges)[0])                           # the ideal graph state
    eng.upload([qb.Circuit(n, ops) for ops in batch])
    got = qb.group_expectations(eng, g, precision=precision).cpu().numpy()
    assert got.shape == (4, 2 ** n)
    for i, ops in enumerate(batch):
        rho = sim.density_matrix(ops, n).reshape(2 ** n, 2 ** n).T                # vec index = x + 2^n y <-> rho[x, y]
        want = ref_glue.group_expectations_dense(rho, group)
        for s in group:
            x = sum(1 << q for q, ch in enumerate(s[:0:-1]) if ch in "XY")
            assert abs(got[i, x] - want[s[1:]]) < TOL[precision], (i, s)
        w2 = qb.fidelity_witness_exact(eng, g, precision=precision, circuit_range=(i, i + 1)).cpu().numpy()[0]
        assert abs(w2 - ref_glue.projector_witness_dense(rho, group, n)) < TOL[precision]
    np.testing.assert_allclose(got[3], 1.0, atol=TOL[precision])                  # every element stabilizes |G>


@pytest.mark.parametrize("n", [10, 14])
def test_group_expectations_at_size_against_pauli_propagation(eng, n):
    """All 2^n elements from one pass over a 4^n density matrix; a random sample of them against the exact Heisenberg-
    picture value of the same observable measured the reference's way (setting circuit + Z-parity)."""
    import qcdenoise_b200 as qb
    rng = np.random.default_rng(n)
    edges = circuits.ref_graph(n, 0)
    g = _graph(n, edges)
    ops = circuits.graph_circuit_amp_damp(n, edges, rng, max_prob=0.2)
    eng.upload([qb.Circuit(n, ops)])
    got = qb.group_expectations(eng, g, precision=32).cpu().numpy()[0]
    assert abs(got[0] - 1.0) < 1e-5                                               # identity: Tr rho
    js = qb.JungStabilizer(g, n)
    signed = js.find_stabilizers(engine=eng)
    picks = rng.choice(2 ** n, size=48, replace=False)
    for sel in picks:
        s = signed[sel]
        label = s[1:]
        sub = []
        for q, ch in enumerate(label[::-1]):
            if ch == "X":
                sub.append(("h", (q,)))
            elif ch == "Y":
                sub += [("sdg", (q,)), ("h", (q,))]
        mask = sum(1 << q for q, ch in enumerate(label[::-1]) if ch != "I")
        want = pauli_prop.expectation(list(ops) + sub, n, mask) * (1 if s[0] == "+" else -1)
        x = sum(1 << q for q, ch in enumerate(label[::-1]) if ch in "XY")
        assert abs(got[x] - want) < 3e-5, (s, got[x], want)
    # the generators are the Toth settings: the same numbers as the per-setting route
    fid = got.mean()
    assert 0.0 < fid < 1.0 and abs(qb.fidelity_witness_exact(eng, g).cpu().numpy()[0] - (0.5 - fid)) < 1e-12


def test_noise_robust_witnesses_through_the_api(eng):
    """16 measurement settings of the full group of P4 sampled through StabilizerSampler; levels 1 and 2 of
    GenuineWitness.estimate against the dense oracle on the exact density matrix, within sampling error; exact on the
    ideal state."""
    import qcdenoise_b200 as qb
    n, edges = 4, [(0, 1), (1, 2), (2, 3)]
    g = _graph(n, edges)
    np.random.seed(11)
    cd = qb.CXGateCircuit(n_qubits=n, stochastic=False).build(g)
    stab = qb.JungStabilizer(g, n_qubits=n).build()
    assert len(stab) == 16
    shots = 200_000
    ss = qb.StabilizerSampler(None, n_shots=shots, method="statevector", precision=64, seed=5)
    counts = ss.sample(stab, cd["circuit"])
    wit = qb.GenuineWitness(n, stab, counts)
    assert abs(wit.estimate(g, noise_robust=2).value + 0.5) < 1e-12 and wit.estimate(g, noise_robust=2).variance == 0
    assert abs(wit.estimate(g, noise_robust=1).value + 1.0) < 1e-12
    # noisy: amplitude damping after every entangling gate (the reference's unitary-noise sites)
    np.random.seed(12)
    cdn = qb.CXGateCircuit(n_qubits=n, stochastic=True).build(g)
    while not cdn["ops"]:
        cdn = qb.CXGateCircuit(n_qubits=n, stochastic=True).build(g)
    us = qb.UnitaryNoiseSampler(None, n_shots=shots, method="statevector", precision=64, seed=6)
    us.build_noise_model(cdn["ops"])
    counts = ss.sample(stab, cdn["circuit"], noise_model=us.noise_model)
    wit = qb.GenuineWitness(n, stab, counts)
    noisy_ops = us.noise_model.instantiate(cdn["circuit"])[0].ops
    rho = sim.density_matrix(noisy_ops, n).reshape(2 ** n, 2 ** n).T
    gens = ref_glue.get_generators(n, list(g.nodes()), {v: list(g.neighbors(v)) for v in g.nodes()})
    group = ref_glue.full_stabilizer_group(gens, n)
    w2 = wit.estimate(g, noise_robust=2)
    want2 = ref_glue.projector_witness_dense(rho, group, n)
    assert w2.variance > 0 and abs(w2.value - want2) < 5 * w2.variance and want2 > -0.5 + 1e-3
    colour = nx.bipartite.color(g)
    w1 = wit.estimate(g, noise_robust=1)
    want1 = ref_glue.two_colour_witness_dense(rho, gens, [colour[v] for v in g.nodes()], n)
    assert abs(w1.value - want1) < 5 * w1.variance and want1 >= 2 * want2 - 1e-12
    # the exact route reads the same fidelity off the density matrix, no sampling
    eng.upload([us.noise_model.instantiate(cdn["circuit"])[0]])
    assert abs(qb.fidelity_witness_exact(eng, g, precision=64).cpu().numpy()[0] - want2) < 1e-10
    # a level that needs settings which were not measured says so
    toth = qb.TothStabilizer(g, n_qubits=n).build()
    wt = qb.GenuineWitness(n, toth, ss.sample(toth, cd["circuit"]))
    assert abs(wt.estimate(g, noise_robust=0).value + 1.0) < 1e-12      # level 0 sums what it is given: the Toth settings
    with pytest.raises(KeyError):
        wt.estimate(g, noise_robust=2)
    tri = _graph(3, [(0, 1), (1, 2), (2, 0)])
    with pytest.raises(ValueError):
        qb.GenuineWitness(n, stab, counts).estimate(tri, noise_robust=1)
