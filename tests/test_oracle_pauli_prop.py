"""oracle/pauli_prop.py (exact <S> of noisy Clifford circuits at any size, Heisenberg picture) pinned against the dense
numpy density-matrix oracle (oracle/sim.py) at sizes the latter reaches."""
import numpy as np
import pytest

from oracle import noise, pauli_prop, ref_glue, sim
from tests import circuits


def _dense_expectation(ops, n, mask):
    p = sim.density_matrix_probs(ops, n)
    x = np.arange(2 ** n)
    par = np.zeros(2 ** n, dtype=np.int64)
    for q in range(n):
        if (mask >> q) & 1:
            par ^= (x >> q) & 1
    return float((p * (1 - 2 * par)).sum())


def _toth_settings(n, edges):
    nbrs = {v: [b if a == v else a for a, b in edges if v in (a, b)] for v in range(n)}
    if n <= 10:
        names = ref_glue.get_generators(n, range(n), nbrs) + ["I" * n]
    else:       # stabilizers.py:115-121 without its n <= 10 guard (Q8): X on the vertex, Z on its neighbours
        names = ["".join("X" if q == v else ("Z" if q in nbrs[v] else "I") for q in range(n))[::-1] for v in range(n)]
        names.append("I" * n)
    out = []
    for nm in names:
        sub, mask = [], 0
        for q, ch in enumerate(nm[::-1]):
            if ch == "X":
                sub.append(("h", (q,)))
            elif ch == "Y":
                sub += [("sdg", (q,)), ("h", (q,))]
            if ch != "I":
                mask |= 1 << q
        out.append((sub, mask))
    return out


@pytest.mark.parametrize("n,gi", [(4, 0), (5, 2), (7, 1), (8, 3)])
def test_toth_expectations_match_dense_oracle(n, gi):
    rng = np.random.default_rng(n * 10 + gi)
    edges = circuits.ref_graph(n, gi)
    ops = circuits.graph_circuit_amp_damp(n, edges, rng, max_prob=0.3, all_sites=True)
    settings = _toth_settings(n, edges)
    got = pauli_prop.toth_expectations(ops, n, settings)
    for (sub, mask), g in zip(settings, got):
        assert g == pytest.approx(_dense_expectation(ops + sub, n, mask), abs=1e-12)
    assert got[-1] == pytest.approx(1.0, abs=1e-14)                   # identity setting


def test_ideal_graph_state_has_unit_stabilizers():
    n = 20
    edges = circuits.ref_graph(20, 0)
    ops, _ = ref_glue.cx_gate_circuit(n, edges)
    for v in pauli_prop.toth_expectations(ops, n, _toth_settings(n, edges)):
        assert v == 1.0


def test_device_noise_and_random_clifford_match_dense_oracle():
    rng = np.random.default_rng(3)
    n = 5
    ops, _ = circuits.device_noise_circuit(n, circuits.ref_graph(5, 1), rng, gate_error=0.2, gate_len_ns=(500, 5000))
    for mask in (0b00001, 0b10110, 0b11111):
        assert pauli_prop.expectation(ops, n, mask) == pytest.approx(_dense_expectation(ops, n, mask), abs=1e-12)
    one = ["h", "x", "y", "z", "s", "sdg", "id"]
    for seed in range(6):
        rng = np.random.default_rng(100 + seed)
        n = 6
        ops = []
        for _ in range(60):
            r = rng.integers(0, 9)
            q = int(rng.integers(0, n))
            q2 = int((q + 1 + rng.integers(0, n - 1)) % n)
            if r < 3:
                ops.append((one[rng.integers(0, len(one))], (q,)))
            elif r < 6:
                ops.append((["cx", "cz", "swap"][rng.integers(0, 3)], (q, q2)))
            elif r == 6:
                ops.append(("kraus", (q,), noise.amplitude_damping_kraus(float(rng.uniform(0, 0.4)))))
            elif r == 7:
                if rng.integers(0, 2):
                    ops.append(("pauli", (q,), noise.depolarizing_pauli_probs(float(rng.uniform(0, 0.3)), 1)))
                else:
                    pr = rng.dirichlet(np.ones(16))
                    ops.append(("pauli", (q, q2), pr))
            else:
                if rng.integers(0, 2):
                    ops.append(("reset", (q,)))
                else:
                    t1 = float(rng.uniform(20, 100))
                    ops.append(("kraus", (q,), noise.thermal_relaxation_kraus(t1, float(rng.uniform(0.2, 2.0)) * t1,
                                                                              float(rng.uniform(0.1, 30)))))
        for mask in (int(rng.integers(1, 2 ** n)), 2 ** n - 1):
            assert pauli_prop.expectation(ops, n, mask) == pytest.approx(_dense_expectation(ops, n, mask), abs=1e-12)


def test_non_clifford_gate_is_rejected():
    with pytest.raises(ValueError):
        pauli_prop.expectation([("ry", (0,), 0.3)], 2, 1)
