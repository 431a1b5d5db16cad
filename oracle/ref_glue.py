"""Restatement of the reference's OWN Python on the hot path (everything around the Aer call).
TEST INFRASTRUCTURE (see oracle/__init__.py).  Each function cites the reference lines it follows; written
the slow, literal way the reference does it (np.kron chains, string keys) so the product's bitmask / integer
implementations are checked against an independent formulation.
"""
import math
from collections import OrderedDict, namedtuple

import numpy as np

witness_output = namedtuple("Witness", ["W_ij", "value", "variance"])      # witnesses.py:21-22


# ----------------------------------------------------------------------------- graph_circuit.py
def graph_iter(edges, directed_adjacency=None):
    """graph_circuit.py:55-60: undirected -> graph.edges() order."""
    return [tuple(e) for e in edges]


def cx_gate_circuit(n, edges, coins=None):
    """CXGateCircuit.build (graph_circuit.py:81-99): H^n; per edge H(ngbr) CX(node, ngbr) [unitary I4
    labelled 'unitary_{node}_{ngbr}' when the coin is 1] H(ngbr).  coins[i] in {0,1} replaces
    `bool(np.random.choice(2)) and self.stochastic` (None = not stochastic)."""
    ops = [("h", (q,)) for q in range(n)]
    labels = []
    for i, (node, ngbr) in enumerate(edges):
        ops.append(("h", (ngbr,)))
        ops.append(("cx", (node, ngbr)))
        if coins is not None and coins[i]:
            label = f"unitary_{node}_{ngbr}"
            labels.append(label)
            ops.append(("unitary", (node, ngbr), np.identity(4), label))
        ops.append(("h", (ngbr,)))
    return ops, labels


def cz_gate_circuit(n, edges, coins=None):
    """CZGateCircuit.build (graph_circuit.py:114-130): H^n; per edge H(node) CZ(node, ngbr) [U] H(node)."""
    ops = [("h", (q,)) for q in range(n)]
    labels = []
    for i, (node, ngbr) in enumerate(edges):
        ops.append(("h", (node,)))
        ops.append(("cz", (node, ngbr)))
        if coins is not None and coins[i]:
            label = f"unitary_{node}_{ngbr}"
            labels.append(label)
            ops.append(("unitary", (node, ngbr), np.identity(4), label))
        ops.append(("h", (node,)))
    return ops, labels


def cphase_gate_circuit(n, edges, coins=None, Lambda=np.pi, Theta=np.pi / 2):
    """CPhaseGateCircuit.build (graph_circuit.py:151-173): per edge Ry(L/2) CX [U] Ry(-T/2) CX [U] Ry(L/2);
    coins has TWO entries per edge."""
    ops, labels = [], []
    for i, (node, ngbr) in enumerate(edges):
        ops.append(("ry", (ngbr,), Lambda / 2))
        ops.append(("cx", (node, ngbr)))
        if coins is not None and coins[2 * i]:
            labels.append(f"unitary_{node}_{ngbr}")
            ops.append(("unitary", (node, ngbr), np.identity(4), labels[-1]))
        ops.append(("ry", (ngbr,), -Theta / 2))
        ops.append(("cx", (node, ngbr)))
        if coins is not None and coins[2 * i + 1]:
            labels.append(f"unitary_{node}_{ngbr}")
            ops.append(("unitary", (node, ngbr), np.identity(4), labels[-1]))
        ops.append(("ry", (ngbr,), Lambda / 2))
    return ops, labels


def attach_unitary_noise(ops, kraus_by_label):
    """What Aer does with `noise_model.add_all_qubit_quantum_error(error, label)` (samplers.py:404-412):
    after every instruction whose label matches, apply the error on that instruction's qubits.
    kraus_by_label[label] = (kraus_on_first_qubit, kraus_on_second_qubit)."""
    out = []
    for op in ops:
        out.append(op)
        if op[0] == "unitary" and len(op) > 3 and op[3] in kraus_by_label:
            kn, kg = kraus_by_label[op[3]]
            out.append(("kraus", (op[1][0],), kn))
            out.append(("kraus", (op[1][1],), kg))
    return out


# ----------------------------------------------------------------------------- samplers.py
def generate_binary_strings(n):
    """samplers.py:37-68: all n-bit strings in lexicographic (= numeric, MSB first) order."""
    out = []

    def rec(bits, i):
        if i == n:
            out.append("".join(bits))
            return
        bits[i] = "0"
        rec(bits, i + 1)
        bits[i] = "1"
        rec(bits, i + 1)

    rec([None] * n, 0)
    return out


def convert_to_prob_vector(prob_dict, n, n_shots):
    """samplers.py:195-202."""
    full = {key: 0 for key in generate_binary_strings(n)}
    for key, itm in prob_dict.items():
        full[key] = itm / n_shots
    return np.array([val for (_, val) in full.items()])


# ----------------------------------------------------------------------------- stabilizers.py
def get_generators(n, nodes, neighbors):
    """StabilizerCircuit._get_generators (stabilizers.py:104-122). neighbors[idx] = iterable of neighbours."""
    if n > 10:
        raise NotImplementedError            # stabilizers.py:111-113 (Q8)
    gens = []
    for idx in nodes:
        temp = list("I" * n)
        temp[idx] = "X"
        for jdx in neighbors[idx]:
            temp[jdx] = "Z"
        gens.append("".join(temp)[::-1])
    return gens


def toth_stabilizers(generators, n):
    """TothStabilizer.find_stabilizers (stabilizers.py:161-165)."""
    return ["+" + g for g in generators] + ["+" + "I" * n]


_PAULI_MUL = {  # (a, b) -> (phase, c) with a*b = phase*c ; phase as power of i
    ("X", "Y"): (1, "Z"), ("Y", "X"): (3, "Z"), ("Y", "Z"): (1, "X"), ("Z", "Y"): (3, "X"),
    ("Z", "X"): (1, "Y"), ("X", "Z"): (3, "Y"),
}


def sigma_prod(op_str):
    """stabilizers.py:39-79: product of single-qubit Paulis given as a string; the loop pops from the END,
    so the product is formed as  ... * (op[-2] * op[-1]) with `mat2 * mat1` = earlier * later."""
    lst = list(op_str)
    coef = 1 + 0j
    while len(lst) > 1:
        m1 = lst.pop()
        m2 = lst.pop()
        if m1 == m2:
            m3, c = "I", 1
        elif "I" not in (m1, m2):
            ph, m3 = _PAULI_MUL[(m2, m1)]
            c = {1: 1j, 3: -1j}[ph]
        else:
            m3, c = [x for x in (m1, m2) if x != "I"][0], 1
        coef *= c
        lst.append(m3)
    return coef, lst[0]


def full_stabilizer_group(generators, n):
    """Legacy 'Jung' enumeration (circuit_constructors.py:554-590): for every n-bit string (np.binary_repr
    order) multiply the selected generators column by column with sigma_prod."""
    labels = []
    for x in range(2 ** n):
        coefs = [int(c) for c in np.binary_repr(x, n)]
        rows = [list(generators[j]) if coefs[j] else list("I" * n) for j in range(n)]
        cf, lb = [], []
        for k in range(n):
            c, l = sigma_prod("".join(r[k] for r in rows))
            cf.append(c)
            lb.append(l)
        val = np.prod(cf)
        assert abs(val.imag) < 1e-12
        labels.append(("+" if val.real > 0 else "-") + "".join(lb))
    return labels


def stabilizer_basis_change(name):
    """StabilizerCircuit.build (stabilizers.py:136-156): ops of the sub-circuit for a (sign-less) string."""
    ops = []
    for idx, ch in enumerate(list(name)[::-1]):
        if ch == "X":
            ops.append(("h", (idx,)))
        elif ch == "Y":
            ops.append(("sdg", (idx,)))
            ops.append(("h", (idx,)))
        else:
            ops.append(("id", (idx,)))
    return ops


# ----------------------------------------------------------------------------- witnesses.py
def construct_diagonal(name):
    """Witness._construct_diagonals (witnesses.py:40-57): kron chain, sign forced '+'."""
    a = [1.0, 1.0]
    b = [1.0, -1.0]
    ops = list(name)
    pop = ops.pop()
    D = a if pop == "I" else b
    while ops:
        pop = ops.pop()
        d = a if pop == "I" else b
        D = np.kron(d, D)
    return np.multiply(1.0, D)


def expectation_value(counts, diagonal):
    """qiskit-ignis 0.5 `mitigation.expectation_value(counts, diagonal)` (witnesses.py:65): formula pinned
    by golden G1:  mean = sum_x p_x d_x ;  std = sqrt((sum_x p_x d_x^2 - mean^2) / shots)."""
    shots = sum(counts.values())
    keys = list(counts.keys())
    probs = np.array([counts[k] for k in keys], dtype=float) / shots
    coeffs = np.array([diagonal[int(k, 2)] for k in keys], dtype=float)
    mean = float(probs.dot(coeffs))
    sq = float(probs.dot(coeffs ** 2))
    var = (sq - mean * mean) / shots
    return mean, math.sqrt(var) if var > 0 else 0.0


def stabilizer_measurements(names, counts_list):
    """Witness._compute_expval (witnesses.py:59-67)."""
    out = OrderedDict()
    for idx, name in enumerate(names):
        out[name] = expectation_value(counts_list[idx], construct_diagonal(name))
    return out


def genuine_witness(meas, n_nodes, n):
    """GenuineWitness.estimate (witnesses.py:113-144); `variance` is sqrt(sum std^2) (uncertainties'
    linear propagation of independent terms), as the reference returns it."""
    id_key = "I" * n
    id_val, _ = meas[id_key]
    vals = [m for m, _ in meas.values()]
    stds = [s for _, s in meas.values()]
    w = (n_nodes - 1) * id_val - (sum(vals) - id_val)
    # witnesses.py:136-144: `unumpy.uarray(witness_terms, var_vals).sum()` -- wit_coef multiplies the nominal values
    # only (witness_terms = wit_coef * meas_vals); the std-devs are var_vals UNSCALED, and uncertainties' linear
    # propagation of a plain sum of independent terms gives sqrt(sum std_k^2), whatever the order of the dict
    sigma = math.sqrt(sum(s * s for s in stds))
    return witness_output(W_ij=None, value=w, variance=sigma)


def bisep_witnesses(meas, edges):
    """BiSeparableWitness.estimate (witnesses.py:74-109)."""
    res = {}
    for idx, (e0, e1) in enumerate(edges):
        gi = [k for k in meas if k[::-1][e0] == "X"][0]
        gj = [k for k in meas if k[::-1][e1] == "X"][0]
        val = 1 - meas[gi][0] - meas[gj][0]
        sigma = math.sqrt(meas[gi][1] ** 2 + meas[gj][1] ** 2)
        res[idx] = witness_output(W_ij=(e0, e1), value=val, variance=sigma)
    return res


# ------------------------------------------------------------------ full-group ("noise-robust") witnesses
# The reference leaves these levels unimplemented ("WIP", circuit_constructors.py:677-680; GenuineWitness.estimate
# asserts noise_robust == 0, witnesses.py:114-116) -- PARITY UNPINNED: there is no reference output to hold them to.
# They are restated here from the paper the legacy docstrings cite (Toth & Guehne, PRA 72, 022340) with DENSE matrices,
# independently of the product's bit arithmetic.
_PAULI = {"I": np.eye(2, dtype=complex), "X": np.array([[0, 1], [1, 0]], dtype=complex),
          "Y": np.array([[0, -1j], [1j, 0]]), "Z": np.diag([1.0 + 0j, -1.0])}


def pauli_matrix(label):
    """Dense matrix of a Pauli string; character j acts on qubit n-1-j (the reference's reversed labels)."""
    m = np.ones((1, 1), dtype=complex)
    for ch in label:
        m = np.kron(m, _PAULI[ch])
    return m


def group_expectations_dense(rho, signed_group):
    """{label: Tr(rho * sign * P)} for the signed elements ('+XZII', '-ZYXY', ...), rho [2^n, 2^n]."""
    return OrderedDict((s[1:], float(np.real(np.trace(rho @ pauli_matrix(s[1:]))) * (1 if s[0] == "+" else -1)))
                       for s in signed_group)


def graph_state_projector(signed_group, n):
    return sum((1 if s[0] == "+" else -1) * pauli_matrix(s[1:]) for s in signed_group) / 2 ** n


def projector_witness_dense(rho, signed_group, n):
    """<W> for W = 1/2 - |G><G|."""
    return float(np.real(np.trace(rho @ (np.eye(2 ** n) / 2 - graph_state_projector(signed_group, n)))))


def two_colour_witness_dense(rho, generators, colours, n):
    """<W> for W = 3 - 2 [prod_{k in A} (1 + g_k)/2 + prod_{k in B} (1 + g_k)/2]; generators[k] belongs to vertex k,
    colours[k] in {0, 1}."""
    w = 3 * np.eye(2 ** n, dtype=complex)
    for c in (0, 1):
        p = np.eye(2 ** n, dtype=complex)
        for k, g in enumerate(generators):
            if colours[k] == c:
                p = p @ (np.eye(2 ** n) + pauli_matrix(g)) / 2
        w = w - 2 * p
    return float(np.real(np.trace(rho @ w)))


# ------------------------------------------------------------------ adjacency tensor
def adjacency_tensor(ops, encoding, tensor_shape=(32, 32, 32), fixed_size=True, directed=False):
    """generate_adjacency_tensor (samplers.py:119-192) over a gate list instead of a qiskit DAG: `dag.gate_nodes()` yields
    the gates in a topological order, and the plane a gate lands in depends only on the earlier gates that hit the SAME
    [q, q] / [q1, q2] slot, whose relative order every topological order keeps.  Literal restatement, quirks included: a
    slot counts as free while it holds 0, so a gate whose code is 0 never occupies one; an undirected two-qubit gate tests
    [q1, q2] only and writes both [q1, q2] and [q2, q1]; gates beyond the last plane are dropped."""
    adj = np.zeros(tensor_shape)
    for op in ops:
        name, qubits = op[0], op[1]
        if name in ("barrier", "measure_all", "kraus", "pauli", "reset"):      # not gate nodes / not in the circuit
            continue
        if len(qubits) == 1:
            row = col = qubits[0]
        else:
            row, col = qubits
        plane = 0
        while plane < adj.shape[0]:
            if adj[plane, row, col] == 0:
                adj[plane, row, col] = encoding[name]
                if len(qubits) == 2 and not directed:
                    adj[plane, col, row] = encoding[name]
                break
            plane += 1
    if not fixed_size:
        zero = np.zeros_like(adj[0])
        adj = np.array([pl for pl in adj if np.any(pl != zero)])
    return adj
