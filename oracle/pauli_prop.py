"""Exact stabilizer expectation values of noisy CLIFFORD circuits at any qubit count, by Heisenberg-picture Pauli
propagation.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Why it exists: the reference estimates <S_k> from the counts of `graph_circuit + basis change + measure_all`
(stabilizers.py:183-215, witnesses.py:40-67).  Its graph circuits are Clifford (h, cx / cz: graph_circuit.py:81-130)
and its noise is a list of single-qubit channels (amplitude damping pairs, samplers.py:404-451; depolarizing /
thermal relaxation, :577-727), so the EXACT value the counts converge to can be computed without ever holding a
2^n state: the measured observable D = prod_{q in mask} Z_q (witnesses.py:40-57: D[x] = (-1)^popcount(x & mask)) is
pulled back through the circuit, gate by gate in reverse,

    unitary U :  P -> U^dag P U           (a signed Pauli again: Aaronson-Gottesman update rules)
    channel E :  P -> E^dag(P) = sum_j K_j^dag P K_j = sum_Q M[Q, P] Q      (4x4 real transfer matrix; terms split)

and <D> = sum over the surviving terms with no X / Y factor of their coefficients (<0..0| P |0..0>).  Amplitude damping
only splits Z -> (1 - g) Z + g I, so a circuit with K damping channels yields at most 2^K terms (duplicates are merged):
milliseconds for the benchmark's 20- and 28-qubit workloads.  This is the at-size oracle for the noisy configurations
the dense oracles (oracle/sim.py: 4^n numpy density matrix) cannot reach; tests/test_oracle_pauli_prop.py pins it
against that dense oracle at small n.

Gate list format: oracle/sim.py's.  Supported: id h x y z s sdg cx cz swap, `unitary` that is the identity (the
reference's labelled markers), kraus / pauli / reset on ONE qubit, pauli on two qubits.  Anything else raises.
"""
import numpy as np

from . import sim

MAX_TERMS = 1 << 22
_I2 = np.eye(2, dtype=np.complex128)
_PAULI = [_I2, np.array([[0, 1], [1, 0]], dtype=np.complex128), np.array([[0, -1j], [1j, 0]], dtype=np.complex128),
          np.array([[1, 0], [0, -1]], dtype=np.complex128)]          # index = x + 2 z with Y at (x, z) = (1, 1) -> see _LETTER
# letter code c = x | (z << 1):  0 = I, 1 = X, 2 = Z, 3 = Y
_LETTER = {0: _PAULI[0], 1: _PAULI[1], 2: _PAULI[3], 3: _PAULI[2]}


def adjoint_transfer_matrix(kraus):
    """M[Q, P] = tr(Q E^dag(P)) / 2 with E^dag(P) = sum_j K_j^dag P K_j, letters coded c = x | z << 1."""
    m = np.zeros((4, 4))
    for p in range(4):
        img = sum(k.conj().T @ _LETTER[p] @ k for k in kraus)
        for q in range(4):
            v = np.trace(_LETTER[q] @ img) / 2.0
            if abs(v.imag) > 1e-12:
                raise ValueError("channel is not Hermiticity preserving")
            m[q, p] = v.real
    return m


class _Terms:
    """sum_t coef[t] * (tensor product of letters given by bit t of the x / z masks)."""

    def __init__(self, x, z, coef):
        self.x, self.z, self.coef = x, z, coef

    def merge(self):
        key = np.stack([self.x, self.z], axis=1)
        uniq, inv = np.unique(key, axis=0, return_inverse=True)
        coef = np.zeros(len(uniq))
        np.add.at(coef, inv.reshape(-1), self.coef)
        keep = np.abs(coef) > 0
        self.x, self.z, self.coef = uniq[keep, 0].copy(), uniq[keep, 1].copy(), coef[keep]

    def bit(self, arr, q):
        return ((arr >> np.uint64(q)) & np.uint64(1)).astype(bool)

    def flip(self, arr, q, where):
        arr ^= where.astype(np.uint64) << np.uint64(q)

    # ---- U^dag P U for the gate U (Aaronson-Gottesman rules; H, CX, CZ, SWAP, X, Y, Z are self-inverse)
    def h(self, q):
        xq, zq = self.bit(self.x, q), self.bit(self.z, q)
        self.coef[xq & zq] *= -1.0                                  # H Y H = -Y
        sw = xq ^ zq
        self.flip(self.x, q, sw)
        self.flip(self.z, q, sw)

    def s(self, q):          # S^dag X S = -Y, S^dag Y S = X, Z fixed
        xq, zq = self.bit(self.x, q), self.bit(self.z, q)
        self.coef[xq & ~zq] *= -1.0
        self.flip(self.z, q, xq)

    def sdg(self, q):        # S X S^dag = Y, S Y S^dag = -X
        xq, zq = self.bit(self.x, q), self.bit(self.z, q)
        self.coef[xq & zq] *= -1.0
        self.flip(self.z, q, xq)

    def cx(self, a, b):
        xa, za, xb, zb = self.bit(self.x, a), self.bit(self.z, a), self.bit(self.x, b), self.bit(self.z, b)
        self.coef[xa & zb & ~(xb ^ za)] *= -1.0
        self.flip(self.x, b, xa)
        self.flip(self.z, a, zb)

    def pauli_gate(self, name, q):
        xq, zq = self.bit(self.x, q), self.bit(self.z, q)
        anti = {"x": zq, "z": xq, "y": xq ^ zq}[name]
        self.coef[anti] *= -1.0

    def swap(self, a, b):
        for arr in (self.x, self.z):
            d = self.bit(arr, a) ^ self.bit(arr, b)
            self.flip(arr, a, d)
            self.flip(arr, b, d)

    def channel1(self, q, m):
        """Every term's letter P on qubit q is replaced by sum_Q m[Q, P] Q."""
        code = self.bit(self.x, q).astype(np.int64) | (self.bit(self.z, q).astype(np.int64) << 1)
        clear = ~(np.uint64(1) << np.uint64(q))
        xs, zs, cs = [], [], []
        for newc in range(4):
            w = m[newc, code]
            keep = w != 0.0
            if not keep.any():
                continue
            xs.append((self.x[keep] & clear) | (np.uint64(newc & 1) << np.uint64(q)))
            zs.append((self.z[keep] & clear) | (np.uint64(newc >> 1) << np.uint64(q)))
            cs.append(self.coef[keep] * w[keep])
        if not xs:       # every term was annihilated (e.g. reset on a qubit carrying X)
            self.x, self.z, self.coef = np.zeros(0, dtype=np.uint64), np.zeros(0, dtype=np.uint64), np.zeros(0)
            return
        self.x, self.z, self.coef = np.concatenate(xs), np.concatenate(zs), np.concatenate(cs)
        self.merge()
        if len(self.coef) > MAX_TERMS:
            raise MemoryError(f"pauli_prop: more than {MAX_TERMS} Pauli terms (too many channels split this observable)")

    def pauli_mixture(self, qubits, probs):
        """sum_j p_j P_j (.) P_j: diagonal in the Pauli basis, factor sum_j p_j (+-1 by commutation)."""
        probs = np.asarray(probs, dtype=np.float64)
        # Pauli index as in include/qcdsim.h: 0 I, 1 X, 2 Y, 3 Z per qubit, index = P(q0) + 4 P(q1)
        xz = {0: (0, 0), 1: (1, 0), 2: (1, 1), 3: (0, 1)}
        fac = np.zeros(len(self.coef))
        for j, pj in enumerate(probs):
            if pj == 0.0:
                continue
            anti = np.zeros(len(self.coef), dtype=bool)
            for h, q in enumerate(qubits):
                px, pz = xz[(j >> (2 * h)) & 3]
                # two letters anticommute iff their symplectic product is 1
                anti ^= (self.bit(self.x, q) & bool(pz)) ^ (self.bit(self.z, q) & bool(px))
            fac += np.where(anti, -pj, pj)
        self.coef = self.coef * fac


def expectation(ops, n, mask):
    """Exact <prod_{q in mask} Z_q> on the output of the noisy Clifford gate list `ops` started in |0...0>."""
    if n > 63:
        raise ValueError("at most 63 qubits")
    t = _Terms(np.zeros(1, dtype=np.uint64), np.array([mask], dtype=np.uint64), np.ones(1))
    for op in reversed(list(ops)):
        name = op[0]
        if name in ("barrier", "measure_all", "id"):
            continue
        qs = op[1]
        if name == "h":
            t.h(qs[0])
        elif name == "s":
            t.s(qs[0])
        elif name == "sdg":
            t.sdg(qs[0])
        elif name in ("x", "y", "z"):
            t.pauli_gate(name, qs[0])
        elif name == "cx":
            t.cx(qs[0], qs[1])
        elif name == "cz":
            t.h(qs[1])
            t.cx(qs[0], qs[1])
            t.h(qs[1])
        elif name == "swap":
            t.swap(qs[0], qs[1])
        elif name == "unitary":
            m = np.asarray(op[2])
            if not np.allclose(m, np.eye(m.shape[0]), atol=1e-14):
                raise ValueError("pauli_prop: only identity `unitary` markers are supported")
        elif name in ("kraus", "reset"):
            if len(qs) != 1:
                raise ValueError("pauli_prop: two-qubit Kraus channels are not supported")
            t.channel1(qs[0], adjoint_transfer_matrix(sim.op_kraus(op)))
        elif name == "pauli":
            t.pauli_mixture(qs, op[2])
        else:
            raise ValueError(f"pauli_prop: op {name!r} is not Clifford / not supported")
    return float(t.coef[t.x == 0].sum())


def toth_expectations(graph_ops, n, settings):
    """settings: list of (basis_change_ops, mask).  -> exact <S_k> per setting, measured the reference's way: graph
    circuit, then the setting's basis-change sub-circuit, then all qubits in the computational basis."""
    return [expectation(list(graph_ops) + list(sub), n, mask) for sub, mask in settings]
