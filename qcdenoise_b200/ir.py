"""Gate-list IR: the qiskit-free stand-in for the `QuantumCircuit` objects the reference builds
(graph_circuit.py:40-65, stabilizers.py:128-157) -- an ordered op list with qiskit's method names, so the
reference's builder code reads the same -- and its encoding into the C-ABI `qcd_op` array (include/qcdsim.h).

Op tuples:  (name, qubits[, param[, label]])
    ("h"|"x"|"y"|"z"|"s"|"sdg"|"t"|"tdg"|"id", (q,))      ("rx"|"ry"|"rz", (q,), theta)
    ("cx"|"cz"|"swap", (a, b))                             cx: a = control
    ("unitary", (q0[, q1]), matrix, label)                 little-endian: basis index = bit(q0) + 2 bit(q1)
    ("kraus", (q0[, q1]), [K_0, ...])   ("pauli", (q0[, q1]), probs)   ("reset", (q,))
    ("barrier",)  ("measure_all",)
"""
import numpy as np

from ._lib import OP_DTYPE, OPCODES

_FIXED_1Q = ("id", "h", "x", "y", "z", "s", "sdg", "t", "tdg")
_ROT = ("rx", "ry", "rz")
_FIXED_2Q = ("cx", "cz", "swap")


class Circuit:
    """Ordered op list on `n_qubits` qubits (qubit q <-> state index bit q <-> character n-1-q of a Counts key)."""

    def __init__(self, n_qubits, ops=None, name=None):
        self.n_qubits = int(n_qubits)
        self.num_qubits = self.n_qubits          # qiskit spelling
        self.ops = list(ops) if ops else []
        self.name = name

    # ---- qiskit-style builders
    def _q(self, q):
        q = int(q)
        if not 0 <= q < self.n_qubits:
            raise ValueError(f"qubit {q} out of range for a {self.n_qubits}-qubit circuit")
        return q

    def _add1(self, name, q):
        self.ops.append((name, (self._q(q),)))
        return self

    def id(self, q): return self._add1("id", q)
    def h(self, q): return self._add1("h", q)
    def x(self, q): return self._add1("x", q)
    def y(self, q): return self._add1("y", q)
    def z(self, q): return self._add1("z", q)
    def s(self, q): return self._add1("s", q)
    def sdg(self, q): return self._add1("sdg", q)
    def t(self, q): return self._add1("t", q)
    def tdg(self, q): return self._add1("tdg", q)

    def rx(self, theta, q):
        self.ops.append(("rx", (self._q(q),), float(theta)))
        return self

    def ry(self, theta, q):
        self.ops.append(("ry", (self._q(q),), float(theta)))
        return self

    def rz(self, theta, q):
        self.ops.append(("rz", (self._q(q),), float(theta)))
        return self

    def _add2(self, name, a, b):
        a, b = self._q(a), self._q(b)
        if a == b:
            raise ValueError("two-qubit gate on a single qubit")
        self.ops.append((name, (a, b)))
        return self

    def cx(self, control, target): return self._add2("cx", control, target)
    def cz(self, a, b): return self._add2("cz", a, b)
    def swap(self, a, b): return self._add2("swap", a, b)

    def unitary(self, matrix, qubits, label=None):
        qubits = tuple(self._q(q) for q in (qubits if np.iterable(qubits) else (qubits,)))
        m = np.asarray(matrix, dtype=np.complex128)
        if m.shape != (2 ** len(qubits),) * 2 or len(qubits) not in (1, 2):
            raise ValueError("unitary must be 2x2 on one qubit or 4x4 on two")
        self.ops.append(("unitary", qubits, m, label))
        return self

    def kraus(self, kraus_ops, qubits):
        qubits = tuple(self._q(q) for q in (qubits if np.iterable(qubits) else (qubits,)))
        ks = [np.asarray(k, dtype=np.complex128) for k in kraus_ops]
        if len(qubits) not in (1, 2) or any(k.shape != (2 ** len(qubits),) * 2 for k in ks) or not ks:
            raise ValueError("kraus needs 2x2 (one qubit) or 4x4 (two qubits) operators")
        self.ops.append(("kraus", qubits, ks))
        return self

    def pauli(self, probs, qubits):
        qubits = tuple(self._q(q) for q in (qubits if np.iterable(qubits) else (qubits,)))
        p = np.asarray(probs, dtype=np.float64)
        if p.shape != (4 ** len(qubits),) or len(qubits) not in (1, 2):
            raise ValueError("pauli needs 4 (one qubit) or 16 (two qubits) probabilities")
        self.ops.append(("pauli", qubits, p))
        return self

    def reset(self, q):
        self.ops.append(("reset", (self._q(q),)))
        return self

    def barrier(self, *_):
        self.ops.append(("barrier",))
        return self

    def measure_all(self):
        self.ops.append(("measure_all",))
        return self

    # ---- composition
    def copy(self, name=None):
        return Circuit(self.n_qubits, self.ops, name or self.name)

    def compose(self, other, inplace=False):
        tgt = self if inplace else self.copy()
        if other.n_qubits != self.n_qubits:
            raise ValueError("compose: qubit counts differ")
        tgt.ops.extend(other.ops)
        return tgt

    def append(self, other, qargs=None):
        """qiskit's `circuit.append(instruction, qargs)` for a whole-register sub-circuit (stabilizers.py:197)."""
        if qargs is not None and list(qargs) != list(range(self.n_qubits)):
            raise ValueError("append: only whole-register sub-circuits are supported")
        return self.compose(other, inplace=True)

    def to_gate(self, label=None):
        return self

    def count_ops(self):
        out = {}
        for op in self.ops:
            out[op[0]] = out.get(op[0], 0) + 1
        return out

    def __len__(self):
        return len(self.ops)

    def __repr__(self):
        return f"Circuit(n_qubits={self.n_qubits}, ops={len(self.ops)})"


def encode_programs(circuits):
    """list[Circuit | (n, ops)] -> (ops: np.ndarray[OP_DTYPE], offsets: uint32[B+1], params: float64[...], n)."""
    n = None
    recs, offsets, params = [], [0], []

    memo = {}      # id(payload object) -> param offset: a noise model attaches the SAME channel object many times

    def put(vals):
        off = len(params)
        params.extend(np.asarray(vals, dtype=np.float64).ravel().tolist())
        return off

    def put_once(obj, make):
        key = id(obj)
        if key not in memo:
            memo[key] = (put(make(obj)), obj)      # keep obj alive so the id stays unique
        return memo[key][0]

    def kraus_flat(ks):
        return np.concatenate([[float(len(ks))]] + [cflat(k) for k in ks])

    def cflat(m):
        m = np.asarray(m, dtype=np.complex128).reshape(-1)
        out = np.empty(2 * m.size)
        out[0::2] = m.real
        out[1::2] = m.imag
        return out

    for circ in circuits:
        cn, ops = (circ.n_qubits, circ.ops) if isinstance(circ, Circuit) else circ
        if n is None:
            n = cn
        elif n != cn:
            raise ValueError("all circuits of one upload must have the same qubit count")
        for op in ops:
            name = op[0]
            if name in ("barrier", "measure_all"):
                continue
            qs = op[1]
            if name in _FIXED_1Q:
                recs.append((OPCODES[name], qs[0], 0xFF, 0))
            elif name in _ROT:
                recs.append((OPCODES[name], qs[0], 0xFF, put([op[2]])))
            elif name in _FIXED_2Q:
                recs.append((OPCODES[name], qs[0], qs[1], 0))
            elif name == "unitary":
                if len(qs) == 1:
                    recs.append((OPCODES["u1q"], qs[0], 0xFF, put_once(op[2], cflat)))
                else:
                    recs.append((OPCODES["u2q"], qs[0], qs[1], put_once(op[2], cflat)))
            elif name == "kraus":
                recs.append((OPCODES["kraus1q" if len(qs) == 1 else "kraus2q"], qs[0],
                             qs[1] if len(qs) > 1 else 0xFF, put_once(op[2], kraus_flat)))
            elif name == "pauli":
                recs.append((OPCODES["pauli1q" if len(qs) == 1 else "pauli2q"], qs[0],
                             qs[1] if len(qs) > 1 else 0xFF, put_once(op[2], lambda p: p)))
            elif name == "reset":
                recs.append((OPCODES["reset"], qs[0], 0xFF, 0))
            else:
                raise ValueError(f"unknown op {name!r}")
        offsets.append(len(recs))
    if n is None:
        raise ValueError("no circuits")
    arr = np.array(recs, dtype=OP_DTYPE) if recs else np.zeros(0, dtype=OP_DTYPE)
    return arr, np.asarray(offsets, dtype=np.uint32), np.asarray(params, dtype=np.float64), n
