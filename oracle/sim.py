"""numpy complex128 restatement of the simulation the reference delegates to qiskit-aer through
`qk.execute(circuit, AerSimulator, noise_model, shots)` (samplers.py:308-326).  TEST INFRASTRUCTURE
(see oracle/__init__.py).  Gate-at-a-time, exactly the structure of Aer's `statevector` (one Kraus trajectory
per shot, `apply_kraus`: accumulate ||K_j psi||^2 in operator order until it exceeds r) and `density_matrix`
(vectorised rho, unitary as U (x) conj(U), channel as superoperator) methods.

Gate list format (the same op vocabulary as include/qcdsim.h):
    ("h"|"x"|"y"|"z"|"s"|"sdg"|"t"|"tdg"|"id", (q,))
    ("rx"|"ry"|"rz", (q,), theta)          ("cx"|"cz"|"swap", (a, b))   # cx: a = control, b = target
    ("unitary", (q0[, q1]), matrix)        # little-endian: basis index = bit(q0) + 2 bit(q1)
    ("kraus", (q0[, q1]), [K_0, K_1, ...]) ("pauli", (q0[, q1]), probs)  ("reset", (q,))
Bit order: state index bit q <-> qubit q; histogram bin x <-> qiskit key format(x, '0{n}b').
"""
import numpy as np

from . import philox

SQ2 = 1.0 / np.sqrt(2.0)
_FIXED = {
    "id": np.eye(2),
    "h": np.array([[SQ2, SQ2], [SQ2, -SQ2]]),
    "x": np.array([[0, 1], [1, 0]]),
    "y": np.array([[0, -1j], [1j, 0]]),
    "z": np.array([[1, 0], [0, -1]]),
    "s": np.array([[1, 0], [0, 1j]]),
    "sdg": np.array([[1, 0], [0, -1j]]),
    "t": np.array([[1, 0], [0, np.exp(1j * np.pi / 4)]]),
    "tdg": np.array([[1, 0], [0, np.exp(-1j * np.pi / 4)]]),
    # two-qubit, basis index = bit(first) + 2*bit(second)
    "cx": np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0]]),   # first = control
    "cz": np.diag([1, 1, 1, -1]),
    "swap": np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]]),
}
STOCHASTIC = ("kraus", "pauli", "reset")
_RESET_KRAUS = [np.array([[1, 0], [0, 0]], dtype=np.complex128), np.array([[0, 1], [0, 0]], dtype=np.complex128)]


def gate_matrix(op):
    name = op[0]
    if name in _FIXED:
        return np.asarray(_FIXED[name], dtype=np.complex128)
    if name == "unitary":
        return np.asarray(op[2], dtype=np.complex128)
    th = op[2]
    c, s = np.cos(th / 2), np.sin(th / 2)
    if name == "rx":
        return np.array([[c, -1j * s], [-1j * s, c]], dtype=np.complex128)
    if name == "ry":
        return np.array([[c, -s], [s, c]], dtype=np.complex128)
    if name == "rz":
        return np.array([[np.exp(-1j * th / 2), 0], [0, np.exp(1j * th / 2)]], dtype=np.complex128)
    raise ValueError(f"unknown gate {name}")


def op_kraus(op):
    """Kraus list of a stochastic op."""
    from .noise import pauli_kraus
    if op[0] == "kraus":
        return [np.asarray(k, dtype=np.complex128) for k in op[2]]
    if op[0] == "pauli":
        return pauli_kraus(op[2])
    if op[0] == "reset":
        return _RESET_KRAUS
    raise ValueError(op[0])


def apply_matrix(vec, mat, qubits, nbits):
    """vec' = (mat on `qubits`) vec, vec of length 2**nbits, index bit q <-> qubit q."""
    k = len(qubits)
    axes = [nbits - 1 - q for q in reversed(qubits)]          # most-significant operator bit first
    t = np.tensordot(mat.reshape([2] * (2 * k)), vec.reshape([2] * nbits), axes=(list(range(k, 2 * k)), axes))
    return np.moveaxis(t, list(range(k)), axes).reshape(-1)


def is_noise(op):
    return op[0] in STOCHASTIC


def ideal_statevector(ops, n):
    psi = np.zeros(2 ** n, dtype=np.complex128)
    psi[0] = 1.0
    for op in ops:
        if op[0] in ("measure_all", "barrier") or is_noise(op):
            continue
        psi = apply_matrix(psi, gate_matrix(op), op[1], n)
    return psi


def density_matrix(ops, n):
    """Evolve vec(rho) (index = x + 2**n * y  <->  rho[x, y]) through gates and channels."""
    rho = np.zeros(4 ** n, dtype=np.complex128)
    rho[0] = 1.0
    for op in ops:
        if op[0] in ("measure_all", "barrier"):
            continue
        col = tuple(q + n for q in op[1])
        if is_noise(op):
            acc = np.zeros_like(rho)
            for k in op_kraus(op):
                acc += apply_matrix(apply_matrix(rho, k, op[1], 2 * n), np.conj(k), col, 2 * n)
            rho = acc
        else:
            u = gate_matrix(op)
            rho = apply_matrix(apply_matrix(rho, u, op[1], 2 * n), np.conj(u), col, 2 * n)
    return rho


def density_matrix_probs(ops, n):
    rho = density_matrix(ops, n)
    return np.real(rho.reshape(2 ** n, 2 ** n).diagonal()).copy()


def apply_readout_to_probs(probs, n, confusion):
    """confusion[q] = [[P(0|0), P(1|0)], [P(0|1), P(1|1)]] (row = true, col = measured)."""
    p = probs.astype(np.float64)
    for q in range(n):
        p = apply_matrix(p, np.asarray(confusion[q], dtype=np.float64).T, (q,), n)
    return p


def _select(cum, u):
    """first j with cum[j] > u*cum[-1] (Aer apply_kraus / sample_measure inverse CDF)."""
    j = int(np.searchsorted(cum, u * cum[-1], side="right"))
    return min(j, len(cum) - 1)


def run_trajectories(ops, n, shots, seed, circuit=0, traj_begin=0, readout=None, dedup=True):
    """One Kraus trajectory per shot (Aer statevector method), Philox streams keyed
    (seed; trajectory, site ordinal, circuit).  Returns outcomes[shots] (uint64 basis indices, after the
    optional classical readout error).  dedup=True shares the state between trajectories with the same
    branch history (identical results, fewer sweeps); dedup=False is the literal per-shot loop."""
    ops = [op for op in ops if op[0] not in ("measure_all", "barrier")]
    trajs = np.arange(traj_begin, traj_begin + shots, dtype=np.uint64)
    out = np.zeros(shots, dtype=np.uint64)
    site_of = {}
    s = 0
    for i, op in enumerate(ops):
        if is_noise(op):
            site_of[i] = s
            s += 1

    def finish(psi, idx):
        cum = np.cumsum(np.abs(psi) ** 2)
        u = philox.uniform53(seed, circuit, trajs[idx], philox.STREAM_MEASURE)
        x = np.searchsorted(cum, u * cum[-1], side="right")
        out[idx] = np.minimum(x, len(cum) - 1)

    def run(i, psi, idx):
        while i < len(ops):
            op = ops[i]
            if not is_noise(op):
                psi = apply_matrix(psi, gate_matrix(op), op[1], n)
                i += 1
                continue
            cands = [apply_matrix(psi, k, op[1], n) for k in op_kraus(op)]
            if op[0] == "pauli":
                p = np.array([pp for pp in op[2] if pp > 0], dtype=np.float64)
            else:
                p = np.array([np.vdot(c, c).real for c in cands])
            cum = np.cumsum(p)
            u = philox.uniform53(seed, circuit, trajs[idx], site_of[i])
            br = np.minimum(np.searchsorted(cum, u * cum[-1], side="right"), len(p) - 1)
            for j in np.unique(br):
                sub = idx[br == j]
                nrm = np.sqrt(np.vdot(cands[j], cands[j]).real)
                run(i + 1, cands[j] / nrm, sub)
            return
        finish(psi, idx)

    psi0 = np.zeros(2 ** n, dtype=np.complex128)
    psi0[0] = 1.0
    if dedup:
        run(0, psi0, np.arange(shots))
    else:
        for t in range(shots):
            run(0, psi0, np.array([t]))
    if readout is not None:
        out = apply_readout_to_outcomes(out, n, readout, seed, circuit, trajs)
    return out


def apply_readout_to_outcomes(outcomes, n, confusion, seed, circuit, trajs):
    """Classical bit flips: qubit q of trajectory t uses 32-bit word (q % 4) of stream STREAM_READOUT + q//4;
    a measured 0 flips with P(1|0), a measured 1 with P(0|1)."""
    out = outcomes.copy()
    for blk in range((n + 3) // 4):
        w = philox.words(seed, circuit, trajs, philox.STREAM_READOUT + blk)
        for j in range(4):
            q = 4 * blk + j
            if q >= n:
                break
            u = w[j].astype(np.float64) * 2.0 ** -32
            bit = (out >> np.uint64(q)) & np.uint64(1)
            pflip = np.where(bit == 0, confusion[q][0][1], confusion[q][1][0])
            out = np.where(u < pflip, out ^ (np.uint64(1) << np.uint64(q)), out)
    return out


def sample_from_probs(probs, shots, seed, circuit=0, traj_begin=0, readout=None):
    """Multinomial sampling of an exact probability vector by inverse CDF with the STREAM_MEASURE draws
    (Aer density_matrix `sample_measure`)."""
    n = int(np.log2(len(probs)))
    trajs = np.arange(traj_begin, traj_begin + shots, dtype=np.uint64)
    cum = np.cumsum(np.asarray(probs, dtype=np.float64))
    u = philox.uniform53(seed, circuit, trajs, philox.STREAM_MEASURE)
    x = np.minimum(np.searchsorted(cum, u * cum[-1], side="right"), len(cum) - 1).astype(np.uint64)
    if readout is not None:
        x = apply_readout_to_outcomes(x, n, readout, seed, circuit, trajs)
    return x


def histogram(outcomes, n):
    return np.bincount(outcomes.astype(np.int64), minlength=2 ** n).astype(np.uint32)


def counts_dict(hist, n):
    """qiskit `Counts`-style dict: key = format(x, '0{n}b') (MSB = qubit n-1), zero bins omitted."""
    return {format(int(x), f"0{n}b"): int(c) for x, c in enumerate(hist) if c}
