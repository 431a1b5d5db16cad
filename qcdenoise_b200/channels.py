"""Noise-channel parameters -> explicit Kraus / Pauli-mixture ops of the gate-list IR.

The reference never builds channels itself: it hands parameters to qiskit-aer's constructors
(`errors.phase_amplitude_damping_error`, `amplitude_damping_error`: samplers.py:404-451; `depolarizing_error`:
circuit_constructors.py:118-131; `NoiseModel.from_backend`: samplers.py:588,608).  Host-side, tiny matrices, once
per site -- no device work here; the device work is applying them (csrc/sweep.cu)."""
import math

import numpy as np

_SQ = math.sqrt


def _mat(rows):
    return np.array(rows, dtype=np.complex128)


def damping_kraus(gamma_amp, gamma_phase=0.0, p_excited=0.0):
    """Generalised amplitude + phase damping on one qubit.  gamma_amp = P(|1> -> |0>), gamma_phase = pure dephasing
    weight, p_excited = thermal population of |1> (0 -> decay towards |0> only).  Zero operators are left out."""
    if gamma_amp < 0 or gamma_phase < 0 or gamma_amp + gamma_phase > 1 + 1e-12 or not 0 <= p_excited <= 1:
        raise ValueError("damping parameters out of range")
    keep = _SQ(max(0.0, 1.0 - gamma_amp - gamma_phase))
    a, ph = _SQ(gamma_amp), _SQ(gamma_phase)
    down, up = _SQ(1.0 - p_excited), _SQ(p_excited)
    cand = [down * _mat([[1, 0], [0, keep]]), down * _mat([[0, a], [0, 0]]), down * _mat([[0, 0], [0, ph]]),
            up * _mat([[keep, 0], [0, 1]]), up * _mat([[0, 0], [a, 0]]), up * _mat([[ph, 0], [0, 0]])]
    return [k for k in cand if np.abs(k).max() > 1e-12]


def depolarizing_probs(p, n_qubits=1):
    """rho -> (1 - p) rho + p I/2^k  as a Pauli mixture; index = P(q0) + 4 P(q1), P in (I, X, Y, Z)."""
    terms = 4 ** n_qubits
    if p < 0 or p > terms / (terms - 1):
        raise ValueError("depolarizing parameter out of range")
    out = np.full(terms, p / terms)
    out[0] = 1.0 - p * (terms - 1) / terms
    return out


def thermal_relaxation_kraus(t1, t2, duration, p_excited=0.0):
    """T1/T2 relaxation over `duration` (same time unit as t1, t2).  T2 <= T1: mixture {identity, Z, reset};
    T1 < T2 <= 2 T1: the non-unital channel's Kraus operators from its Choi matrix."""
    if t1 <= 0 or t2 <= 0 or duration < 0:
        raise ValueError("relaxation parameters must be positive")
    if t2 > 2 * t1 * (1 + 1e-12):
        raise ValueError("T2 cannot exceed 2 T1")
    p_reset = 1.0 - math.exp(-duration / t1) if math.isfinite(t1) else 0.0
    coh = math.exp(-duration / t2) if math.isfinite(t2) else 1.0
    to0, to1 = p_reset * (1.0 - p_excited), p_reset * p_excited
    if t2 <= t1:
        surv = 1.0 - p_reset
        pz = 0.5 * surv * (1.0 - coh / surv) if surv > 0 else 0.0
        pid = 1.0 - pz - to0 - to1
        cand = [_SQ(max(pid, 0.0)) * np.eye(2, dtype=np.complex128), _SQ(max(pz, 0.0)) * _mat([[1, 0], [0, -1]]),
                _SQ(to0) * _mat([[1, 0], [0, 0]]), _SQ(to0) * _mat([[0, 1], [0, 0]]),
                _SQ(to1) * _mat([[0, 0], [0, 1]]), _SQ(to1) * _mat([[0, 0], [1, 0]])]
        return [k for k in cand if np.abs(k).max() > 1e-15]
    # Choi matrix in the basis |i><j| (x) E(|i><j|); the coherent 2x2 block couples |00> and |11>
    a, d = 1.0 - to1, 1.0 - to0
    ops = []
    tr, det = a + d, a * d - coh * coh
    disc = _SQ(max(tr * tr / 4 - det, 0.0))
    for lam in (tr / 2 + disc, tr / 2 - disc):
        if lam <= 1e-14:
            continue
        v = np.array([coh, lam - a]) if abs(coh) > 1e-300 else (np.array([1.0, 0.0]) if abs(lam - a) < abs(lam - d)
                                                                 else np.array([0.0, 1.0]))
        v = v / np.linalg.norm(v)
        ops.append(_SQ(lam) * _mat([[v[0], 0], [0, v[1]]]))
    if to0 > 1e-15:
        ops.append(_SQ(to0) * _mat([[0, 1], [0, 0]]))
    if to1 > 1e-15:
        ops.append(_SQ(to1) * _mat([[0, 0], [1, 0]]))
    return ops


def _process_fidelity(kraus):
    d = kraus[0].shape[0]
    return sum(abs(k[0, 0] + k[1, 1]) ** 2 for k in kraus) / d ** 2 if d == 2 else \
        sum(abs(np.trace(k)) ** 2 for k in kraus) / d ** 2


def device_gate_error(n_gate_qubits, gate_error, gate_length_ns, t1_us, t2_us):
    """Channel a calibration-derived noise model attaches to one gate: a depolarizing Pauli mixture followed by
    thermal relaxation of each gate qubit, with the depolarizing strength chosen so that the COMBINED average gate
    infidelity equals `gate_error`.  Returns (pauli_probs | None, [kraus per qubit] | None)."""
    relax, f_relax = None, 1.0
    if gate_length_ns and gate_length_ns > 0:
        relax = []
        for t1, t2 in zip(t1_us, t2_us):
            t1n = t1 * 1e3
            relax.append(thermal_relaxation_kraus(t1n, min(t2 * 1e3, 2 * t1n), gate_length_ns))
        # process fidelity of a tensor product of channels is the product of the factors' (tr(A (x) B) = tr A tr B)
        f_pro = 1.0
        for ks in relax:
            f_pro *= _process_fidelity(ks)
        dim_all = 2 ** n_gate_qubits
        f_relax = (dim_all * f_pro + 1.0) / (dim_all + 1.0)
    probs = None
    if gate_error and gate_error > 0:
        dim = 2 ** n_gate_qubits
        target = min(gate_error, dim / (dim + 1))
        if 1.0 - f_relax <= target:
            terms = 4 ** n_gate_qubits
            p = min(dim * (target - (1.0 - f_relax)) / (dim * f_relax - 1.0), terms / (terms - 1))
            if p > 0:
                probs = depolarizing_probs(p, n_gate_qubits)
    return probs, relax


def readout_matrix(p_meas1_prep0, p_meas0_prep1):
    """2x2 assignment matrix, row = prepared bit, column = measured bit."""
    return np.array([[1.0 - p_meas1_prep0, p_meas1_prep0], [p_meas0_prep1, 1.0 - p_meas0_prep1]])
