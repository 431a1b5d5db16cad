"""Noise-channel constructors restating qiskit-aer's `noise.errors.standard_errors` and
`noise.device.basic_device_gate_errors` (Aer 0.8/0.9, the versions the reference's imports need: SURVEY Q13).
TEST INFRASTRUCTURE (see oracle/__init__.py).  qiskit-aer is NOT vendored in /root/reference and not
installable here; these are restated from its published algorithm and checked by analytic identities only
(tests/test_oracle_channels.py).  Reference call sites: samplers.py:404-451 (amplitude damping pairs),
samplers.py:588,608 (NoiseModel.from_backend), circuit_constructors.py:118-131 (depolarizing_error).

A channel is a list of numpy complex128 Kraus matrices (2^k x 2^k).  Basis index of a k-qubit operator acting
on qubits [q0, q1, ...] is  s = bit(q0) + 2*bit(q1) + ...  (qiskit little-endian: first listed qubit = LSB).
"""
import numpy as np

I2 = np.eye(2, dtype=np.complex128)
X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
Y = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)
PAULIS = (I2, X, Y, Z)


def phase_amplitude_damping_kraus(param_amp, param_phase=0.0, excited_state_population=0.0):
    """aer `phase_amplitude_damping_error(param_amp, param_phase, excited_state_population)` Kraus set
    (before Aer's canonicalisation, which does not change the channel).  Zero operators are dropped."""
    if param_amp < 0 or param_phase < 0 or param_amp + param_phase > 1:
        raise ValueError("invalid damping parameters")
    c0 = np.sqrt(1 - excited_state_population)
    c1 = np.sqrt(excited_state_population)
    param = 1 - param_amp - param_phase
    ops = [
        c0 * np.array([[1, 0], [0, np.sqrt(param)]], dtype=np.complex128),
        c0 * np.array([[0, np.sqrt(param_amp)], [0, 0]], dtype=np.complex128),
        c0 * np.array([[0, 0], [0, np.sqrt(param_phase)]], dtype=np.complex128),
        c1 * np.array([[np.sqrt(param), 0], [0, 1]], dtype=np.complex128),
        c1 * np.array([[0, 0], [np.sqrt(param_amp), 0]], dtype=np.complex128),
        c1 * np.array([[np.sqrt(param_phase), 0], [0, 0]], dtype=np.complex128),
    ]
    return [k for k in ops if np.linalg.norm(k) > 1e-12]


def amplitude_damping_kraus(param_amp, excited_state_population=0.0):
    """aer `amplitude_damping_error(param_amp)` = phase_amplitude_damping_error(param_amp, 0)."""
    return phase_amplitude_damping_kraus(param_amp, 0.0, excited_state_population)


def reference_unitary_noise_site(draws, error_type="phase_amplitude_damping_error"):
    """The reference's per-label channel (samplers.py:414-451, quirk Q1).

    draws = the two U(0, max_prob) numbers.  For 'phase_amplitude_damping_error' the call is
    func(phases[i], amps[i]) with amps[i] = min(0, 0.5 - phases[i]) -> param_amp = draw, param_phase = amps.
    Returns (kraus_on_node, kraus_on_ngbr): `q_error_1.tensor(q_error_2)` bound to [node, ngbr] puts
    q_error_2 (draws[1]) on the first listed qubit (node) and q_error_1 (draws[0]) on ngbr."""
    if error_type == "phase_amplitude_damping_error":
        amps = [min(0, 0.5 - p) for p in draws]
        e1 = phase_amplitude_damping_kraus(draws[0], amps[0])
        e2 = phase_amplitude_damping_kraus(draws[1], amps[1])
    elif error_type == "amplitude_damping_error":
        e1 = amplitude_damping_kraus(draws[0])
        e2 = amplitude_damping_kraus(draws[1])
    else:
        raise ValueError(error_type)
    return e2, e1


def depolarizing_pauli_probs(param, num_qubits):
    """aer `depolarizing_error(param, num_qubits)`: probs over Paulis, index = P(q0) + 4*P(q1) (I,X,Y,Z)."""
    num_terms = 4 ** num_qubits
    max_param = num_terms / (num_terms - 1)
    if param < 0 or param > max_param:
        raise ValueError("depolarizing parameter out of range")
    probs = np.full(num_terms, param / num_terms)
    probs[0] = 1 - param / max_param
    return probs


def pauli_kraus(probs):
    """Kraus form sqrt(p_P) P of a Pauli mixture over 1 or 2 qubits."""
    k = {4: 1, 16: 2}[len(probs)]
    ops = []
    for idx, p in enumerate(probs):
        if p <= 0:
            continue
        m = PAULIS[idx & 3]
        if k == 2:
            m = np.kron(PAULIS[idx >> 2], m)      # q1 is the more significant bit
        ops.append(np.sqrt(p) * m)
    return ops


def thermal_relaxation_kraus(t1, t2, time, excited_state_population=0.0):
    """aer `thermal_relaxation_error(t1, t2, time)` as a Kraus set (same channel as Aer's mixture of
    {I, Z, reset} for t2 <= t1 and as its Choi matrix for t1 < t2 <= 2 t1)."""
    if t1 <= 0 or t2 <= 0 or time < 0:
        raise ValueError("invalid relaxation parameters")
    if t2 - 2 * t1 > 0:
        raise ValueError("T2 > 2 T1")
    p_reset = 1 - np.exp(-time / t1) if np.isfinite(t1) else 0.0
    exp_t2 = np.exp(-time / t2) if np.isfinite(t2) else 1.0
    p0 = 1 - excited_state_population
    p1 = excited_state_population
    if t2 > t1:
        choi = np.array([[1 - p1 * p_reset, 0, 0, exp_t2],
                         [0, p1 * p_reset, 0, 0],
                         [0, 0, p0 * p_reset, 0],
                         [exp_t2, 0, 0, 1 - p0 * p_reset]], dtype=np.complex128)
        return choi_to_kraus(choi)
    exp_t1 = 1 - p_reset
    p_z = (1 - p_reset) * (1 - exp_t2 / exp_t1) / 2 if exp_t1 > 0 else 0.0
    p_r0 = p_reset * p0
    p_r1 = p_reset * p1
    p_id = 1 - p_z - p_r0 - p_r1
    ops = [np.sqrt(max(p_id, 0.0)) * I2, np.sqrt(max(p_z, 0.0)) * Z]
    ket0 = np.array([[1, 0], [0, 0]], dtype=np.complex128)
    low = np.array([[0, 1], [0, 0]], dtype=np.complex128)
    ket1 = np.array([[0, 0], [0, 1]], dtype=np.complex128)
    rai = np.array([[0, 0], [1, 0]], dtype=np.complex128)
    ops += [np.sqrt(p_r0) * ket0, np.sqrt(p_r0) * low, np.sqrt(p_r1) * ket1, np.sqrt(p_r1) * rai]
    return [k for k in ops if np.linalg.norm(k) > 1e-15]


def choi_to_kraus(choi, atol=1e-12):
    """Column-stacking Choi (qiskit convention: Choi = sum |i><j| (x) E(|i><j|)) -> Kraus via eigh."""
    d = int(round(np.sqrt(choi.shape[0])))
    w, v = np.linalg.eigh(choi)
    ops = []
    for val, vec in zip(w[::-1], v.T[::-1]):
        if val > atol:
            ops.append(np.sqrt(val) * vec.reshape(d, d).T)   # vec = vec(K), column stacking
    return ops


def kraus_to_superop(kraus):
    """S = sum_j conj(K_j) (x) K_j acting on column-stacked vec(rho) (index = row + d*col)."""
    return sum(np.kron(np.conj(k), k) for k in kraus)


def process_fidelity(kraus):
    d = kraus[0].shape[0]
    return float(sum(abs(np.trace(k)) ** 2 for k in kraus).real) / d ** 2


def average_gate_fidelity(kraus):
    d = kraus[0].shape[0]
    return (d * process_fidelity(kraus) + 1) / (d + 1)


def tensor_kraus(kraus_q1, kraus_q0):
    """Kraus set of E(q1) (x) E(q0): q0 = least-significant qubit."""
    return [np.kron(a, b) for a in kraus_q1 for b in kraus_q0]


def device_gate_channels(n_gate_qubits, gate_error, gate_length_ns, t1_us, t2_us, thermal=True):
    """aer `basic_device_gate_errors` for ONE gate (Aer 0.8: `depol_error.compose(relax_error)`):
    returns (pauli_probs or None, [relax kraus per qubit] or None) to be applied in that order.

    t1_us / t2_us are sequences (one per gate qubit) in microseconds, gate_length in ns; T2 is truncated to
    2*T1; the depolarizing parameter is chosen so that the COMBINED error has average gate infidelity
    `gate_error`:  p = d (e - (1 - F_relax)) / (d F_relax - 1),  d = 2^k, clipped to [0, 4^k/(4^k-1)]."""
    relax = None
    relax_fid = 1.0
    if thermal and gate_length_ns is not None and gate_length_ns > 0:
        relax = []
        for t1, t2 in zip(t1_us, t2_us):
            t1n = t1 * 1e3
            t2n = min(t2 * 1e3, 2 * t1n)
            relax.append(thermal_relaxation_kraus(t1n, t2n, gate_length_ns))
        full = relax[0] if n_gate_qubits == 1 else tensor_kraus(relax[1], relax[0])
        relax_fid = average_gate_fidelity(full)
    probs = None
    if gate_error is not None and gate_error > 0:
        dim = 2 ** n_gate_qubits
        error_param = min(gate_error, dim / (dim + 1))
        relax_infid = 1 - relax_fid
        if relax_infid <= error_param:
            depol = dim * (error_param - relax_infid) / (dim * relax_fid - 1)
            num_terms = 4 ** n_gate_qubits
            depol = min(depol, num_terms / (num_terms - 1))
            if depol > 0:
                probs = depolarizing_pauli_probs(depol, n_gate_qubits)
    return probs, relax


def readout_confusion(prob_meas1_prep0, prob_meas0_prep1):
    """aer `basic_device_readout_errors`: row = prepared, column = measured."""
    return np.array([[1 - prob_meas1_prep0, prob_meas1_prep0],
                     [prob_meas0_prep1, 1 - prob_meas0_prep1]])
