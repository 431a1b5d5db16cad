"""Analytic identities for the Aer-side half of the oracle (the boundary the reference leaves unpinned,
SURVEY.md section 8c 'Known-answer tests available')."""
import numpy as np
import pytest

from oracle import noise, ref_glue, sim


def _cptp(kraus):
    d = kraus[0].shape[0]
    return np.allclose(sum(k.conj().T @ k for k in kraus), np.eye(d), atol=1e-12)


def test_channels_are_trace_preserving():
    assert _cptp(noise.amplitude_damping_kraus(0.3))
    assert _cptp(noise.phase_amplitude_damping_kraus(0.2, 0.1, 0.05))
    assert _cptp(noise.pauli_kraus(noise.depolarizing_pauli_probs(0.1, 1)))
    assert _cptp(noise.pauli_kraus(noise.depolarizing_pauli_probs(0.1, 2)))
    assert _cptp(noise.thermal_relaxation_kraus(50e3, 30e3, 100.0))
    assert _cptp(noise.thermal_relaxation_kraus(50e3, 80e3, 100.0))     # T1 < T2 <= 2 T1: Choi path
    assert _cptp(noise.thermal_relaxation_kraus(50e3, 30e3, 100.0, 0.1))


def test_thermal_relaxation_two_forms_agree_at_t2_eq_t1():
    a = noise.kraus_to_superop(noise.thermal_relaxation_kraus(40e3, 40e3, 500.0))
    b = noise.kraus_to_superop(noise.thermal_relaxation_kraus(40e3, 40e3 * (1 + 1e-12), 500.0))
    np.testing.assert_allclose(a, b, atol=1e-9)


def test_amplitude_damping_on_one_gives_gamma():
    ops = [("x", (0,)), ("kraus", (0,), noise.amplitude_damping_kraus(0.37))]
    p = sim.density_matrix_probs(ops, 1)
    np.testing.assert_allclose(p, [0.37, 0.63], atol=1e-15)


def test_gamma_zero_is_ideal(ref_graphs):
    edges = [tuple(e) for e in ref_graphs["graphs"]["5"][0]]
    ops, labels = ref_glue.cx_gate_circuit(5, edges, coins=[1] * len(edges))
    kb = {l: noise.reference_unitary_noise_site([0.0, 0.0]) for l in labels}
    noisy = sim.density_matrix_probs(ref_glue.attach_unitary_noise(ops, kb), 5)
    ideal = np.abs(sim.ideal_statevector(ops, 5)) ** 2
    np.testing.assert_allclose(noisy, ideal, atol=1e-14)
    np.testing.assert_allclose(ideal, np.full(32, 1 / 32), atol=1e-14)   # graph states are flat in Z basis


def test_depolarizing_contracts_pauli_expectations():
    rng = np.random.default_rng(7)
    n = 3
    psi = rng.normal(size=8) + 1j * rng.normal(size=8)
    psi /= np.linalg.norm(psi)
    p = 0.23
    ops_pre = [("unitary", (0, 1), np.linalg.qr(rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)))[0]),
               ("h", (2,)), ("cx", (2, 0))]
    base = sim.density_matrix_probs(ops_pre, n)
    dep = sim.density_matrix_probs(ops_pre + [("pauli", (0, 1), noise.depolarizing_pauli_probs(p, 2))], n)
    x = np.arange(8)
    for mask in range(1, 8):
        d = 1 - 2 * (np.array([bin(v & mask).count("1") for v in x]) & 1)
        e0, e1 = base.dot(d), dep.dot(d)
        factor = (1 - p) if mask & 0b011 else 1.0
        assert e1 == pytest.approx(factor * e0, abs=1e-13)


def test_trajectory_mean_matches_density_matrix(ref_graphs):
    n = 4
    edges = [tuple(e) for e in ref_graphs["P4"]]
    ops, labels = ref_glue.cx_gate_circuit(n, edges, coins=[1, 1, 1])
    rng = np.random.default_rng(3)
    kb = {l: noise.reference_unitary_noise_site(rng.uniform(0, 0.35, 2)) for l in labels}
    noisy_ops = ref_glue.attach_unitary_noise(ops, kb) + [("pauli", (1,), noise.depolarizing_pauli_probs(0.1, 1))]
    exact = sim.density_matrix_probs(noisy_ops, n)
    assert exact.sum() == pytest.approx(1.0, abs=1e-13)
    shots = 20000
    out = sim.run_trajectories(noisy_ops, n, shots, seed=1234)
    h = sim.histogram(out, n)
    chi2 = float((((h - shots * exact) ** 2) / (shots * exact)).sum())
    assert chi2 < 50.0      # 15 dof: P(chi2 > 50) ~ 1e-5
    # dedup and the literal per-shot loop are the same function of (seed, trajectory)
    np.testing.assert_array_equal(out[:300], sim.run_trajectories(noisy_ops, n, 300, seed=1234, dedup=False))
    # a slice [a, b) of the trajectory index space reproduces the same outcomes (multi-GPU partition)
    np.testing.assert_array_equal(out[100:180], sim.run_trajectories(noisy_ops, n, 80, seed=1234, traj_begin=100))


def test_readout_error_probs_and_samples():
    n = 2
    conf = [noise.readout_confusion(0.1, 0.2), noise.readout_confusion(0.05, 0.3)]
    probs = np.array([0.5, 0.0, 0.0, 0.5])
    noisy = sim.apply_readout_to_probs(probs, n, conf)
    assert noisy.sum() == pytest.approx(1.0)
    assert noisy[0] == pytest.approx(0.5 * 0.9 * 0.95 + 0.5 * 0.2 * 0.3)
    x = sim.sample_from_probs(probs, 40000, seed=5, readout=conf)
    h = sim.histogram(x, n) / 40000
    np.testing.assert_allclose(h, noisy, atol=0.01)
