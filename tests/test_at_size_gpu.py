"""Parity AT THE SIZES THE BENCHMARK NUMBERS ARE QUOTED ON (BASELINE.json configs[2..4]): the CUDA path through the C ABI
against the oracle on the benchmark's own workloads.

    C4  n = 20 noisy trajectories: bench.py's programs, complex128 outcomes shot for shot against the compiled oracle
        (oracle/cpu_sim.c), complex64 within float rounding of a CDF boundary; sampled <S_k> and the genuine witness
        against the EXACT values (oracle/pauli_prop.py) within 5 sigma.
    C3  n = 10, 12 density matrices against the dense numpy oracle (1e-5 complex64 / 1e-10 complex128); n = 14 (28 index
        bits, superoperators on (q, q + 14)) exact <S_k> against oracle/pauli_prop.py and against the mean of compiled-
        oracle trajectories.
    C5  n = 24 shot for shot (also with the depth-first schedule forced), n = 28 noisy: 2 shots against the compiled
        oracle, sampled <S_k> against the exact values.
Reference seam: samplers.py:308-326 (simulate_circuit), stabilizers.py:183-215 (StabilizerSampler.sample)."""
import numpy as np
import pytest

from oracle import cpu_build, pauli_prop, sim
from tests import circuits

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import qcdenoise_b200 as qb
    e = qb.Engine(0)
    yield e
    e.close()


def _circ(n, ops):
    import qcdenoise_b200 as qb
    return qb.Circuit(n, ops)


def _enc(n, ops):
    from qcdenoise_b200.ir import encode_programs
    return encode_programs([_circ(n, ops)])


def _reset_opts(eng):
    for k, v in (("tile_bits", 0), ("max_slots", 0), ("table_cap_log2", 22), ("max_chunk", 1 << 25), ("use_direct", 1)):
        eng.set_option(k, v)


def _toth(n, edges):
    """(name, basis-change ops, mask) of the n + 1 Toth settings (stabilizers.py:115-121, :136-156), no n <= 10 cap."""
    nbrs = {v: [b if a == v else a for a, b in edges if v in (a, b)] for v in range(n)}
    out = []
    for v in list(range(n)) + [None]:
        name = "".join("X" if q == v else ("Z" if v is not None and q in nbrs[v] else "I") for q in range(n))[::-1]
        sub = [("h", (v,))] if v is not None else []
        mask = sum(1 << q for q, ch in enumerate(name[::-1]) if ch != "I")
        out.append((name, sub, mask))
    return out


def _parity_mean(outcomes, mask):
    x = np.asarray(outcomes, dtype=np.uint64) & np.uint64(mask)
    par = np.zeros(len(x), dtype=np.int64)
    while x.any():
        par ^= (x & np.uint64(1)).astype(np.int64)
        x >>= np.uint64(1)
    return float((1 - 2 * par).mean())


# ------------------------------------------------------------------------------------------- C4: n = 20
def test_c4_bench_workload_shot_for_shot(eng):
    """The exact programs bench.py times (graph circuit, first and last Toth setting), 256 shots each."""
    import bench
    from qcdenoise_b200.ir import encode_programs
    _reset_opts(eng)
    work = bench.build_workload(0, 256)
    progs = [work["programs"][i] for i in (0, 1, 21)]
    assert progs[0].n_qubits == 20 and sum(1 for op in progs[0].ops if op[0] == "kraus") == 2 * work["n_sites"] > 0
    eng.upload(progs)
    shots = 256
    _, o64 = eng.run_trajectories(shots, seed=bench.SEED, precision=64, want_outcomes=True)
    _, o32 = eng.run_trajectories(shots, seed=bench.SEED, precision=32, want_outcomes=True)
    o64 = o64.cpu().numpy().reshape(3, shots).astype(np.int64)
    o32 = o32.cpu().numpy().reshape(3, shots).astype(np.int64)
    for c, p in enumerate(progs):
        want = cpu_build.run_trajectories(encode_programs([p]), shots, bench.SEED, circuit=c).astype(np.int64)
        assert np.array_equal(o64[c], want), f"circuit {c}: {np.count_nonzero(o64[c] != want)} of {shots} outcomes differ"
        assert np.count_nonzero(o32[c] != want) <= 3


def test_c4_bench_workload_witness_against_exact_values(eng):
    """All 22 circuits of a bench step, 200 000 complex64 shots each: every sampled <S_k> within 5 sigma of the exact
    value, and the genuine witness W = (n - 1) - sum_k <g_k> within 5 sigma of the exact one."""
    import bench
    _reset_opts(eng)
    shots = 200_000
    work = bench.build_workload(0, shots)
    n = 20
    exact = bench.exact_stabilizer_values(work)
    eng.upload(work["programs"])
    hist, _ = eng.run_trajectories(shots, seed=bench.SEED, precision=32)
    ss, tot = eng.parity_counts(hist, work["masks"])
    mean = ss.cpu().numpy()[:, 0] / tot.cpu().numpy()
    assert tot.cpu().tolist() == [shots] * 22
    sig = np.sqrt(np.maximum(1 - mean ** 2, 0) / shots)
    # circuit 0 is the dataset row (mask 0), circuits 1..21 the settings in TothStabilizer order (identity last)
    for k in range(1, 22):
        assert abs(mean[k] - exact[k - 1]) <= 5 * sig[k] + 1e-12, (k, mean[k], exact[k - 1], sig[k])
    w_gpu = (n - 1) * mean[21] - mean[1:21].sum()
    w_exact = (n - 1) * exact[20] - sum(exact[:20])
    assert abs(w_gpu - w_exact) <= 5 * np.sqrt((sig[1:] ** 2).sum())
    assert 0.5 < w_exact < 19      # the noise is doing something: an ideal graph state has W = -1


# ------------------------------------------------------------------------------------------- C3: density matrices
@pytest.mark.parametrize("n", [10, 12])
def test_c3_density_matrix_against_dense_oracle(eng, n):
    _reset_opts(eng)
    rng = np.random.default_rng(300 + n)
    g = circuits.ref_graph(10, 2) + ([(9, 10), (10, 11), (2, 11)] if n == 12 else [])
    ops = circuits.graph_circuit_amp_damp(n, g, rng, max_prob=0.1)
    assert sum(1 for op in ops if op[0] == "kraus") >= 6
    want = sim.density_matrix_probs(ops, n)                   # 4^n complex128 in numpy: 16 MB / 268 MB
    eng.upload([_circ(n, ops)])
    p64 = eng.run_density_matrix(64).cpu().numpy()[0]
    p32 = eng.run_density_matrix(32).cpu().numpy()[0]
    assert np.abs(p64 - want).max() < 1e-10
    assert np.abs(p32 - want).max() < 1e-5
    # the tile kernel alone
    eng.set_option("use_direct", 0)
    t64 = eng.run_density_matrix(64).cpu().numpy()[0]
    eng.set_option("use_direct", 1)
    assert np.abs(t64 - want).max() < 1e-10


def test_c3_n14_density_matrix_stabilizers(eng):
    """C3's own size: 14 qubits = 28 index bits, 2 GiB (complex64) / 4 GiB (complex128) per density matrix."""
    import torch
    _reset_opts(eng)
    n = 14
    rng = np.random.default_rng(314)
    edges = circuits.ref_graph(14, 3)
    ops = circuits.graph_circuit_amp_damp(n, edges, rng, max_prob=0.1)
    toth = _toth(n, edges)
    exact = pauli_prop.toth_expectations(ops, n, [(sub, mask) for _, sub, mask in toth])
    assert min(exact[:-1]) < 0.97 and exact[-1] == pytest.approx(1.0, abs=1e-14)
    picks = [0, 5, 13, 14]          # settings simulated as circuits of their own (the reference's way)
    eng.upload([_circ(n, ops + toth[k][1]) for k in picks])
    for precision, tol in ((64, 1e-10), (32, 1e-5)):
        probs = eng.run_density_matrix(precision)
        masks = np.array([[toth[k][2]] for k in picks], dtype=np.uint32)
        ev = eng.parity_probs(probs, masks).cpu().numpy()[:, 0]
        assert abs(float(probs.sum(dim=1).max()) - 1) < 10 * tol
        for j, k in enumerate(picks):
            assert ev[j] == pytest.approx(exact[k], abs=20 * tol), (precision, k)
        del probs
        torch.cuda.empty_cache()
    # every setting read off ONE density matrix of the graph circuit (qcd_run_density_matrix_settings)
    import qcdenoise_b200 as qb
    eng.upload([_circ(n, ops)])
    H = np.array([[1, 1], [1, -1]], dtype=np.complex128) / np.sqrt(2)
    sq = np.array([[k if k < n else 0xFF for k in range(n + 1)]], dtype=np.uint8)
    sop = np.broadcast_to(np.kron(H.conj(), H), (1, n + 1, 4, 4)).copy()
    ps = eng.run_density_matrix_settings(sq, sop, precision=64)[0]
    ev = eng.parity_probs(ps, np.array([[m] for _, _, m in toth], dtype=np.uint32), ).cpu().numpy()
    for k in range(n + 1):
        assert ev[k, 0] == pytest.approx(exact[k], abs=1e-9), k
    # the same numbers from compiled-oracle trajectories (one state per shot): 2000 shots per picked setting
    shots = 2000
    for k in (0, 13):
        out = cpu_build.run_trajectories(_enc(n, ops + toth[k][1]), shots, 77, circuit=k)
        m = _parity_mean(out, toth[k][2])
        assert abs(m - exact[k]) <= 5 * np.sqrt(max(1 - exact[k] ** 2, 1e-9) / shots), (k, m, exact[k])
    assert qb is not None


# ------------------------------------------------------------------------------------------- C5: n = 24 / 28
def test_c5_n24_noisy_shot_for_shot_and_depth_first(eng):
    _reset_opts(eng)
    n = 24
    rng = np.random.default_rng(524)
    edges = circuits.ref_graph(20, 1) + [(19, 20), (20, 21), (21, 22), (22, 23), (3, 23)]
    ops = circuits.graph_circuit_amp_damp(n, edges, rng, max_prob=0.2)
    n_kraus = sum(1 for op in ops if op[0] == "kraus")
    assert n_kraus >= 10
    toth = _toth(n, edges)
    batch = [ops, ops + toth[23][1]]
    eng.upload([_circ(n, o) for o in batch])
    shots, seed = 24, 99
    _, o64 = eng.run_trajectories(shots, seed=seed, precision=64, want_outcomes=True)
    st = eng.stats()
    o64 = o64.cpu().numpy().reshape(2, shots).astype(np.int64)
    for c, o in enumerate(batch):
        want = cpu_build.run_trajectories(_enc(n, o), shots, seed, circuit=c, wide=True).astype(np.int64)
        assert np.array_equal(o64[c], want), f"circuit {c}: {np.count_nonzero(o64[c] != want)} of {shots} differ"
    _, o32 = eng.run_trajectories(shots, seed=seed, precision=32, want_outcomes=True)
    assert np.count_nonzero(o32.cpu().numpy().reshape(2, shots) != o64) <= 2
    # depth-first schedule: barely more slots than the deepest path needs
    eng.set_option("max_slots", n_kraus // 2 + 4)
    _, d64 = eng.run_trajectories(shots, seed=seed, precision=64, want_outcomes=True)
    assert eng.stats()["max_live_states"] <= n_kraus // 2 + 4 and st["max_live_states"] >= 1
    eng.set_option("max_slots", 0)
    assert np.array_equal(d64.cpu().numpy().reshape(2, shots).astype(np.int64), o64)


def test_c5_n28_noisy_against_compiled_oracle_and_exact_values(eng):
    """BASELINE configs[4]: 28 qubits, noisy trajectories.  2 complex128 shots equal the compiled oracle; 96 complex64
    shots of three Toth settings reproduce the exact <S_k> within 5 sigma."""
    import torch
    _reset_opts(eng)
    n = 28
    rng = np.random.default_rng(528)
    edges = circuits.ref_graph(28, 1)
    ops = circuits.graph_circuit_amp_damp(n, edges, rng, max_prob=0.1)
    assert sum(1 for op in ops if op[0] == "kraus") >= 10
    toth = _toth(n, edges)
    eng.upload([_circ(n, ops)])
    _, o64 = eng.run_trajectories(2, seed=5, precision=64, want_outcomes=True)
    want = cpu_build.run_trajectories(_enc(n, ops), 2, 5, circuit=0, wide=True)
    assert np.array_equal(o64.cpu().numpy().astype(np.int64) & 0xFFFFFFFF, want.astype(np.int64))
    torch.cuda.empty_cache()
    picks = [2, 17, 27]
    exact = pauli_prop.toth_expectations(ops, n, [(toth[k][1], toth[k][2]) for k in picks])
    eng.upload([_circ(n, ops + toth[k][1]) for k in picks])
    shots = 96
    _, out = eng.run_trajectories(shots, seed=6, precision=32, want_outcomes=True)
    out = out.reshape(len(picks), shots)
    ss = eng.parity_outcomes(out, np.array([[toth[k][2]] for k in picks], dtype=np.uint32)).cpu().numpy()[:, 0]
    for j, k in enumerate(picks):
        m = ss[j] / shots
        assert abs(m - exact[j]) <= 5 * np.sqrt(max(1 - exact[j] ** 2, 1e-6) / shots), (k, m, exact[j])
        assert m == _parity_mean(out[j].cpu().numpy().astype(np.int64) & 0xFFFFFFFF, toth[k][2])
