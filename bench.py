#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native noisy-circuit sampling engine.

Workload (BASELINE.json configs[3], "C4" in SURVEY.md 8d -- the largest single-GPU configuration the metric's
20-28 qubit range is quoted on): 20-qubit reference-sampled graph states, CXGateCircuit gate lists with stochastic
labelled sites, two amplitude-damping Kraus channels per site (max_prob 0.1), 1,000,000 noisy statevector-trajectory
shots of the graph circuit AND of each of its n + 1 = 21 Toth stabilizer settings (22 circuits x 1M shots per graph),
dense 2^20 histograms, bitmask-parity <S> for the 21 settings -> genuine witness.  complex64 amplitudes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--graphs G] [--scaling strong|weak]
                    [--shots S] [--literal] [--precision 32|64]

The JOB is G = 8 graphs (176 circuits, 176 M shots) per step.  One step = one pass of the path over the job; `value` =
noisy shots/s with the lowered programs resident on the GPU (CUDA events, max over ranks); `e2e` = the same through
the reference-facing Python API (UnitaryNoiseSampler.sample + StabilizerSampler.sample + GenuineWitness.estimate per
graph) from host gate lists to host numpy results, every step including encode, upload + lowering and the
device->host copy of the dataset histograms and parity sums.  `roofline` = algorithmic bytes moved by the state-sweep
kernel / its CUDA-event time, against the measured HBM copy bandwidth.  `cpu_baseline` / `--impl reference` = the
reference's CPU implementation of the path (real qiskit-aer if importable, else the compiled port oracle/cpu_sim.c) on
a bounded number of shots of the same circuits (baseline/run_cpu.py; that arm never imports the product).

N > 1 (torchrun), default `--scaling weak`: the job is N x G graphs (fixed work per GPU), DISTINCT graphs by global graph
index, ONE seed, Philox counters on the global circuit index -- a partition of one dataset-generation job, not N
copies of it.  Every rank holds the programs of the whole job and pulls the next graph from one shared counter
(distributed.WorkQueue: an atomic add on the process group's store, heaviest graph first) because graphs differ ~10x in
cost; the assembled results are bit-identical to a single-GPU run, which rank 0 asserts in warm-up by re-simulating
graphs other ranks ran.  No data-path collective; ONE NCCL all_reduce assembles the dataset rows and the parity sums.
At N > 1 `other_configs.C4_strong` times the FIXED single-GPU job (G graphs) over the N ranks as well (strong scaling;
`--scaling strong` makes that the headline).
Outside the timed region the sampled <S_k> and the witness of every graph are checked against their EXACT values
(oracle/pauli_prop.py).  `other_configs` carries C2 / C3 / C5 and the complex128 / literal variants of C4.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_QUBITS = 20
SEED = 1234


# ------------------------------------------------------------------------------------------------ workload
def build_workload(graph_index, shots):
    """Gate lists of the step for graph `graph_index` of tests/golden/graphs.json (reference-sampled, n = 20)."""
    import networkx as nx

    import qcdenoise_b200 as qcd
    with open(os.path.join(ROOT, "tests", "golden", "graphs.json")) as f:
        graphs = json.load(f)["graphs"][str(N_QUBITS)]
    edges = [tuple(e) for e in graphs[graph_index % len(graphs)]]
    g = nx.Graph()
    g.add_nodes_from(range(N_QUBITS))
    g.add_edges_from(edges)
    np.random.seed(SEED + graph_index)                       # coins and channel draws of the reference API use np.random
    cd = qcd.CXGateCircuit(n_qubits=N_QUBITS, stochastic=True).build(g)
    rng_state = np.random.get_state()      # the end-to-end leg re-draws the SAME damping parameters from here
    sampler = qcd.UnitaryNoiseSampler(None, n_shots=shots, noise_specs=qcd.unitary_noise_spec, method="statevector",
                                      precision=32, seed=SEED)
    sampler.build_noise_model(cd["ops"])
    stab = qcd.TothStabilizer(g, n_qubits=N_QUBITS).build()
    circuits = [cd["circuit"]] + [cd["circuit"].copy(name=nm).compose(sc) for nm, sc in stab.items()]
    programs = [sampler.noise_model.instantiate(c)[0] for c in circuits]
    masks = np.array([[0]] + [[qcd.pauli_mask(nm)] for nm in stab], dtype=np.uint32)
    return {"graph": g, "edges": edges, "circuit_dict": cd, "stab": stab, "programs": programs, "masks": masks,
            "noise_model": sampler.noise_model, "n_sites": len(cd["ops"]), "rng_state": rng_state}


def exact_stabilizer_values(work):
    """CHECKER (oracle/pauli_prop.py, outside every timed region): the exact <S_k> of the step's 21 Toth settings under
    the step's own noise channels, by Heisenberg-picture Pauli propagation of the measured parity observable through
    the noisy Clifford gate list.  Order = work["stab"] (identity last)."""
    import qcdenoise_b200 as qcd
    from oracle import pauli_prop
    noisy = work["programs"][0].ops
    return pauli_prop.toth_expectations(noisy, N_QUBITS, [(sc.ops, qcd.pauli_mask(nm)) for nm, sc in work["stab"].items()])


def check_witness(work, signed, totals, label):
    """The sampled <S_k> (parity sums of the step's histograms) and the genuine witness against the exact values."""
    exact = np.array(exact_stabilizer_values(work))
    mean = signed[1:] / totals[1:]
    sig = np.sqrt(np.maximum(1 - mean ** 2, 0) / totals[1:])
    worst = float(np.max(np.abs(mean - exact) / np.maximum(sig, 1e-12)))
    w_gpu = (N_QUBITS - 1) * mean[-1] - mean[:-1].sum()
    w_exact = (N_QUBITS - 1) * exact[-1] - exact[:-1].sum()
    w_sig = float(np.sqrt((sig ** 2).sum()))
    if worst > 6.0 or abs(w_gpu - w_exact) > 6.0 * w_sig:
        raise AssertionError(f"{label}: sampled stabilizer values disagree with the exact ones (worst {worst:.1f} sigma; "
                             f"witness {w_gpu:.5f} vs exact {w_exact:.5f} +- {w_sig:.5f})")
    return {"witness_sampled": float(w_gpu), "witness_exact": float(w_exact), "witness_sigma": w_sig,
            "worst_setting_sigmas": worst, "checker": "oracle/pauli_prop.py (exact Pauli propagation), outside the timed region"}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def host_threads():
    """Host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which is not what the CPU arm wants)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


METRIC = "noisy shots/sec at 20 qubits (statevector trajectories + stabilizer witnesses)"
CIRCUITS_PER_GRAPH = N_QUBITS + 2


def workload_config(n_graphs, n_sites, shots, precision, scaling):
    return {"workload": "C4: 20-qubit graph-state stochastic-trajectory sampling, 1M shots, stabilizer witnesses "
                        "(BASELINE.json configs[3])",
            "n_qubits": N_QUBITS, "graphs_per_step": n_graphs, "circuits_per_step": n_graphs * CIRCUITS_PER_GRAPH,
            "noise_sites_per_graph": n_sites, "kraus_channels_per_graph": [None if s is None else 2 * s for s in n_sites], "max_prob": 0.1,
            "shots_per_circuit": shots, "precision": "complex64" if precision == 32 else "complex128",
            "partition": "self-scheduled: every rank pulls the next graph (22 circuits) from one shared counter, heaviest "
                         "first; one seed, Philox on the global circuit index (results independent of the partition)",
            "cache": "working set per tree level (hundreds to thousands of 8 MiB states) >> 126 MB L2; "
                     "histograms re-zeroed every step"}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_sample(graph_index, shots_per_circuit, threads=0, circuits=None):
    """Bounded sample of the CPU implementation on graph `graph_index` of the job (baseline/run_cpu.py: real qiskit-aer
    if importable, else the compiled port).  The gate lists come from oracle/workload.py, not from the product."""
    from baseline import run_cpu
    from oracle import workload
    return run_cpu.run(workload.c4_workload(graph_index), shots_per_circuit, threads, circuits)


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import workload
    cores = host_threads()
    per_circuit = max(cores // 2, 2)
    n_graphs = args.graphs * (world if args.scaling == "weak" else 1)
    for _ in range(args.warmup):
        cpu_sample(0, 1, circuits=[0, 1])
    t, shots, kind = 0.0, 0, "port"
    for k in range(args.steps):
        r = cpu_sample(k % n_graphs, per_circuit)
        t += r["seconds"]
        shots += r["shots"]
        kind = r["kind"]
    v = shots / t
    one = cpu_sample(0, 2, threads=1, circuits=[0])          # the reference's tests pin OMP_NUM_THREADS=1
    sample = (f"per step {per_circuit} shots of EACH of the 22 circuits of one graph of the job (graph = step mod "
              f"{n_graphs}), one trajectory per shot, complex128, gate at a time; no extrapolation inside the sample")
    sites = [workload.c4_workload(g)["n_sites"] for g in range(n_graphs)]
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "shots/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "c128", "data": "synthetic",
            "config": workload_config(n_graphs, sites, args.shots, 64, args.scaling),
            "cpu_baseline": {"value": v, "unit": "shots/s", "cores": cores, "kind": kind, "sample": sample,
                             "value_1thread": one["shots"] / one["seconds"],
                             "qiskit_aer_importable": kind == "reference",
                             "product_imported": "qcdenoise_b200" in sys.modules},
            "e2e": {"value": v, "unit": "shots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ secondary configs
def _timed(torch, fn, reps=1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def _kind_rates(st, peak):
    """Per launch kind (prep / inner / leaf / mixed): algorithmic GB/s and fraction of the measured copy bandwidth."""
    out = {}
    for i, nm in enumerate(("prep", "inner", "leaf", "mixed")):
        if st["kind_launches"][i] and st["kind_ms"][i] > 0:
            gbs = st["kind_bytes"][i] / (st["kind_ms"][i] * 1e-3) / 1e9
            out[nm] = {"launches": int(st["kind_launches"][i]), "ms": st["kind_ms"][i], "algorithmic_GBps": gbs,
                       "frac": gbs / peak}
    return out


def _add_stats(agg, st):
    for k, v in st.items():
        if isinstance(v, list):
            agg[k] = [a + b for a, b in zip(agg.get(k, [0] * len(v)), v)]
        elif k == "max_live_states":
            agg[k] = max(agg.get(k, 0), v)
        else:
            agg[k] = agg.get(k, 0) + v
    return agg


def c2_end_to_end(qcd, torch, n_circ=10000):
    """BASELINE.json configs[1] (8-qubit random graph states, 10k circuits, device noise model, count histograms)
    end to end through qcd.simulate(): host stages (graph sampling, gate lists, calibration draws, channels, lowering),
    upload, simulation and the histograms back on the host, all inside the timed call."""
    import random
    random.seed(SEED)
    spec = qcd.NoiseSpec(specs={"qubit": {"T1": 100.0, "T2": 100.0, "readout_error": 0.25, "prob_meas0_prep1": 0.1,
                                          "prob_meas1_prep0": 0.1},
                                "cx": {"gate_error": 0.25, "gate_length": 500.0},
                                "h": {"gate_error": 0.05, "gate_length": 50.0}})
    sim_specs = qcd.SimulationSpecs(n_qubits=8, n_circuits=n_circ, data=qcd.GraphData(),
                                    circuit=qcd.CircuitSpecs(builder=qcd.CXGateCircuit, stochastic=False))
    sampler_specs = qcd.SamplingSpecs(sampler=qcd.DeviceNoiseSampler, noise_specs=spec, n_shots=1024)
    best = None
    for it in range(3):       # first call pays allocations; report the best of the next two
        diag = {}
        t0 = time.perf_counter()
        res = qcd.simulate(sim_specs, sampler_specs, qcd.GraphSpecs(max_num_edges=7), diagnostics=diag)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        assert len(res) == n_circ and abs(res[n_circ - 1]["prob_vec"].sum() - 1) < 1e-9
        if it > 0 and (best is None or dt < best[0]):
            best = (dt, dict(diag["timing"]))
    dt, stages = best
    return {"workload": f"C2: 8-qubit random graph states, {n_circ} circuits, device noise model, 1024-shot count "
                        "histograms (BASELINE.json configs[1]), through qcd.simulate()",
            "circuits_per_s_e2e": n_circ / dt, "shots_per_s_e2e": 1024 * n_circ / dt, "s_per_call": dt,
            "host_stages_s": {k: round(v, 4) for k, v in stages.items()}}


def _noisy_programs(qcd, n, graph_indices, seed):
    """Reference-shaped noisy graph circuits at n qubits: CXGateCircuit (stochastic sites) + UnitaryNoiseSampler draws."""
    import networkx as nx
    with open(os.path.join(ROOT, "tests", "golden", "graphs.json")) as f:
        graphs = json.load(f)["graphs"][str(n)]
    np.random.seed(seed)
    sampler = qcd.UnitaryNoiseSampler(None, n_shots=1024, method="statevector")
    progs, masks, sites = [], [], []
    for gi in graph_indices:
        g = nx.Graph()
        g.add_nodes_from(range(n))
        g.add_edges_from([tuple(e) for e in graphs[gi % len(graphs)]])
        cd = qcd.CXGateCircuit(n_qubits=n, stochastic=True).build(g)
        sampler.build_noise_model(cd["ops"])
        progs.append(sampler.noise_model.instantiate(cd["circuit"])[0])
        # <g_k> of the UNROTATED state's Z-type part is not a stabilizer; the masks only exercise the parity reduction
        masks.append([qcd.pauli_mask(nm) for nm in qcd.TothStabilizer(g, n).build()])
        sites.append(len(cd["ops"]))
    return progs, np.array(masks, dtype=np.uint32), sites


def c3_density_matrix(qcd, torch, eng, peak, n_circ=16):
    """BASELINE.json configs[2]: 14-qubit density matrices (2 GiB complex64 each), custom-unitary Kraus sites, 1024-shot
    multinomial sampling from the exact distribution, 15 Toth <S> + genuine witness per circuit (settings read off the
    circuit's density matrix)."""
    n = 14
    progs, _, sites = _noisy_programs(qcd, n, range(n_circ), SEED + 14)
    eng.upload(progs)
    eng.set_option("time_sweeps", 1)
    H = np.array([[1, 1], [1, -1]], dtype=np.complex128) / np.sqrt(2)
    sq = np.broadcast_to(np.array([k if k < n else 0xFF for k in range(n + 1)], dtype=np.uint8), (n_circ, n + 1)).copy()
    sop = np.broadcast_to(np.kron(H.conj(), H), (n_circ, n + 1, 4, 4)).copy()

    def probs_only():
        p = eng.run_density_matrix(32)
        st = eng.stats()
        h = eng.sample_from_probs(p, 1024, seed=SEED)
        return p, h, st

    probs_only()
    ms, (p, h, st) = _timed(torch, probs_only)
    assert abs(float(p.sum(dim=1).min()) - 1) < 1e-4 and h.sum(dim=1).tolist() == [1024] * n_circ
    del p, h

    def with_witness():
        ps = eng.run_density_matrix_settings(sq, sop, precision=32)             # [B, n + 1, 2^n]
        hs = eng.sample_from_probs(ps.reshape(-1, 2 ** n), 1024, seed=SEED + 1)
        return eng.parity_counts(hs, np.array([[1 << k] if k < n else [0] for k in range(n + 1)] * n_circ, dtype=np.uint32))

    with_witness()
    ms_w, _ = _timed(torch, with_witness)
    eng.set_option("time_sweeps", 0)
    kinds = _kind_rates(st, peak)
    gbs = st["sweep_bytes"] / (st["sweep_ms"] * 1e-3) / 1e9
    return {"workload": f"C3: {n}-qubit density-matrix simulation (4^14 complex64 = 2 GiB per circuit), {n_circ} circuits, "
                        "amplitude-damping Kraus sites (max_prob 0.1), 1024 shots from the exact distribution "
                        "(BASELINE.json configs[2])",
            "circuits_per_s": n_circ / (ms * 1e-3), "ms_per_batch": ms, "noise_sites": sites,
            "circuits_per_s_with_15_settings_and_witness": n_circ / (ms_w * 1e-3),
            "state_passes_per_circuit": st["node_sweeps"] / n_circ,
            "roofline": {"bound": "hbm", "kernel": "qcd::sweep_direct_kernel<float> (density-matrix passes)",
                         "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                         "per_kind": kinds, "sweep_share_of_batch": st["sweep_ms"] / ms,
                         "algorithmic_bytes": "passes x 2 x 4^14 x 8 B per circuit (the first pass writes only)"}}


def c5_trajectories(qcd, torch, dist, eng, peak, rank, world, local):
    """BASELINE.json configs[4]: 28-qubit noisy statevector trajectories (2 GiB complex64 per state), 64 circuits x 8
    shots sharded by circuit index over the ranks (N = 1: one GPU's share of 8 circuits), sparse outcome lists + 29
    parity sums per circuit, one all_gather."""
    from qcdenoise_b200 import distributed as qd
    n, shots = 28, 8
    total = 64 if world > 1 else 8
    progs, masks, sites = _noisy_programs(qcd, n, range(total), SEED + 28)
    lo, hi = qd.circuit_slice(total, rank, world)
    eng.upload(progs[lo:hi])
    eng.set_option("time_sweeps", 1)

    def step():
        _, out = eng.run_trajectories(shots, seed=SEED, precision=32, want_outcomes=True, first_circuit=lo)
        st = eng.stats()
        out = out.reshape(hi - lo, shots)
        ss = eng.parity_outcomes(out, masks[lo:hi])
        rows = torch.cat([out.to(torch.int64), ss], dim=1).contiguous()
        return st, (qd.gather_rows(rows) if world > 1 else rows)

    step()
    if world > 1:
        dist.barrier()
    ms, (st, rows) = _timed(torch, step)
    eng.set_option("time_sweeps", 0)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=torch.device("cuda", local))
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    assert rows.shape == (total, shots + n + 1) and int(rows[:, -1].min()) == shots      # identity setting: +1 always
    gbs = st["sweep_bytes"] / (st["sweep_ms"] * 1e-3) / 1e9
    return {"workload": f"C5: {n}-qubit noisy statevector trajectories, {total} circuits x {shots} shots, sharded by circuit "
                        f"index over {world} GPU(s) (BASELINE.json configs[4])",
            "shots_per_s": total * shots / (ms * 1e-3), "circuits_per_s": total / (ms * 1e-3), "ms_per_step": ms,
            "n_gpus": world, "state_sweeps_rank0": int(st["node_sweeps"]), "live_states_rank0": int(st["max_live_states"]),
            "noise_sites": sites[:8],
            "roofline_rank0": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                               "per_kind": _kind_rates(st, peak), "sweep_share_of_step": st["sweep_ms"] / ms}}


def c4_variant(torch, eng, work, peak, shots, precision, literal):
    """One graph of the job at complex128, or executed literally (one state per shot, the way Aer runs it)."""
    B = len(work["programs"])
    eng.upload(work["programs"])
    eng.set_option("time_sweeps", 1)
    hist = torch.zeros((B, 2 ** N_QUBITS), dtype=torch.int32, device=eng.device)

    def step():
        hist.zero_()
        eng.run_trajectories(shots, seed=SEED, precision=precision, hist=hist, dedup=not literal)
        return eng.stats()

    eng.run_trajectories(max(shots // 20, 64), seed=SEED, precision=precision, hist=hist, dedup=not literal)
    ms, st = _timed(torch, step)
    eng.set_option("time_sweeps", 0)
    assert hist.sum(dim=1).tolist() == [shots] * B
    gbs = st["sweep_bytes"] / (st["sweep_ms"] * 1e-3) / 1e9
    out = {"shots_per_s": B * shots / (ms * 1e-3), "ms_per_step": ms, "shots_per_circuit": shots,
           "states_swept": int(st["node_sweeps"]),
           "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                        "per_kind": _kind_rates(st, peak), "sweep_share_of_step": st["sweep_ms"] / ms}}
    if literal:
        # SURVEY.md 8d's own per-shot accounting: A_unit = N b (1 + 2 K + 1), K = state-dependent site groups
        k_sites = work["n_sites"]
        a_unit = (2 ** N_QUBITS) * (8 if precision == 32 else 16) * (2 + 2 * k_sites)
        out["survey_8d"] = {"A_unit_bytes_per_shot": a_unit, "K_sites": k_sites,
                            "achieved_GBps": out["shots_per_s"] * a_unit / 1e9,
                            "frac": out["shots_per_s"] * a_unit / 1e9 / peak,
                            "note": "whole-step wall time, per-shot algorithmic bytes N b (2 + 2K) of SURVEY.md 8d"}
    return out


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--graphs", type=int, default=8, help="graphs of the job (per GPU with --scaling weak)")
    ap.add_argument("--scaling", default="weak", choices=["strong", "weak"])
    ap.add_argument("--shots", type=int, default=1_000_000, help="shots per circuit (1M = the BASELINE config)")
    ap.add_argument("--literal", action="store_true", help="one state per shot (no trajectory-tree sharing)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other", action="store_true", help="skip the secondary configurations")
    ap.add_argument("--precision", type=int, default=32, choices=[32, 64], help="32 = complex64 (default), 64 = complex128")
    ap.add_argument("--max-chunk", type=int, default=0, help="engine option max_chunk (trajectories per tree); 0 = default")
    ap.add_argument("--nostore-max", type=int, default=-1, help="engine option nostore_max (A/B measurements); -1 = default")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import qcdenoise_b200 as qcd
    from qcdenoise_b200 import distributed as qd
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    # ---- the job: G graphs x 22 circuits.  Every rank holds the programs of the WHOLE job (a few KB per circuit) and
    # pulls one graph at a time from a shared counter (qd.WorkQueue), heaviest first: a static slice per rank balances
    # counts, not cost (a 17-site graph sweeps ~10x the states of a 7-site one)
    G = args.graphs * (world if args.scaling == "weak" else 1)
    CPG = CIRCUITS_PER_GRAPH
    works = [build_workload(g, args.shots) for g in range(G)]
    masks_all = np.concatenate([w["masks"] for w in works])
    N = 2 ** N_QUBITS
    WIDTH = N + CPG * 4                                     # int32 words per graph: dataset row + [22][2] int64 sums
    eng = qcd.get_engine(local)
    eng.upload([p for w in works for p in w["programs"]])
    eng.set_option("time_sweeps", 1)
    if args.max_chunk:
        eng.set_option("max_chunk", args.max_chunk)
    if args.nostore_max >= 0:
        eng.set_option("nostore_max", args.nostore_max)
    for key in ("nostore_max", "use_direct", "no_slots", "tile_bits", "share_prefix"):        # tuning experiments only
        if os.environ.get("QCD_" + key.upper()):
            eng.set_option(key, float(os.environ["QCD_" + key.upper()]))
    hist = torch.zeros((CPG, N), dtype=torch.int32, device=dev)          # one graph's 22 histograms at a time

    def heaviest_first(n_units):
        return sorted(range(n_units), key=lambda g: -works[g]["n_sites"])

    def run_graph(g, out_hist):
        """Graph g of the job: its 22 circuits x shots trajectories -> histograms, 21 parity sums."""
        out_hist.zero_()
        eng.run_trajectories(args.shots, seed=SEED, precision=args.precision, hist=out_hist, dedup=not args.literal,
                             shot_range=(g * CPG * args.shots, (g + 1) * CPG * args.shots), hist_row0=g * CPG)
        st = eng.stats()
        ss, tot = eng.parity_counts(out_hist, masks_all[g * CPG:(g + 1) * CPG])
        return st, torch.stack([ss[:, 0], tot], dim=1)                    # [22, 2] int64

    def step(queue):
        """One pass over the queue's graphs.  Assembly: ONE all_reduce of the [graphs, dataset row + parity sums] buffer
        in which every rank filled the rows of the graphs it ran."""
        full = torch.zeros((queue.n_units, WIDTH), dtype=torch.int32, device=dev)
        st_sum, ran = {}, []
        t0 = time.perf_counter()
        for g in queue.epoch():
            st, sums_g = run_graph(g, hist)
            full[g, :N] = hist[0]                                         # dataset row: the noisy graph circuit
            full[g, N:] = sums_g.reshape(-1).view(torch.int32)
            _add_stats(st_sum, st)
            ran.append(g)
        torch.cuda.synchronize()
        busy = time.perf_counter() - t0
        qd.assemble_rows(full)
        sums = full[:, N:].contiguous().view(torch.int64).reshape(queue.n_units * CPG, 2)
        return st_sum, full[:, :N], sums, ran, busy

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_steps(queue, k):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        agg, busy = {}, 0.0
        ev0.record()
        for _ in range(k):
            st, rows, sums, ran, b = step(queue)
            _add_stats(agg, st)
            busy += b
        ev1.record()
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1), 1e3 * busy, float(len(ran)), float(agg.get("sweep_bytes", 0)),
                          float(agg.get("node_sweeps", 0))], dtype=torch.float64, device=dev)
        if world > 1:
            every = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(every, t)
            t = torch.stack(every)
        else:
            t = t[None]
        job = {"sweep_bytes_per_step": float(t[:, 3].sum().item()) / k, "states_swept_per_step": float(t[:, 4].sum().item()) / k}
        return float(t[:, 0].max().item()), agg, rows, sums, (t[:, 1] / k).tolist(), [int(x) for x in t[:, 2].tolist()], job

    wq = qd.WorkQueue(G, order=heaviest_first(G), name="qcd_c4")
    rows = sums = ran = None
    for _ in range(max(args.warmup, 1)):
        st, rows, sums, ran, _ = step(wq)
    # ---- warm-up checks (outside the timed region)
    partition_check = None
    if world > 1:
        # the partitioned job equals the single-GPU job: rank 0 re-simulates graphs that OTHER ranks ran in the last
        # warm-up step (same seed, same global circuit indices) and compares with what they contributed -- bit for bit
        theirs = [g for g in range(G) if g not in ran]
        if rank == 0 and theirs:
            picks = sorted({theirs[0], theirs[-1], max(theirs, key=lambda g: works[g]["n_sites"])})
            h2 = torch.zeros_like(hist)
            same_rows = same_sums = True
            for g in picks:
                _, s2 = run_graph(g, h2)
                same_rows &= bool(torch.equal(h2[0], rows[g]))
                same_sums &= bool(torch.equal(s2, sums[g * CPG:(g + 1) * CPG]))
            del h2
            if not (same_rows and same_sums):
                raise AssertionError("partitioned run differs from the single-GPU run of the same circuits "
                                     f"(dataset rows identical: {same_rows}, parity sums identical: {same_sums})")
            partition_check = {"graphs_rerun_on_rank0": picks, "ran_by_other_ranks": True,
                               "dataset_rows_identical": same_rows, "parity_sums_identical": same_sums}
    witness_check = None
    if not args.literal or args.shots >= 100_000:
        sm = sums.cpu().numpy().astype(np.float64)
        res = {g: check_witness(works[g], sm[g * CPG:(g + 1) * CPG, 0], sm[g * CPG:(g + 1) * CPG, 1], f"graph {g}")
               for g in ran}                                  # every rank checks the graphs it ran
        if world > 1:
            parts = [None] * world
            dist.all_gather_object(parts, res)
            res = {g: r for p in parts for g, r in p.items()}
        assert sorted(res) == list(range(G))
        witness_check = {"graphs": len(res), "worst_setting_sigmas": max(r["worst_setting_sigmas"] for r in res.values()),
                         "witness_sampled": [res[g]["witness_sampled"] for g in range(G)],
                         "witness_exact": [res[g]["witness_exact"] for g in range(G)],
                         "witness_sigma": [res[g]["witness_sigma"] for g in range(G)], "checker": res[0]["checker"]}
    with ClockSampler(local) as clk:
        ms, agg, rows, sums, rank_busy_ms, rank_graphs, job = timed_steps(wq, args.steps)
    assert sums[:, 1].cpu().tolist() == [args.shots] * (G * CPG), "histograms do not hold every shot"
    shots_per_step = G * CPG * args.shots
    value = shots_per_step * args.steps / (ms * 1e-3)

    # ---- the FIXED job of --graphs graphs on all ranks (strong scaling), next to the weak-scaling headline
    strong = None
    if world > 1 and args.scaling == "weak":
        sq = qd.WorkQueue(args.graphs, order=heaviest_first(args.graphs), name="qcd_c4_strong")
        step(sq)
        ms_s, _, _, sums_s, busy_s, graphs_s, _ = timed_steps(sq, 2)
        assert torch.equal(sums_s, sums[: args.graphs * CPG]), "strong-scaling pass differs from the weak-scaling pass"
        strong = {"workload": f"the single-GPU job ({args.graphs} graphs x {CPG} circuits x {args.shots} shots) "
                              f"self-scheduled over {world} GPUs at graph granularity",
                  "scaling": "strong", "shots_per_s": args.graphs * CPG * args.shots * 2 / (ms_s * 1e-3),
                  "ms_per_step": ms_s / 2, "rank_busy_ms": busy_s, "graphs_per_rank": graphs_s,
                  "noise_sites_per_graph": [w["n_sites"] for w in works[: args.graphs]],
                  "limiter": "load imbalance at graph granularity: the step lasts as long as the rank that drew the "
                             "heaviest graph(s) (rank_busy_ms); the all_reduce of the assembled rows is < 1 ms"}

    # ---- end to end through the reference-facing API (host gate lists in, host numpy out), self-scheduled the same way
    e2e = None
    if not args.no_e2e:
        sampler = qcd.UnitaryNoiseSampler(None, n_shots=args.shots, method="statevector", precision=args.precision, seed=SEED)
        ssam = qcd.StabilizerSampler(None, n_shots=args.shots, method="statevector", precision=args.precision, seed=SEED + 1)
        eq = qd.WorkQueue(G, order=heaviest_first(G), name="qcd_c4_e2e")

        def e2e_step(counts_to_host=False):
            out = {}
            for gi in eq.epoch():
                w = works[gi]
                g, cd, stab = w["graph"], w["circuit_dict"], w["stab"]
                np.random.set_state(w["rng_state"])      # sample() draws its noise model: the device leg's channels
                p = sampler.sample(cd["circuit"], cd["ops"])                               # np.ndarray (2^n,) float64
                counts = ssam.sample(stab, cd["circuit"], noise_model=sampler.noise_model)   # 21 Counts
                if counts_to_host:
                    for c in counts:
                        c.hist                                                             # host histograms, as Aer returns
                wit = qcd.GenuineWitness(N_QUBITS, stab, counts).estimate(g)
                out[gi] = (p, wit)
            return out

        eng.set_option("time_sweeps", 0)
        e2e_step()
        barrier()
        k_e2e = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            res = e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        barrier()
        t0 = time.perf_counter()
        e2e_step(counts_to_host=True)
        torch.cuda.synchronize()
        dt_full = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt, dt_full], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt, dt_full = float(t[0].item()), float(t[1].item())
        from qcdenoise_b200.ir import encode_programs
        h2d = sum(int(a.nbytes) for w in works for a in encode_programs(w["programs"])[:3])
        d2h = G * int(N * 4 + (CPG - 1) * 16)
        e2e = {"value": shots_per_step * k_e2e / dt, "unit": "shots/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": k_e2e,
               "witness_values_rank0": {str(g): float(w.value) for g, (_, w) in sorted(res.items())},
               "api": "per graph: UnitaryNoiseSampler.sample + StabilizerSampler.sample + GenuineWitness.estimate",
               "stabilizer_histograms": "stay on the GPU behind lazy Counts (the witness reduces them on the device); "
                                        "`with_counts_on_host` copies all 21 per graph to the host as well, as Aer "
                                        "returns them",
               "with_counts_on_host": {"value": shots_per_step / dt_full, "d2h_bytes_per_step": G * CPG * N * 4}}

    # ---- secondary configurations (C5 runs on every rank count; the others at N = 1)
    peak, peak_src = measured_hbm_peak()
    other = None
    if not args.no_other:
        other = {}
        try:
            other["C5"] = c5_trajectories(qcd, torch, dist, eng, peak, rank, world, local)
        except Exception as exc:     # a secondary measurement must not lose the headline
            other["C5"] = {"failed": repr(exc)}
        if strong is not None:
            other["C4_strong"] = strong
        if world == 1:
            for name, fn in (("C3", lambda: c3_density_matrix(qcd, torch, eng, peak)),
                             # round 1's headline job (graph 0 alone, complex64), for continuity with BENCH_r01
                             ("C4_graph0", lambda: c4_variant(torch, eng, works[0], peak, args.shots, 32, False)),
                             ("C4_complex128", lambda: c4_variant(torch, eng, works[0], peak, args.shots, 64, False)),
                             ("C4_literal", lambda: c4_variant(torch, eng, works[0], peak, 4000, 32, True)),
                             ("C2", lambda: c2_end_to_end(qcd, torch))):
                try:
                    other[name] = fn()
                except Exception as exc:
                    other[name] = {"failed": repr(exc)}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    for k in ("sweep_bytes", "sweep_ms", "sweep_launches", "aux_launches", "node_sweeps", "leaves"):
        agg.setdefault(k, 0)
    for k in ("kind_launches", "kind_bytes", "kind_ms"):
        agg.setdefault(k, [0, 0, 0, 0])
    achieved = agg["sweep_bytes"] / (agg["sweep_ms"] * 1e-3) / 1e9 if agg["sweep_ms"] > 0 else None
    kinds = _kind_rates(agg, peak)
    traffic, traffic_src, dram_frac, tj_inner = None, None, None, None
    tp = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tp):   # DRAM bytes per algorithmic byte PER LAUNCH KIND, from the committed ncu --set full capture
        with open(tp) as f:
            tj = json.load(f)["dram_bytes_per_algorithmic_byte"]
        tj_inner = tj.get("inner")
        dram_bytes = sum(tj.get(nm, tj.get("inner", 1.0)) * agg["kind_bytes"][i]
                         for i, nm in enumerate(("prep", "inner", "leaf", "mixed")))
        traffic = dram_bytes / max(agg["sweep_launches"], 1)
        traffic_src = "profiles/r02_traffic.json: ncu dram__bytes_read.sum + dram__bytes_write.sum per algorithmic byte, " \
                      "per launch kind, x this run's bytes per kind"
        dram_frac = dram_bytes / (agg["sweep_ms"] * 1e-3) / 1e9 / peak
        for nm, k in kinds.items():
            k["dram_frac"] = k["frac"] * tj.get(nm, tj.get("inner", 1.0))
    # The dominant kernel is the register-only sweep of the INNER tree levels (about half of the step, HBM-bound): the
    # top-level figures are its own.  The leaf level (sweep_direct_kernel<.,.,2> for leaves that keep their own sweep +
    # leaf_cross_kernel for the shared passes of sibling leaves) is bounded by instruction issue / latency, not by HBM:
    # its entry under per_kind says how far below the HBM roofline that leaves it; `all_sweeps` aggregates every launch.
    inner = kinds.get("inner") or {}
    n_inner = max(agg["kind_launches"][1], 1)
    if "leaf" in kinds and agg["kind_ms"][2] > 0:
        state_bytes = (1 << N_QUBITS) * (8 if args.precision == 32 else 16)
        kinds["leaf"]["leaves_per_ms"] = agg["leaves"] / agg["kind_ms"][2]
        # what the same leaves would have requested at one read of the parent per leaf (the accounting before sibling groups)
        kinds["leaf"]["one_read_per_leaf_equivalent_GBps"] = agg["leaves"] * state_bytes / (agg["kind_ms"][2] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "qcd::sweep_direct_kernel<float, true, 1> (register-only sweep, inner tree levels)",
                "achieved": inner.get("algorithmic_GBps"), "peak": peak, "unit": "GB/s", "frac": inner.get("frac"),
                "traffic": (tj_inner * agg["kind_bytes"][1] / n_inner) if tj_inner is not None else None,
                "traffic_source": traffic_src, "peak_source": peak_src,
                "bytes_per_launch": agg["kind_bytes"][1] / n_inner, "avg_launch_ms": agg["kind_ms"][1] / n_inner,
                # the same launches in DRAM bytes (ncu capture): sibling states re-read their parent from L2, so `frac` can
                # exceed 1 while the DRAM interface stays below its peak
                "dram_frac": inner.get("dram_frac"),
                "share_of_step": agg["kind_ms"][1] / ms,
                "sweep_share_of_step": agg["sweep_ms"] / ms, "per_kind": kinds,
                "all_sweeps": {"achieved": achieved, "frac": achieved / peak if achieved else None, "dram_frac": dram_frac,
                               "dram_bytes_per_launch": traffic,
                               "bytes_per_launch": agg["sweep_bytes"] / max(agg["sweep_launches"], 1),
                               "avg_launch_ms": agg["sweep_ms"] / max(agg["sweep_launches"], 1)},
                "algorithmic_bytes": "2 x 2^20 x 8 B per (state, sweep) [write-only for the prep sweep, read-only for a "
                                     "leaf that is not stored, one read of the parent per shared pass of a sibling group]; "
                                     "one sweep per state-dependent noise-site group"}
    cpu = None
    if not args.no_cpu and world == 1:      # reported at N = 1 only
        cores = host_threads()
        try:
            r = cpu_sample(0, max(cores // 2, 2))
            one = cpu_sample(0, 2, threads=1, circuits=[0])
            cpu = {"value": r["shots"] / r["seconds"], "unit": "shots/s", "cores": r["cores"], "kind": r["kind"],
                   "value_1thread": one["shots"] / one["seconds"],
                   "sample": f"{r['shots']} shots: {max(cores // 2, 2)} of each of the 22 circuits of graph 0 of the job, one "
                             f"trajectory per shot, complex128, {r['seconds']:.1f} s (oracle/cpu_sim.c via "
                             f"baseline/run_cpu.py; gate lists from oracle/workload.py)"}
        except Exception as exc:     # a CPU-baseline problem must not lose the GPU measurement
            cpu = {"value": None, "unit": "shots/s", "cores": cores, "kind": "port", "sample": f"failed: {exc!r}"}
    sites = [w["n_sites"] for w in works]
    line = {"metric": METRIC, "value": value, "unit": "shots/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "c64" if args.precision == 32 else "c128", "data": "synthetic",
            "config": workload_config(G, sites, args.shots, args.precision, args.scaling),
            "circuits_per_s": G * CPG * args.steps / (ms * 1e-3),
            "trajectory_sharing": not args.literal,
            "states_swept_per_step_rank0": agg["node_sweeps"] / args.steps,
            "leaf_states_per_step_rank0": agg["leaves"] / args.steps,
            "states_swept_per_graph_rank0": agg["node_sweeps"] / args.steps / max(rank_graphs[0], 1),
            # leaf level: unstored leaves of one parent and branch word whose circuits differ in one trailing Hadamard take
            # their chunk sums from shared passes over the parent (states_swept counts those passes, not the leaves)
            "sibling_groups_rank0": {"leaves_per_step": agg.get("cross_leaves", 0) / args.steps,
                                     "shared_passes_per_step": agg.get("cross_passes", 0) / args.steps,
                                     "leaves_with_own_sweep_per_step": (agg["leaves"] - agg.get("cross_leaves", 0)) / args.steps},
            "rank_busy_ms_per_step": rank_busy_ms, "graphs_per_rank_last_step": rank_graphs,
            # the job's cost depends on which graphs it holds (noise sites per graph), so shots/s of different job sizes
            # are not directly comparable; the bytes the sweeps of ALL ranks moved per second are
            "whole_job": {"states_swept_per_step": job["states_swept_per_step"],
                          "algorithmic_sweep_GB_per_step": job["sweep_bytes_per_step"] / 1e9,
                          "aggregate_algorithmic_GBps": job["sweep_bytes_per_step"] / (ms / args.steps * 1e-3) / 1e9,
                          "per_gpu_algorithmic_GBps": job["sweep_bytes_per_step"] / (ms / args.steps * 1e-3) / 1e9 / world},
            "partition_check": partition_check, "witness_check": witness_check,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "other_configs": other,
            "gpu_launches": int(agg["sweep_launches"] + agg["aux_launches"]),
            "clocks": clk.summary()}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
