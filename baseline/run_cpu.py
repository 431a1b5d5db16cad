"""CPU arm of the benchmark: the reference's own CPU implementation of the hot path, timed on the box's
host cores.  Never imports the product package (qcdenoise_b200) and never maps libqcdsim.so.

Order of preference:
  1. the REAL thing -- qiskit-aer importable: the workload's gate lists are rebuilt as qiskit circuits with an explicit
     Aer NoiseModel (the Kraus channels of every labelled site) and run through
     `qk.execute(circuit, AerSimulator(method="statevector"), noise_model=..., shots=...)`, the call the reference makes
     (samplers.py:308-326).  ("kind": "reference").  qiskit is neither vendored in the reference checkout nor present in
     this image or its offline wheelhouse, so this branch has never run here; it is kept so that a box that has qiskit
     reports the real number.
  2. the compiled port of Aer's statevector-trajectory algorithm, oracle/cpu_sim.c ("kind": "port"): complex128, gate at
     a time, one Kraus trajectory per shot, OpenMP over shots.
"""
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def have_qiskit_aer():
    try:
        import qiskit  # noqa: F401
        import qiskit_aer  # noqa: F401
        return True
    except Exception:
        return False


def _run_aer(work, shots_per_circuit, threads):
    """Real qiskit-aer on the same gate lists (untested here: qiskit is absent from this image)."""
    import numpy as np
    import qiskit as qk
    from qiskit_aer import AerSimulator
    from qiskit_aer.noise import NoiseModel
    from qiskit_aer.noise.errors import kraus_error
    n = work["n"]
    t0 = time.perf_counter()
    done = 0
    for ops in work["circuits"]:
        circ = qk.QuantumCircuit(n)
        model = NoiseModel()
        pending = None
        for op in ops:
            name, qs = op[0], op[1]
            if name == "unitary":
                pending = op[3]
                circ.unitary(np.asarray(op[2]), list(qs), label=pending)
                errs = []
            elif name == "kraus":
                errs.append(kraus_error([np.asarray(k) for k in op[2]]))
                if len(errs) == 2:      # (node, ngbr): `e_ngbr.tensor(e_node)` bound to [node, ngbr]
                    model.add_all_qubit_quantum_error(errs[1].tensor(errs[0]), pending)
            elif name != "id":
                getattr(circ, name)(*qs)
        model.add_basis_gates(["unitary"])
        circ.measure_all()
        sim = AerSimulator(method="statevector", max_parallel_threads=threads)
        sim.run(circ, noise_model=model, shots=shots_per_circuit).result().get_counts()
        done += shots_per_circuit
    return done, time.perf_counter() - t0


def run_port(work, shots_per_circuit, threads, circuits=None, seed=1234):
    """`shots_per_circuit` trajectories of each listed circuit of the step (default: all 22), two circuits at a time on
    threads / 2 OpenMP threads each.  -> (shots done, seconds)."""
    from oracle import cpu_build, encode
    cpu_build.load()
    idx = list(range(len(work["circuits"]))) if circuits is None else list(circuits)
    encs = {c: encode.encode_circuit(work["n"], work["circuits"][c]) for c in idx}
    lanes = 2 if threads >= 4 and len(idx) > 1 else 1
    per = max(1, threads // lanes)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=lanes) as pool:
        list(pool.map(lambda c: cpu_build.run_trajectories(encs[c], shots_per_circuit, seed, circuit=c, n_threads=per), idx))
    return shots_per_circuit * len(idx), time.perf_counter() - t0


def run(work, shots_per_circuit, threads=0, circuits=None):
    """-> dict(kind, shots, seconds, cores)."""
    threads = threads or host_threads()
    if have_qiskit_aer():
        try:
            shots, dt = _run_aer(work, shots_per_circuit, threads)
            return {"kind": "reference", "shots": shots, "seconds": dt, "cores": threads}
        except Exception as exc:       # fall through to the port, saying why
            note = f"qiskit-aer present but failed ({exc}); "
    else:
        note = ""
    shots, dt = run_port(work, shots_per_circuit, threads, circuits)
    return {"kind": "port", "shots": shots, "seconds": dt, "cores": threads, "note": note}
