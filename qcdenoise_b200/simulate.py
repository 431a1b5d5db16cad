"""Orchestration with the reference's spec dataclasses and `simulate()` signature (qcdenoise/simulate.py:43-326).
The reference loops over circuits, one qiskit call each; here all circuits of a run are built first and simulated
in ONE batched engine call per stage (sampling, stabilizer settings)."""
import contextlib
import gc
import random
from copy import deepcopy
from dataclasses import dataclass
from typing import Any, Dict

import numpy as np

from .graph_circuit import CPhaseGateCircuit, CXGateCircuit, CZGateCircuit
from .graph_state import GraphDB, GraphState
from .samplers import CircuitSampler, DeviceNoiseSampler, NoiseSpec, UnitaryNoiseSampler, get_engine
from .stabilizers import StabilizerSampler, TothStabilizer
from .witnesses import GenuineWitness

__all__ = ["GraphSpecs", "SamplingSpecs", "CircuitSpecs", "AdjTensorSpecs", "EntanglementSpecs", "SimulationSpecs",
           "simulate", "circuit_sampling_op", "estimate_entanglement_op", "adjacency_tensor_op"]


@dataclass(init=True)
class GraphSpecs:
    min_num_edges: int = 2
    max_num_edges: int = 3
    directed: bool = False
    min_vertex_cover: int = 1
    max_itr: int = 1


@dataclass(init=True)
class SamplingSpecs:
    sampler: Any
    noise_specs: NoiseSpec
    sample_options: dict = None
    transpile_options: dict = None
    n_shots: int = 1024


@dataclass(init=True)
class CircuitSpecs:
    builder: Any
    stochastic: bool
    build_options: dict = None
    gate_type: str = "none"


@dataclass(init=True)
class AdjTensorSpecs:
    tensor_shape: tuple
    fixed_size: bool = True
    directed: bool = False
    encoding: dict = None     # gate name -> code; None: the gate names that occur, enumerated from 1


@dataclass(init=True)
class EntanglementSpecs:
    witness: Any
    stabilizer: Any
    stabilizer_sampler: Any = StabilizerSampler
    n_shots: int = 1024
    noisy: bool = False     # reference behaviour (quirk Q7): stabilizer settings are simulated WITHOUT noise


@dataclass(init=True)
class SimulationSpecs:
    n_qubits: int
    n_circuits: int
    data: Any
    circuit: CircuitSpecs
    backend: Any = None


def adjacency_tensor_op(circuit, backend, sampler, sampler_specs, adj_tensor_specs):
    """Adjacency-tensor encoding of one circuit (reference simulate.py:105-134).  Without a transpiler the walk is
    over the gate list as written; `adj_tensor_specs.encoding` (or the gate names that occur, from 1) gives the codes."""
    from .dataset import generate_adjacency_tensor
    enc = adj_tensor_specs.encoding
    if enc is None:
        names = sorted({op[0] for op in circuit.ops if op[0] not in ("barrier", "measure_all")})
        enc = {nm: i + 1 for i, nm in enumerate(names)}
    return generate_adjacency_tensor(circuit, enc, adj_tensor_specs.tensor_shape, adj_tensor_specs.fixed_size,
                                     adj_tensor_specs.directed)


def circuit_sampling_op(graph, circ_builder, circ_specs, sampler, sampler_specs) -> Dict:
    circ_dict = circ_builder.build(graph, **circ_specs.build_options) if circ_specs.build_options \
        else circ_builder.build(graph)
    if isinstance(sampler, UnitaryNoiseSampler):
        prob_vec = sampler.sample(circ_dict["circuit"], circ_dict["ops"])
    else:
        prob_vec = sampler.sample(circ_dict["circuit"], **(sampler_specs.sample_options or {}))
    return {"prob_vec": prob_vec, "graph": graph, "circuit": circ_dict["circuit"],
            "noise_model": deepcopy(sampler.noise_model)}


def estimate_entanglement_op(graph, graph_circuit, entangle_specs, n_qubits, backend=None, noise_model=None):
    stabilizer = entangle_specs.stabilizer(graph, n_qubits)
    stab_circs = stabilizer.build()
    sampler = entangle_specs.stabilizer_sampler(n_shots=entangle_specs.n_shots, backend=backend)
    counts = sampler.sample(stabilizer_circuits=stab_circs, graph_circuit=graph_circuit,
                            noise_model=noise_model if entangle_specs.noisy else None)
    witness = entangle_specs.witness(n_qubits=n_qubits, stabilizer_circuits=stab_circs, stabilizer_counts=counts)
    return witness.estimate(graph=graph)


_DAMPING_KINDS = ("phase_amplitude_damping_error", "amplitude_damping_error")


def _array_route_ok(sim_specs, sampler, builder, graph_specs, entangle_specs):
    """The whole-batch array route (batch.py) covers the three graph-circuit builders on undirected graphs, with either
    random device noise (DeviceNoiseSampler, backend None, no stochastic sites) or the labelled damping sites of
    UnitaryNoiseSampler; an entanglement stage if it is Toth settings + genuine / bi-separable witness in the
    density-matrix regime (n_shots > 2^n, n <= 12)."""
    if type(builder) not in (CXGateCircuit, CZGateCircuit, CPhaseGateCircuit) or graph_specs.directed \
            or not sampler.noisy:
        return False
    if entangle_specs is not None:
        # the Toth settings are read off one density-matrix simulation per circuit: needs that method to apply
        from .witnesses import BiSeparableWitness
        n = sim_specs.n_qubits
        if entangle_specs.stabilizer is not TothStabilizer or entangle_specs.stabilizer_sampler is not StabilizerSampler \
                or entangle_specs.witness not in (GenuineWitness, BiSeparableWitness) \
                or not (entangle_specs.n_shots > 2 ** n and n <= StabilizerSampler.dm_max_qubits):
            return False
    if isinstance(sampler, DeviceNoiseSampler):
        return sim_specs.backend is None and not sim_specs.circuit.stochastic
    if isinstance(sampler, UnitaryNoiseSampler):
        specs = sampler.noise_specs.specs
        return specs.get("type", _DAMPING_KINDS[0]) in _DAMPING_KINDS and set(specs) <= {"type", "max_prob"}
    return False


@contextlib.contextmanager
def _gc_paused():
    """Bulk creation of ~10^5 small objects (edge tuples, result dicts per 10 000 circuits): the cyclic collector would
    walk all of them, and the caller's previous results, several times -- 50 ms pauses were measured.  Nothing created
    here is cyclic."""
    was_on = gc.isenabled()
    gc.disable()
    try:
        yield
    finally:
        if was_on:
            gc.enable()


def _simulate_arrays(sim_specs, sampler_specs, graph_specs, adj_tensor_specs, graph_state, builder, sampler,
                     distributed=False, entangle_specs=None, diagnostics=None):
    """simulate() without per-circuit Python objects: graphs, gate lists, calibration draws and adjacency tensors are
    built for the whole batch (natively / in numpy), the noise channels are written natively at upload, one engine
    call simulates everything.  Draws come from numpy / Philox generators seeded from `random`, so `random.seed()`
    still fixes the result -- but the stream differs from the object route's."""
    from . import batch
    from .samplers import _to_host, get_engine
    import time
    timing, t0 = {}, time.perf_counter()

    def lap(name):
        nonlocal t0
        t1 = time.perf_counter()
        timing[name] = timing.get(name, 0.0) + t1 - t0
        t0 = t1

    n, B = sim_specs.n_qubits, sim_specs.n_circuits
    # Sharded run (one process per GPU): every rank builds the (cheap) host arrays of the WHOLE batch from common
    # seeds, lowers / simulates only its contiguous slice of circuits, and one all_gather assembles the histograms.
    # In density-matrix mode the result is independent of the number of ranks, bit for bit.
    import torch.distributed as dist
    from . import distributed as qdist
    world = dist.get_world_size() if (distributed and dist.is_available() and dist.is_initialized()) else 1
    seeds = [random.getrandbits(63), random.getrandbits(63), sampler._next_seed()]
    if world > 1:
        dist.broadcast_object_list(seeds, src=0)
    lo, hi = qdist.circuit_slice(B) if world > 1 else (0, B)
    edges, n_edges = graph_state.sample_batch(B, min_vertex_cover=graph_specs.min_vertex_cover,
                                              max_itr=graph_specs.max_itr, seed=seeds[0])
    lap("graphs")
    rng = np.random.default_rng(seeds[1])
    eng = get_engine(sampler.device)
    build_options = sim_specs.circuit.build_options or {}
    if isinstance(sampler, DeviceNoiseSampler):
        enc = batch.build_gate_lists(builder, edges, n_edges, **build_options)
        lap("gate_lists")
        props = batch.draw_device_props(sampler.noise_specs, enc, rng)
        t1t2, gp, readout_all = batch.device_noise_inputs(props)
        sampler.backend_props = props
        lap("calibration_draws")
        noisy = True
    else:
        # UnitaryNoiseSampler: a coin per potential site, two U(0, max_prob) damping draws per site
        n_slots = (int(n_edges.sum()), batch.sites_per_edge(builder))
        coins = rng.integers(0, 2, size=n_slots) if sim_specs.circuit.stochastic else np.zeros(n_slots, dtype=np.int64)
        gammas = rng.uniform(0.0, sampler.noise_specs.specs.get("max_prob", 0.1), size=(int(coins.sum()), 2))
        enc = batch.build_gate_lists(builder, edges, n_edges, coins=coins, gammas=gammas, **build_options)
        lap("gate_lists")
        t1t2 = gp = readout_all = None
        noisy = bool(coins.any())

    def upload_range(e, c0, c1):
        """Circuits [c0, c1) into engine e -> their readout rows."""
        mine, (o_lo, o_hi) = batch.slice_programs(enc, c0, c1)
        if t1t2 is not None:
            e.upload_device_noise(mine, t1t2[c0:c1], gp[o_lo:o_hi])
        else:
            e.upload(mine)
        return readout_all[c0:c1] if readout_all is not None else None

    seed = seeds[2]
    method = sampler._pick_method(n, noisy)
    # Large density-matrix batches run in chunks on two alternating contexts: while one chunk's kernels run, the host
    # lowers the next one (lowering, not the GPU, bounds this path).  Counters use global circuit indices, so the
    # result does not depend on the chunking.
    n_chunks = 4 if (method == "density_matrix" and entangle_specs is None and hi - lo >= PIPELINE_MIN) else 1
    if hi == lo:
        import torch
        readout = None
        hist = torch.zeros((0, 2 ** n), dtype=torch.int32, device=eng.device)
    elif n_chunks > 1:
        import torch
        engines = [eng, get_engine(sampler.device, slot=1)]
        hist = torch.zeros((hi - lo, 2 ** n), dtype=torch.int32, device=eng.device)
        sampler.last_probs = None
        for k in range(n_chunks):
            c0, c1 = lo + (hi - lo) * k // n_chunks, lo + (hi - lo) * (k + 1) // n_chunks
            e = engines[k % 2]
            ro = upload_range(e, c0, c1)
            probs = e.run_density_matrix(sampler.precision)
            e.sample_from_probs(probs, sampler.n_shots, seed=seed, first_circuit=c0, readout=ro, hist=hist[c0 - lo:c1 - lo])
        readout = readout_all[lo:hi] if readout_all is not None else None
    else:
        readout = upload_range(eng, lo, hi)
        lap("upload")
        if method == "density_matrix":
            probs = eng.run_density_matrix(sampler.precision)
            sampler.last_probs = probs
            hist = eng.sample_from_probs(probs, sampler.n_shots, seed=seed, first_circuit=lo, readout=readout)
        else:
            sampler.last_probs = None
            # the rank uploaded only its slice [lo, hi) of the circuits: first_circuit keeps the Philox counters on the
            # GLOBAL circuit index, so the assembled histograms do not depend on the number of ranks
            hist, _ = eng.run_trajectories(sampler.n_shots, seed=seed, precision=sampler.precision, readout=readout,
                                           first_circuit=lo)
    lap("lower_and_enqueue")       # channels + lowering on host threads, program upload, kernel launches
    # host-only stages while the GPU works on what was just enqueued
    adj = None
    if adj_tensor_specs is not None:
        enc_names = adj_tensor_specs.encoding
        if enc_names is None:
            names = sorted({batch._GATE_NAMES[int(c)] for c in np.unique(enc[0]["opcode"]) if int(c) in batch._GATE_NAMES})
            enc_names = {nm: i + 1 for i, nm in enumerate(names)}
        adj = batch.adjacency_tensors_batch(enc, enc_names, adj_tensor_specs.tensor_shape, adj_tensor_specs.directed)
        lap("adjacency_tensors")
    with _gc_paused():
        graph_lists = batch.edge_lists(edges, n_edges)
    lap("graph_lists")
    if world > 1:
        hist = qdist.gather_rows(hist.contiguous())
        lap("all_gather")
    ent = None
    if entangle_specs is not None:
        # ---- stage 2: the n + 1 Toth settings of every circuit, read off ONE density-matrix simulation per circuit
        import torch
        S = n + 1
        stab_sampler = entangle_specs.stabilizer_sampler(n_shots=entangle_specs.n_shots, backend=sim_specs.backend)
        sop = np.empty((hi - lo, S, 4, 4), dtype=np.complex128)
        sop[:] = np.kron(batch._H.conj(), batch._H)
        ro2 = None
        if not entangle_specs.noisy:      # ideal settings: the noise-free gate lists
            clean = enc if isinstance(sampler, DeviceNoiseSampler) else batch.build_gate_lists(
                builder, edges, n_edges, coins=np.zeros(n_slots, dtype=np.int64), **build_options)
            eng.upload(batch.slice_programs(clean, lo, hi)[0])
        elif isinstance(sampler, DeviceNoiseSampler):
            # the final H of generator k carries the error of this circuit's (h, (k,)) calibration entry
            if not isinstance(builder, CPhaseGateCircuit):          # CX / CZ lists start with H on every qubit
                first = enc[1][lo:hi].astype(np.int64)[:, None] + np.arange(n)[None, :]
                sop[:, :n] = batch.device_gate_superops_1q(batch._H, gp[first, 0], gp[first, 1], t1t2[lo:hi, :, 0],
                                                           t1t2[lo:hi, :, 1])
            ro2 = np.repeat(readout, S, axis=0) if readout is not None else None
        if hi > lo:
            q = np.tile(np.append(np.arange(n, dtype=np.uint8), np.uint8(0xFF)), (hi - lo, 1))
            sp = eng.run_density_matrix_settings(q, sop, stab_sampler.precision).reshape((hi - lo) * S, 2 ** n)
            h2 = eng.sample_from_probs(sp, stab_sampler.n_shots, seed=seeds[2] ^ 0x5DEECE66D, first_circuit=lo * S,
                                       readout=ro2)
            masks = batch.toth_masks(edges[lo:hi], n_edges[lo:hi], n).reshape(-1, 1)
            ss, tot = eng.parity_counts(h2, masks)
            par = torch.stack([ss[:, 0], tot], dim=1).reshape(hi - lo, 2 * S)
        else:
            par = torch.zeros((0, 2 * S), dtype=torch.int64, device=eng.device)
        if world > 1:
            par = qdist.gather_rows(par.contiguous())
        par = par.cpu().numpy().reshape(B, S, 2)
        kind = "genuine" if entangle_specs.witness is GenuineWitness else "biseparable"
        ent = batch.toth_witnesses(kind, par[..., 0], par[..., 1], edges, n_edges, n)
        lap("entanglement")
    prob = _to_host(hist).astype(np.float64) / sampler_specs.n_shots
    lap("gpu_and_fetch")
    results = {}
    with _gc_paused():
        for num, graph_data in enumerate(graph_lists):
            results[num] = {"graph_data": graph_data, "prob_vec": prob[num]}
            if ent is not None:
                results[num]["entanglement"] = {"value": ent[0][num], "variance": ent[1][num]}
            if adj is not None:
                a = adj[num]
                results[num]["adj_tensor"] = a if adj_tensor_specs.fixed_size else a[np.any(a != 0, axis=(1, 2))]
    lap("results")
    if diagnostics is not None:      # per call, owned by the caller: nothing is published through module state
        batch = {"edges": edges, "n_edges": n_edges, "programs": enc}
        if isinstance(sampler, DeviceNoiseSampler):
            batch["device_props"] = props
        else:
            batch.update(coins=coins, gammas=gammas)
        diagnostics["timing"] = timing
        diagnostics["batch"] = batch
    return results


PIPELINE_MIN = 2048    # density-matrix batches of at least this many circuits run as 4 pipelined chunks


def simulate(sim_specs, sampler_specs, graph_specs, adj_tensor_specs=None, entangle_specs=None, *, route="auto",
             distributed=False, diagnostics=None) -> Dict:
    """-> {i: {"graph_data": edges, "prob_vec": np.ndarray (2^n,), ["adj_tensor"], ["entanglement": {"value",
    "variance"}]}}.  route: "auto" (whole-batch array route where it applies, see _array_route_ok) or "objects"
    (always build a Circuit + NoiseModel per circuit, like the reference's loop, simulate.py:276-322).
    distributed=True (array route, torch.distributed initialised, one process per GPU): every rank simulates its
    contiguous slice of the circuits and all ranks return the complete result (one all_gather) -- the reference's
    mp.Pool.map + np.concatenate (older_version/circuit_sampling_utils.py:81-89).
    diagnostics: an optional dict the array route fills with "timing" (seconds per host stage) and "batch" (the drawn
    inputs: graphs, encoded programs, calibration numbers / coins + dampings) of THIS call."""
    if not isinstance(sim_specs, SimulationSpecs):
        raise AssertionError("sim_specs must be an instance of SimulationSpecs")
    if not isinstance(sampler_specs, SamplingSpecs):
        raise AssertionError("sampler_specs must be an instance of SamplingSpecs")
    if not isinstance(graph_specs, GraphSpecs):
        raise AssertionError("graph_specs must be an instance of GraphSpecs")
    if adj_tensor_specs is not None and not isinstance(adj_tensor_specs, AdjTensorSpecs):
        raise AssertionError("adj_tensor_specs must be an instance of AdjTensorSpecs")
    if entangle_specs is not None and not isinstance(entangle_specs, EntanglementSpecs):
        raise AssertionError("entangle_specs must be an instance of EntanglementSpecs")
    n = sim_specs.n_qubits
    graph_db = GraphDB(graph_data=sim_specs.data, directed=graph_specs.directed)
    graph_state = GraphState(graph_db, n_qubits=n, min_num_edges=graph_specs.min_num_edges,
                             max_num_edges=graph_specs.max_num_edges, directed=graph_specs.directed)
    builder = sim_specs.circuit.builder(n_qubits=n, stochastic=sim_specs.circuit.stochastic)
    sampler = sampler_specs.sampler(n_shots=sampler_specs.n_shots, backend=sim_specs.backend,
                                    noise_specs=sampler_specs.noise_specs)
    if sim_specs.circuit.stochastic and not isinstance(sampler, UnitaryNoiseSampler):
        raise AssertionError("stochastic gates need sampler=UnitaryNoiseSampler")

    if route not in ("auto", "objects"):
        raise ValueError("route must be 'auto' or 'objects'")
    if route == "auto" and _array_route_ok(sim_specs, sampler, builder, graph_specs, entangle_specs):
        return _simulate_arrays(sim_specs, sampler_specs, graph_specs, adj_tensor_specs, graph_state, builder, sampler,
                                distributed=distributed, entangle_specs=entangle_specs, diagnostics=diagnostics)
    if distributed:
        raise NotImplementedError("distributed=True is implemented for the array route (see _array_route_ok)")

    results = {i: {} for i in range(sim_specs.n_circuits)}
    graphs, circuits, models = [], [], []
    for num in range(sim_specs.n_circuits):
        graph = graph_state.sample(min_vertex_cover=graph_specs.min_vertex_cover, max_itr=graph_specs.max_itr)
        results[num]["graph_data"] = list(graph.edges())
        cd = builder.build(graph, **sim_specs.circuit.build_options) if sim_specs.circuit.build_options \
            else builder.build(graph)
        if isinstance(sampler, UnitaryNoiseSampler):
            sampler.build_noise_model(cd["ops"])
        else:
            sampler.transpile_circuit(cd["circuit"], {})
            sampler.build_noise_model(n, **(sampler_specs.sample_options or {}))
        graphs.append(graph)
        circuits.append(cd["circuit"])
        models.append(sampler.noise_model)      # a fresh object per build_noise_model call

    if adj_tensor_specs is not None:
        # gate-list walk (no transpiler DAG); codes enumerate the gate names that occur, from 1 (0 = empty slot)
        from .dataset import generate_adjacency_tensor
        enc = adj_tensor_specs.encoding
        if enc is None:
            names = sorted({op[0] for c in circuits for op in c.ops if op[0] not in ("barrier", "measure_all")})
            enc = {nm: i + 1 for i, nm in enumerate(names)}
        for num in range(sim_specs.n_circuits):
            results[num]["adj_tensor"] = generate_adjacency_tensor(
                circuits[num], enc, adj_tensor_specs.tensor_shape, adj_tensor_specs.fixed_size, adj_tensor_specs.directed)

    # ---- stage 1: all noisy circuits in one engine call
    prob_vecs = _batched_counts(sampler, circuits, models, fetch=True)
    for num in range(sim_specs.n_circuits):
        results[num]["prob_vec"] = prob_vecs[num].hist.astype(np.float64) / sampler_specs.n_shots

    # ---- stage 2: all stabilizer settings of all circuits in one engine call
    if entangle_specs is not None:
        stab_sampler = entangle_specs.stabilizer_sampler(n_shots=entangle_specs.n_shots, backend=sim_specs.backend)
        all_circs, all_models, spans, stab_dicts = [], [], [], []
        for num in range(sim_specs.n_circuits):
            stab_circs = entangle_specs.stabilizer(graphs[num], n).build()
            stab_dicts.append(stab_circs)
            spans.append((len(all_circs), len(all_circs) + len(stab_circs)))
            for name, sc in stab_circs.items():
                all_circs.append(circuits[num].copy(name=name).compose(sc))
                all_models.append(models[num] if entangle_specs.noisy else None)
        stab_sampler.noisy = entangle_specs.noisy
        counts = None
        if entangle_specs.noisy and stab_sampler.settings_from_one_state and \
                stab_sampler._pick_method(n, True) == "density_matrix":
            # density-matrix route: one simulation per circuit, all its settings read off the density matrix
            from .stabilizers import run_settings_batch, single_qubit_settings
            settings = [single_qubit_settings(sd, m, n) for sd, m in zip(stab_dicts, models)]
            if all(st is not None for st in settings) and len({len(st[0]) for st in settings}) == 1:
                counts = [c for per in run_settings_batch(stab_sampler, circuits, models, settings) for c in per]
        if counts is None:
            counts = _batched_counts(stab_sampler, all_circs, all_models)
        for num in range(sim_specs.n_circuits):
            lo, hi = spans[num]
            wit = entangle_specs.witness(n_qubits=n, stabilizer_circuits=stab_dicts[num],
                                         stabilizer_counts=counts[lo:hi])
            w = wit.estimate(graph=graphs[num])
            if isinstance(w, dict):
                results[num]["entanglement"] = {"value": [x.value for x in w.values()],
                                                "variance": [x.variance for x in w.values()]}
            else:
                results[num]["entanglement"] = {"value": w.value, "variance": w.variance}
    return results


def _batched_counts(sampler: CircuitSampler, circuits, models, fetch=False):
    """Instantiate each circuit with ITS noise model and simulate the whole batch at once."""
    from .samplers import run_batch
    return run_batch(sampler, circuits, models, fetch=fetch)
