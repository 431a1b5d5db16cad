"""Dataset rows straight from the engine: the `(N, 2^n, 2)` [noisy, ideal] probability
arrays the reference's legacy samplers produce one circuit at a time (older_version/circuit_samplers.py:351-372,
circuit_sampling_utils.py:60-61) and its torch datasets consume (dataset.py:88-89, :371-372), plus the gate-list
version of the adjacency-tensor circuit encoding (samplers.py:119-192) that AdjTModel / AdjTAsymModel take as second
input.  Rows stay on the GPU as a torch tensor (the denoiser trains there); `.cpu().numpy()` gives the reference's
`.npy` layout."""
import numpy as np
import torch

from . import distributed as qdist
from .ir import Circuit
from .samplers import UnitaryNoiseSampler, get_engine, upload_batch

__all__ = ["sample_prob_rows", "generate_adjacency_tensor", "encode_basis_gates", "adjacency_tensors"]


def _run_uploaded(eng, n, noisy, readout, sampler, seed):
    if sampler._pick_method(n, noisy) == "density_matrix":
        probs = eng.run_density_matrix(sampler.precision)
        return eng.sample_from_probs(probs, sampler.n_shots, seed=seed, readout=readout)
    return eng.run_trajectories(sampler.n_shots, seed=seed, precision=sampler.precision, readout=readout)[0]


def sample_prob_rows(circuit_dicts, sampler, ideal=True, distributed=False):
    """circuit_dicts: list of builder outputs {"circuit", "ops"}; sampler: a UnitaryNoiseSampler (fresh noise model
    per circuit, like the reference) or any CircuitSampler with a fixed `noise_model`.
    -> float32 tensor on the sampler's GPU, shape (N, 2^n, 2) with [..., 0] = noisy and [..., 1] = ideal
    probabilities (counts / n_shots, both SAMPLED like the reference's two executions), or (N, 2^n) if not ideal.
    distributed=True: every rank builds the same list, simulates its contiguous slice of circuits and the rows are
    all-gathered (one collective)."""
    n_all = len(circuit_dicts)
    lo, hi = qdist.circuit_slice(n_all) if distributed else (0, n_all)
    mine = circuit_dicts[lo:hi]
    eng = get_engine(sampler.device)
    n = circuit_dicts[0]["circuit"].n_qubits
    if not mine:
        rows = torch.zeros((0, 2 ** n * (2 if ideal else 1)), dtype=torch.float32, device=eng.device)
    else:
        models = []
        for cd in mine:
            if isinstance(sampler, UnitaryNoiseSampler):
                sampler.build_noise_model(cd["ops"])
            models.append(sampler.noise_model if (sampler.noise_model is not None and sampler.noisy) else None)
        # every circuit under ITS model in one upload (calibration-defined models: channels written natively)
        readout, is_noisy = upload_batch(eng, [cd["circuit"] for cd in mine], models)
        noisy = _run_uploaded(eng, n, is_noisy, readout, sampler, sampler._next_seed()).to(torch.float32) / sampler.n_shots
        if ideal:
            clean = [Circuit(c["circuit"].n_qubits, c["circuit"].ops) for c in mine]
            eng.upload(clean)
            ide = _run_uploaded(eng, n, False, None, sampler, sampler._next_seed()).to(torch.float32) / sampler.n_shots
            rows = torch.stack([noisy, ide], dim=-1).reshape(len(mine), -1)
        else:
            rows = noisy
    if distributed:
        rows = qdist.gather_rows(rows.contiguous())
    return rows.reshape(n_all, 2 ** n, 2) if ideal else rows.reshape(n_all, 2 ** n)


def encode_basis_gates(basis_gates):
    """gate name -> integer code, enumerated from 0 exactly like the reference (samplers.py:115-116) -- note that the
    first gate therefore encodes as 0, which the tensor cannot tell from "empty" (reference behaviour, kept)."""
    return {gate: num for num, gate in enumerate(basis_gates)}


def generate_adjacency_tensor(circuit, encoding, tensor_shape=(32, 32, 32), fixed_size=True, directed=False):
    """Adjacency-tensor encoding of a gate list: walk the gates in order; a 1-qubit gate on q writes its code at
    [p, q, q], a 2-qubit gate on (a, b) at [p, a, b] (and [p, b, a] unless directed), p = first plane whose slot is
    still 0.  Same integer scatter as the reference's DAG walk (samplers.py:119-192); without a transpiler the walk
    is over the circuit as written.  Gates missing from `encoding` raise KeyError, as there."""
    adj = np.zeros(tensor_shape)
    planes = adj.shape[0]
    for op in circuit.ops:
        if op[0] in ("barrier", "measure_all", "kraus", "pauli", "reset"):
            continue
        qs = op[1]
        code = encoding[op[0]]
        r, c = (qs[0], qs[0]) if len(qs) == 1 else (qs[0], qs[1])
        for p in range(planes):
            if adj[p, r, c] == 0:
                adj[p, r, c] = code
                if len(qs) == 2 and not directed:
                    adj[p, c, r] = code
                break
    if not fixed_size:
        adj = np.array([pl for pl in adj if np.any(pl != 0)])
    return adj


def adjacency_tensors(circuit_dicts, encoding, tensor_shape=(32, 32, 32), directed=False):
    """(N, planes, q, q) stack for a batch of builder outputs."""
    return np.stack([generate_adjacency_tensor(cd["circuit"], encoding, tensor_shape, True, directed)
                     for cd in circuit_dicts])
