"""Counts -> stabilizer expectation values -> entanglement witnesses (reference qcdenoise/witnesses.py:21-144).
The +-1 diagonals are never materialised for the computation: <S> = sum_x c_x (-1)^popcount(x & mask) / N is one
bitmask-parity reduction on the GPU (csrc/kernels.cu parity_counts_kernel), bit-exact integer arithmetic; the
`diagonals` attribute is built lazily for callers that ask for it."""
import math
from collections import OrderedDict, namedtuple

import networkx as nx
import numpy as np
import torch

from .samplers import Counts, get_engine
from .stabilizers import JungStabilizer, graph_adjacency_masks, pauli_mask

__all__ = ["Witness", "BiSeparableWitness", "GenuineWitness", "witness_output", "group_expectations",
           "fidelity_witness_exact"]

witness_output = namedtuple("Witness", ["W_ij", "value", "variance"])


class Witness:
    def __init__(self, n_qubits, stabilizer_circuits, stabilizer_counts, device=None):
        self.n_qubits = n_qubits
        self.stabilizer_circuits = stabilizer_circuits
        self.stabilizer_counts = stabilizer_counts
        self._device = device
        self._diagonals = None
        self.stabilizer_measurements = self._compute_expval()

    @property
    def diagonals(self):
        if self._diagonals is None:
            self._diagonals = self._construct_diagonals()
        return self._diagonals

    def _construct_diagonals(self):
        x = np.arange(2 ** self.n_qubits, dtype=np.uint64)
        out = OrderedDict()
        for name in self.stabilizer_circuits.keys():
            m = np.uint64(pauli_mask(name))
            v = x & m
            par = np.zeros_like(v)
            while v.any():
                par ^= v & np.uint64(1)
                v >>= np.uint64(1)
            out[name] = 1.0 - 2.0 * par.astype(np.float64)
        return out

    def _compute_expval(self):
        """(mean, std) per setting: mean = sum_x p_x d_x, std = sqrt((sum_x p_x d_x^2 - mean^2) / shots) with
        d_x^2 = 1 -- the qiskit-ignis `expectation_value` the reference calls (witnesses.py:59-67)."""
        names = list(self.stabilizer_circuits.keys())
        n = self.n_qubits
        eng = get_engine(self._device)
        rows = []
        for c in self.stabilizer_counts:
            c = c if isinstance(c, Counts) else Counts.from_dict(c, n)
            if c.device_hist is not None:
                rows.append(c.device_hist.reshape(1, -1))
            else:
                rows.append(torch.as_tensor(c.hist.astype(np.int32)).reshape(1, -1).to(eng.device))
        hist = torch.cat(rows, dim=0).contiguous()
        masks = np.array([[pauli_mask(nm)] for nm in names], dtype=np.uint32)
        ss, tot = eng.parity_counts(hist, masks)
        ss, tot = ss.cpu().numpy()[:, 0], tot.cpu().numpy()
        out = OrderedDict()
        for i, nm in enumerate(names):
            shots = int(tot[i])
            mean = int(ss[i]) / shots
            var = (1.0 - mean * mean) / shots
            out[nm] = (mean, math.sqrt(var) if var > 0 else 0.0)
        return out

    def estimate(self, graph):
        raise NotImplementedError

    @staticmethod
    def _edges(graph):
        if isinstance(graph, nx.DiGraph):
            return [(u, v) for u, nbrs in graph.adjacency() for v in nbrs]
        return list(graph.edges())


class BiSeparableWitness(Witness):
    """W_ij = 1 - <g_i> - <g_j> per edge; sigma = sqrt(std_i^2 + std_j^2)."""

    def estimate(self, graph):
        meas = self.stabilizer_measurements
        res = {}
        for idx, (e0, e1) in enumerate(self._edges(graph)):
            gi = next(k for k in meas if k[::-1][e0] == "X")
            gj = next(k for k in meas if k[::-1][e1] == "X")
            val = 1 - meas[gi][0] - meas[gj][0]
            res[idx] = witness_output(W_ij=(e0, e1), value=val,
                                      variance=math.sqrt(meas[gi][1] ** 2 + meas[gj][1] ** 2))
        return res


class GenuineWitness(Witness):
    """Genuine multipartite entanglement witnesses of a graph state from measured stabilizer expectation values.

    noise_robust = 0  W = (n - 1) <I> - sum_k <g_k>: the generators only (Toth & Guehne, PRA 72, 022340, Eq. 21); this is
                      the one level the reference implements (witnesses.py:112-144).
    noise_robust = 1  two-setting witness of a two-colourable graph, W = 3 - 2 [P_A + P_B] with
                      P_C = prod_{k in C} (1 + g_k) / 2 = 2^-|C| sum_{T subset C} prod_{k in T} g_k over the colour
                      classes A, B (ibid. Eq. 23 for the GHZ / cluster case, what the legacy docstring
                      circuit_constructors.py:1309-1316 points at): needs the 2^|A| + 2^|B| - 1 group elements
                      generated inside one colour class.
    noise_robust = 2  projector witness W = 1/2 - |G><G| with |G><G| = 2^-n sum_S S over the FULL stabilizer group
                      (ibid. Eq. 16; legacy "all B vertex sets", circuit_constructors.py:537-542): needs all 2^n
                      elements (JungStabilizer).  It tolerates the most white noise of the three.
    Levels 1 and 2 are "WIP" in the reference (circuit_constructors.py:677-680), so they are pinned by identities only
    (ideal state, maximally mixed state, W_1 >= 2 W_2) and by the dense-matrix oracle.  `variance` is the propagated
    standard deviation; for levels 1 and 2 the coefficients scale the std-devs as usual."""

    def _signed_group(self, graph):
        signed = JungStabilizer(graph, self.n_qubits).find_stabilizers()
        by_x = {}
        for s in signed:          # graph states: the x mask identifies the element
            by_x[sum(1 << q for q, ch in enumerate(s[:0:-1]) if ch in "XY")] = (s[1:], 1.0 if s[0] == "+" else -1.0)
        return by_x

    def _term(self, label):
        if label not in self.stabilizer_measurements:
            raise KeyError(f"no measurement of stabilizer {label}: this noise_robust level needs more settings "
                           "(build them with JungStabilizer)")
        return self.stabilizer_measurements[label]

    def estimate(self, graph, noise_robust=0):
        if noise_robust not in (0, 1, 2):
            raise AssertionError("noise_robust must be 0, 1 or 2")
        if noise_robust == 2:
            group = self._signed_group(graph)
            dim = float(2 ** self.n_qubits)
            fid = var = 0.0
            for label, sign in group.values():
                m, s = self._term(label)
                fid += sign * m
                var += s * s
            return witness_output(W_ij=None, value=0.5 - fid / dim, variance=math.sqrt(var) / dim)
        if noise_robust == 1:
            und = graph.to_undirected() if isinstance(graph, nx.DiGraph) else graph
            if not nx.is_bipartite(und):
                raise ValueError("noise_robust=1 needs a two-colourable graph")
            colour = nx.bipartite.color(und)
            group = self._signed_group(graph)
            value, var = 3.0, 0.0
            for c in (0, 1):
                verts = [v for v in und.nodes() if colour[v] == c]
                if not verts:
                    value -= 2.0          # empty product = identity
                    continue
                w = 2.0 / float(2 ** len(verts))
                for sub in range(2 ** len(verts)):
                    x = sum(1 << verts[i] for i in range(len(verts)) if (sub >> i) & 1)
                    label, sign = group[x]
                    m, s = self._term(label)
                    value -= w * sign * m
                    var += (w * s) ** 2
            return witness_output(W_ij=None, value=value, variance=math.sqrt(var))
        meas = self.stabilizer_measurements
        id_key = "I" * self.n_qubits
        id_val, _ = meas[id_key]
        if hasattr(self.stabilizer_circuits, "move_to_end"):
            self.stabilizer_circuits.move_to_end(id_key)
        vals = [m for m, _ in meas.values()]
        stds = [s for _, s in meas.values()]
        w = (graph.order() - 1) * id_val - (sum(vals) - id_val)
        # reference: unumpy.uarray(wit_coef * meas_vals, var_vals).sum() -- the coefficients scale the NOMINAL values
        # only, the standard deviations enter unscaled, so sigma = sqrt(sum std_k^2) for every ordering of the
        # measurements (witnesses.py:136-144)
        sigma = math.sqrt(sum(s * s for s in stds))
        return witness_output(W_ij=None, value=w, variance=sigma)


def group_expectations(engine, graph, precision=32, circuit_range=None):
    """Exact <S> of every element of the graph's stabilizer group for the circuits uploaded to `engine`, read off each
    density matrix in one pass (no per-setting simulation, no sampling) -> float64 tensor [count, 2^n] indexed by the
    element's x mask: entry sum(1 << k for k in T) is < prod_{k in T} g_k >."""
    return engine.run_density_matrix_group(graph_adjacency_masks(graph, engine.n_qubits), precision=precision,
                                           circuit_range=circuit_range)


def fidelity_witness_exact(engine, graph, precision=32, circuit_range=None):
    """Projector witness <W> = 1/2 - <G|rho|G> (noise_robust = 2) of every uploaded circuit from its exact density
    matrix -> float64 tensor [count]."""
    return 0.5 - group_expectations(engine, graph, precision, circuit_range).mean(dim=1)
