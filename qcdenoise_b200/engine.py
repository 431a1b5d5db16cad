"""Thin Python handle on a libqcdsim context: torch owns device memory and the stream, the library owns the
kernels.  Everything that computes goes through the C ABI (include/qcdsim.h)."""
import ctypes
import json

import numpy as np
import torch

from . import _lib
from .ir import encode_programs


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        return ctypes.c_void_p(a.data_ptr())
    return a.ctypes.data_as(ctypes.c_void_p)


class Engine:
    """One context per (process, device).  Not thread-safe."""

    def __init__(self, device=0, mem_budget_bytes=0):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("qcdenoise_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.zeros(1, device=self.device)   # make sure the primary context exists before the library uses it
        h = ctypes.c_void_p()
        rc = self.lib.qcd_create(self.device.index, mem_budget_bytes, ctypes.byref(h))
        if rc:
            raise _lib.QcdError(rc, self.lib.qcd_last_error(None).decode())
        self._h = h
        self.n_qubits = 0
        self.n_circuits = 0

    def close(self):
        if getattr(self, "_h", None):
            self.lib.qcd_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise _lib.QcdError(rc, self.lib.qcd_last_error(self._h).decode())

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_option(self, key, value):
        self._check(self.lib.qcd_set_option(self._h, key.encode(), float(value)))

    def stats(self):
        s = _lib.Stats()
        self._check(self.lib.qcd_get_stats(self._h, ctypes.byref(s)))
        return s.as_dict()

    # ---------------------------------------------------------------- programs
    def upload(self, circuits):
        """circuits: list of ir.Circuit (same qubit count), or the tuple returned by ir.encode_programs."""
        enc = circuits if isinstance(circuits, tuple) else encode_programs(circuits)
        ops, offsets, params, n = enc
        self._check(self.lib.qcd_upload_programs(self._h, _ptr(ops) if len(ops) else None, _ptr(offsets),
                                                 _ptr(params) if len(params) else None, len(params),
                                                 len(offsets) - 1, n))
        self.n_qubits = n
        self.n_circuits = len(offsets) - 1
        return self

    def upload_device_noise(self, circuits, qubit_t1_t2, op_gate_props):
        """Noise-free circuits (list of ir.Circuit or an encode_programs tuple) + calibration numbers; the library
        writes the device-noise channels after every gate natively (qcd_upload_device_noise_programs).
        qubit_t1_t2: float64 [B, n, 2] (us); op_gate_props: float64 [total ops, 2] (gate_error, gate_length ns;
        NaN row = no error on that op)."""
        enc = circuits if isinstance(circuits, tuple) else encode_programs(circuits)
        ops, offsets, params, n = enc
        B = len(offsets) - 1
        tq = np.ascontiguousarray(qubit_t1_t2, dtype=np.float64)
        gp = np.ascontiguousarray(op_gate_props, dtype=np.float64)
        if tq.shape != (B, n, 2) or gp.shape != (len(ops), 2):
            raise ValueError(f"qubit_t1_t2 must be ({B}, {n}, 2) and op_gate_props ({len(ops)}, 2)")
        self._check(self.lib.qcd_upload_device_noise_programs(
            self._h, _ptr(ops) if len(ops) else None, _ptr(offsets), _ptr(params) if len(params) else None,
            len(params), B, n, _ptr(tq), _ptr(gp)))
        self.n_qubits = n
        self.n_circuits = B
        return self

    @staticmethod
    def _readout_array(readout, n_vec, n):
        if readout is None:
            return None
        r = np.ascontiguousarray(np.asarray(readout, dtype=np.float64))
        if r.shape == (n, 2, 2):
            r = np.ascontiguousarray(np.broadcast_to(r, (n_vec, n, 2, 2)))
        if r.shape != (n_vec, n, 2, 2):
            raise ValueError(f"readout must have shape ({n_vec}, {n}, 2, 2)")
        return r

    # ---------------------------------------------------------------- simulation
    def run_trajectories(self, shots, seed=1234, precision=32, shot_range=None, readout=None, hist=None,
                         want_outcomes=False, dedup=True, first_circuit=0, hist_row0=0):
        """Statevector trajectories.  Returns (hist [B, 2^n] int32 or None, outcomes [range] int32 or None).
        shot_range = (begin, end) in the flattened [circuit][shot] space (default: everything).
        first_circuit: index of uploaded circuit 0 in the whole batch (a rank that uploaded only its slice of the
        circuits draws exactly what a single-GPU run of the whole batch draws).
        hist_row0: uploaded circuit that row 0 of `hist` stands for (a caller running one slice of the uploaded circuits
        passes a histogram with rows for that slice only)."""
        n, B = self.n_qubits, self.n_circuits
        begin, end = shot_range if shot_range is not None else (0, B * shots)
        if hist_row0 and hist is None:
            raise ValueError("hist_row0 needs a caller-owned hist")
        if hist is not None and begin < end and \
                not (hist_row0 <= begin // shots and (end - 1) // shots < hist_row0 + hist.shape[0]):
            raise ValueError("hist has no rows for some circuits of the shot range")
        with torch.cuda.device(self.device):
            if hist is None and n <= 26:
                hist = torch.zeros((B, 2 ** n), dtype=torch.int32, device=self.device)
            outcomes = torch.empty((max(end - begin, 1),), dtype=torch.int32, device=self.device) \
                if (want_outcomes or hist is None) else None
            r = self._readout_array(readout, B, n)
            self.set_option("hist_row0", hist_row0)
            try:
                self._check(self.lib.qcd_run_sv_trajectories(
                    self._h, precision, shots, seed, begin, end, first_circuit, _ptr(r), _ptr(hist), _ptr(outcomes),
                    0 if dedup else _lib.QCD_RUN_NO_DEDUP, self._stream()))
            finally:
                self.set_option("hist_row0", 0)
        if outcomes is not None:
            outcomes = outcomes[: end - begin]
        return hist, outcomes

    def run_density_matrix(self, precision=32, circuit_range=None, out=None):
        """Exact outcome distributions (diagonal of rho): tensor [count, 2^n] float32/float64."""
        n, B = self.n_qubits, self.n_circuits
        cb, ce = circuit_range if circuit_range is not None else (0, B)
        dt = torch.float64 if precision == 64 else torch.float32
        with torch.cuda.device(self.device):
            if out is None:
                out = torch.empty((ce - cb, 2 ** n), dtype=dt, device=self.device)
            self._check(self.lib.qcd_run_density_matrix(self._h, precision, cb, ce, _ptr(out), self._stream()))
        return out

    def run_density_matrix_settings(self, setting_qubits, setting_superops, precision=32, circuit_range=None):
        """Outcome distributions of several measurement settings per circuit from ONE density-matrix simulation each
        (qcd_run_density_matrix_settings).  setting_qubits: uint8 [count, S] (0xFF = measure as is);
        setting_superops: complex128 [count, S, 4, 4], the final single-qubit step of each setting as a superoperator
        on (ket, bra), index = ket + 2 * bra.  -> tensor [count, S, 2^n] float32 / float64."""
        n, B = self.n_qubits, self.n_circuits
        cb, ce = circuit_range if circuit_range is not None else (0, B)
        q = np.ascontiguousarray(setting_qubits, dtype=np.uint8)
        sop = np.ascontiguousarray(setting_superops, dtype=np.complex128)
        if q.ndim != 2 or q.shape[0] != ce - cb or sop.shape != q.shape + (4, 4):
            raise ValueError(f"setting_qubits must be ({ce - cb}, S) and setting_superops ({ce - cb}, S, 4, 4)")
        S = q.shape[1]
        dt = torch.float64 if precision == 64 else torch.float32
        with torch.cuda.device(self.device):
            out = torch.empty((ce - cb, S, 2 ** n), dtype=dt, device=self.device)
            self._check(self.lib.qcd_run_density_matrix_settings(self._h, precision, cb, ce, S, _ptr(q),
                                                                 _ptr(sop.view(np.float64)), _ptr(out), self._stream()))
        return out

    def run_density_matrix_group(self, adjacency, precision=32, circuit_range=None):
        """Exact <S> of ALL 2^n elements of a graph state's stabilizer group, read off each circuit's density matrix in
        one pass (qcd_run_density_matrix_group).  adjacency: uint32 [n], bit j of row k = edge (k, j).
        -> tensor [count, 2^n] float64 indexed by the element's x mask (element = product of the generators g_k over the
        set bits); column 0 is the identity, the row mean is the fidelity with the ideal graph state."""
        n, B = self.n_qubits, self.n_circuits
        cb, ce = circuit_range if circuit_range is not None else (0, B)
        adj = np.ascontiguousarray(adjacency, dtype=np.uint32)
        if adj.shape != (n,):
            raise ValueError(f"adjacency must hold {n} rows")
        with torch.cuda.device(self.device):
            out = torch.empty((ce - cb, 2 ** n), dtype=torch.float64, device=self.device)
            self._check(self.lib.qcd_run_density_matrix_group(self._h, precision, cb, ce, _ptr(adj), _ptr(out),
                                                              self._stream()))
        return out

    def stabilizer_group(self, gen_x, gen_z, begin=0, end=None):
        """Elements [begin, end) of the signed stabilizer group generated by (gen_x[j], gen_z[j]), enumerated on the
        device (qcd_stabilizer_group) -> (x uint32, z uint32, sign uint8) tensors; sign 0 = '+', 1 = '-'."""
        gx = np.ascontiguousarray(gen_x, dtype=np.uint32)
        gz = np.ascontiguousarray(gen_z, dtype=np.uint32)
        if gx.ndim != 1 or gx.shape != gz.shape:
            raise ValueError("gen_x and gen_z must be 1-d arrays of the same length")
        end = 2 ** len(gx) if end is None else end
        with torch.cuda.device(self.device):
            x = torch.empty((max(end - begin, 0),), dtype=torch.int32, device=self.device)
            z = torch.empty_like(x)
            s = torch.empty((max(end - begin, 0),), dtype=torch.uint8, device=self.device)
            self._check(self.lib.qcd_stabilizer_group(self._h, _ptr(gx), _ptr(gz), len(gx), begin, end, _ptr(x), _ptr(z),
                                                      _ptr(s), self._stream()))
        if bool((s == 2).any()):
            raise ArithmeticError("non-Hermitian product in a stabilizer group (the generators do not commute)")
        return x, z, s

    def apply_readout(self, probs, readout):
        n_vec, N = probs.shape
        n = N.bit_length() - 1
        r = self._readout_array(readout, n_vec, n)
        prec = 64 if probs.dtype == torch.float64 else 32
        with torch.cuda.device(self.device):
            self._check(self.lib.qcd_apply_readout_probs(self._h, prec, _ptr(probs), n_vec, n, _ptr(r), self._stream()))
        return probs

    def sample_from_probs(self, probs, shots, seed=1234, first_circuit=0, shot_range=None, readout=None, hist=None):
        n_vec, N = probs.shape
        n = N.bit_length() - 1
        begin, end = shot_range if shot_range is not None else (0, shots)
        prec = 64 if probs.dtype == torch.float64 else 32
        with torch.cuda.device(self.device):
            if hist is None:
                hist = torch.zeros((n_vec, N), dtype=torch.int32, device=self.device)
            r = self._readout_array(readout, n_vec, n)
            self._check(self.lib.qcd_sample_from_probs(self._h, prec, _ptr(probs), n_vec, n, begin, end, seed,
                                                       first_circuit, _ptr(r), _ptr(hist), self._stream()))
        return hist

    # ---------------------------------------------------------------- stabilizer reductions
    @staticmethod
    def _masks(masks, n_vec):
        m = np.ascontiguousarray(np.asarray(masks, dtype=np.uint32))
        if m.ndim == 1:
            return m, m.shape[0], 0
        if m.shape[0] != n_vec:
            raise ValueError("per-vector masks must have one row per vector")
        return m, m.shape[1], 1

    def parity_counts(self, hist, masks):
        """-> (signed_sums int64 [B, M], totals int64 [B]); bit-exact integer arithmetic."""
        n_vec, N = hist.shape
        n = N.bit_length() - 1
        m, n_masks, per = self._masks(masks, n_vec)
        with torch.cuda.device(self.device):
            ss = torch.empty((n_vec, n_masks), dtype=torch.int64, device=self.device)
            tot = torch.empty((n_vec,), dtype=torch.int64, device=self.device)
            self._check(self.lib.qcd_parity_counts(self._h, _ptr(hist), n_vec, n, _ptr(m), n_masks, per, _ptr(ss),
                                                   _ptr(tot), self._stream()))
        return ss, tot

    def parity_outcomes(self, outcomes, masks):
        """outcomes: int32 tensor [B, shots] of basis indices -> signed_sums int64 [B, M] (sparse form of
        parity_counts for registers too wide for dense histograms)."""
        n_vec, shots = outcomes.shape
        m, n_masks, per = self._masks(masks, n_vec)
        with torch.cuda.device(self.device):
            ss = torch.empty((n_vec, n_masks), dtype=torch.int64, device=self.device)
            self._check(self.lib.qcd_parity_outcomes(self._h, _ptr(outcomes), n_vec, shots, _ptr(m), n_masks, per,
                                                     _ptr(ss), self._stream()))
        return ss

    def parity_probs(self, probs, masks):
        n_vec, N = probs.shape
        n = N.bit_length() - 1
        m, n_masks, per = self._masks(masks, n_vec)
        prec = 64 if probs.dtype == torch.float64 else 32
        with torch.cuda.device(self.device):
            ev = torch.empty((n_vec, n_masks), dtype=torch.float64, device=self.device)
            self._check(self.lib.qcd_parity_probs(self._h, prec, _ptr(probs), n_vec, n, _ptr(m), n_masks, per,
                                                  _ptr(ev), self._stream()))
        return ev


def debug_schedule(circuit, mode=0, precision=32, tile_bits=0):
    """Host-only: the sweep program the scheduler emits for one circuit, as a dict (no GPU needed)."""
    lib = _lib.load()
    ops, offsets, params, n = encode_programs([circuit])
    cap = 1 << 20
    while True:
        buf = ctypes.create_string_buffer(cap)
        need = ctypes.c_size_t()
        rc = lib.qcd_debug_schedule_json(_ptr(ops) if len(ops) else None, len(ops),
                                         _ptr(params) if len(params) else None, len(params), n, mode, precision,
                                         tile_bits, buf, cap, ctypes.byref(need))
        if rc == _lib.QCD_E_NOMEM and need.value > cap:
            cap = need.value + 16
            continue
        d = json.loads(buf.value.decode())
        if rc:
            raise _lib.QcdError(rc, d.get("error", "schedule failed"))
        return d


def debug_sibling_runs(circuits, precision=32):
    """Host-only: the sibling runs the engine forms for this batch in statevector mode (option "leaf_cross") -> list of
    dicts {lead, members, n_high, reg_bits, nops, kbit, u} (see include/qcdsim.h; no GPU needed)."""
    lib = _lib.load()
    ops, offsets, params, n = encode_programs(circuits)
    cap = 1 << 16
    while True:
        buf = ctypes.create_string_buffer(cap)
        need = ctypes.c_size_t()
        rc = lib.qcd_debug_sibling_runs_json(_ptr(ops), _ptr(offsets), _ptr(params) if len(params) else None, len(params),
                                             len(circuits), n, precision, buf, cap, ctypes.byref(need))
        if rc == _lib.QCD_E_NOMEM and need.value > cap:
            cap = need.value + 16
            continue
        d = json.loads(buf.value.decode())
        if rc:
            raise _lib.QcdError(rc, d.get("error", "lowering failed"))
        return d["runs"]


def debug_dm_schedule(circuit):
    """Host-only: the pass schedule of the on-chip density-matrix interpreter for one circuit ->
    [((q0, q1 or None), [op indices in execution order]), ...] (no GPU needed)."""
    lib = _lib.load()
    ops, offsets, params, n = encode_programs([circuit])
    cap = max(len(ops), 1)
    pq = np.zeros((cap, 2), dtype=np.uint8)
    pn = np.zeros(cap, dtype=np.uint32)
    order = np.zeros(cap, dtype=np.uint32)
    n_p, n_o = ctypes.c_size_t(), ctypes.c_size_t()
    rc = lib.qcd_debug_dm_schedule(_ptr(ops) if len(ops) else None, len(ops), _ptr(params) if len(params) else None,
                                   len(params), n, _ptr(pq), _ptr(pn), cap, _ptr(order), cap, ctypes.byref(n_p),
                                   ctypes.byref(n_o))
    if rc:
        raise _lib.QcdError(rc, lib.qcd_last_error(None).decode())
    out, pos = [], 0
    for p in range(n_p.value):
        k = int(pn[p])
        out.append(((int(pq[p, 0]), None if pq[p, 1] == 0xFF else int(pq[p, 1])), [int(i) for i in order[pos:pos + k]]))
        pos += k
    assert pos == n_o.value
    return out


def expand_device_noise(enc, qubit_t1_t2, op_gate_props):
    """Host-only: the explicit programs the library builds from noise-free gate lists + calibration numbers
    (qcd_expand_device_noise) -> (ops, offsets, params, n) like ir.encode_programs.  No GPU needed."""
    lib = _lib.load()
    ops, offsets, params, n = enc
    B = len(offsets) - 1
    tq = np.ascontiguousarray(qubit_t1_t2, dtype=np.float64)
    gp = np.ascontiguousarray(op_gate_props, dtype=np.float64)
    if tq.shape != (B, n, 2) or gp.shape != (len(ops), 2):
        raise ValueError(f"qubit_t1_t2 must be ({B}, {n}, 2) and op_gate_props ({len(ops)}, 2)")
    n_ops, n_par = ctypes.c_size_t(), ctypes.c_size_t()
    out_ops = np.zeros(0, dtype=_lib.OP_DTYPE)
    out_par = np.zeros(0, dtype=np.float64)
    out_off = np.zeros(B + 1, dtype=np.uint32)
    for _ in range(2):
        rc = lib.qcd_expand_device_noise(_ptr(ops) if len(ops) else None, _ptr(offsets),
                                         _ptr(params) if len(params) else None, len(params), B, n, _ptr(tq), _ptr(gp),
                                         _ptr(out_ops) if len(out_ops) else None, len(out_ops), _ptr(out_off),
                                         _ptr(out_par) if len(out_par) else None, len(out_par),
                                         ctypes.byref(n_ops), ctypes.byref(n_par))
        if rc != _lib.QCD_E_NOMEM:
            break
        out_ops = np.zeros(n_ops.value, dtype=_lib.OP_DTYPE)
        out_par = np.zeros(n_par.value, dtype=np.float64)
    if rc:
        raise _lib.QcdError(rc, lib.qcd_last_error(None).decode())
    return out_ops, out_off, out_par, n
