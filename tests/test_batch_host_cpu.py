"""The whole-batch host stages (qcdenoise_b200/batch.py, GraphState.sample_batch -> qcd_sample_graph_states,
qcd_adjacency_tensors) against the object route they replace.  Host-only: no GPU."""
import random

import networkx as nx
import numpy as np
import pytest
from networkx.algorithms.approximation import vertex_cover

import qcdenoise_b200 as qcd
from oracle import philox
from qcdenoise_b200 import batch
from qcdenoise_b200.dataset import generate_adjacency_tensor
from qcdenoise_b200.engine import expand_device_noise
from qcdenoise_b200.ir import encode_programs
from qcdenoise_b200.samplers import device_noise_arrays


class _Stream:
    """Word j of Philox4x32-10(key = seed, counter = (graph lo, graph hi, attempt, j // 4)) -- the draw convention of
    qcd_sample_graph_states (include/qcdsim.h)."""

    def __init__(self, seed, graph, attempt):
        self.seed, self.graph, self.attempt, self.j = seed, graph, attempt, 0

    def below(self, k):
        blk = philox.philox4x32_10(self.graph & 0xFFFFFFFF, self.graph >> 32, self.attempt, self.j // 4,
                                   self.seed & 0xFFFFFFFF, self.seed >> 32)
        w = int(blk[self.j % 4])
        self.j += 1
        return (w * k) >> 32


def _model_sample(gs, seed, index, min_vertex_cover, max_itr):
    """GraphState.sample restated with networkx operations (disjoint_union, add_edge, min_weighted_vertex_cover)
    and the native sampler's draw order."""
    g = None
    for attempt in range(max(1, max_itr)):
        rng = _Stream(seed, index, attempt)
        comb = gs.graph_combs[rng.below(len(gs.graph_combs))]
        subs = [gs.all_graphs[k][rng.below(len(gs.all_graphs[k]))] for k in comb]
        g = nx.Graph()
        for sg in subs:
            have = g.order()
            if have > 1:
                first = rng.below(have)
                targets = [have + rng.below(sg.order() - 1) for _ in range(sg.order())]
                g = nx.disjoint_union(g, sg)
                g.add_weighted_edges_from((first, t, 1.0) for t in targets)
            else:
                g = nx.disjoint_union(g, sg)
        if len(vertex_cover.min_weighted_vertex_cover(g, weight="weight")) // 2 >= min_vertex_cover:
            break
    return g


@pytest.mark.parametrize("n,cover,itr", [(4, 1, 1), (8, 1, 1), (9, 3, 4), (14, 5, 3), (20, 1, 10)])
def test_native_graph_batch_equals_networkx_model(n, cover, itr):
    gs = qcd.GraphState(qcd.GraphDB(), n_qubits=n, min_num_edges=2, max_num_edges=7)
    seed = 0x1234_5678_9ABC_DEF0 >> 1
    edges, n_edges = gs.sample_batch(40, min_vertex_cover=cover, max_itr=itr, seed=seed, first_graph=3)
    for i in range(40):
        g = _model_sample(gs, seed, 3 + i, cover, itr)
        assert [tuple(e) for e in edges[i, :n_edges[i]].tolist()] == [tuple(e) for e in g.edges()], i
        assert g.order() == n and nx.is_connected(g)
    # any split of the batch gives the same graphs
    e2, k2 = gs.sample_batch(15, min_vertex_cover=cover, max_itr=itr, seed=seed, first_graph=3 + 25)
    assert np.array_equal(k2, n_edges[25:]) and np.array_equal(e2, edges[25:])


def test_graph_batch_statistics_match_object_route():
    """Same distribution as GraphState.sample (different generator): edge-count statistics agree."""
    random.seed(3)
    np.random.seed(3)
    gs = qcd.GraphState(qcd.GraphDB(), n_qubits=8)
    obj = np.array([gs.sample(max_itr=1).number_of_edges() for _ in range(1500)])
    _, ne = gs.sample_batch(20000, max_itr=1, seed=9)
    assert abs(obj.mean() - ne.mean()) < 4 * obj.std() / np.sqrt(len(obj))
    assert abs(obj.std() - ne.std()) < 0.25


def _graphs_from(edges, n_edges, n):
    out = []
    for rows, k in zip(edges, n_edges):
        g = nx.Graph()
        g.add_nodes_from(range(n))
        g.add_edges_from((int(u), int(v)) for u, v in rows[:k])
        out.append(g)
    return out


@pytest.mark.parametrize("cls,opts", [(qcd.CXGateCircuit, {}), (qcd.CZGateCircuit, {}),
                                      (qcd.CPhaseGateCircuit, {}), (qcd.CPhaseGateCircuit, {"Lambda": 1.1, "Theta": 0.4})])
def test_batched_gate_lists_equal_builder(cls, opts):
    n = 7
    gs = qcd.GraphState(qcd.GraphDB(), n_qubits=n)
    edges, n_edges = gs.sample_batch(25, seed=2)
    builder = cls(n_qubits=n, stochastic=False)
    ops, off, par, nn = batch.build_gate_lists(builder, edges, n_edges, **opts)
    want = encode_programs([cls(n_qubits=n, stochastic=False).build(g, **opts)["circuit"]
                            for g in _graphs_from(edges, n_edges, n)])
    assert nn == n and np.array_equal(off, want[1]) and len(ops) == len(want[0])
    for f in ("opcode", "q0", "q1"):
        assert np.array_equal(ops[f], want[0][f])
    has_par = ops["opcode"] == qcd._lib.OPCODES["ry"]
    assert np.array_equal(par[ops["param_off"][has_par]], want[2][want[0]["param_off"][has_par]])
    with pytest.raises(ValueError):
        batch.build_gate_lists(cls(n_qubits=n, stochastic=True), edges, n_edges)


def test_batched_device_props_feed_the_same_programs_as_the_object_route():
    n = 6
    spec = qcd.NoiseSpec(specs={"qubit": {"T1": 100.0, "T2": 150.0, "readout_error": 0.2, "prob_meas0_prep1": 0.1,
                                          "prob_meas1_prep0": 0.05},
                                "cx": {"gate_error": 0.25, "gate_length": 500.0},
                                "h": {"gate_error": 0.05}})          # h has no gate_length -> 1e-6 (quirk Q4)
    gs = qcd.GraphState(qcd.GraphDB(), n_qubits=n)
    edges, n_edges = gs.sample_batch(70, seed=4)
    enc = batch.build_gate_lists(qcd.CXGateCircuit(n_qubits=n, stochastic=False), edges, n_edges)
    props = batch.draw_device_props(spec, enc, np.random.default_rng(1))
    t1t2, gp, ro = batch.device_noise_inputs(props)
    assert (props["gate"]["gate_length"][(props["gate_key"] >> 16) == qcd._lib.OPCODES["h"]] == 1e-6).all()
    assert (props["qubit"]["T2"] <= 150.0).all() and (props["gate"]["gate_error"] <= 0.25).all()
    native = expand_device_noise(enc, t1t2, gp)
    # object route with the same numbers: DeviceNoiseSampler(backend=<dict>) per circuit
    circuits, models = [], []
    for i, g in enumerate(_graphs_from(edges, n_edges, n)):
        c = qcd.CXGateCircuit(n_qubits=n, stochastic=False).build(g)["circuit"]
        s = qcd.DeviceNoiseSampler(batch.device_props_dict(props, i), n_shots=16, noise_specs=spec)
        s.transpile_circuit(c)
        s.build_noise_model(n)
        circuits.append(c)
        models.append(s.noise_model)
    t1t2_o, gp_o, ro_o = device_noise_arrays(circuits, models)
    assert np.array_equal(t1t2, t1t2_o) and np.array_equal(gp, gp_o) and np.array_equal(ro, ro_o)
    obj = expand_device_noise(encode_programs(circuits), t1t2_o, gp_o)
    for x, y in zip(native, obj):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("directed", [False, True])
def test_native_adjacency_tensors_equal_python_walk(directed):
    n = 6
    gs = qcd.GraphState(qcd.GraphDB(), n_qubits=n)
    edges, n_edges = gs.sample_batch(70, seed=6)
    encoding = {"h": 1, "cx": 2}
    enc = batch.build_gate_lists(qcd.CXGateCircuit(n_qubits=n, stochastic=False), edges, n_edges)
    got = batch.adjacency_tensors_batch(enc, encoding, (9, 8, 8), directed)
    for i, g in enumerate(_graphs_from(edges, n_edges, n)):
        c = qcd.CXGateCircuit(n_qubits=n, stochastic=False).build(g)["circuit"]
        assert np.array_equal(got[i], generate_adjacency_tensor(c, encoding, (9, 8, 8), True, directed)), i
    with pytest.raises(KeyError):
        batch.adjacency_tensors_batch(enc, {"h": 1}, (9, 8, 8))
    with pytest.raises(IndexError):
        batch.adjacency_tensors_batch(enc, encoding, (9, 4, 4))


@pytest.mark.parametrize("cls,kind", [(qcd.CXGateCircuit, "phase_amplitude_damping_error"),
                                      (qcd.CZGateCircuit, "amplitude_damping_error"),
                                      (qcd.CPhaseGateCircuit, "phase_amplitude_damping_error")])
def test_batched_stochastic_gate_lists_equal_object_route(cls, kind, monkeypatch):
    """Stochastic builders + UnitaryNoiseSampler: feed the object route the SAME coins and damping draws (by
    patching np.random.choice / np.random.uniform) and compare the explicit programs op for op."""
    n = 6
    gs = qcd.GraphState(qcd.GraphDB(), n_qubits=n)
    edges, n_edges = gs.sample_batch(30, seed=8)
    rng = np.random.default_rng(2)
    builder = cls(n_qubits=n, stochastic=True)
    S = batch.sites_per_edge(builder)
    coins = rng.integers(0, 2, size=(int(n_edges.sum()), S))
    gammas = rng.uniform(0, 0.3, size=(int(coins.sum()), 2))
    got = batch.build_gate_lists(builder, edges, n_edges, coins=coins, gammas=gammas)
    coin_iter, gamma_iter = iter(coins.ravel().tolist()), iter(gammas)
    monkeypatch.setattr(np.random, "choice", lambda k: next(coin_iter))
    monkeypatch.setattr(np.random, "uniform", lambda high, size: next(gamma_iter).copy())
    s = qcd.UnitaryNoiseSampler(None, n_shots=16, noise_specs=qcd.NoiseSpec(specs={"type": kind, "max_prob": 0.3}))
    progs = []
    for g in _graphs_from(edges, n_edges, n):
        cd = cls(n_qubits=n, stochastic=True).build(g)
        s.build_noise_model(cd["ops"])
        progs.append(s.noise_model.instantiate(cd["circuit"])[0])
    want = encode_programs(progs)
    assert next(coin_iter, None) is None and next(gamma_iter, None) is None
    (ops, off, par, _), (wops, woff, wpar, _) = got, want
    assert np.array_equal(off, woff)
    for f in ("opcode", "q0", "q1"):
        assert np.array_equal(ops[f], wops[f])
    K, U, RY = (qcd._lib.OPCODES[k] for k in ("kraus1q", "u2q", "ry"))
    for i in np.nonzero(ops["opcode"] == K)[0]:
        assert np.array_equal(par[ops["param_off"][i]:ops["param_off"][i] + 17], wpar[wops["param_off"][i]:wops["param_off"][i] + 17])
    for i in np.nonzero(ops["opcode"] == U)[0]:
        assert np.array_equal(par[ops["param_off"][i]:ops["param_off"][i] + 32], wpar[wops["param_off"][i]:wops["param_off"][i] + 32])
    ry = ops["opcode"] == RY
    assert np.array_equal(par[ops["param_off"][ry]], wpar[wops["param_off"][ry]])
    assert (ops["opcode"] == K).sum() >= 2 * coins.sum()


def test_slice_programs_is_the_sub_batch():
    n = 6
    gs = qcd.GraphState(qcd.GraphDB(), n_qubits=n)
    edges, n_edges = gs.sample_batch(12, seed=3)
    builder = qcd.CPhaseGateCircuit(n_qubits=n, stochastic=False)
    enc = batch.build_gate_lists(builder, edges, n_edges)
    (ops, off, par, nn), (o_lo, o_hi) = batch.slice_programs(enc, 4, 9)
    want = batch.build_gate_lists(builder, edges[4:9], n_edges[4:9])
    assert nn == n and np.array_equal(off, want[1]) and np.array_equal(ops, want[0]) and np.array_equal(par, want[2])
    assert (o_lo, o_hi) == (int(enc[1][4]), int(enc[1][9]))
    empty, _ = batch.slice_programs(enc, 5, 5)
    assert len(empty[0]) == 0 and empty[1].tolist() == [0]


def test_toth_masks_and_noisy_h_superops_equal_object_route():
    n = 6
    spec = qcd.NoiseSpec(specs={"qubit": {"T1": 100.0, "T2": 150.0, "readout_error": 0.1},
                                "cx": {"gate_error": 0.25, "gate_length": 500.0}, "h": {"gate_error": 0.05, "gate_length": 60.0}})
    gs = qcd.GraphState(qcd.GraphDB(), n_qubits=n)
    edges, n_edges = gs.sample_batch(20, seed=12)
    masks = batch.toth_masks(edges, n_edges, n)
    enc = batch.build_gate_lists(qcd.CXGateCircuit(n_qubits=n, stochastic=False), edges, n_edges)
    props = batch.draw_device_props(spec, enc, np.random.default_rng(4))
    t1t2, gp, _ = batch.device_noise_inputs(props)
    first = enc[1][:-1].astype(np.int64)[:, None] + np.arange(n)[None, :]
    sops = batch.device_gate_superops_1q(batch._H, gp[first, 0], gp[first, 1], t1t2[..., 0], t1t2[..., 1])
    assert sops.shape == (20, n, 4, 4)
    for i, g in enumerate(_graphs_from(edges, n_edges, n)):
        stab = qcd.TothStabilizer(g, n_qubits=n).build()
        assert [qcd.pauli_mask(nm) for nm in stab] == masks[i].tolist()
        c = qcd.CXGateCircuit(n_qubits=n, stochastic=False).build(g)["circuit"]
        s = qcd.DeviceNoiseSampler(batch.device_props_dict(props, i), n_shots=16, noise_specs=spec)
        s.transpile_circuit(c)
        s.build_noise_model(n)
        q, want = qcd.single_qubit_settings(stab, s.noise_model, n)
        assert q.tolist() == list(range(n)) + [0xFF]
        np.testing.assert_allclose(sops[i], want[:n], rtol=0, atol=1e-14)
        assert s.noise_model.native                       # reading the settings did not materialise the model
    # no calibration entry -> the bare gate
    bare = batch.device_gate_superops_1q(batch._H, np.full(3, np.nan), np.full(3, np.nan), np.ones(3), np.ones(3))
    np.testing.assert_allclose(bare, np.broadcast_to(np.kron(batch._H.conj(), batch._H), (3, 4, 4)), atol=1e-15)


def test_toth_witness_formulas():
    n = 4
    edges = np.array([[[0, 1], [1, 2], [2, 3], [0, 0], [0, 0], [0, 0]]], dtype=np.uint8)
    n_edges = np.array([3], dtype=np.uint32)
    tot = np.full((1, n + 1), 1000, dtype=np.int64)
    # a (non-physical) last column with a non-zero std: the std-devs enter the genuine witness UNSCALED
    # (uncertainties' propagation of uarray(coef * means, stds).sum(), witnesses.py:136-144), the (n - 1) coefficient
    # multiplies the nominal value only
    signed = np.array([[800, 600, 900, 1000, 980]], dtype=np.int64)
    val, sig = batch.toth_witnesses("genuine", signed, tot, edges, n_edges, n)
    mean = signed[0] / 1000
    std = np.sqrt((1 - mean ** 2) / 1000)
    assert abs(val[0] - (3 * mean[4] - mean[:4].sum())) < 1e-15 and abs(sig[0] - np.sqrt((std ** 2).sum())) < 1e-15
    vals, sigs = batch.toth_witnesses("biseparable", signed, tot, edges, n_edges, n)
    assert len(vals[0]) == 3
    assert abs(vals[0][1] - (1 - mean[1] - mean[2])) < 1e-15 and abs(sigs[0][2] - np.sqrt(std[2] ** 2 + std[3] ** 2)) < 1e-15


def test_native_batch_entry_points_reject_bad_arguments():
    import ctypes
    lib = qcd._lib.load()
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)       # noqa: E731
    gs = qcd.GraphState(qcd.GraphDB(), n_qubits=6)
    le, lo, lord, oo, og, cs, co = gs._library_tables()
    edges = np.zeros((4, 15, 2), dtype=np.uint8)
    ne = np.zeros(4, dtype=np.uint32)

    def sample(**kw):
        a = dict(le=le, lo=lo, lord=lord, oo=oo, og=og, cs=cs, co=co, n=6, max_edges=15)
        a.update(kw)
        return lib.qcd_sample_graph_states(p(a["le"]), p(a["lo"]), p(a["lord"]), len(a["lord"]), p(a["oo"]), p(a["og"]),
                                           len(a["oo"]) - 2, p(a["cs"]), p(a["co"]), len(a["co"]) - 1, a["n"], 4, 1, 1, 7, 0,
                                           p(edges), a["max_edges"], p(ne))
    assert sample() == 0 and (ne > 0).all()
    assert sample(max_edges=5) == qcd._lib.QCD_E_INVALID                   # output rows too short for 6 vertices
    assert sample(n=7) == qcd._lib.QCD_E_INVALID                           # combinations do not add up to n_qubits
    bad = le.copy()
    bad[0] = bad[0][::-1]                                                   # an edge with u > v
    assert sample(le=bad) == qcd._lib.QCD_E_INVALID
    assert lib.qcd_sample_graph_states(None, p(lo), p(lord), len(lord), p(oo), p(og), len(oo) - 2, p(cs), p(co),
                                       len(co) - 1, 6, 4, 1, 1, 7, 0, p(edges), 15, p(ne)) == qcd._lib.QCD_E_INVALID
    # adjacency tensors / superoperators: NULL pointers and bad shapes
    out = np.zeros((1, 2, 4, 4))
    codes = np.zeros(len(qcd._lib.OPCODES), dtype=np.int32)
    off = np.zeros(2, dtype=np.uint32)
    assert lib.qcd_adjacency_tensors(None, p(off), 1, p(codes), 2, 4, 4, 0, p(out)) == 0      # an empty circuit is fine
    assert lib.qcd_adjacency_tensors(None, p(off), 1, None, 2, 4, 4, 0, p(out)) == qcd._lib.QCD_E_INVALID
    assert lib.qcd_adjacency_tensors(None, p(off), 1, p(codes), 0, 4, 4, 0, p(out)) == qcd._lib.QCD_E_INVALID
    with pytest.raises(qcd._lib.QcdError):
        batch.device_gate_superops_1q(batch._H, np.array([0.1]), np.array([10.0]), np.array([np.nan]), np.array([1.0]))
    # T2 > 2 T1 is clipped like the Python route does, not rejected
    s = batch.device_gate_superops_1q(batch._H, np.array([0.1]), np.array([10.0]), np.array([1.0]), np.array([50.0]))
    assert np.isfinite(s).all()
    np.testing.assert_allclose(s[0][0] + s[0][3], [1, 0, 0, 1], atol=1e-12)      # trace preserving: rows (0,0) + (1,1)


def test_stochastic_builder_with_no_coin_set_equals_plain_builder():
    n = 5
    gs = qcd.GraphState(qcd.GraphDB(), n_qubits=n)
    edges, n_edges = gs.sample_batch(9, seed=1)
    for cls in (qcd.CXGateCircuit, qcd.CZGateCircuit, qcd.CPhaseGateCircuit):
        sto = cls(n_qubits=n, stochastic=True)
        zero = np.zeros((int(n_edges.sum()), batch.sites_per_edge(sto)), dtype=np.int64)
        a = batch.build_gate_lists(sto, edges, n_edges, coins=zero, gammas=np.zeros((0, 2)))
        b = batch.build_gate_lists(cls(n_qubits=n, stochastic=False), edges, n_edges)
        assert all(np.array_equal(x, y) for x, y in zip(a[:3], b[:3])) and a[3] == b[3]
    with pytest.raises(ValueError):
        batch.build_gate_lists(qcd.CXGateCircuit(n_qubits=n, stochastic=True), edges, n_edges)       # coins missing
    with pytest.raises(ValueError):
        one = np.ones((int(n_edges.sum()), 1), dtype=np.int64)
        batch.build_gate_lists(qcd.CXGateCircuit(n_qubits=n, stochastic=True), edges, n_edges, coins=one,
                               gammas=np.full((int(one.sum()), 2), 1.5))                             # gamma > 1


def test_native_gate_entries_equal_numpy_unique():
    """qcd_gate_entries (the distinct (gate, qubits) calibration entries per circuit) lists exactly what
    numpy.unique((circuit << 24) | opcode << 16 | q0 << 8 | q1) lists, in the same order, for every builder -- the
    calibration draws therefore did not change when the sort over the whole batch was replaced."""
    for cls, n, B in ((qcd.CXGateCircuit, 8, 700), (qcd.CZGateCircuit, 5, 300), (qcd.CPhaseGateCircuit, 6, 257)):
        gs = qcd.GraphState(qcd.GraphDB(), n_qubits=n)
        edges, n_edges = gs.sample_batch(B, seed=11)
        enc = batch.build_gate_lists(cls(n_qubits=n, stochastic=False), edges, n_edges)
        ops, offsets, _, _ = enc
        spec = qcd.NoiseSpec(specs={"qubit": {"T1": 100.0, "T2": 100.0}, "cx": {"gate_error": 0.25, "gate_length": 500.0},
                                    "h": {"gate_error": 0.05, "gate_length": 50.0}})
        props = batch.draw_device_props(spec, enc, np.random.default_rng(1))
        circ = np.repeat(np.arange(B, dtype=np.int64), np.diff(offsets.astype(np.int64)))
        key = (ops["opcode"].astype(np.int64) << 16) | (ops["q0"].astype(np.int64) << 8) | ops["q1"].astype(np.int64)
        full, inv = np.unique((circ << 24) | key, return_inverse=True)
        assert np.array_equal(props["gate_key"], (full & 0xFFFFFF).astype(np.uint32))
        assert np.array_equal(props["gate_circuit"], full >> 24) and np.array_equal(props["op_entry"], inv)


@pytest.mark.parametrize("directed", [False, True])
def test_adjacency_tensors_equal_the_oracle_restatement(directed):
    """Python walk (dataset.generate_adjacency_tensor) and native batch (qcd_adjacency_tensors) against the oracle's
    restatement of samplers.py:119-192 on random circuits: repeated gates on one slot (several planes), both orders of a
    qubit pair, more gates than planes (truncation), a gate encoded as 0 (never occupies a slot), trimmed output."""
    from oracle import ref_glue
    from tests import circuits
    rng = np.random.default_rng(12)
    names = ["h", "x", "s", "cx", "cz", "swap", "id"]
    enc = {nm: i for i, nm in enumerate(names)}            # "h" -> 0: the reference's enumerate-from-0 quirk
    n = 5
    circs = []
    for _ in range(12):
        ops = []
        for _ in range(int(rng.integers(5, 60))):
            nm = names[rng.integers(0, len(names))]
            q = int(rng.integers(0, n))
            if nm in ("cx", "cz", "swap"):
                ops.append((nm, (q, int((q + 1 + rng.integers(0, n - 1)) % n))))
            else:
                ops.append((nm, (q,)))
        circs.append(ops)
    for shape in ((4, n, n), (32, n, n)):
        for ops in circs:
            want = ref_glue.adjacency_tensor(ops, enc, shape, True, directed)
            got = generate_adjacency_tensor(qcd.Circuit(n, ops), enc, shape, True, directed)
            assert np.array_equal(got, want)
            trimmed = generate_adjacency_tensor(qcd.Circuit(n, ops), enc, shape, False, directed)
            assert np.array_equal(trimmed, ref_glue.adjacency_tensor(ops, enc, shape, False, directed))
        encp = encode_programs([qcd.Circuit(n, ops) for ops in circs])
        native = batch.adjacency_tensors_batch(encp, enc, shape, directed)
        for i, ops in enumerate(circs):
            assert np.array_equal(native[i], ref_glue.adjacency_tensor(ops, enc, shape, True, directed)), i
