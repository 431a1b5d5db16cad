"""numpy interpreter of the sweep program emitted by the host scheduler (qcd_debug_schedule_json), mirroring what
the CUDA engine does with it -- per-sweep fused ops, population reduction, Philox branch decisions, rescaling,
inverse-CDF sampling -- in complex128 on the CPU.  TEST INFRASTRUCTURE: it lets the CPU suite check the host logic
(lowering, fusion, sweep segmentation, tiling, site groups) against the oracle without a GPU."""
import numpy as np

from oracle import philox
from oracle.sim import apply_matrix

K_M1, K_M2, K_M4, K_SIGNS, K_D1 = range(5)
SW_PREP, SW_FINAL = 1, 2


def _c(v):
    return np.array([complex(a, b) for a, b in v])


def run_sweep(state, sw, n_bits, branch):
    N = 1 << n_bits
    if sw["flags"] & SW_PREP:
        prep = _c(sw["prep"]).reshape(n_bits, 2)
        state = np.ones(1, dtype=np.complex128)
        for b in range(n_bits):
            state = np.kron(prep[b], state)          # bit b more significant than the bits before it
    L, high = sw["L"], sw["high"]
    k = sw["k"]
    assert k == L + len(high) and k <= n_bits
    assert all(h >= L for h in high) and sorted(set(high)) == high
    loc2glob = list(range(L)) + list(high)
    idx = np.arange(N)
    for ps in sw["passes"]:
        G = [loc2glob[b] for b in ps["rbits"]]
        assert len(set(G)) == len(G)
        for op in ps["ops"]:
            kind = op["kind"]
            if kind == K_SIGNS:
                for m in op["masks"]:
                    state = np.where((idx & m) == m, -state, state)
                continue
            sel = (branch >> op["sel_shift"]) & op["sel_mask"]
            mats = _c(op["mats"])
            if kind == K_D1:
                d = mats[:2]
                state = state * np.where((idx >> op["gbit"]) & 1, d[1], d[0])
                continue
            dim = {K_M1: 2, K_M2: 4, K_M4: 16}[kind]
            m = mats[sel * dim * dim:(sel + 1) * dim * dim].reshape(dim, dim)
            if kind == K_M1:
                qs = (G[op["t0"]],)
            elif kind == K_M2:
                assert op["t0"] < op["t1"]
                qs = (G[op["t0"]], G[op["t1"]])
            else:
                qs = tuple(G)
            for q in qs:   # every non-diagonal target must be a tile bit
                assert q in loc2glob
            state = apply_matrix(state, m, qs, n_bits)
    return state


def populations(state, red_q):
    """joint populations P[b(q0) + 2 b(q1)] of the sweep's red_q bits."""
    idx = np.arange(state.size)
    w = np.abs(state) ** 2
    q0, q1 = red_q
    P = np.zeros(4)
    if q0 < 0:
        return P
    b = (idx >> q0) & 1
    if q1 >= 0:
        b = b | (((idx >> q1) & 1) << 1)
    for v in range(4):
        P[v] = w[b == v].sum()
    return P


def resolve_group(group, Pin, seed, circuit, shot, force=None):
    """Port of resolve_group (kernels.cu).  Returns (word, scale)."""
    P = np.array([1.0, 0.0, 0.0, 0.0]) if group["q"][0] < 0 else np.array(Pin, dtype=np.float64)
    word, norm2 = 0, 1.0
    for ss in group["subs"]:
        nb = ss["nbranch"]
        data = np.array(ss["data"]).reshape(nb, 20)
        w, T = data[:, :4], data[:, 4:].reshape(nb, 4, 4)
        pj = w @ P
        if force is None:
            target = philox.uniform53(seed, circuit, np.uint64(shot), ss["stream"]) * pj.sum()
            cum = np.cumsum(pj)
            sel = nb - 1
            for j in range(nb):
                if cum[j] > target:
                    sel = j
                    break
        else:
            nbits = max(1, int(np.ceil(np.log2(nb)))) if nb > 1 else 0
            sel = (force >> ss["shift"]) & ((1 << nbits) - 1)
        psum = P.sum()
        if not ss["unit"]:
            norm2 *= pj[sel] / psum
        Q = T[sel] @ P
        P = Q * (psum / pj[sel]) if pj[sel] > 0 else Q * 0
        word |= sel << ss["shift"]
    n0 = Pin.sum() if group["q"][0] >= 0 else 1.0
    return word, 1.0 / np.sqrt(norm2 * n0)


def run_dm(prog):
    nb = prog["n_bits"]
    state = None
    for sw in prog["sweeps"]:
        assert "group" not in sw
        state = run_sweep(state, sw, nb, 0)
    n = prog["n_qubits"]
    return np.real(state.reshape(2 ** n, 2 ** n).diagonal()).copy()


def run_trajectory(prog, seed, circuit, shot):
    """One trajectory, the way the engine walks it.  Returns (final normalised state, branch words)."""
    nb = prog["n_bits"]
    sweeps = prog["sweeps"]
    words = []
    P = np.array([1.0, 0.0, 0.0, 0.0])
    state = None
    for i, sw in enumerate(sweeps):
        word, scale = 0, 1.0
        if "group" in sw:
            word, scale = resolve_group(sw["group"], P, seed, circuit, shot)
        words.append(word)
        if state is not None:
            state = state * scale
        state = run_sweep(state, sw, nb, word)
        P = populations(state, sw["red_q"])
        if i + 1 < len(sweeps) and "group" in sweeps[i + 1] and sweeps[i + 1]["group"]["q"][0] >= 0:
            assert sw["red_q"] == sweeps[i + 1]["group"]["q"]
    return state, words


def sample_outcome(state, seed, circuit, shot):
    cum = np.cumsum(np.abs(state) ** 2)
    u = philox.uniform53(seed, circuit, np.uint64(shot), philox.STREAM_MEASURE)
    return int(min(np.searchsorted(cum, u * cum[-1], side="right"), cum.size - 1))


def run_shots(prog, seed, circuit, shots):
    cache = {}
    out = np.zeros(shots, dtype=np.uint64)
    for s in range(shots):
        # branch words first (cheap replay is not possible without the states, so cache by word history lazily)
        state, words = run_trajectory_cached(prog, seed, circuit, s, cache)
        out[s] = sample_outcome(state, seed, circuit, s)
    return out


def run_trajectory_cached(prog, seed, circuit, shot, cache):
    """Same as run_trajectory but shares states between identical branch histories (the engine's tree)."""
    nb = prog["n_bits"]
    sweeps = prog["sweeps"]
    key = ()
    state, P = None, np.array([1.0, 0.0, 0.0, 0.0])
    for i, sw in enumerate(sweeps):
        word, scale = 0, 1.0
        if "group" in sw:
            word, scale = resolve_group(sw["group"], P, seed, circuit, shot)
        key = key + (word,)
        if key in cache:
            state, P = cache[key]
            continue
        if state is not None:
            state = state * scale
        state = run_sweep(state, sw, nb, word)
        P = populations(state, sw["red_q"])
        cache[key] = (state, P)
    return state, key
