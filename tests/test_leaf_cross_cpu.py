"""The identity behind the shared leaf passes (csrc/sweep_direct.cu leaf_cross_kernel + the compose step of
scan_l1_kernel), in numpy: for a state phi and a real 2x2 u on qubit k >= 5, the probability mass of u_k phi in the
32-amplitude chunk c is

    u00^2 C0[c] + u01^2 C0[c ^ k] + 2 u00 u01 X_k[c]     (bit k of c clear; u10, u11 when set),
    C0[c] = sum_{x in c} |phi(x)|^2,   X_k[c] = sum_{x in c} Re(conj(phi(x)) phi(x ^ 2^k)),

and C0[c] itself when u is orthogonal and k < 5 (the 2x2 acts inside the chunks).  The reference appends exactly such an op
to the graph circuit per Toth setting: one Hadamard on the generator's vertex (stabilizers.py:136-156, :194-199)."""
import numpy as np
import pytest

CHUNK = 5


def _apply_1q(phi, k, u):
    n = int(np.log2(phi.size))
    t = phi.reshape([2] * n)                 # axis 0 = qubit n - 1 ... axis n - 1 = qubit 0
    ax = n - 1 - k
    t = np.moveaxis(t, ax, 0)
    out = np.tensordot(np.asarray(u, dtype=complex), t, axes=([1], [0]))
    return np.moveaxis(out, 0, ax).reshape(-1)


def _chunk_sums(p):
    return p.reshape(-1, 1 << CHUNK).sum(axis=1)


@pytest.mark.parametrize("n", [7, 10])
def test_member_chunk_sums_from_shared_sums(n):
    rng = np.random.default_rng(n)
    phi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    phi[rng.random(1 << n) < 0.3] = 0                      # graph states have many exact zeros
    c0 = _chunk_sums(np.abs(phi) ** 2)
    idx = np.arange(1 << n)
    h = np.array([[1, 1], [1, -1]]) / np.sqrt(2)
    th = 0.7
    ry = np.array([[np.cos(th / 2), -np.sin(th / 2)], [np.sin(th / 2), np.cos(th / 2)]])
    skew = np.array([[0.9, 0.3], [-0.2, 1.1]])              # the formula does not need u to be orthogonal for k >= 5
    for k in range(n):
        for u in (h, ry, skew):
            want = _chunk_sums(np.abs(_apply_1q(phi, k, u)) ** 2)
            if k < CHUNK:
                if u is skew:
                    continue                                 # not norm preserving inside the chunk: no shortcut (the host refuses it)
                np.testing.assert_allclose(want, c0, rtol=1e-12, atol=1e-12)
                continue
            x_k = _chunk_sums(np.real(np.conj(phi) * phi[idx ^ (1 << k)]))
            kb = 1 << (k - CHUNK)
            c = np.arange(c0.size)
            lo, hi = c & ~kb, c | kb
            got = np.where(c & kb,
                           u[1, 0] ** 2 * c0[lo] + u[1, 1] ** 2 * c0[hi] + 2 * u[1, 0] * u[1, 1] * x_k[c],
                           u[0, 0] ** 2 * c0[lo] + u[0, 1] ** 2 * c0[hi] + 2 * u[0, 0] * u[0, 1] * x_k[c])
            np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-12)
            assert np.allclose(x_k[lo], x_k[hi])             # both partner chunks carry the same cross sum


def test_component_split_of_the_cross_term():
    """The two lanes of a pair split Re(conj(a) b) = a.x b.x + a.y b.y by component (the lower lane forms a.x b.x, the upper
    one a.y b.y) and add their halves afterwards: one exchange per amplitude instead of two."""
    rng = np.random.default_rng(1)
    a = rng.normal(size=64) + 1j * rng.normal(size=64)
    b = rng.normal(size=64) + 1j * rng.normal(size=64)
    lower = (a.real * b.real).sum()          # lower lane: own = a.x, received = b.x
    upper = (b.imag * a.imag).sum()          # upper lane: own = b.y, received = a.y
    assert np.isclose(lower + upper, np.real(np.conj(a) * b).sum())


# ------------------------------------------------------------------------------------------- host side: which runs qualify
def _settings(n, edges, rng, extra=(), identity_last=True, dataset_first=True):
    from tests import circuits
    ops = circuits.graph_circuit_amp_damp(n, edges, rng, max_prob=0.3, all_sites=True)
    batch = [ops] if dataset_first else []
    batch += [ops + [("h", (v,))] for v in range(n)]
    if identity_last:
        batch.append(list(ops))
    batch += [ops + list(sub) for sub in extra]
    return batch


def _runs(n, batch, precision=32):
    import qcdenoise_b200 as qb
    from qcdenoise_b200.engine import debug_sibling_runs
    return debug_sibling_runs([qb.Circuit(n, ops) for ops in batch], precision)


@pytest.mark.parametrize("precision", [32, 64])
def test_sibling_run_of_a_graph_and_its_toth_settings(precision):
    """Dataset circuit + one Hadamard per vertex + identity setting: every member qualifies; vertices below the chunk size and
    the circuits without a basis change take the shared C0 (kbit -1), the others their vertex with u = H -- also the one
    whose Hadamard the host compiler fused into the last site's own matrices (and moved behind the sign layer)."""
    from tests import circuits
    n = 12
    rng = np.random.default_rng(3)
    batch = _settings(n, circuits.ref_graph(14, 1)[:9] + [(10, 11)], rng)
    runs = _runs(n, batch, precision)
    assert len(runs) == 1 and runs[0]["lead"] == 0 and runs[0]["members"] == n + 2
    r = runs[0]
    assert r["kbit"] == [-1] + [-1] * 5 + list(range(5, n)) + [-1]
    u = np.array(r["u"]).reshape(-1, 4)
    h = np.array([1, 1, 1, -1]) / np.sqrt(2)
    for j, k in enumerate(r["kbit"]):
        if k >= 0:
            np.testing.assert_allclose(u[j], h, atol=1e-6 if precision == 32 else 1e-14)
    # register bits of the shared passes: the last site's qubits + the lowest other bits; bit 0 always (128-bit loads)
    assert r["reg_bits"][0] == 0 and r["n_high"] == sum(1 for b in r["reg_bits"] if b >= 5) and len(set(r["reg_bits"])) == 4
    # the order the reference API produces (settings first, identity last, no dataset circuit) qualifies as well
    r2 = _runs(n, batch[1:], precision)[0]
    assert r2["members"] == n + 1 and r2["kbit"] == r["kbit"][1:]


def test_members_that_keep_their_own_sweep():
    """A Y-type basis change (complex 2x2), a second member on an already taken qubit and a trailing two-qubit gate do not
    qualify (kbit -2); a real rotation on a high qubit does, with its own 2x2; an unrelated circuit ends the run."""
    from tests import circuits
    n = 10
    rng = np.random.default_rng(8)
    th = 0.7
    batch = _settings(n, circuits.ref_graph(10, 0), rng,
                      extra=([("sdg", (9,)), ("h", (9,))], [("h", (9,))], [("ry", (8,), th)], [("cx", (7, 9))]))
    batch.append(circuits.random_circuit(n, 30, rng))
    r = _runs(n, batch)
    assert len(r) == 1 and r[0]["members"] == len(batch) - 1
    kb = r[0]["kbit"]
    assert kb[:n + 2] == [-1] + [-1] * 5 + list(range(5, n)) + [-1]
    assert kb[n + 2] == -2 and kb[n + 3] == -2 and kb[n + 5] == -2      # Y setting, duplicate of vertex 9, trailing CX
    assert kb[n + 4] == -2                                               # ry on qubit 8: vertex 8 already has a member
    # without the Hadamard settings the rotation is the only member on qubit 8 and qualifies with its own matrix
    ops = batch[0]
    r = _runs(n, [ops, ops + [("ry", (8,), th)], ops + [("h", (6,))], list(ops)])
    assert r[0]["kbit"] == [-1, 8, 6, -1]
    np.testing.assert_allclose(np.array(r[0]["u"]).reshape(-1, 4)[1],
                               [np.cos(th / 2), -np.sin(th / 2), np.sin(th / 2), np.cos(th / 2)], atol=1e-6)


def test_no_sibling_runs_without_a_common_prefix():
    from tests import circuits
    n = 10
    rng = np.random.default_rng(1)
    batch = [circuits.graph_circuit_amp_damp(n, circuits.ref_graph(10, i), rng, max_prob=0.3) for i in range(3)]
    assert _runs(n, batch) == []
