"""Pin the oracle against every results-level golden the reference holds for this path
(fixtures made by tests/golden/make_golden.py from the reference notebooks)."""
import numpy as np
import pytest

from oracle import philox, ref_glue, sim


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    f = lambda *a: [int(x) for x in philox.philox4x32_10(*a)]
    assert f(0, 0, 0, 0, 0, 0) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert f(*([0xffffffff] * 6)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert f(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_G1_diagonals_and_expvals(goldens):
    g = goldens["G1"]
    names = g["stabilizer_order"]
    for (name, diag) in g["diagonals"]:
        np.testing.assert_array_equal(ref_glue.construct_diagonal(name), np.array(diag))
    meas = ref_glue.stabilizer_measurements(names, g["counts"])
    for name, (mean, std) in g["stabilizer_measurements"]:
        assert meas[name][0] == mean                      # bit-exact: dyadic rationals
        assert meas[name][1] == pytest.approx(std, rel=1e-14, abs=1e-18)


def test_G2_genuine_witness(goldens):
    g = goldens["G1"]
    meas = ref_glue.stabilizer_measurements(g["stabilizer_order"], g["counts"])
    w = ref_glue.genuine_witness(meas, n_nodes=4, n=4)
    assert w.value == goldens["G2"]["genuine"]["value"] == goldens["G2"]["genuine_value_c53"]
    assert w.variance == pytest.approx(goldens["G2"]["genuine"]["variance"], abs=5e-9)


def test_genuine_witness_sigma_identity_not_last(goldens):
    """witnesses.py:136-144: arr = unumpy.uarray(wit_coef * meas_vals, var_vals); sigma = std_devs(arr.sum()).  The
    `uncertainties` rule for a plain sum of independent variables x_k +- s_k is sigma = sqrt(sum_k s_k^2): the
    coefficients were folded into the nominal values, the std-devs enter UNSCALED, so the result cannot depend on
    where the identity sits in the (hash-ordered, quirk Q9) measurement dict."""
    g = goldens["G1"]
    names, counts = list(g["stabilizer_order"]), list(g["counts"])
    stds = dict((nm, ms[1]) for nm, ms in g["stabilizer_measurements"])
    expected = sum(s * s for s in stds.values()) ** 0.5          # hand evaluation of the rule above on G1's stds
    assert expected == pytest.approx(goldens["G2"]["genuine"]["variance"], abs=5e-9)
    i = names.index("IIII")
    for order in ([i] + [j for j in range(len(names)) if j != i],            # identity first
                  [1, i, 0] + [j for j in range(len(names)) if j not in (0, 1, i)]):
        meas = ref_glue.stabilizer_measurements([names[j] for j in order], [counts[j] for j in order])
        w = ref_glue.genuine_witness(meas, n_nodes=4, n=4)
        assert w.value == goldens["G2"]["genuine"]["value"]
        assert w.variance == pytest.approx(expected, rel=1e-14)


def test_G3_biseparable_witness(goldens):
    g = goldens["G1"]
    meas = ref_glue.stabilizer_measurements(g["stabilizer_order"], g["counts"])
    res = ref_glue.bisep_witnesses(meas, [tuple(e) for e in g["edges"]])
    for idx, exp in enumerate(goldens["G3"]["bisep"]):
        assert list(res[idx].W_ij) == exp["W_ij"]
        assert res[idx].value == exp["value"]
        assert res[idx].variance == pytest.approx(exp["variance"], abs=5e-9)


def test_G4_ideal_support_and_witness(goldens, ref_graphs):
    g = goldens["G4"]
    n = 4
    edges = [tuple(e) for e in ref_graphs["P4"]]
    names = g["stabilizer_order_c18"]
    nbrs = {v: [b if a == v else a for a, b in edges if v in (a, b)] for v in range(n)}
    gens = ref_glue.get_generators(n, range(n), nbrs)
    assert sorted(gens + ["I" * n]) == sorted(names)
    ops, _ = ref_glue.cx_gate_circuit(n, edges)
    psi = sim.ideal_statevector(ops, n)
    # closed form of a graph state: 2^{-n/2} (-1)^{sum_{(a,b)} x_a x_b}
    x = np.arange(2 ** n)
    sign = np.ones(2 ** n)
    for a, b in edges:
        sign *= 1 - 2 * (((x >> a) & 1) & ((x >> b) & 1))
    np.testing.assert_allclose(psi, sign / 4, atol=1e-15)
    # recorded counts (#c34) are aligned with the #c18 key order: support must equal the exact support
    for name, counts in zip(names, g["ideal_counts"]):
        phi = sim.ideal_statevector(ops + ref_glue.stabilizer_basis_change(name), n)
        support = {format(i, "04b") for i in np.nonzero(np.abs(phi) ** 2 > 1e-12)[0]}
        assert set(counts.keys()) == support
        assert sum(counts.values()) == g["n_shots"]
        mean, std = ref_glue.expectation_value(counts, ref_glue.construct_diagonal(name))
        assert (mean, std) == (1.0, 0.0)
    meas = ref_glue.stabilizer_measurements(names, g["ideal_counts"])
    w = ref_glue.genuine_witness(meas, n_nodes=n, n=n)
    assert w.value == g["genuine"]["value"] == -1.0 and w.variance == 0.0
    for _, (mean, std) in g["stabilizer_measurements"]:
        assert (mean, std) == (1.0, 0.0)


def test_G5_generators_and_group(goldens):
    for entry in goldens["G5"]:
        gens = entry["generators"]
        n = len(gens)
        # recover adjacency from the generator strings and rebuild them
        # (generator i belongs to the node holding its X; node iteration order = order of the list)
        nodes = [g[::-1].index("X") for g in gens]
        nbrs = {v: [j for j, ch in enumerate(g[::-1]) if ch == "Z"] for v, g in zip(nodes, gens)}
        assert ref_glue.get_generators(n, nodes, nbrs) == gens
        assert ref_glue.full_stabilizer_group(gens, n) == entry["stabilizers"]


def test_Q1_draws_are_pure_amplitude_damping(goldens):
    from oracle import noise
    for phases, amps in goldens["Q1"]["printed_params"]:
        assert amps == [0, 0]
        k_node, k_ngbr = noise.reference_unitary_noise_site(phases)
        assert len(k_node) == 2 and len(k_ngbr) == 2
        np.testing.assert_allclose(k_node[0], np.diag([1, np.sqrt(1 - phases[1])]))
        np.testing.assert_allclose(k_ngbr[1], [[0, np.sqrt(phases[0])], [0, 0]])


def test_prob_vector_matches_bin_order():
    counts = {"0101": 3, "1000": 5}
    v = ref_glue.convert_to_prob_vector(counts, 4, 8)
    assert v[5] == 3 / 8 and v[8] == 5 / 8 and v.sum() == 1.0
    assert ref_glue.generate_binary_strings(3) == [format(i, "03b") for i in range(8)]
