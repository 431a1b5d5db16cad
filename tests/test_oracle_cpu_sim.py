"""The compiled CPU baseline (oracle/cpu_sim.c) reproduces the numpy oracle shot for shot."""
import numpy as np
import pytest

from oracle import cpu_build, sim
from qcdenoise_b200.ir import Circuit, encode_programs
from tests import circuits


@pytest.mark.parametrize("seed", range(4))
def test_cpu_sim_matches_numpy_oracle(seed):
    rng = np.random.default_rng(seed)
    n = 6
    if seed == 0:
        ops = circuits.graph_circuit_amp_damp(n, circuits.ref_graph(5, 1) + [(3, 5)], rng, max_prob=0.3, all_sites=True)
    elif seed == 1:
        ops, _ = circuits.device_noise_circuit(n, circuits.ref_graph(5, 1) + [(3, 5)], rng, gate_error=0.2,
                                               gate_len_ns=(500, 5000))
    else:
        ops = circuits.random_circuit(n, 40, rng)
    shots = 300
    got = cpu_build.run_trajectories(encode_programs([Circuit(n, ops)]), shots, seed=9, circuit=2, n_threads=2)
    want = sim.run_trajectories(ops, n, shots, 9, circuit=2)
    assert np.count_nonzero(got != want) == 0
