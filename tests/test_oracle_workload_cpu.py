"""The oracle-side restatement of the benchmark workload and of the qcd_op encoding (what the CPU reference arm runs
without importing the product) equals what bench.py uploads to the GPU, gate for gate and byte for byte."""
import numpy as np

import bench
from oracle import encode, workload
from qcdenoise_b200.ir import encode_programs


def _same_op(a, b):
    if a[0] != b[0] or tuple(a[1]) != tuple(b[1]):
        return False
    if a[0] == "unitary":
        return np.array_equal(np.asarray(a[2]), np.asarray(b[2]))
    if a[0] == "kraus":
        return len(a[2]) == len(b[2]) and all(np.array_equal(np.asarray(x), np.asarray(y)) for x, y in zip(a[2], b[2]))
    return True


def test_oracle_workload_is_the_bench_workload():
    for gi in (0, 3):
        work = bench.build_workload(gi, 1000)
        ow = workload.c4_workload(gi)
        assert ow["edges"] == [tuple(e) for e in work["graph"].edges()]
        assert ow["n_sites"] == work["n_sites"] and len(ow["circuits"]) == len(work["programs"]) == 22
        assert ow["masks"] == work["masks"][:, 0].tolist()
        assert ow["names"] == list(work["stab"].keys())
        for mine, theirs in zip(ow["circuits"], work["programs"]):
            assert len(mine) == len(theirs.ops)
            assert all(_same_op(a, b) for a, b in zip(mine, theirs.ops))


def test_oracle_encoder_matches_the_product_encoder():
    work = bench.build_workload(1, 1000)
    ow = workload.c4_workload(1)
    for c in (0, 5, 21):
        mine = encode.encode_circuit(20, ow["circuits"][c])
        # the product encoder stores a channel object once per upload; compare op by op through the parameters
        ops, offsets, params, n = encode_programs([work["programs"][c]])
        assert n == mine[3] and len(ops) == len(mine[0]) and offsets.tolist() == mine[1].tolist()
        assert ops["opcode"].tolist() == mine[0]["opcode"].tolist()
        assert ops["q0"].tolist() == mine[0]["q0"].tolist() and ops["q1"].tolist() == mine[0]["q1"].tolist()
        for a, b in zip(ops, mine[0]):
            if a["opcode"] == encode.OPCODES["u2q"]:
                assert np.array_equal(params[a["param_off"]:a["param_off"] + 32], mine[2][b["param_off"]:b["param_off"] + 32])
            if a["opcode"] == encode.OPCODES["kraus1q"]:
                assert np.array_equal(params[a["param_off"]:a["param_off"] + 17], mine[2][b["param_off"]:b["param_off"] + 17])
    assert encode.OP_DTYPE == __import__("qcdenoise_b200")._lib.OP_DTYPE
