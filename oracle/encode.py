"""Gate list (oracle/sim.py format) -> the qcd_op array of include/qcdsim.h, restated on the ORACLE side so that the CPU
reference arm (bench.py --impl reference, oracle/cpu_sim.c) never imports the product package or maps libqcdsim.so.
TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py).  tests/test_oracle_workload_cpu.py checks that it produces
the same arrays as the product's encoder (qcdenoise_b200/ir.py)."""
import numpy as np

# enum order of include/qcdsim.h (QCD_OP_*)
OPCODES = {name: i for i, name in enumerate(
    ["id", "h", "x", "y", "z", "s", "sdg", "t", "tdg", "rx", "ry", "rz", "u1q", "cx", "cz", "swap", "u2q",
     "kraus1q", "kraus2q", "pauli1q", "pauli2q", "reset"])}
OP_DTYPE = np.dtype([("opcode", "<u2"), ("q0", "u1"), ("q1", "u1"), ("param_off", "<u4")])     # struct qcd_op


def _cflat(m):
    m = np.asarray(m, dtype=np.complex128).reshape(-1)
    out = np.empty(2 * m.size)
    out[0::2] = m.real
    out[1::2] = m.imag
    return out


def encode_circuit(n, ops):
    """One circuit -> (ops: OP_DTYPE[], offsets: uint32[2], params: float64[], n), the argument layout of
    qcd_upload_programs / cpu_run_trajectories."""
    recs, params = [], []

    def put(vals):
        off = len(params)
        params.extend(np.asarray(vals, dtype=np.float64).ravel().tolist())
        return off

    for op in ops:
        name = op[0]
        if name in ("barrier", "measure_all"):
            continue
        qs = op[1]
        if name in ("id", "h", "x", "y", "z", "s", "sdg", "t", "tdg"):
            recs.append((OPCODES[name], qs[0], 0xFF, 0))
        elif name in ("rx", "ry", "rz"):
            recs.append((OPCODES[name], qs[0], 0xFF, put([op[2]])))
        elif name in ("cx", "cz", "swap"):
            recs.append((OPCODES[name], qs[0], qs[1], 0))
        elif name == "unitary":
            recs.append((OPCODES["u1q" if len(qs) == 1 else "u2q"], qs[0], qs[1] if len(qs) > 1 else 0xFF,
                         put(_cflat(op[2]))))
        elif name == "kraus":
            flat = np.concatenate([[float(len(op[2]))]] + [_cflat(k) for k in op[2]])
            recs.append((OPCODES["kraus1q" if len(qs) == 1 else "kraus2q"], qs[0], qs[1] if len(qs) > 1 else 0xFF,
                         put(flat)))
        elif name == "pauli":
            recs.append((OPCODES["pauli1q" if len(qs) == 1 else "pauli2q"], qs[0], qs[1] if len(qs) > 1 else 0xFF,
                         put(op[2])))
        elif name == "reset":
            recs.append((OPCODES["reset"], qs[0], 0xFF, 0))
        else:
            raise ValueError(f"unknown op {name!r}")
    arr = np.array(recs, dtype=OP_DTYPE) if recs else np.zeros(0, dtype=OP_DTYPE)
    return arr, np.array([0, len(recs)], dtype=np.uint32), np.asarray(params, dtype=np.float64), n
