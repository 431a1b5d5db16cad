"""ctypes binding of libqcdsim.so (include/qcdsim.h).  The library is the product: if it is missing or cannot be
loaded this module raises -- there is no Python / CPU fallback for the simulation path."""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libqcdsim.so")

QCD_ABI_VERSION = 4
QCD_OK, QCD_E_INVALID, QCD_E_CUDA, QCD_E_NOMEM, QCD_E_UNSUPPORTED, QCD_E_STATE = range(6)
QCD_RUN_NO_DEDUP = 1
_ERR_NAMES = {1: "QCD_E_INVALID", 2: "QCD_E_CUDA", 3: "QCD_E_NOMEM", 4: "QCD_E_UNSUPPORTED", 5: "QCD_E_STATE"}

# opcodes, in the order of the enum in include/qcdsim.h
OPCODES = {name: i for i, name in enumerate(
    ["id", "h", "x", "y", "z", "s", "sdg", "t", "tdg", "rx", "ry", "rz", "u1q", "cx", "cz", "swap", "u2q",
     "kraus1q", "kraus2q", "pauli1q", "pauli2q", "reset"])}

OP_DTYPE = np.dtype([("opcode", "<u2"), ("q0", "u1"), ("q1", "u1"), ("param_off", "<u4")])
assert OP_DTYPE.itemsize == 8


class QcdError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"{_ERR_NAMES.get(code, code)}: {message}")
        self.code = code


class Stats(ctypes.Structure):
    _fields_ = [("sweep_launches", ctypes.c_uint64), ("aux_launches", ctypes.c_uint64),
                ("node_sweeps", ctypes.c_uint64), ("sweep_bytes", ctypes.c_uint64), ("leaves", ctypes.c_uint64),
                ("trajectories", ctypes.c_uint64), ("max_live_states", ctypes.c_uint64),
                ("sweep_ms", ctypes.c_double), ("kind_launches", ctypes.c_uint64 * 4),
                ("kind_bytes", ctypes.c_uint64 * 4), ("kind_ms", ctypes.c_double * 4),
                ("cross_passes", ctypes.c_uint64), ("cross_leaves", ctypes.c_uint64)]

    def as_dict(self):
        return {k: (list(getattr(self, k)) if k.startswith("kind_") else getattr(self, k)) for k, _ in self._fields_}


_P = ctypes.c_void_p
_PROTOS = {
    "qcd_abi_version": (ctypes.c_int, []),
    "qcd_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_size_t, ctypes.POINTER(_P)]),
    "qcd_destroy": (None, [_P]),
    "qcd_last_error": (ctypes.c_char_p, [_P]),
    "qcd_set_option": (ctypes.c_int, [_P, ctypes.c_char_p, ctypes.c_double]),
    "qcd_get_stats": (ctypes.c_int, [_P, ctypes.POINTER(Stats)]),
    "qcd_upload_programs": (ctypes.c_int, [_P, _P, _P, _P, ctypes.c_size_t, ctypes.c_int, ctypes.c_int]),
    "qcd_upload_device_noise_programs": (ctypes.c_int, [_P, _P, _P, _P, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, _P,
                                                        _P]),
    "qcd_expand_device_noise": (ctypes.c_int, [_P, _P, _P, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, _P, _P, _P,
                                               ctypes.c_size_t, _P, _P, ctypes.c_size_t,
                                               ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t)]),
    "qcd_run_sv_trajectories": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64,
                                               ctypes.c_uint64, ctypes.c_uint32, _P, _P, _P, ctypes.c_uint32, _P]),
    "qcd_run_density_matrix": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int, ctypes.c_int, _P, _P]),
    "qcd_run_density_matrix_settings": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _P,
                                                       _P, _P, _P]),
    "qcd_apply_readout_probs": (ctypes.c_int, [_P, ctypes.c_int, _P, ctypes.c_int, ctypes.c_int, _P, _P]),
    "qcd_sample_from_probs": (ctypes.c_int, [_P, ctypes.c_int, _P, ctypes.c_int, ctypes.c_int, ctypes.c_uint64,
                                             ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32, _P, _P, _P]),
    "qcd_parity_counts": (ctypes.c_int, [_P, _P, ctypes.c_int, ctypes.c_int, _P, ctypes.c_int, ctypes.c_int, _P, _P,
                                         _P]),
    "qcd_parity_outcomes": (ctypes.c_int, [_P, _P, ctypes.c_int, ctypes.c_uint64, _P, ctypes.c_int, ctypes.c_int, _P, _P]),
    "qcd_parity_probs": (ctypes.c_int, [_P, ctypes.c_int, _P, ctypes.c_int, ctypes.c_int, _P, ctypes.c_int,
                                        ctypes.c_int, _P, _P]),
    "qcd_run_density_matrix_group": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int, ctypes.c_int, _P, _P, _P]),
    "qcd_stabilizer_group": (ctypes.c_int, [_P, _P, _P, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, _P, _P, _P, _P]),
    "qcd_device_gate_superops_1q": (ctypes.c_int, [ctypes.c_int, _P, _P, _P, _P, _P, _P]),
    "qcd_sample_graph_states": (ctypes.c_int, [_P, _P, _P, ctypes.c_int, _P, _P, ctypes.c_int, _P, _P, ctypes.c_int,
                                               ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint64,
                                               ctypes.c_uint64, _P, ctypes.c_uint32, _P]),
    "qcd_adjacency_tensors": (ctypes.c_int, [_P, _P, ctypes.c_int, _P, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_int, _P]),
    "qcd_gate_entries": (ctypes.c_int, [_P, _P, ctypes.c_int, _P, _P, _P, _P]),
    "qcd_debug_dm_schedule": (ctypes.c_int, [_P, ctypes.c_uint32, _P, ctypes.c_size_t, ctypes.c_int, _P, _P, ctypes.c_size_t,
                                             _P, ctypes.c_size_t, _P, _P]),
    "qcd_debug_schedule_json": (ctypes.c_int, [_P, ctypes.c_uint32, _P, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_int, ctypes.c_int, _P, ctypes.c_size_t,
                                               ctypes.POINTER(ctypes.c_size_t)]),
    "qcd_debug_sibling_runs_json": (ctypes.c_int, [_P, _P, _P, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                   _P, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]),
}
EXPORTS = tuple(_PROTOS)
_lib = None


def load():
    """Load libqcdsim.so (once).  Raises if it has not been built: run `python qcdenoise_b200/build.py`."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: build the CUDA engine first (python qcdenoise_b200/build.py); "
                          "qcdenoise_b200 has no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)     # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.qcd_abi_version() != QCD_ABI_VERSION:
        raise ImportError("libqcdsim.so ABI version mismatch: rebuild it")
    _lib = lib
    return lib
