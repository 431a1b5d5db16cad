"""Build oracle/cpu_sim.c -> oracle/_build/libcpusim.so (gcc, OpenMP) and bind it with ctypes.
TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cpu_sim.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libcpusim.so")
_lib = None


def _host_has_avx2():
    try:
        with open("/proc/cpuinfo") as f:
            flags = f.read()
        return all(w in flags for w in (" avx2", " fma", " bmi2"))
    except OSError:
        return False


def build(force=False):
    """The AVX2 build (x86-64-v3) is made in the build container and travels to the GPU box; a host without AVX2 gets
    a generic build compiled on the spot (separate file name, so the two never overwrite each other)."""
    hdr = os.path.join(HERE, "..", "include", "qcdsim.h")
    avx2 = _host_has_avx2()
    lib = LIB if avx2 else os.path.join(OUT_DIR, "libcpusim_generic.so")
    if not force and os.path.exists(lib) and os.path.getmtime(lib) >= max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        return lib
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["gcc", "-O3"] + (["-march=x86-64-v3"] if avx2 else []) + \
          ["-ffast-math", "-fno-finite-math-only", "-fopenmp", "-shared", "-fPIC", "-std=c11", SRC, "-o", lib, "-lm"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("gcc failed:\n" + r.stdout)
    return lib


def load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        lib.cpu_run_trajectories.restype = ctypes.c_int
        lib.cpu_run_trajectories.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int,
                                             ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32,
                                             ctypes.c_void_p, ctypes.c_int]
        lib.cpu_run_trajectories_wide.restype = ctypes.c_int
        lib.cpu_run_trajectories_wide.argtypes = lib.cpu_run_trajectories.argtypes
        lib.cpu_max_threads.restype = ctypes.c_int
        _lib = lib
    return _lib


def run_trajectories(encoded, shots, seed, circuit=0, shot_begin=0, n_threads=0, wide=False):
    """encoded = (ops, offsets, params, n) of ONE circuit in the C-ABI encoding (qcdenoise_b200.ir.encode_programs
    output format; the encoder itself is pure numpy host code).  Returns outcomes uint64[shots].
    wide=True: shots one after the other, each state's amplitude loops spread over the threads (large n, few shots)."""
    ops, offsets, params, n = encoded
    lib = load()
    out = np.empty(shots, dtype=np.uint64)
    params = np.ascontiguousarray(params if len(params) else np.zeros(1))
    fn = lib.cpu_run_trajectories_wide if wide else lib.cpu_run_trajectories
    rc = fn(ops.ctypes.data, len(ops), params.ctypes.data, n, shot_begin, shot_begin + shots, seed,
                                  circuit, out.ctypes.data, n_threads)
    if rc:
        raise RuntimeError("cpu_run_trajectories failed")
    return out


def max_threads():
    return load().cpu_max_threads()
