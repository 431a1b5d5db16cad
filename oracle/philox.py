"""Philox4x32-10 counter-based RNG (Salmon et al., SC'11 "Parallel random numbers: as easy as 1, 2, 3"),
vectorised with numpy.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Stream convention shared with the CUDA engine (include/qcdsim.h "RNG"):
    key     = (seed & 0xffffffff, seed >> 32)
    counter = (traj & 0xffffffff, traj >> 32, stream, circuit)
    stream  = ordinal of the stochastic site in the circuit's op list, or STREAM_MEASURE / STREAM_READOUT+k
    u53     = ((w1 << 32 | w0) >> 11) * 2**-53          (uniform double in [0,1))
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

STREAM_MEASURE = 0xFFFFFFFF
STREAM_READOUT = 0xFFFF0000   # + k for the k-th block of four 32-bit draws


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """All arguments broadcastable integer arrays (values < 2**32). Returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0)
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def words(seed, circuit, traj, stream):
    traj = np.asarray(traj, dtype=np.uint64)
    return philox4x32_10(traj & MASK, traj >> np.uint64(32), stream, circuit,
                         seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)


def uniform53(seed, circuit, traj, stream):
    w0, w1, _, _ = words(seed, circuit, traj, stream)
    x = (w1.astype(np.uint64) << np.uint64(32)) | w0.astype(np.uint64)
    return (x >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)
