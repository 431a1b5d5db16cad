// dm_cluster.cu -- small density matrices simulated ON CHIP: the gate list is interpreted directly, no host compiler.
//
// Replaces, for n <= 9 qubits (complex64) / n <= 8 (complex128), the lower.cpp -> sweep-kernel route of the AerSimulator
// `density_matrix` method (samplers.py:308-326 with a noise model attached, samplers.py:577-727).  The reference's C2
// workload (8-qubit graph states, device noise, 10 000 circuits) spent 80 % of its time in the host compiler: 300 us per
// circuit per core to fuse ~130 ops into 16x16 superoperators, against ~40 us of GPU work.
//
//   layout      vec(rho), index = x + 2^n y, lives in the SHARED MEMORY OF A THREAD-BLOCK CLUSTER: 2^14 complex64 (128 KB)
//               per CTA, 1 / 4 / 16 CTAs for n <= 7 / 8 / 9; the top index bits name the CTA, remote parts are read and
//               written through distributed shared memory (cluster.map_shared_rank).
//   pass        one group of ops on one qubit pair {a, b} (scheduled on the host, O(ops), no matrices): a thread owns the
//               4x4 block of rho spanned by row bits a, b and column bits a + n, b + n (16 amplitudes in registers),
//               applies EVERY op of the pass there -- the gate, its depolarizing mixture, the relaxation channels of
//               both qubits, pending single-qubit gates -- and writes the block back in place.  Blocks are disjoint, so
//               a pass needs no double buffering, only a barrier (cluster-wide when a gate bit is a CTA bit).
//   ops         1q: everything (unitaries, Kraus lists, Pauli mixtures, reset) becomes a 4x4 superoperator on
//               (b00, b01, b10, b11), built once per pass in shared memory; most channels only couple the diagonal with
//               the diagonal and the anti-diagonal with the anti-diagonal ("X shape": 8 of 16 entries).
//               2q: CX / CZ / SWAP are index permutations / signs of the register block; Pauli mixtures reduce to 16 real
//               weights W[flip][r ^ c]; U2Q acts on rows then (conjugated) on columns; KRAUS2Q sums K B K^dag.
//   bound       shared-memory bandwidth: every pass reads and writes 4^n amplitudes once (n = 8: 1 MB per pass over 4 SMs).
#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "dm_cluster.h"
#include "kernels.cuh"

namespace cg = cooperative_groups;

namespace qcd {

namespace {

#ifndef DMI_F32_THREADS
#define DMI_F32_THREADS 512
#endif
constexpr int IR_THREADS = 256;
constexpr int IR_MAX_OPS = DMI_MAX_PASS_OPS;

enum RecType : int { R_SKIP = 0, R_SUPER1, R_SUPER1_X, R_SUPER1_R, R_SUPER1_XR, R_CX, R_CZ, R_SWAP, R_U2, R_PAULI2, R_DEPOL2, R_KRAUS2 };

template <typename T>
struct OpRec {
    int type;
    int slot;          // 1q: slot of the qubit; CX: slot of the control
    uint32_t param;    // KRAUS2Q: offset of its parameters
    int perm;          // 2q ops given as (q1, q0) of the pass: 1 = swap the two slot bits of every index
    int pass_q;        // pass of the current group the record belongs to
    T aux[2];          // uniform two-qubit depolarizing: B' = aux[0] B + aux[1] tr(B) 1
    T mr[16];          // real-valued superoperators, packed (4 x 128-bit shared loads instead of 16 x 64-bit)
    typename Cplx<T>::type m[16];   // superoperator / 4x4 unitary; PAULI2: W in .x
};

template <typename C> __device__ __forceinline__ C cmul(C a, C b) { return C{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
template <typename C> __device__ __forceinline__ C cmulc(C a, C b) { return C{a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y}; }  // a conj(b)
template <typename C> __device__ __forceinline__ void cfma(C& acc, C a, C b) {
    acc.x += a.x * b.x - a.y * b.y;
    acc.y += a.x * b.y + a.y * b.x;
}

// 2x2 unitary of a gate opcode, row-major (re, im)
__device__ void gate_matrix(const qcd_op& op, const double* params, double2* u) {
    const double s = 0.70710678118654752440;
    double2 z{0, 0}, one{1, 0};
    u[0] = one; u[1] = z; u[2] = z; u[3] = one;
    switch (op.opcode) {
        case QCD_OP_H: u[0] = {s, 0}; u[1] = {s, 0}; u[2] = {s, 0}; u[3] = {-s, 0}; break;
        case QCD_OP_X: u[0] = z; u[1] = one; u[2] = one; u[3] = z; break;
        case QCD_OP_Y: u[0] = z; u[1] = {0, -1}; u[2] = {0, 1}; u[3] = z; break;
        case QCD_OP_Z: u[3] = {-1, 0}; break;
        case QCD_OP_S: u[3] = {0, 1}; break;
        case QCD_OP_SDG: u[3] = {0, -1}; break;
        case QCD_OP_T: u[3] = {s, s}; break;
        case QCD_OP_TDG: u[3] = {s, -s}; break;
        case QCD_OP_RX: { double c = cos(params[op.param_off] / 2), t = sin(params[op.param_off] / 2);
                          u[0] = {c, 0}; u[1] = {0, -t}; u[2] = {0, -t}; u[3] = {c, 0}; break; }
        case QCD_OP_RY: { double c = cos(params[op.param_off] / 2), t = sin(params[op.param_off] / 2);
                          u[0] = {c, 0}; u[1] = {-t, 0}; u[2] = {t, 0}; u[3] = {c, 0}; break; }
        case QCD_OP_RZ: { double c = cos(params[op.param_off] / 2), t = sin(params[op.param_off] / 2);
                          u[0] = {c, -t}; u[3] = {c, t}; break; }
        case QCD_OP_U1Q: for (int i = 0; i < 4; i++) u[i] = {params[op.param_off + 2 * i], params[op.param_off + 2 * i + 1]}; break;
        default: break;
    }
}

// entry (i, j) of the superoperator sum_k K_k (x) conj(K_k) on vec(b) = (b00, b01, b10, b11), index = 2 r + c
__device__ double2 super_entry(const qcd_op& op, const double* params, int i, int j) {
    const int r = i >> 1, c = i & 1, r2 = j >> 1, c2 = j & 1;
    double2 acc{0, 0};
    auto add = [&](double2 a, double2 b) {   // a conj(b)
        acc.x += a.x * b.x + a.y * b.y;
        acc.y += a.y * b.x - a.x * b.y;
    };
    if (op.opcode == QCD_OP_KRAUS1Q) {
        const int m = (int)params[op.param_off];
        for (int k = 0; k < m; k++) {
            const double* K = params + op.param_off + 1 + 8 * k;
            add(double2{K[2 * (2 * r + r2)], K[2 * (2 * r + r2) + 1]}, double2{K[2 * (2 * c + c2)], K[2 * (2 * c + c2) + 1]});
        }
    } else if (op.opcode == QCD_OP_PAULI1Q) {
        const double* p = params + op.param_off;   // I, X, Y, Z:  (P b P)[r][c] = (+-) b[r ^ f][c ^ f], sign (-1)^{z (r ^ c)}
        for (int k = 0; k < 4; k++) {
            const int f = (k == 1 || k == 2), zz = (k == 2 || k == 3);
            if ((r ^ f) == r2 && (c ^ f) == c2) acc.x += ((zz && (r ^ c)) ? -p[k] : p[k]);
        }
    } else if (op.opcode == QCD_OP_RESET) {
        if (i == 0 && (j == 0 || j == 3)) acc.x = 1.0;         // b00' = b00 + b11, everything else 0
    } else {
        double2 u[4];
        gate_matrix(op, params, u);
        add(u[2 * r + r2], u[2 * c + c2]);
    }
    return acc;
}

template <typename T, int NS>
struct Block {
    static constexpr int D = 1 << NS;
    typename Cplx<T>::type b[D * D];   // b[r * D + c]
};

// 1q superoperator on slot S of the register block; REAL: all 16 entries are real (Hadamard, Pauli mixtures, damping) and
// come packed in mr[]
template <typename T, int NS, int S, bool XSHAPE, bool REAL>
__device__ __forceinline__ void apply_super1(Block<T, NS>& B, const typename Cplx<T>::type* m, const T* mr) {
    typedef typename Cplx<T>::type C;
    constexpr int D = 1 << NS, M = 1 << S;
    C k[16];
    if (REAL) {
#pragma unroll
        for (int i = 0; i < 16; i++) k[i] = C{mr[i], (T)0};
    } else {
#pragma unroll
        for (int i = 0; i < 16; i++)
            if (!XSHAPE || i == 0 || i == 3 || i == 5 || i == 6 || i == 9 || i == 10 || i == 12 || i == 15) k[i] = m[i];
    }
    auto mul = [](C a, C v) { return REAL ? C{a.x * v.x, a.x * v.y} : cmul(a, v); };
    auto fma = [](C& acc, C a, C v) {
        if (REAL) { acc.x += a.x * v.x; acc.y += a.x * v.y; }
        else cfma(acc, a, v);
    };
#pragma unroll
    for (int ro = 0; ro < D; ro++) {
        if (ro & M) continue;
#pragma unroll
        for (int co = 0; co < D; co++) {
            if (co & M) continue;
            const C v0 = B.b[ro * D + co], v1 = B.b[ro * D + (co | M)], v2 = B.b[(ro | M) * D + co], v3 = B.b[(ro | M) * D + (co | M)];
            C w0, w1, w2, w3;
            if (XSHAPE) {
                w0 = mul(k[0], v0); fma(w0, k[3], v3);
                w1 = mul(k[5], v1); fma(w1, k[6], v2);
                w2 = mul(k[9], v1); fma(w2, k[10], v2);
                w3 = mul(k[12], v0); fma(w3, k[15], v3);
            } else {
                w0 = mul(k[0], v0); fma(w0, k[1], v1); fma(w0, k[2], v2); fma(w0, k[3], v3);
                w1 = mul(k[4], v0); fma(w1, k[5], v1); fma(w1, k[6], v2); fma(w1, k[7], v3);
                w2 = mul(k[8], v0); fma(w2, k[9], v1); fma(w2, k[10], v2); fma(w2, k[11], v3);
                w3 = mul(k[12], v0); fma(w3, k[13], v1); fma(w3, k[14], v2); fma(w3, k[15], v3);
            }
            B.b[ro * D + co] = w0; B.b[ro * D + (co | M)] = w1; B.b[(ro | M) * D + co] = w2; B.b[(ro | M) * D + (co | M)] = w3;
        }
    }
}

// index permutation pi applied to rows and columns: B'[r][c] = B[pi(r)][pi(c)]  (pi an involution here)
template <typename T, int CTRL>
__device__ __forceinline__ void apply_cx(Block<T, 2>& B) {
    typedef typename Cplx<T>::type C;
    constexpr int TG = 1 - CTRL;
    C t[16];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int pr = r ^ (((r >> CTRL) & 1) << TG), pc = c ^ (((c >> CTRL) & 1) << TG);
            t[r * 4 + c] = B.b[pr * 4 + pc];
        }
#pragma unroll
    for (int i = 0; i < 16; i++) B.b[i] = t[i];
}
template <typename T>
__device__ __forceinline__ void apply_swap(Block<T, 2>& B) {
    typedef typename Cplx<T>::type C;
    C t[16];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int pr = ((r & 1) << 1) | (r >> 1), pc = ((c & 1) << 1) | (c >> 1);
            t[r * 4 + c] = B.b[pr * 4 + pc];
        }
#pragma unroll
    for (int i = 0; i < 16; i++) B.b[i] = t[i];
}
template <typename T>
__device__ __forceinline__ void apply_cz(Block<T, 2>& B) {
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++)
            if ((r == 3) != (c == 3)) { B.b[r * 4 + c].x = -B.b[r * 4 + c].x; B.b[r * 4 + c].y = -B.b[r * 4 + c].y; }
}
// B' = U B U^dag, U 4x4 in slot order
template <typename T>
__device__ __forceinline__ void apply_u2(Block<T, 2>& B, const typename Cplx<T>::type* u) {
    typedef typename Cplx<T>::type C;
#pragma unroll
    for (int c = 0; c < 4; c++) {          // rows: column c of B <- U * column
        C v[4];
#pragma unroll
        for (int r = 0; r < 4; r++) v[r] = B.b[r * 4 + c];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            C w = cmul(u[r * 4], v[0]);
            cfma(w, u[r * 4 + 1], v[1]); cfma(w, u[r * 4 + 2], v[2]); cfma(w, u[r * 4 + 3], v[3]);
            B.b[r * 4 + c] = w;
        }
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {          // columns: row r of B <- row * U^dag
        C v[4];
#pragma unroll
        for (int c = 0; c < 4; c++) v[c] = B.b[r * 4 + c];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            C w = cmulc(v[0], u[c * 4]);
            C t1 = cmulc(v[1], u[c * 4 + 1]), t2 = cmulc(v[2], u[c * 4 + 2]), t3 = cmulc(v[3], u[c * 4 + 3]);
            w.x += t1.x + t2.x + t3.x;
            w.y += t1.y + t2.y + t3.y;
            B.b[r * 4 + c] = w;
        }
    }
}
// two-qubit Pauli mixture: B'[r][c] = sum_f W[f][r ^ c] B[r ^ f][c ^ f]
template <typename T>
__device__ __forceinline__ void apply_pauli2(Block<T, 2>& B, const typename Cplx<T>::type* w) {
    typedef typename Cplx<T>::type C;
    C t[16];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            C acc{0, 0};
#pragma unroll
            for (int f = 0; f < 4; f++) {
                const T ww = w[f * 4 + (r ^ c)].x;
                acc.x += ww * B.b[(r ^ f) * 4 + (c ^ f)].x;
                acc.y += ww * B.b[(r ^ f) * 4 + (c ^ f)].y;
            }
            t[r * 4 + c] = acc;
        }
#pragma unroll
    for (int i = 0; i < 16; i++) B.b[i] = t[i];
}
// uniform two-qubit depolarizing (all 15 non-identity Paulis equally likely): sum_P P B P = 4 tr(B) 1
template <typename T>
__device__ __forceinline__ void apply_depol2(Block<T, 2>& B, T c0, T c1) {
    const T tx = c1 * (B.b[0].x + B.b[5].x + B.b[10].x + B.b[15].x), ty = c1 * (B.b[0].y + B.b[5].y + B.b[10].y + B.b[15].y);
#pragma unroll
    for (int i = 0; i < 16; i++) {
        B.b[i].x *= c0;
        B.b[i].y *= c0;
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
        B.b[r * 5].x += tx;
        B.b[r * 5].y += ty;
    }
}
// by value: a reference handed to a non-inlined function would make the caller's register block addressable and send
// EVERY access to it through local memory (measured: 10 % of all instructions were local loads / stores)
template <typename T>
__device__ __noinline__ Block<T, 2> apply_kraus2(Block<T, 2> B, const double* params, uint32_t off, int perm) {
    typedef typename Cplx<T>::type C;
    const int m = (int)params[off];
    C acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = C{0, 0};
    for (int k = 0; k < m; k++) {
        const double* K = params + off + 1 + 32 * k;
        Block<T, 2> tmp = B;
        C u[16];
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int pr = perm ? (((r & 1) << 1) | (r >> 1)) : r, pc = perm ? (((c & 1) << 1) | (c >> 1)) : c;
                u[r * 4 + c] = C{(T)K[2 * (pr * 4 + pc)], (T)K[2 * (pr * 4 + pc) + 1]};
            }
        apply_u2<T>(tmp, u);
#pragma unroll
        for (int i = 0; i < 16; i++) { acc[i].x += tmp.b[i].x; acc[i].y += tmp.b[i].y; }
    }
#pragma unroll
    for (int i = 0; i < 16; i++) B.b[i] = acc[i];
    return B;
}

template <typename T, int NS>
__device__ __forceinline__ void apply_rec(Block<T, NS>& B, const OpRec<T>& rc, const double* params) {
    const bool hi = NS == 2 && rc.slot == 1;
    switch (rc.type) {
        case R_SUPER1:
            if (!hi) apply_super1<T, NS, 0, false, false>(B, rc.m, rc.mr); else apply_super1<T, NS, NS - 1, false, false>(B, rc.m, rc.mr);
            break;
        case R_SUPER1_X:
            if (!hi) apply_super1<T, NS, 0, true, false>(B, rc.m, rc.mr); else apply_super1<T, NS, NS - 1, true, false>(B, rc.m, rc.mr);
            break;
        case R_SUPER1_R:
            if (!hi) apply_super1<T, NS, 0, false, true>(B, rc.m, rc.mr); else apply_super1<T, NS, NS - 1, false, true>(B, rc.m, rc.mr);
            break;
        case R_SUPER1_XR:
            if (!hi) apply_super1<T, NS, 0, true, true>(B, rc.m, rc.mr); else apply_super1<T, NS, NS - 1, true, true>(B, rc.m, rc.mr);
            break;
        default: break;
    }
    if constexpr (NS == 2) {
        switch (rc.type) {
            case R_CX: if (rc.slot == 0) apply_cx<T, 0>(B); else apply_cx<T, 1>(B); break;
            case R_CZ: apply_cz<T>(B); break;
            case R_SWAP: apply_swap<T>(B); break;
            case R_U2: apply_u2<T>(B, rc.m); break;
            case R_PAULI2: apply_pauli2<T>(B, rc.m); break;
            case R_DEPOL2: apply_depol2<T>(B, rc.aux[0], rc.aux[1]); break;
            case R_KRAUS2: B = apply_kraus2<T>(B, params, rc.param, rc.perm); break;
            default: break;
        }
    }
}

// insert a zero bit at position p
__device__ __forceinline__ uint32_t gap(uint32_t f, int p) { return ((f >> p) << (p + 1)) | (f & ((1u << p) - 1u)); }

// Shared-memory index swizzle: a gate bit below the bank-index width pins an address bit for the whole warp and would
// serialise the access 2-4x; XOR-ing the bank bits with the next higher index bits makes them vary with the lane again.
template <typename C>
__device__ __forceinline__ uint32_t swz(uint32_t e, uint32_t on) {
    constexpr int W = sizeof(C) == 8 ? 4 : 3;     // 16 (8) elements of 8 (16) bytes span the 32 banks
    return e ^ (((e >> W) & ((1u << W) - 1u)) & on);
}

template <typename T, int NS>
__device__ __forceinline__ void run_pass(const DmInterpArgs& a, const DmPass ps, const OpRec<T>* recs,
                                         typename Cplx<T>::type* const* part, typename Cplx<T>::type* loc, unsigned rank,
                                         unsigned csize, uint32_t swz_on) {
    typedef typename Cplx<T>::type C;
    constexpr int D = 1 << NS, NB = 2 * NS;
    const int n = a.n_qubits, LB = a.local_bits;
    // gate bit positions: rows q0 (, q1), columns q0 + n (, q1 + n); s0 < s1 (< s2 < s3) the same, sorted -- scalars only,
    // nothing here may end up in local memory
    const int p0 = ps.q0, p1 = NS == 2 ? ps.q1 : 0;
    const int lo = NS == 2 ? min(p0, p1) : p0, hi = NS == 2 ? max(p0, p1) : p0;
    const int s0 = lo, s1 = NS == 2 ? hi : lo + n, s2 = lo + n, s3 = hi + n;      // rows < n <= columns: already sorted
    // element (r, c) of a block: swizzled local offset and CTA-rank bits, relative to the block base.  The swizzle is
    // XOR-linear and base / offset bits are disjoint, so swz(base | off) = swz(base) ^ swz(off): one XOR per element
    const uint32_t lmask = (1u << LB) - 1u;
    uint32_t soff[D * D], roff[D * D];
#pragma unroll
    for (int r = 0; r < D; r++)
#pragma unroll
        for (int c = 0; c < D; c++) {
            uint32_t o = (((uint32_t)r & 1u) << p0) | (((uint32_t)c & 1u) << (p0 + n));
            if (NS == 2) o |= (((uint32_t)r >> 1) << p1) | (((uint32_t)c >> 1) << (p1 + n));
            soff[r * D + c] = swz<C>(o & lmask, swz_on);
            roff[r * D + c] = o >> LB;
        }
    const int free_bits = 2 * n - NB;
    const bool remote = csize > 1 && (NS == 2 ? s3 : s1) >= LB;      // a gate bit names the CTA: blocks span CTAs
    uint32_t f, f_end, f_step, f_or;
    if (!remote) {
        const int local_free = free_bits - (2 * n - LB);        // the CTA bits are the top free bits
        f = threadIdx.x; f_end = 1u << local_free; f_step = blockDim.x; f_or = rank << local_free;
    } else {
        f = rank * blockDim.x + threadIdx.x; f_end = 1u << free_bits; f_step = csize * blockDim.x; f_or = 0;
    }
    for (; f < f_end; f += f_step) {
        uint32_t base = gap(gap(f | f_or, s0), s1);
        if (NS == 2) base = gap(gap(base, s2), s3);
        const uint32_t sb = swz<C>(base & lmask, swz_on), rb = base >> LB;
        Block<T, NS> B;
        if (!remote) {
#pragma unroll
            for (int i = 0; i < D * D; i++) B.b[i] = loc[sb ^ soff[i]];
        } else {
#pragma unroll
            for (int i = 0; i < D * D; i++) B.b[i] = part[rb ^ roff[i]][sb ^ soff[i]];
        }
        for (int k = 0; k < ps.n_ops; k++) apply_rec<T, NS>(B, recs[k], a.params);
        if (!remote) {
#pragma unroll
            for (int i = 0; i < D * D; i++) loc[sb ^ soff[i]] = B.b[i];
        } else {
#pragma unroll
            for (int i = 0; i < D * D; i++) part[rb ^ roff[i]][sb ^ soff[i]] = B.b[i];
        }
    }
}

// The dominant pass shape of noisy graph circuits after composition: [S0] [S1] CX|CZ DEPOL [S0'] [S1'] with real
// superoperators (Hadamards, Pauli mixtures, relaxation).  Such a pass runs as ONE straight-line function per block --
// same single shared-memory round trip as the generic loop, without the per-record dispatch and the register shuffling
// it causes (22 % MOV in the generic loop, ncu).
struct FusedPass {
    int ok;              // 0 = generic loop
    int ctrl;            // slot of the CX control; 2 = the gate is a CZ
    int pre0, pre1, post0, post1, depol;   // record indices inside the pass, -1 = absent
};

template <typename T, int S>
__device__ __forceinline__ void fused_super(Block<T, 2>& B, const OpRec<T>* r) {
    if (r == nullptr) return;
    if (r->type == R_SUPER1_XR) apply_super1<T, 2, S, true, true>(B, r->m, r->mr);
    else apply_super1<T, 2, S, false, true>(B, r->m, r->mr);
}

template <typename T, int CTRL>
__device__ __forceinline__ void run_pass_fused(const DmInterpArgs& a, const DmPass ps, const OpRec<T>* recs, const FusedPass fp,
                                               typename Cplx<T>::type* const* part, typename Cplx<T>::type* loc,
                                               unsigned rank, unsigned csize, uint32_t swz_on) {
    typedef typename Cplx<T>::type C;
    const int n = a.n_qubits, LB = a.local_bits;
    const int p0 = ps.q0, p1 = ps.q1;
    const int s0 = min(p0, p1), s1 = max(p0, p1), s2 = s0 + n, s3 = s1 + n;
    const uint32_t lmask = (1u << LB) - 1u;
    uint32_t soff[16], roff[16];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const uint32_t o = (((uint32_t)r & 1u) << p0) | (((uint32_t)c & 1u) << (p0 + n)) | (((uint32_t)r >> 1) << p1) |
                               (((uint32_t)c >> 1) << (p1 + n));
            soff[r * 4 + c] = swz<C>(o & lmask, swz_on);
            roff[r * 4 + c] = o >> LB;
        }
    const int free_bits = 2 * n - 4;
    const bool remote = csize > 1 && s3 >= LB;
    uint32_t f, f_end, f_step, f_or;
    if (!remote) {
        const int local_free = free_bits - (2 * n - LB);
        f = threadIdx.x; f_end = 1u << local_free; f_step = blockDim.x; f_or = rank << local_free;
    } else {
        f = rank * blockDim.x + threadIdx.x; f_end = 1u << free_bits; f_step = csize * blockDim.x; f_or = 0;
    }
    const OpRec<T>* pre0 = fp.pre0 >= 0 ? recs + fp.pre0 : nullptr;
    const OpRec<T>* pre1 = fp.pre1 >= 0 ? recs + fp.pre1 : nullptr;
    const OpRec<T>* post0 = fp.post0 >= 0 ? recs + fp.post0 : nullptr;
    const OpRec<T>* post1 = fp.post1 >= 0 ? recs + fp.post1 : nullptr;
    const bool depol = fp.depol >= 0;
    const T c0 = depol ? recs[fp.depol].aux[0] : (T)1, c1 = depol ? recs[fp.depol].aux[1] : (T)0;
    for (; f < f_end; f += f_step) {
        const uint32_t base = gap(gap(gap(gap(f | f_or, s0), s1), s2), s3);
        const uint32_t sb = swz<C>(base & lmask, swz_on), rb = base >> LB;
        Block<T, 2> B;
        if (!remote) {
#pragma unroll
            for (int i = 0; i < 16; i++) B.b[i] = loc[sb ^ soff[i]];
        } else {
#pragma unroll
            for (int i = 0; i < 16; i++) B.b[i] = part[rb ^ roff[i]][sb ^ soff[i]];
        }
        fused_super<T, 0>(B, pre0);
        fused_super<T, 1>(B, pre1);
        // CX as a renaming of the stores: element (r, c) of the result is element (pi r, pi c) of the block
        if (depol) apply_depol2<T>(B, c0, c1);         // depolarizing commutes with the permutation (trace and scaling)
        Block<T, 2> P;
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) {
                if (CTRL == 2) {           // CZ: element (r, c) changes sign when exactly one of r, c is |11>
                    const bool neg = (r == 3) != (c == 3);
                    P.b[r * 4 + c] = C{neg ? -B.b[r * 4 + c].x : B.b[r * 4 + c].x, neg ? -B.b[r * 4 + c].y : B.b[r * 4 + c].y};
                } else {
                    constexpr int CT = CTRL == 2 ? 0 : CTRL;
                    const int pr = r ^ (((r >> CT) & 1) << (1 - CT)), pc = c ^ (((c >> CT) & 1) << (1 - CT));
                    P.b[r * 4 + c] = B.b[pr * 4 + pc];
                }
            }
        fused_super<T, 0>(P, post0);
        fused_super<T, 1>(P, post1);
        if (!remote) {
#pragma unroll
            for (int i = 0; i < 16; i++) loc[sb ^ soff[i]] = P.b[i];
        } else {
#pragma unroll
            for (int i = 0; i < 16; i++) part[rb ^ roff[i]][sb ^ soff[i]] = P.b[i];
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(sizeof(T) == 4 ? DMI_F32_THREADS : IR_THREADS, 1) dm_interp_kernel(const DmInterpArgs a) {
    typedef typename Cplx<T>::type C;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* loc = reinterpret_cast<C*>(smem_raw);
    __shared__ C* part[16];
    // op records of a GROUP of consecutive passes (as many as fit): built once, then the group's passes run back to back.
    // Building per pass left 15 of 16 warps waiting at a barrier for the sequential composition step (ncu: "barrier" was
    // the top stall, 3.1 warps per issue slot).
    constexpr int GP = 64;                    // passes per group
    __shared__ DmPass spass[GP];
    __shared__ uint32_t srec[GP + 1];         // first record of pass q of the group
    __shared__ int s_gcount;
    __shared__ FusedPass sfused[GP];
    OpRec<T>* recs = reinterpret_cast<OpRec<T>*>(smem_raw + ((size_t)1 << a.local_bits) * sizeof(C));
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank(), csize = cluster.num_blocks();
    const unsigned cluster_id = blockIdx.x / csize, n_clusters = gridDim.x / csize;
    const int n = a.n_qubits, LB = a.local_bits;
    const uint32_t loc_elems = 1u << LB, N = 1u << n;
    const uint32_t swz_on = LB >= 8 ? ~0u : 0u;      // the swizzle reads index bits up to 7
    if (threadIdx.x < csize) part[threadIdx.x] = cluster.map_shared_rank(loc, threadIdx.x);
    auto barrier = [&]() {
        if (csize > 1) cluster.sync();
        else __syncthreads();
    };
    for (unsigned c = cluster_id; c < (unsigned)a.n_circuits; c += n_clusters) {
        for (uint32_t i = threadIdx.x; i < loc_elems; i += blockDim.x) loc[i] = C{(T)((rank == 0 && i == 0) ? 1 : 0), (T)0};   // swz(0) = 0
        barrier();
        const uint32_t p_begin = a.pass_off[a.circuit0 + c], p_end = a.pass_off[a.circuit0 + c + 1];
        for (uint32_t p = p_begin; p < p_end;) {
            // ---- the group: passes p .. p + g_count whose records fit the buffer
            if (threadIdx.x < GP && p + threadIdx.x < p_end) spass[threadIdx.x] = a.passes[p + threadIdx.x];
            __syncthreads();
            if (threadIdx.x == 0) {
                uint32_t tot = 0;
                int g = 0;
                const int lim = (int)min((uint32_t)GP, p_end - p);
                while (g < lim && (g == 0 || tot + spass[g].n_ops <= (uint32_t)a.rec_cap)) {
                    srec[g] = tot;
                    tot += spass[g].n_ops;
                    g++;
                }
                srec[g] = tot;
                s_gcount = g;
            }
            __syncthreads();
            const int g_count = s_gcount;
            const int g_ops = (int)srec[g_count];
            if (threadIdx.x < g_count)
                for (uint32_t j = srec[threadIdx.x]; j < srec[threadIdx.x + 1]; j++) recs[j].pass_q = threadIdx.x;
            __syncthreads();
            // ---- phase A: one record per op, two ops per warp, all ops of the group
            for (int k0 = (threadIdx.x >> 5) * 2; k0 < g_ops; k0 += (blockDim.x >> 5) * 2) {
                const int k = k0 + ((threadIdx.x >> 4) & 1), e = threadIdx.x & 15;
                const bool valid = k < g_ops;
                const DmPass ps = spass[valid ? recs[k].pass_q : 0];
                qcd_op op;
                op.opcode = QCD_OP_ID; op.q0 = ps.q0; op.q1 = 0xFF; op.param_off = 0;
                if (valid) op = a.ops[a.sched[spass[0].first + k]];      // the group's ops are contiguous in the schedule
                const bool two = op.opcode == QCD_OP_CX || op.opcode == QCD_OP_CZ || op.opcode == QCD_OP_SWAP ||
                                 op.opcode == QCD_OP_U2Q || op.opcode == QCD_OP_KRAUS2Q || op.opcode == QCD_OP_PAULI2Q;
                const int slot0 = op.q0 == ps.q0 ? 0 : 1;       // slot of the op's first qubit
                // 1q ops: entry e of the 4x4 superoperator; X shape = nothing outside the diagonal / anti-diagonal pattern
                double2 s{0, 0};
                if (valid && !two) s = super_entry(op, a.params, e >> 2, e & 3);
                const bool inside = e == 0 || e == 3 || e == 5 || e == 6 || e == 9 || e == 10 || e == 12 || e == 15;
                const unsigned bad = __ballot_sync(0xffffffffu, !inside && (s.x != 0.0 || s.y != 0.0));   // whole warp
                if (!valid) continue;
                OpRec<T>& rc = recs[k];
                if (!two) {
                    rc.m[e] = C{(T)s.x, (T)s.y};
                    if (e == 0) {
                        rc.type = ((bad >> (threadIdx.x & 16)) & 0xffffu) ? R_SUPER1 : R_SUPER1_X;
                        rc.slot = slot0;
                    }
                } else if (op.opcode == QCD_OP_U2Q) {
                    // header basis index = bit(q0) + 2 bit(q1); slot basis index = bit(slot 0) + 2 bit(slot 1)
                    const int r = e >> 2, cc = e & 3;
                    const int pr = slot0 ? (((r & 1) << 1) | (r >> 1)) : r, pc = slot0 ? (((cc & 1) << 1) | (cc >> 1)) : cc;
                    rc.m[e] = C{(T)a.params[op.param_off + 2 * (pr * 4 + pc)], (T)a.params[op.param_off + 2 * (pr * 4 + pc) + 1]};
                    if (e == 0) rc.type = R_U2;
                } else if (op.opcode == QCD_OP_PAULI2Q) {
                    // W[f][d] = sum over Paulis with flip pattern f of p (-1)^{popc(z & d)}, f, d, z in slot order
                    const int f = e >> 2, d = e & 3;
                    double w = 0;
                    for (int k0p = 0; k0p < 4; k0p++)
                        for (int k1p = 0; k1p < 4; k1p++) {
                            const int fa = (k0p == 1 || k0p == 2), za = (k0p == 2 || k0p == 3);    // op.q0
                            const int fb = (k1p == 1 || k1p == 2), zb = (k1p == 2 || k1p == 3);    // op.q1
                            const int ff = slot0 ? (fb | (fa << 1)) : (fa | (fb << 1));
                            const int zz = slot0 ? (zb | (za << 1)) : (za | (zb << 1));
                            const double pr = a.params[op.param_off + k0p + 4 * k1p];
                            if (ff == f) w += (__popc(zz & d) & 1) ? -pr : pr;
                        }
                    rc.m[e] = C{(T)w, (T)0};
                    if (e == 0) {
                        const double* pp = a.params + op.param_off;
                        bool uniform = true;
                        for (int q = 2; q < 16; q++) uniform = uniform && pp[q] == pp[1];
                        rc.type = uniform ? R_DEPOL2 : R_PAULI2;
                        rc.aux[0] = (T)(pp[0] - pp[1]);
                        rc.aux[1] = (T)(4.0 * pp[1]);
                    }
                } else if (e == 0) {
                    rc.type = op.opcode == QCD_OP_CX ? R_CX : (op.opcode == QCD_OP_CZ ? R_CZ : (op.opcode == QCD_OP_SWAP ? R_SWAP : R_KRAUS2));
                    rc.slot = slot0;           // CX: slot of the control
                    rc.param = op.param_off;
                    rc.perm = slot0;
                }
            }
            __syncthreads();
            // ---- phase B: consecutive single-qubit records on one slot collapse into ONE superoperator (later op on the
            // left); records on the other slot commute with them, a two-qubit record ends the run.  One half-warp per PASS
            // (the passes of the group in parallel), sequential inside a pass.
            {
                const int e = threadIdx.x & 15, i = e >> 2, j = e & 3;
                const unsigned hmask = 0xffffu << (threadIdx.x & 16);
                for (int q = threadIdx.x >> 4; q < g_count; q += blockDim.x >> 4) {
                    OpRec<T>* pr = recs + srec[q];
                    const int n_ops = (int)(srec[q + 1] - srec[q]);
                    int head0 = -1, head1 = -1;
                    for (int k = 0; k < n_ops; k++) {
                        OpRec<T>& rc = pr[k];
                        const int ty = rc.type;
                        if (ty != R_SUPER1 && ty != R_SUPER1_X) {
                            head0 = head1 = -1;
                            continue;
                        }
                        const int sl = rc.slot;
                        const int hd = sl == 0 ? head0 : head1;
                        if (hd < 0) {
                            if (sl == 0) head0 = k; else head1 = k;
                            continue;
                        }
                        OpRec<T>& h = pr[hd];
                        C acc = cmul(rc.m[i * 4], h.m[j]);
                        cfma(acc, rc.m[i * 4 + 1], h.m[4 + j]);
                        cfma(acc, rc.m[i * 4 + 2], h.m[8 + j]);
                        cfma(acc, rc.m[i * 4 + 3], h.m[12 + j]);
                        __syncwarp(hmask);
                        h.m[e] = acc;
                        if (e == 0) {
                            if (ty == R_SUPER1) h.type = R_SUPER1;
                            rc.type = R_SKIP;
                        }
                        __syncwarp(hmask);
                    }
                    for (int k = 0; k < n_ops; k++) {       // real-arithmetic variants
                        OpRec<T>& rc = pr[k];
                        const int ty = rc.type;
                        const bool one = ty == R_SUPER1 || ty == R_SUPER1_X;
                        const unsigned nz = __ballot_sync(hmask, one && rc.m[e].y != (T)0);
                        __syncwarp(hmask);
                        if (one && nz == 0) {
                            rc.mr[e] = rc.m[e].x;
                            if (e == 0) rc.type = ty == R_SUPER1 ? R_SUPER1_R : R_SUPER1_XR;
                        }
                    }
                    __syncwarp(hmask);
                    if (e == 0) {      // does the pass have the fused shape [S0] [S1] CX [DEPOL] [S0'] [S1'] with real S?
                        FusedPass fp{0, 0, -1, -1, -1, -1, -1};
                        int stage = 0;             // 0 before the CX, 1 after the CX, 2 after the Pauli mixture
                        bool ok = spass[q].q1 != 0xFF;
                        for (int k = 0; k < n_ops && ok; k++) {
                            const int ty = pr[k].type;
                            if (ty == R_SKIP) continue;
                            if (ty == R_SUPER1_R || ty == R_SUPER1_XR) {
                                int& dst = stage == 0 ? (pr[k].slot == 0 ? fp.pre0 : fp.pre1) : (pr[k].slot == 0 ? fp.post0 : fp.post1);
                                ok = dst < 0;
                                dst = k;
                                if (stage == 1) stage = 2;      // a superoperator after the CX: a Pauli mixture may no longer follow
                            } else if ((ty == R_CX || ty == R_CZ) && stage == 0) {
                                fp.ctrl = ty == R_CZ ? 2 : pr[k].slot;
                                stage = 1;
                            } else if (ty == R_DEPOL2 && stage == 1) {
                                fp.depol = k;
                                stage = 2;
                            } else {
                                ok = false;
                            }
                        }
                        fp.ok = ok && stage >= 1;
                        sfused[q] = fp;
                    }
                }
            }
            __syncthreads();
            // ---- the group's passes, back to back
            // (a stage-specialised variant -- straight-line code per [supers][gate][Pauli] group, one shared-memory round
            // trip per stage -- measured slower: unchanged at n = 7, 0.6x at n = 8, 0.45x at n = 9)
            for (int q = 0; q < g_count; q++) {
                const DmPass ps = spass[q];
                const FusedPass fp = sfused[q];
                if (ps.q1 == 0xFF) run_pass<T, 1>(a, ps, recs + srec[q], part, loc, rank, csize, swz_on);
                else if (fp.ok && fp.ctrl == 0) run_pass_fused<T, 0>(a, ps, recs + srec[q], fp, part, loc, rank, csize, swz_on);
                else if (fp.ok && fp.ctrl == 1) run_pass_fused<T, 1>(a, ps, recs + srec[q], fp, part, loc, rank, csize, swz_on);
                else if (fp.ok) run_pass_fused<T, 2>(a, ps, recs + srec[q], fp, part, loc, rank, csize, swz_on);
                else run_pass<T, 2>(a, ps, recs + srec[q], part, loc, rank, csize, swz_on);
                barrier();
            }
            p += g_count;
        }
        // ---- results: the diagonal (outcome probabilities) and / or the whole matrix
        if (a.probs) {
            T* out = reinterpret_cast<T*>(a.probs) + (uint64_t)c * N;
            for (uint32_t x = threadIdx.x; x < N; x += blockDim.x) {
                const uint32_t e = x + (x << n);
                if ((e >> LB) == rank) out[x] = loc[swz<C>(e & (loc_elems - 1u), swz_on)].x;
            }
        }
        if (a.rho_out) {
            C* out = reinterpret_cast<C*>(a.rho_out) + (uint64_t)c * a.slot_stride + (uint64_t)rank * loc_elems;
            for (uint32_t i = threadIdx.x; i < loc_elems; i += blockDim.x) out[i] = loc[swz<C>(i, swz_on)];
        }
        __syncthreads();    // the next circuit's first barrier orders everything else
    }
}

size_t op_param_len(const qcd_op& op, const double* params, size_t n_params) {
    switch (op.opcode) {
        case QCD_OP_RX: case QCD_OP_RY: case QCD_OP_RZ: return 1;
        case QCD_OP_U1Q: return 8;
        case QCD_OP_U2Q: return 32;
        case QCD_OP_PAULI1Q: return 4;
        case QCD_OP_PAULI2Q: return 16;
        case QCD_OP_KRAUS1Q: case QCD_OP_KRAUS2Q: {
            if (op.param_off >= n_params) return (size_t)-1;
            const double m = params[op.param_off];
            if (!(m >= 1 && m <= 64) || m != std::floor(m)) return (size_t)-1;
            return 1 + (size_t)m * (op.opcode == QCD_OP_KRAUS1Q ? 8 : 32);
        }
        default: return 0;
    }
}

}  // namespace

size_t dm_interp_record_bytes(int precision) { return precision == 64 ? sizeof(OpRec<double>) : sizeof(OpRec<float>); }

int dm_interp_rec_cap(int n_qubits, int precision, uint32_t max_circuit_ops) {
    const int lb = std::min(2 * n_qubits, precision == 64 ? 13 : 14);
    const size_t amp = ((size_t)1 << lb) * (precision == 64 ? 16 : 8);
    const size_t room = (size_t)227 * 1024 - 3072 - amp;            // static shared memory + slack
    const size_t fit = room / dm_interp_record_bytes(precision);
    return (int)std::min<size_t>(fit, std::max<uint32_t>(max_circuit_ops, (uint32_t)IR_MAX_OPS));
}

int dm_interp_cluster_size(int n_qubits, int precision) {
    const int lb = precision == 64 ? 13 : 14;               // 128 KB of amplitudes per CTA
    if (n_qubits < 1 || 2 * n_qubits > lb + 4) return 0;    // at most 16 CTAs per cluster
    return 2 * n_qubits <= lb ? 1 : 1 << (2 * n_qubits - lb);
}

int dm_interp_schedule(const qcd_op* ops, uint32_t n_ops, uint32_t op_base, const double* params, size_t n_params,
                       int n_qubits, std::vector<DmPass>& passes, std::vector<uint32_t>& sched, std::string& err) {
    const size_t pass0 = passes.size();
    std::vector<std::vector<uint32_t>> pending(n_qubits);
    std::vector<long long> last(n_qubits, -1);     // index (into passes) of the last pass touching the qubit
    std::vector<std::vector<uint32_t>> pass_ops;    // ops of the passes of THIS circuit
    auto new_pass = [&](int q0, int q1) {
        DmPass p;
        p.first = 0; p.n_ops = 0; p.q0 = (uint8_t)q0; p.q1 = (uint8_t)q1;
        passes.push_back(p);
        pass_ops.emplace_back();
        return (long long)(passes.size() - 1);
    };
    for (uint32_t i = 0; i < n_ops; i++) {
        const qcd_op& op = ops[i];
        if (op.opcode >= QCD_OP_COUNT) { err = "op " + std::to_string(i) + ": unknown opcode"; return QCD_E_INVALID; }
        const bool two = op.opcode == QCD_OP_CX || op.opcode == QCD_OP_CZ || op.opcode == QCD_OP_SWAP ||
                         op.opcode == QCD_OP_U2Q || op.opcode == QCD_OP_KRAUS2Q || op.opcode == QCD_OP_PAULI2Q;
        if (op.q0 >= n_qubits || (two && (op.q1 >= n_qubits || op.q1 == op.q0))) {
            err = "op " + std::to_string(i) + ": qubit index out of range";
            return QCD_E_INVALID;
        }
        const size_t plen = op_param_len(op, params, n_params);
        if (plen == (size_t)-1 || (plen && (size_t)op.param_off + plen > n_params)) {
            err = "op " + std::to_string(i) + ": op parameters run past params[]";
            return QCD_E_INVALID;
        }
        if (op.opcode == QCD_OP_ID) continue;
        if (!two) {
            // after a pass has touched the qubit, later passes either touch it too (then `last` moved on) or commute with
            // this op: it can join the last pass that touched the qubit
            if (last[op.q0] >= 0 && pass_ops[last[op.q0] - pass0].size() < (size_t)IR_MAX_OPS)
                pass_ops[last[op.q0] - pass0].push_back(op_base + i);
            else if (last[op.q0] >= 0) {
                last[op.q0] = new_pass(op.q0, 0xFF);
                pass_ops.back().push_back(op_base + i);
            } else
                pending[op.q0].push_back(op_base + i);
            continue;
        }
        const int a = op.q0, b = op.q1;
        long long p = -1;
        if (last[a] >= 0 && last[a] == last[b]) {
            const DmPass& lp = passes[last[a]];
            if (lp.q1 != 0xFF && pass_ops[last[a] - pass0].size() < (size_t)IR_MAX_OPS) p = last[a];
        }
        if (p < 0) {
            // single-qubit ops waiting on a or b run first, in passes of their own if they do not fit
            for (int q : {a, b}) {
                while (pending[q].size() > (size_t)IR_MAX_OPS / 2 - 1) {
                    const long long pp = new_pass(q, 0xFF);
                    const size_t take = std::min(pending[q].size(), (size_t)IR_MAX_OPS);
                    pass_ops[pp - pass0].assign(pending[q].begin(), pending[q].begin() + take);
                    pending[q].erase(pending[q].begin(), pending[q].begin() + take);
                }
            }
            p = new_pass(a, b);
            for (int q : {a, b}) {
                pass_ops[p - pass0].insert(pass_ops[p - pass0].end(), pending[q].begin(), pending[q].end());
                pending[q].clear();
            }
            last[a] = last[b] = p;
        }
        pass_ops[p - pass0].push_back(op_base + i);
    }
    // qubits no two-qubit gate ever touched: their op lists commute, so two of them share one pass
    auto flush1 = [&](int q) {
        while (!pending[q].empty()) {
            const long long p = new_pass(q, 0xFF);
            const size_t take = std::min(pending[q].size(), (size_t)IR_MAX_OPS);
            pass_ops[p - pass0].assign(pending[q].begin(), pending[q].begin() + take);
            pending[q].erase(pending[q].begin(), pending[q].begin() + take);
        }
    };
    std::vector<int> left;
    for (int q = 0; q < n_qubits; q++)
        if (!pending[q].empty()) left.push_back(q);
    size_t li = 0;
    for (; li + 1 < left.size(); li += 2) {
        const int qa = left[li], qb = left[li + 1];
        if (pending[qa].size() + pending[qb].size() <= (size_t)IR_MAX_OPS) {
            const long long p = new_pass(qa, qb);
            pass_ops[p - pass0] = pending[qa];
            pass_ops[p - pass0].insert(pass_ops[p - pass0].end(), pending[qb].begin(), pending[qb].end());
            pending[qa].clear();
            pending[qb].clear();
        } else {
            flush1(qa);
            flush1(qb);
        }
    }
    if (li < left.size()) flush1(left[li]);
    for (size_t k = 0; k < pass_ops.size(); k++) {
        passes[pass0 + k].first = (uint32_t)sched.size();
        passes[pass0 + k].n_ops = (uint16_t)pass_ops[k].size();
        sched.insert(sched.end(), pass_ops[k].begin(), pass_ops[k].end());
    }
    return QCD_OK;
}

template <typename T>
cudaError_t launch_dm_interp(const DmInterpArgs& a, int cluster_size, int max_clusters, cudaStream_t st) {
    typedef typename Cplx<T>::type C;
    const size_t smem = ((size_t)1 << a.local_bits) * sizeof(C) + (size_t)a.rec_cap * dm_interp_record_bytes(sizeof(T) == 8 ? 64 : 32);
    cudaError_t e = cudaFuncSetAttribute(dm_interp_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (cluster_size > 8) {
        e = cudaFuncSetAttribute(dm_interp_kernel<T>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(sizeof(T) == 4 ? DMI_F32_THREADS : IR_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster_size;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n_clusters = max_clusters;
    if (n_clusters <= 0) {
        // resident clusters for this (device, precision, cluster size, shared memory): asked once, not per launch (the object
        // route of simulate() launches one circuit at a time)
        struct Key { int dev, cs; size_t smem; int n; };
        static Key cache[16];
        static int n_cached = 0;
        int dev = 0;
        cudaGetDevice(&dev);
        for (int i = 0; i < n_cached; i++)
            if (cache[i].dev == dev && cache[i].cs == cluster_size * (sizeof(T) == 8 ? -1 : 1) && cache[i].smem == smem) n_clusters = cache[i].n;
        if (n_clusters <= 0) {
            cfg.gridDim = dim3((unsigned)cluster_size);
            e = cudaOccupancyMaxActiveClusters(&n_clusters, dm_interp_kernel<T>, &cfg);
            if (e != cudaSuccess) return e;
            if (n_clusters < 1) return cudaErrorLaunchOutOfResources;
            if (n_cached < 16) cache[n_cached++] = Key{dev, cluster_size * (sizeof(T) == 8 ? -1 : 1), smem, n_clusters};
        }
    }
    n_clusters = std::min(n_clusters, a.n_circuits);
    cfg.gridDim = dim3((unsigned)(n_clusters * cluster_size));
    return cudaLaunchKernelEx(&cfg, dm_interp_kernel<T>, a);
}

template cudaError_t launch_dm_interp<float>(const DmInterpArgs&, int, int, cudaStream_t);
template cudaError_t launch_dm_interp<double>(const DmInterpArgs&, int, int, cudaStream_t);

}  // namespace qcd
