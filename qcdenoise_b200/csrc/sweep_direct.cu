// sweep_direct.cu -- register-only state sweep for sweeps that consist of ONE pass over <= 4 register bits (the common
// case in trajectory mode: the Kraus variants of one site group + the few gates up to the next site; in
// density-matrix mode every pass of a multi-pass sweep is launched as such a sweep).
//
// No shared-memory tile: a thread owns one "register group" = the 16 amplitudes that differ only in the pass's 4
// register bits, loads them straight from global memory (128-bit loads when index bit 0 is a register bit, otherwise
// 64-bit loads that neighbouring lanes complete to full sectors), applies every fused op of the sweep in registers and
// stores them back: exactly one read + one write of the state, with all 16 loads of a thread in flight at once and
// occupancy not limited by a 64 KB tile.  A CTA owns a contiguous, aligned range of register groups of one state
// (>= 256), so that
//   * the joint populations of the next site's qubits reduce per thread -> warp shuffle -> one double4 per CTA
//     (fixed order: deterministic), and
//   * on the final sweep every 32-amplitude probability chunk lies inside ONE WARP: it is summed over the thread's own
//     values (register bits below 5) and by xor-shuffles over the lanes that hold the other low index bits.
// The sweep's description (register bits, ops, the matrices of this node's branch) is staged once per CTA in shared
// memory: three dependent global reads (node -> descriptor -> matrices), then the streaming loop.
//
// Leaf states that only a few trajectories sample from are NOT written (NODE_NOSTORE): the final sweep then only emits
// the chunk sums, and sample_nostore_kernel recomputes the one 32-amplitude chunk each trajectory lands in from the
// parent state with the same staging + apply code (identical arithmetic).
#include "kernels.cuh"

#include "../../include/qcdsim.h"

namespace qcd {
namespace {

#ifndef DIRECT_DT
#define DIRECT_DT 128
#endif
constexpr int DT = DIRECT_DT;  // threads per CTA
#ifndef DIRECT_MINB
#define DIRECT_MINB 8
#endif
constexpr int MAX_SIGNS = 48;  // sign masks staged per sweep (more: the sweep is not direct-eligible)

template <typename T> struct Cx;
template <> struct Cx<float> { typedef float2 C; };
template <> struct Cx<double> { typedef double2 C; };

template <typename C>
__device__ __forceinline__ C cmul(C a, C b) {
    C r;
    r.x = a.x * b.x - a.y * b.y;
    r.y = a.x * b.y + a.y * b.x;
    return r;
}
template <typename C>
__device__ __forceinline__ void cfma(C& acc, C a, C b) {
    acc.x += a.x * b.x - a.y * b.y;
    acc.y += a.x * b.y + a.y * b.x;
}

template <int TB, typename C>
__device__ __forceinline__ void m1(C (&a)[16], const C* m) {
    const C m00 = m[0], m01 = m[1], m10 = m[2], m11 = m[3];
#pragma unroll
    for (int s = 0; s < 16; s++) {
        if (s & (1 << TB)) continue;
        const C x0 = a[s], x1 = a[s | (1 << TB)];
        C y0 = cmul(m00, x0);
        cfma(y0, m01, x1);
        C y1 = cmul(m10, x0);
        cfma(y1, m11, x1);
        a[s] = y0;
        a[s | (1 << TB)] = y1;
    }
}
template <int T0, int T1, typename C>
__device__ __forceinline__ void m2(C (&a)[16], const C* m) {
#pragma unroll
    for (int s = 0; s < 16; s++) {
        if (s & ((1 << T0) | (1 << T1))) continue;
        const C x[4] = {a[s], a[s | (1 << T0)], a[s | (1 << T1)], a[s | (1 << T0) | (1 << T1)]};
        C y[4];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            y[r] = cmul(m[r * 4], x[0]);
            cfma(y[r], m[r * 4 + 1], x[1]);
            cfma(y[r], m[r * 4 + 2], x[2]);
            cfma(y[r], m[r * 4 + 3], x[3]);
        }
        a[s] = y[0];
        a[s | (1 << T0)] = y[1];
        a[s | (1 << T1)] = y[2];
        a[s | (1 << T0) | (1 << T1)] = y[3];
    }
}
// real diagonal (no-jump damping Kraus, Z-like): two multipliers
template <int TB, typename C>
__device__ __forceinline__ void m1d(C (&a)[16], const C* m) {
    const auto d0 = m[0].x, d1 = m[3].x;
#pragma unroll
    for (int s = 0; s < 16; s++) {
        const auto d = (s & (1 << TB)) ? d1 : d0;
        a[s].x *= d;
        a[s].y *= d;
    }
}
// real-coefficient variants (H, Pauli X/Z, damping Kraus, Ry, ...): half the multiplies
template <int TB, typename C>
__device__ __forceinline__ void m1r(C (&a)[16], const C* m) {
    const auto m00 = m[0].x, m01 = m[1].x, m10 = m[2].x, m11 = m[3].x;
#pragma unroll
    for (int s = 0; s < 16; s++) {
        if (s & (1 << TB)) continue;
        const C x0 = a[s], x1 = a[s | (1 << TB)];
        C y0, y1;
        y0.x = m00 * x0.x + m01 * x1.x;
        y0.y = m00 * x0.y + m01 * x1.y;
        y1.x = m10 * x0.x + m11 * x1.x;
        y1.y = m10 * x0.y + m11 * x1.y;
        a[s] = y0;
        a[s | (1 << TB)] = y1;
    }
}
template <int T0, int T1, typename C>
__device__ __forceinline__ void m2r(C (&a)[16], const C* m) {
#pragma unroll
    for (int s = 0; s < 16; s++) {
        if (s & ((1 << T0) | (1 << T1))) continue;
        const C x[4] = {a[s], a[s | (1 << T0)], a[s | (1 << T1)], a[s | (1 << T0) | (1 << T1)]};
        C y[4];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            y[r].x = m[r * 4].x * x[0].x + m[r * 4 + 1].x * x[1].x + m[r * 4 + 2].x * x[2].x + m[r * 4 + 3].x * x[3].x;
            y[r].y = m[r * 4].x * x[0].y + m[r * 4 + 1].x * x[1].y + m[r * 4 + 2].x * x[2].y + m[r * 4 + 3].x * x[3].y;
        }
        a[s] = y[0];
        a[s | (1 << T0)] = y[1];
        a[s | (1 << T1)] = y[2];
        a[s | (1 << T0) | (1 << T1)] = y[3];
    }
}

__device__ __forceinline__ void ldv(const float2* p, float2& a, float2& b) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    a = make_float2(v.x, v.y);
    b = make_float2(v.z, v.w);
}
__device__ __forceinline__ void ldv(const double2* p, double2& a, double2& b) {
    a = p[0];
    b = p[1];
}
__device__ __forceinline__ void stv(float2* p, float2 a, float2 b) {
    *reinterpret_cast<float4*>(p) = make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void stv(double2* p, double2 a, double2 b) {
    p[0] = a;
    p[1] = b;
}

// ---------------------------------------------------------------------------------------------------------------
// pre-decoded op codes (low 8 bits; K_SIGNS: count in bits 8..15, slot offset in bits 16..23; K_D1: global bit in 8..15)
enum : uint32_t {
    C_M1D = 0,    // +t0   real diagonal 2x2
    C_M1R = 4,    // +t0   real 2x2
    C_M1C = 8,    // +t0   complex 2x2
    C_M2R = 12,   // +pair real 4x4, pair index: (0,1) (0,2) (0,3) (1,2) (1,3) (2,3)
    C_M2C = 18,   // +pair complex 4x4
    C_SIGNS = 24,
    C_D1 = 25,
    C_M4 = 26     // 16x16 on the 4 register bits (density-matrix superoperators); bit 8: real coefficients
};
__device__ __forceinline__ uint32_t pair_index(int t0, int t1) {
    return t0 == 0 ? (uint32_t)(t1 - 1) : (t0 == 1 ? (uint32_t)(t1 + 1) : 5u);
}

// Per-sweep staging area (one per CTA in the sweep kernel, one per warp in the recompute-sampling kernel)
template <typename T>
struct Staged {
    typedef typename Cx<T>::C C;
    DirectSweep D;
    C mats[MAX_DIRECT_OPS][16];      // the selected variant of every matrix op
    C m4[256];                       // the (single) 16x16 op of the pass, row-major
    uint32_t sign_base[MAX_SIGNS];   // part of each sign mask outside the register bits
    uint32_t sign_pat[MAX_SIGNS];    // per-value pattern of the part inside the register bits
    uint32_t kind_flags[MAX_DIRECT_OPS];  // 0 complex, 1 real, 2 real diagonal
    uint32_t code[MAX_DIRECT_OPS];        // pre-decoded op: one jump-table dispatch per op in the streaming loop
    int scale_op;                    // matrix op that carries the node's rescale factor (-1: none)
    uint32_t slots;                  // slotted fast path: 8 nibbles = op index of the real 2x2 on register bit r before
                                     // (nibbles 0..3) / after (4..7) the sign layer, 0xF = none;  0 = shape not slotted
    uint32_t sign_code;              // code word of the sign layer of the slotted shape (0: none)
    C prep[64];                      // prep sweeps: (alpha, beta) per index bit
    C prep_reg[16];                  // ... and the 16 products over the register bits
};

template <bool WARP>
__device__ __forceinline__ void group_sync() {
    if (WARP) __syncwarp();
    else __syncthreads();
}

// Stage descriptor `desc`, the matrices of branch word `branch` (rescale folded into the first matrix op) and the
// sign masks.  Executed by NT cooperating threads (tid in [0, NT)).
template <typename T, int NT, bool WARP>
__device__ __forceinline__ void stage_sweep(Staged<T>& S, const DevProgram<T>& prog, uint32_t desc, uint32_t branch,
                                            double scale, uint32_t tid) {
    typedef typename Cx<T>::C C;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(prog.direct + desc);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&S.D);
        for (uint32_t e = tid; e < sizeof(DirectSweep) / 4; e += NT) dst[e] = src[e];
    }
    group_sync<WARP>();
    const int nops = S.D.nops, n_bits = S.D.n_bits;
    uint32_t go[4];
#pragma unroll
    for (int i = 0; i < 4; i++) go[i] = 1u << S.D.rbit[i];
    const uint32_t regmask = go[0] | go[1] | go[2] | go[3];
    for (int o = 0; o < nops; o++) {
        const KOp op = S.D.ops[o];
        if (op.kind == K_SIGNS) {
            for (uint32_t e = tid; e < op.count; e += NT) {
                const uint32_t mk = prog.masks[op.off + e];
                const uint32_t mr = mk & regmask;
                uint32_t pat = 0;
#pragma unroll
                for (int s = 0; s < 16; s++) {
                    const uint32_t g = ((s & 1) ? go[0] : 0u) | ((s & 2) ? go[1] : 0u) | ((s & 4) ? go[2] : 0u) |
                                       ((s & 8) ? go[3] : 0u);
                    pat |= ((g & mr) == mr ? 1u : 0u) << s;
                }
                // K_SIGNS ops of one sweep share the staging arrays: `t1` carries the op's slot offset (see host)
                S.sign_base[op.t1 + e] = mk & ~regmask;
                S.sign_pat[op.t1 + e] = pat;
            }
        } else {
            const C* m = prog.pool + op.off + ((branch >> op.sel_shift) & op.sel_mask) * op.stride;
            if (op.kind == K_M4) {
                for (uint32_t e = tid; e < 256u; e += NT) S.m4[e] = m[e];
            } else {
                const uint32_t cnt = op.kind == K_M1 ? 4u : (op.kind == K_M2 ? 16u : 2u);
                for (uint32_t e = tid; e < cnt; e += NT) S.mats[o][e] = m[e];
            }
        }
    }
    const bool is_prep = (S.D.flags & SW_PREP) != 0;
    if (is_prep)
        for (uint32_t e = tid; e < 2u * (uint32_t)n_bits; e += NT) S.prep[e] = prog.pool[S.D.prep_off + e];
    if (tid == 0) {
        int so = -1;
        for (int o = 0; o < nops && so < 0; o++)
            if (S.D.ops[o].kind == K_M1 || S.D.ops[o].kind == K_M2) so = o;
        S.scale_op = so;
    }
    group_sync<WARP>();
    for (uint32_t o = tid; o < (uint32_t)nops; o += NT) {
        // classify the staged matrix; the first matrix op absorbs the node's 1/sqrt(p) rescale
        const int kind = S.D.ops[o].kind;
        const int cnt = kind == K_M1 ? 4 : (kind == K_M2 ? 16 : (kind == K_D1 ? 2 : 0));
        const T f = ((int)o == S.scale_op) ? (T)scale : (T)1;
        bool real = true;
        for (int e = 0; e < cnt; e++) {
            C v = S.mats[o][e];
            v.x *= f;
            v.y *= f;
            S.mats[o][e] = v;
            real = real && v.y == (T)0;
        }
        const bool diag = real && kind == K_M1 && S.mats[o][1].x == (T)0 && S.mats[o][2].x == (T)0;
        S.kind_flags[o] = diag ? 2u : (real ? 1u : 0u);
        const KOp op = S.D.ops[o];
        uint32_t code;
        if (kind == K_M1) code = (diag ? C_M1D : (real ? C_M1R : C_M1C)) + op.t0;
        else if (kind == K_M2) code = (real ? C_M2R : C_M2C) + pair_index(op.t0, op.t1);
        else if (kind == K_SIGNS) code = C_SIGNS | ((uint32_t)op.count << 8) | ((uint32_t)op.t1 << 16);
        else if (kind == K_M4) {
            bool r4 = true;
            for (int e = 0; e < 256; e++) r4 = r4 && S.m4[e].y == (T)0;
            code = C_M4 | (r4 ? 0x100u : 0u);
        } else code = C_D1 | ((uint32_t)op.gbit << 8);
        S.code[o] = code;
    }
    group_sync<WARP>();
    if (tid == 0) {
        // Slotted shape: [real 2x2 ops on DISTINCT register bits] [one sign layer] [real 2x2 ops on distinct register
        // bits].  The ops of one side commute, so the streaming loop can run them in register-bit order with a
        // uniform predicate per slot instead of a jump-table dispatch per op.
        uint32_t slots = 0xFFFFFFFFu, sign_code = 0;
        bool ok = true;
        int phase = 0;
        for (int o = 0; o < nops && ok; o++) {
            const uint32_t code = S.code[o] & 0xffu;
            if (code == C_SIGNS) {
                if (phase == 1) ok = false;
                phase = 1;
                sign_code = S.code[o];
            } else if (code < C_M1C) {   // real diagonal or real 2x2
                const int slot = phase * 4 + (int)(code & 3u);
                if (((slots >> (4 * slot)) & 0xFu) != 0xFu) ok = false;
                slots = (slots & ~(0xFu << (4 * slot))) | ((uint32_t)o << (4 * slot));
            } else {
                ok = false;
            }
        }
        S.slots = ok ? slots : 0u;
        S.sign_code = ok ? sign_code : 0u;
    }
    if (is_prep) {
        for (uint32_t sidx = tid; sidx < 16; sidx += NT) {
            C v;
            v.x = 1;
            v.y = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) v = cmul(v, S.prep[2 * S.D.rbit[i] + ((sidx >> i) & 1)]);
            S.prep_reg[sidx] = v;
        }
    }
    group_sync<WARP>();
}

template <typename T>
__device__ __forceinline__ void apply_signs(typename Cx<T>::C (&a)[16], const Staged<T>& S, uint32_t code, uint32_t gb) {
    const uint32_t cnt = (code >> 8) & 0xffu, off = code >> 16;
    uint32_t par = 0;
    for (uint32_t t = 0; t < cnt; t++) {
        const uint32_t mb = S.sign_base[off + t];
        par ^= ((gb & mb) == mb) ? S.sign_pat[off + t] : 0u;
    }
#pragma unroll
    for (int s = 0; s < 16; s++) {
        const auto sg = ((par >> s) & 1u) ? (decltype(a[s].x))-1 : (decltype(a[s].x))1;
        a[s].x *= sg;
        a[s].y *= sg;
    }
}

// All fused ops of the sweep on one register group (gb = its base index, register bits clear).
template <typename T>
__device__ __forceinline__ void apply_sweep(typename Cx<T>::C (&a)[16], const Staged<T>& S, int nops, uint32_t gb,
                                            const uint32_t (&go)[4], typename Cx<T>::C* ybuf /*[16][stride] or null*/,
                                            uint32_t ystride, uint32_t ytid) {
    typedef typename Cx<T>::C C;
    for (int o = 0; o < nops; o++) {
        const uint32_t code = S.code[o];
        const C* mt = S.mats[o];
        switch (code & 0xffu) {
            case C_M1D + 0: m1d<0>(a, mt); break;
            case C_M1D + 1: m1d<1>(a, mt); break;
            case C_M1D + 2: m1d<2>(a, mt); break;
            case C_M1D + 3: m1d<3>(a, mt); break;
            case C_M1R + 0: m1r<0>(a, mt); break;
            case C_M1R + 1: m1r<1>(a, mt); break;
            case C_M1R + 2: m1r<2>(a, mt); break;
            case C_M1R + 3: m1r<3>(a, mt); break;
            case C_M1C + 0: m1<0>(a, mt); break;
            case C_M1C + 1: m1<1>(a, mt); break;
            case C_M1C + 2: m1<2>(a, mt); break;
            case C_M1C + 3: m1<3>(a, mt); break;
            case C_M2R + 0: m2r<0, 1>(a, mt); break;
            case C_M2R + 1: m2r<0, 2>(a, mt); break;
            case C_M2R + 2: m2r<0, 3>(a, mt); break;
            case C_M2R + 3: m2r<1, 2>(a, mt); break;
            case C_M2R + 4: m2r<1, 3>(a, mt); break;
            case C_M2R + 5: m2r<2, 3>(a, mt); break;
            case C_M2C + 0: m2<0, 1>(a, mt); break;
            case C_M2C + 1: m2<0, 2>(a, mt); break;
            case C_M2C + 2: m2<0, 3>(a, mt); break;
            case C_M2C + 3: m2<1, 2>(a, mt); break;
            case C_M2C + 4: m2<1, 3>(a, mt); break;
            case C_M2C + 5: m2<2, 3>(a, mt); break;
            case C_SIGNS:
                apply_signs<T>(a, S, code, gb);
                break;
            case C_M4: {
                // y = M a with M 16x16: rows one at a time, results parked in this thread's private shared-memory
                // column (dynamic row index), then read back into the registers
                if (ybuf) {
                    const bool r4 = (code & 0x100u) != 0;
                    for (int r = 0; r < 16; r++) {
                        const C* row = S.m4 + r * 16;
                        C acc;
                        acc.x = 0;
                        acc.y = 0;
                        if (r4) {
#pragma unroll
                            for (int c = 0; c < 16; c++) {
                                acc.x += row[c].x * a[c].x;
                                acc.y += row[c].x * a[c].y;
                            }
                        } else {
#pragma unroll
                            for (int c = 0; c < 16; c++) cfma(acc, row[c], a[c]);
                        }
                        ybuf[r * ystride + ytid] = acc;
                    }
#pragma unroll
                    for (int c = 0; c < 16; c++) a[c] = ybuf[c * ystride + ytid];
                }
                break;
            }
            default: {  // C_D1
                const C d0 = mt[0], d1 = mt[1];
                const uint32_t gbit = (code >> 8) & 0xffu;
#pragma unroll
                for (int s = 0; s < 16; s++) {
                    const uint32_t g = gb | ((s & 1) ? go[0] : 0u) | ((s & 2) ? go[1] : 0u) | ((s & 4) ? go[2] : 0u) |
                                       ((s & 8) ? go[3] : 0u);
                    a[s] = cmul(a[s], ((g >> gbit) & 1u) ? d1 : d0);
                }
                break;
            }
        }
    }
}

// Slotted fast path (see stage_sweep): `slots` != 0.
template <typename T>
__device__ __forceinline__ void apply_slotted(typename Cx<T>::C (&a)[16], const Staged<T>& S, uint32_t slots,
                                              uint32_t sign_code, uint32_t gb) {
    // the pre-sign slots hold the Kraus variants: the no-jump operators are diagonal (two multipliers per pair)
#define QCD_SLOT(shift, r)                                                   \
    {                                                                        \
        const uint32_t o = (slots >> (shift)) & 0xFu;                        \
        if (o != 0xFu) {                                                     \
            if (S.kind_flags[o] == 2u) m1d<r>(a, S.mats[o]);                 \
            else m1r<r>(a, S.mats[o]);                                       \
        }                                                                    \
    }
    QCD_SLOT(0, 0) QCD_SLOT(4, 1) QCD_SLOT(8, 2) QCD_SLOT(12, 3)
    if (sign_code) apply_signs<T>(a, S, sign_code, gb);
    QCD_SLOT(16, 0) QCD_SLOT(20, 1) QCD_SLOT(24, 2) QCD_SLOT(28, 3)
#undef QCD_SLOT
}

// The 16 amplitudes of register group gb BEFORE the sweep's ops: product state (prep sweeps) or the parent state.
template <typename T, bool VEC>
__device__ __forceinline__ void load_group(typename Cx<T>::C (&a)[16], const Staged<T>& S,
                                           const typename Cx<T>::C* src, uint32_t gb, const uint32_t (&go)[4],
                                           uint32_t regmask, T sc, bool is_prep, bool need_scale) {
    typedef typename Cx<T>::C C;
    if (is_prep) {
        C base;
        base.x = 1;
        base.y = 0;
        for (int b = 0; b < S.D.n_bits; b++)
            if (!((regmask >> b) & 1u)) base = cmul(base, S.prep[2 * b + ((gb >> b) & 1u)]);
#pragma unroll
        for (int s = 0; s < 16; s++) a[s] = cmul(base, S.prep_reg[s]);
    } else if (VEC) {
#pragma unroll
        for (int s = 0; s < 16; s += 2) {
            const uint32_t g = gb | ((s & 2) ? go[1] : 0u) | ((s & 4) ? go[2] : 0u) | ((s & 8) ? go[3] : 0u);
            ldv(src + g, a[s], a[s + 1]);
        }
    } else {
#pragma unroll
        for (int s = 0; s < 16; s++) {
            const uint32_t g = gb | ((s & 1) ? go[0] : 0u) | ((s & 2) ? go[1] : 0u) | ((s & 4) ? go[2] : 0u) |
                               ((s & 8) ? go[3] : 0u);
            a[s] = src[g];
        }
    }
    if (need_scale) {   // no matrix op to carry the rescale factor
#pragma unroll
        for (int s = 0; s < 16; s++) {
            a[s].x *= sc;
            a[s].y *= sc;
        }
    }
}

// Probability chunk sums of one warp-iteration.  w[s] = |amplitude|^2 of this thread's 16 values.  The 32-element chunk
// bits are R_LOW register bits (the lowest register indices) + the lowest (5 - R_LOW) lane bits.
//   (1) fold the R_LOW register bits in-thread: L = 16 >> R_LOW live partial sums per thread (distinct chunks);
//   (2) butterfly reduce-scatter over the G = 5 - R_LOW lane bits: while more than one value is live, lane bit k decides
//       which half of the values a lane keeps (it adds its partner's copy of that half) -- L-1 shuffles instead of
//       L*G; remaining lane bits are plain xor-reductions;
//   (3) every lane ends up owning max(1, L >> G) chunk sums and stores them (lanes whose unused bits are 0).
template <int R_LOW, typename T>
__device__ __forceinline__ void chunk_sums(T (&w)[16], double* csum, uint32_t gb, const uint32_t (&go)[4], uint32_t lane) {
#pragma unroll
    for (int i = 0; i < R_LOW; i++) {
#pragma unroll
        for (int s = 0; s < 16; s++)
            if (!(s & (1 << i))) w[s] += w[s | (1 << i)];
    }
    constexpr int L = 16 >> R_LOW;       // live values: s = j << R_LOW
    constexpr int G = CHUNK_LOG - R_LOW;  // lane bits inside the chunk
    T v[L];
#pragma unroll
    for (int j = 0; j < L; j++) v[j] = w[j << R_LOW];
    int cur = L;          // compile-time after unrolling
    uint32_t sel = 0;     // which original live index this lane's v[0..cur) block starts at
#pragma unroll
    for (int k = 0; k < G; k++) {
        const bool up = (lane >> k) & 1u;
        if (cur > 1) {
            const int half = cur >> 1;
#pragma unroll
            for (int j = 0; j < L / 2; j++) {
                if (j < half) {
                    const T send = up ? v[j] : v[j + half];
                    const T recv = __shfl_xor_sync(0xffffffffu, send, 1 << k);
                    const T keep = up ? v[j + half] : v[j];
                    v[j] = keep + recv;
                }
            }
            sel += up ? (uint32_t)half : 0u;
            cur = half;
        } else {
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1 << k);
        }
    }
    // lane bits consumed by plain reductions (k >= log2 L) must be zero for the writer
    constexpr int LOG_L = L == 16 ? 4 : (L == 8 ? 3 : (L == 4 ? 2 : (L == 2 ? 1 : 0)));
    constexpr uint32_t plain_mask = G > LOG_L ? (((1u << G) - 1u) & ~((1u << LOG_L) - 1u)) : 0u;
    if ((lane & plain_mask) == 0) {
#pragma unroll
        for (int j = 0; j < L; j++) {
            if (j < cur) {
                const uint32_t sidx = (sel + (uint32_t)j) << R_LOW;   // original value index (low R_LOW bits clear)
                const uint32_t g = (gb & ~((1u << CHUNK_LOG) - 1u)) | ((sidx & 1u) ? go[0] : 0u) | ((sidx & 2u) ? go[1] : 0u) |
                                   ((sidx & 4u) ? go[2] : 0u) | ((sidx & 8u) ? go[3] : 0u);
                csum[g >> CHUNK_LOG] = (double)v[j];
            }
        }
    }
}

// r[(bit I0 of s) | (bit I1 of s) << 1] += w[s]   (I = 4: that reduced bit is not a register bit -> contributes 0)
template <int I0, int I1, typename T>
__device__ __forceinline__ void rel_accumulate(T (&r)[4], const T (&w)[16]) {
#pragma unroll
    for (int s = 0; s < 16; s++) {
        const int b0 = I0 < 4 ? ((s >> I0) & 1) : 0;
        const int b1 = I1 < 4 ? ((s >> I1) & 1) : 0;
        r[b0 | (b1 << 1)] += w[s];
    }
}
template <typename T>
__device__ __forceinline__ void rel_accumulate_dyn(T (&r)[4], const T (&w)[16], int red_case) {
    switch (red_case) {
#define RC(i0, i1) case (i0) * 5 + (i1): rel_accumulate<i0, i1>(r, w); break;
        RC(0, 1) RC(0, 2) RC(0, 3) RC(0, 4) RC(1, 0) RC(1, 2) RC(1, 3) RC(1, 4) RC(2, 0) RC(2, 1) RC(2, 3) RC(2, 4)
        RC(3, 0) RC(3, 1) RC(3, 2) RC(3, 4) RC(4, 0) RC(4, 1) RC(4, 2) RC(4, 3)
#undef RC
        default: rel_accumulate<4, 4>(r, w); break;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// KIND selects which epilogues / sources are compiled in (dead paths cost registers in this kernel):
//   0 PLAIN   no reduction, no chunk sums; product-state source allowed      (density-matrix passes, prep sweeps)
//   1 INNER   population reduction                                           (inner tree levels)
//   2 FINAL   chunk sums                                                     (leaf level)
//   3 GENERIC everything                                                     (mixed launches)
template <typename T, bool VEC, int KIND>
__global__ void __launch_bounds__(DT, sizeof(T) == 4 ? DIRECT_MINB : DIRECT_MINB / 2)
    sweep_direct_kernel(DevProgram<T> prog, SweepArgs args) {
    typedef typename Cx<T>::C C;
    constexpr bool HAS_RED = KIND == 1 || KIND == 3, HAS_CHUNKS = KIND == 2 || KIND == 3,
                   HAS_PREP = KIND == 0 || KIND == 3;
    __shared__ Staged<T> S;
    __shared__ double s_red[DT / 32][4];
    constexpr bool HAS_M4 = KIND == 0 || KIND == 3;
    __shared__ C s_ybuf[HAS_M4 ? 16 : 1][DT];

    const uint32_t tid = threadIdx.x;
    const uint32_t node_idx = blockIdx.x >> args.tiles_log2;
    const uint32_t part = blockIdx.x & ((1u << args.tiles_log2) - 1u);
    const NodeRec N = args.nodes[node_idx];
    stage_sweep<T, DT, false>(S, prog, args.desc_from_pad ? N.pad : N.sweep, N.branch, N.scale, tid);
    const DirectSweep& D = S.D;
    const int n_bits = D.n_bits;
    uint32_t go[4];  // global offset of register bit i
#pragma unroll
    for (int i = 0; i < 4; i++) go[i] = 1u << D.rbit[i];
    const uint32_t regmask = go[0] | go[1] | go[2] | go[3];

    const C* src = reinterpret_cast<const C*>(args.states) + (uint64_t)N.src_slot * args.slot_stride;
    C* dst = reinterpret_cast<C*>(args.states) + (uint64_t)N.dst_slot * args.slot_stride;
    const bool store = !(N.flags & NODE_NOSTORE);
    const T sc = (T)N.scale;
    const bool is_prep = HAS_PREP && (D.flags & SW_PREP) != 0, need_scale = S.scale_op < 0;
    const int nops = D.nops;
    const uint32_t slots = args.no_slots ? 0u : S.slots, sign_code = S.sign_code;
    const uint32_t groups_per_cta = 1u << (n_bits - 4 - (int)args.tiles_log2);
    const uint32_t grp0 = part * groups_per_cta;
    const int rq0 = D.red_q[0], rq1 = D.red_q[1];
    const bool do_red = HAS_RED && rq0 >= 0;
    const bool do_chunks = HAS_CHUNKS && (D.flags & SW_FINAL) != 0 && args.chunk_sums != nullptr;
    int r_low = 0;  // register bits below the chunk size
#pragma unroll
    for (int i = 0; i < 4; i++) r_low += (D.rbit[i] < CHUNK_LOG) ? 1 : 0;
    double* csum = args.chunk_sums + (uint64_t)N.dst_slot * args.chunk_stride;
    // Population reduction: value s falls into bin (bit rq0) | (bit rq1) << 1 of its global index.  A reduced bit that
    // is a register bit depends on s only ("relative bin", compile-time selected by red_case); one that is not
    // contributes a base bin that is constant per thread (red_fast) or changes per iteration.
    T rel[4] = {0, 0, 0, 0};   // red_fast: accumulated over all iterations;  else: per iteration
    T p[4] = {0, 0, 0, 0};
    bool red_fast = true;
    int red_case = 24;
    uint32_t base_mask0 = 0, base_mask1 = 0;   // global-index masks of the reduced bits that are NOT register bits
    if (do_red) {
        const int gl = n_bits - 4 - (int)args.tiles_log2;     // log2(register groups per CTA)
        const int rq[2] = {rq0, rq1};
        int idx[2] = {4, 4};
#pragma unroll
        for (int j = 0; j < 2; j++) {
            if (rq[j] < 0) continue;
            if ((regmask >> rq[j]) & 1u) {
#pragma unroll
                for (int i = 0; i < 4; i++)
                    if ((int)D.rbit[i] == rq[j]) idx[j] = i;
                continue;
            }
            if (j == 0) base_mask0 = 1u << rq[j];
            else base_mask1 = 1u << rq[j];
            const int cp = rq[j] - __popc(regmask & ((1u << rq[j]) - 1u));   // bit position in the group index
            if (cp >= (DT == 256 ? 8 : 7) && cp < gl) red_fast = false;
        }
        red_case = idx[0] * 5 + idx[1];
    }
    uint32_t gb_last = 0;

    for (uint32_t it = tid; it < groups_per_cta; it += DT) {
        uint32_t gb = grp0 + it;   // insert a zero at every register bit (ascending)
        gb = ((gb & ~(go[0] - 1u)) << 1) | (gb & (go[0] - 1u));
        gb = ((gb & ~(go[1] - 1u)) << 1) | (gb & (go[1] - 1u));
        gb = ((gb & ~(go[2] - 1u)) << 1) | (gb & (go[2] - 1u));
        gb = ((gb & ~(go[3] - 1u)) << 1) | (gb & (go[3] - 1u));
        C a[16];
        load_group<T, VEC>(a, S, src, gb, go, regmask, sc, is_prep, need_scale);
        if (slots) apply_slotted<T>(a, S, slots, sign_code, gb);
        else apply_sweep<T>(a, S, nops, gb, go, HAS_M4 ? &s_ybuf[0][0] : nullptr, DT, tid);
        if (store) {
            if (VEC) {
#pragma unroll
                for (int s = 0; s < 16; s += 2) {
                    const uint32_t g = gb | ((s & 2) ? go[1] : 0u) | ((s & 4) ? go[2] : 0u) | ((s & 8) ? go[3] : 0u);
                    stv(dst + g, a[s], a[s + 1]);
                }
            } else {
#pragma unroll
                for (int s = 0; s < 16; s++) {
                    const uint32_t g = gb | ((s & 1) ? go[0] : 0u) | ((s & 2) ? go[1] : 0u) | ((s & 4) ? go[2] : 0u) |
                                       ((s & 8) ? go[3] : 0u);
                    dst[g] = a[s];
                }
            }
        }
        if (do_red || do_chunks) {
            T w[16];
#pragma unroll
            for (int s = 0; s < 16; s++) w[s] = a[s].x * a[s].x + a[s].y * a[s].y;
            gb_last = gb;
            if (do_red) {
                rel_accumulate_dyn<T>(rel, w, red_case);
                if (!red_fast) {   // base bin changes with the iteration: fold the relative bins now
                    const uint32_t bb = ((gb & base_mask0) ? 1u : 0u) | ((gb & base_mask1) ? 2u : 0u);
#pragma unroll
                    for (int j = 0; j < 4; j++) {
#pragma unroll
                        for (int b = 0; b < 4; b++) p[b] += ((j | bb) == (uint32_t)b) ? rel[j] : (T)0;
                        rel[j] = 0;
                    }
                }
            }
            if (do_chunks) {
                switch (r_low) {
                    case 0: chunk_sums<0, T>(w, csum, gb, go, tid & 31u); break;
                    case 1: chunk_sums<1, T>(w, csum, gb, go, tid & 31u); break;
                    case 2: chunk_sums<2, T>(w, csum, gb, go, tid & 31u); break;
                    case 3: chunk_sums<3, T>(w, csum, gb, go, tid & 31u); break;
                    default: chunk_sums<4, T>(w, csum, gb, go, tid & 31u); break;
                }
            }
        }
    }
    if (do_red) {
        if (red_fast) {
            const uint32_t bb = ((gb_last & base_mask0) ? 1u : 0u) | ((gb_last & base_mask1) ? 2u : 0u);
#pragma unroll
            for (int j = 0; j < 4; j++) {
#pragma unroll
                for (int b = 0; b < 4; b++) p[b] += ((j | bb) == (uint32_t)b) ? rel[j] : (T)0;
            }
        }
        double d[4] = {(double)p[0], (double)p[1], (double)p[2], (double)p[3]};
#pragma unroll
        for (int b = 0; b < 4; b++)
            for (int o = 16; o > 0; o >>= 1) d[b] += __shfl_down_sync(0xffffffffu, d[b], o);
        if ((tid & 31u) == 0) {
#pragma unroll
            for (int b = 0; b < 4; b++) s_red[tid >> 5][b] = d[b];
        }
        __syncthreads();
        if (tid < 4) {
            double accd = 0;
            for (int w = 0; w < DT / 32; w++) accd += s_red[w][tid];
            args.partials[((uint64_t)N.dst_slot * args.part_stride + part) * 4u + tid] = accd;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Philox4x32-10 (same stream convention as kernels.cu / include/qcdsim.h)
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ uint4 draw(uint64_t seed, uint32_t circuit, uint64_t shot, uint32_t stream) {
    return philox4x32_10(make_uint4((uint32_t)shot, (uint32_t)(shot >> 32), stream, circuit),
                         make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}
__device__ __forceinline__ double u53(uint4 w) {
    const uint64_t x = ((uint64_t)w.y << 32) | (uint64_t)w.x;
    return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ uint32_t upper_search(const double* arr, uint32_t lo, uint32_t hi, double target) {
    uint32_t a = lo, b = hi;
    while (a < b) {
        const uint32_t mid = (a + b) >> 1;
        if (arr[mid] > target) b = mid;
        else a = mid + 1;
    }
    return a < hi ? a : hi - 1;
}

constexpr int SN_WARPS = 4;

// Measurement sampling for NODE_NOSTORE leaves.  Phase 1 (per thread): locate the trajectory's chunk from the scanned
// chunk sums the final sweep wrote.  Phase 2 (per warp, one trajectory at a time): stage the leaf's sweep, let lane g
// recompute register group g of that chunk from the PARENT state, scatter the |amplitude|^2 of the chunk's 32 elements
// into shared memory in index order, warp-shuffle prefix sum, ballot pick.
template <typename T>
__global__ void __launch_bounds__(32 * SN_WARPS) sample_nostore_kernel(
    DevProgram<T> prog, const NodeRec* leaves, const void* states, uint64_t slot_stride, const double* chunk_sums,
    uint64_t chunk_stride, const double* block_sums, uint64_t block_stride, uint32_t n_chunks, int n_bits,
    const uint32_t* owner, const uint32_t* shots, uint32_t n_traj, uint64_t seed, const double* flip, uint32_t* hist,
    uint32_t* outcomes, uint64_t shots_per_circuit, uint64_t shot_begin) {
    typedef typename Cx<T>::C C;
    __shared__ Staged<T> S_all[SN_WARPS];
    __shared__ double s_w[SN_WARPS][32];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    Staged<T>& S = S_all[wid];

    bool live = i < n_traj;
    NodeRec nd;
    memset(&nd, 0, sizeof(nd));
    uint64_t shot = 0;
    uint32_t c = 0;
    double target = 0.0;
    if (live) {
        nd = leaves[owner[i]];
        live = (nd.flags & NODE_LEAF) && (nd.flags & NODE_NOSTORE);
    }
    if (live) {
        shot = shots[i];
        const double* cs = chunk_sums + (uint64_t)nd.dst_slot * chunk_stride;
        const double* bs = block_sums + (uint64_t)nd.dst_slot * block_stride;
        const uint32_t n_blocks = (n_chunks + (1u << SCAN_BLOCK_LOG) - 1u) >> SCAN_BLOCK_LOG;
        target = u53(draw(seed, nd.circuit, shot, QCD_STREAM_MEASURE)) * bs[n_blocks - 1];
        const uint32_t b = upper_search(bs, 0, n_blocks, target);
        if (b > 0) target -= bs[b - 1];
        const uint32_t cbeg = b << SCAN_BLOCK_LOG;
        const uint32_t cend = min(cbeg + (1u << SCAN_BLOCK_LOG), n_chunks);
        c = upper_search(cs, cbeg, cend, target);
        if (c > cbeg) target -= cs[c - 1];
    }
    uint32_t todo = __ballot_sync(0xffffffffu, live);
    uint32_t e_sel = 0;
    while (todo) {
        const int srcl = __ffs(todo) - 1;
        todo &= todo - 1u;
        const uint32_t sweep_j = __shfl_sync(0xffffffffu, nd.sweep, srcl);
        const uint32_t branch_j = __shfl_sync(0xffffffffu, nd.branch, srcl);
        const double scale_j = __shfl_sync(0xffffffffu, nd.scale, srcl);
        const uint32_t pslot_j = __shfl_sync(0xffffffffu, nd.src_slot, srcl);
        const uint32_t c_j = __shfl_sync(0xffffffffu, c, srcl);
        const double t_j = __shfl_sync(0xffffffffu, target, srcl);
        __syncwarp();
        stage_sweep<T, 32, true>(S, prog, sweep_j, branch_j, scale_j, lane);
        uint32_t go[4];
#pragma unroll
        for (int k = 0; k < 4; k++) go[k] = 1u << S.D.rbit[k];
        const uint32_t regmask = go[0] | go[1] | go[2] | go[3];
        int r_low = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) r_low += (S.D.rbit[k] < CHUNK_LOG) ? 1 : 0;
        const uint32_t n_groups = 1u << (CHUNK_LOG - r_low);
        const uint32_t chunk_base = c_j << CHUNK_LOG;
        if (lane < n_groups) {
            // group base: chunk base with the register bits cleared, lane bits deposited on the non-register low bits
            uint32_t low = 0, lb = lane;
            for (int b = 0; b < CHUNK_LOG; b++)
                if (!((regmask >> b) & 1u)) {
                    low |= (lb & 1u) << b;
                    lb >>= 1;
                }
            const uint32_t gb = (chunk_base & ~regmask) | low;
            const C* src = reinterpret_cast<const C*>(states) + (uint64_t)pslot_j * slot_stride;
            C a[16];
            load_group<T, false>(a, S, src, gb, go, regmask, (T)scale_j, (S.D.flags & SW_PREP) != 0, S.scale_op < 0);
            if (S.slots) apply_slotted<T>(a, S, S.slots, S.sign_code, gb);
            else apply_sweep<T>(a, S, S.D.nops, gb, go, nullptr, 0, 0);
#pragma unroll
            for (int s = 0; s < 16; s++) {
                const uint32_t g = gb | ((s & 1) ? go[0] : 0u) | ((s & 2) ? go[1] : 0u) | ((s & 4) ? go[2] : 0u) |
                                   ((s & 8) ? go[3] : 0u);
                if ((g >> CHUNK_LOG) == c_j) s_w[wid][g & 31u] = (double)(a[s].x * a[s].x + a[s].y * a[s].y);
            }
        }
        __syncwarp();
        // inclusive prefix over the chunk in index order
        double acc = s_w[wid][lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double nb = __shfl_up_sync(0xffffffffu, acc, o);
            if (lane >= (uint32_t)o) acc += nb;
        }
        const uint32_t hit = __ballot_sync(0xffffffffu, acc > t_j);
        const uint32_t pick = hit ? (uint32_t)(__ffs(hit) - 1) : 31u;
        if ((int)lane == srcl) e_sel = pick;
    }
    if (!live) return;
    uint32_t x = (c << CHUNK_LOG) + e_sel;
    if (flip) {
        const double* fc = flip + (uint64_t)nd.circuit * n_bits * 2u;
        for (int blk = 0; blk * 4 < n_bits; blk++) {
            const uint4 w = draw(seed, nd.circuit, shot, QCD_STREAM_READOUT + blk);
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int q = blk * 4 + j;
                if (q < n_bits) {
                    const uint32_t bit = (x >> q) & 1u;
                    if ((double)ww[j] * (1.0 / 4294967296.0) < fc[q * 2 + bit]) x ^= 1u << q;
                }
            }
        }
    }
    if (hist) atomicAdd(hist + ((uint64_t)nd.circuit << n_bits) + x, 1u);
    if (outcomes) outcomes[(uint64_t)nd.circuit * shots_per_circuit + shot - shot_begin] = x;
}

}  // namespace

template <typename T, bool VEC>
static void launch_kind(const DevProgram<T>& prog, const SweepArgs& a, cudaStream_t st) {
    switch (a.direct_kind) {
        case 0: sweep_direct_kernel<T, VEC, 0><<<a.n_items, DT, 0, st>>>(prog, a); break;
        case 1: sweep_direct_kernel<T, VEC, 1><<<a.n_items, DT, 0, st>>>(prog, a); break;
        case 2: sweep_direct_kernel<T, VEC, 2><<<a.n_items, DT, 0, st>>>(prog, a); break;
        default: sweep_direct_kernel<T, VEC, 3><<<a.n_items, DT, 0, st>>>(prog, a); break;
    }
}

template <typename T>
void launch_sweep_direct(const DevProgram<T>& prog, const SweepArgs& a, cudaStream_t st) {
    // direct_vec: index bit 0 is a register bit of EVERY sweep in this launch (128-bit accesses); the scalar variant
    // is correct for any sweep.  direct_kind: see sweep_direct_kernel.
    if (a.direct_vec) launch_kind<T, true>(prog, a, st);
    else launch_kind<T, false>(prog, a, st);
}

template <typename T>
void launch_sample_nostore(const DevProgram<T>& prog, const NodeRec* leaves, const void* states, uint64_t slot_stride,
                           const double* chunk_sums, uint64_t chunk_stride, const double* block_sums,
                           uint64_t block_stride, uint32_t n_chunks, int n_bits, const uint32_t* owner,
                           const uint32_t* shots, uint32_t n_traj, uint64_t seed, const double* flip, uint32_t* hist,
                           uint32_t* outcomes, uint64_t shots_per_circuit, uint64_t shot_begin, cudaStream_t st) {
    if (!n_traj) return;
    const uint32_t bs = 32 * SN_WARPS;
    sample_nostore_kernel<T><<<(n_traj + bs - 1) / bs, bs, 0, st>>>(prog, leaves, states, slot_stride, chunk_sums,
                                                                    chunk_stride, block_sums, block_stride, n_chunks,
                                                                    n_bits, owner, shots, n_traj, seed, flip, hist,
                                                                    outcomes, shots_per_circuit, shot_begin);
}

template void launch_sweep_direct<float>(const DevProgram<float>&, const SweepArgs&, cudaStream_t);
template void launch_sweep_direct<double>(const DevProgram<double>&, const SweepArgs&, cudaStream_t);
template void launch_sample_nostore<float>(const DevProgram<float>&, const NodeRec*, const void*, uint64_t, const double*,
                                           uint64_t, const double*, uint64_t, uint32_t, int, const uint32_t*,
                                           const uint32_t*, uint32_t, uint64_t, const double*, uint32_t*, uint32_t*,
                                           uint64_t, uint64_t, cudaStream_t);
template void launch_sample_nostore<double>(const DevProgram<double>&, const NodeRec*, const void*, uint64_t,
                                            const double*, uint64_t, const double*, uint64_t, uint32_t, int,
                                            const uint32_t*, const uint32_t*, uint32_t, uint64_t, const double*,
                                            uint32_t*, uint32_t*, uint64_t, uint64_t, cudaStream_t);

}  // namespace qcd
