// direct_ops.cuh -- register-level op machinery shared by the register-only sweep kernel (sweep_direct.cu), the
// recompute-sampling kernel and the shared-memory tile kernel (sweep.cu): a pass's ops are staged once per CTA (or warp)
// in shared memory -- the matrices of the node's branch, classified real / real-diagonal / complex, sign masks split
// into a per-thread base test and a per-value pattern, ops pre-decoded into one code word -- and then applied to the 16
// amplitudes a thread holds for the pass's 4 register bits.
#pragma once
#include "kernels.cuh"

namespace qcd {
namespace dops {

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
    uint32_t has_m4;                 // the pass contains a 16x16 op (needs the caller's per-thread y buffer)
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
        uint32_t m4 = 0;
        for (int o = 0; o < nops; o++) m4 |= (S.code[o] & 0xffu) == C_M4 ? 1u : 0u;
        S.has_m4 = m4;
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

// Slotted fast path (see stage_sweep): `slots` != 0.  `diag_bits`: bit i set <=> the op of slot i is real diagonal
// (slot_diag_bits, formed once per CTA: the streaming loop then branches on registers only -- with the classification read
// from shared memory in the loop, every slot cost a dependent LDS -> compare -> branch, the top stall of the leaf kernels).
template <typename T>
__device__ __forceinline__ uint32_t slot_diag_bits(const Staged<T>& S, uint32_t slots) {
    uint32_t bits = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t o = (slots >> (4 * i)) & 0xFu;
        if (o != 0xFu && S.kind_flags[o] == 2u) bits |= 1u << i;
    }
    return bits;
}
template <typename T>
__device__ __forceinline__ void apply_slotted(typename Cx<T>::C (&a)[16], const Staged<T>& S, uint32_t slots,
                                              uint32_t sign_code, uint32_t gb, uint32_t diag_bits) {
    // the pre-sign slots hold the Kraus variants: the no-jump operators are diagonal (two multipliers per pair)
#define QCD_SLOT(i, r)                                                       \
    {                                                                        \
        const uint32_t o = (slots >> (4 * (i))) & 0xFu;                      \
        if (o != 0xFu) {                                                     \
            if ((diag_bits >> (i)) & 1u) m1d<r>(a, S.mats[o]);               \
            else m1r<r>(a, S.mats[o]);                                       \
        }                                                                    \
    }
    QCD_SLOT(0, 0) QCD_SLOT(1, 1) QCD_SLOT(2, 2) QCD_SLOT(3, 3)
    if (sign_code) apply_signs<T>(a, S, sign_code, gb);
    QCD_SLOT(4, 0) QCD_SLOT(5, 1) QCD_SLOT(6, 2) QCD_SLOT(7, 3)
#undef QCD_SLOT
}
template <typename T>
__device__ __forceinline__ void apply_slotted(typename Cx<T>::C (&a)[16], const Staged<T>& S, uint32_t slots,
                                              uint32_t sign_code, uint32_t gb) {
    apply_slotted<T>(a, S, slots, sign_code, gb, slot_diag_bits<T>(S, slots));
}


}  // namespace dops
}  // namespace qcd
