// sweep_direct.cu -- register-only state sweep for sweeps that consist of ONE pass over <= 4 register bits (the common
// case in trajectory mode: the Kraus variants of one site group + the few gates up to the next site).
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
#include "kernels.cuh"

namespace qcd {
namespace {

#ifndef DIRECT_DT
#define DIRECT_DT 128
#endif
constexpr int DT = DIRECT_DT;  // threads per CTA
#ifndef DIRECT_MINB
#define DIRECT_MINB 8
#endif

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

constexpr int MAX_SIGNS = 48;  // sign masks staged per sweep (more: the sweep is not direct-eligible)

template <typename T, bool VEC>
__global__ void __launch_bounds__(DT, sizeof(T) == 4 ? DIRECT_MINB : DIRECT_MINB / 2) sweep_direct_kernel(DevProgram<T> prog, SweepArgs args) {
    typedef typename Cx<T>::C C;
    __shared__ DirectSweep D;
    __shared__ C s_mats[MAX_DIRECT_OPS][16];   // the selected variant of every matrix op
    __shared__ uint32_t s_sign_base[MAX_SIGNS];
    __shared__ uint32_t s_sign_pat[MAX_SIGNS];
    __shared__ double s_red[DT / 32][4];
    __shared__ uint32_t s_real[MAX_DIRECT_OPS];   // matrix op o has purely real coefficients
    __shared__ int s_scale_op;                    // matrix op that carries the node's rescale factor (-1: none)
    __shared__ C s_prep[64];                      // prep sweeps: (alpha, beta) per index bit
    __shared__ C s_prep_reg[16];                  // ... and the 16 products over the register bits

    const uint32_t tid = threadIdx.x;
    const uint32_t node_idx = blockIdx.x >> args.tiles_log2;
    const uint32_t part = blockIdx.x & ((1u << args.tiles_log2) - 1u);
    const NodeRec N = args.nodes[node_idx];
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(prog.direct + (args.desc_from_pad ? N.pad : N.sweep));
        uint32_t* dst = reinterpret_cast<uint32_t*>(&D);
        if (tid < sizeof(DirectSweep) / 4) dst[tid] = src[tid];
    }
    __syncthreads();
    const int n_bits = D.n_bits, nops = D.nops;
    uint32_t go[4];  // global offset of register bit i
#pragma unroll
    for (int i = 0; i < 4; i++) go[i] = 1u << D.rbit[i];
    const uint32_t regmask = go[0] | go[1] | go[2] | go[3];
    // stage matrices of this node's branch + sign masks
    for (int o = 0; o < nops; o++) {
        const KOp op = D.ops[o];
        if (op.kind == K_SIGNS) {
            if (tid < op.count) {
                const uint32_t mk = prog.masks[op.off + tid];
                const uint32_t mr = mk & regmask;
                uint32_t pat = 0;
#pragma unroll
                for (int s = 0; s < 16; s++) {
                    const uint32_t g = ((s & 1) ? go[0] : 0u) | ((s & 2) ? go[1] : 0u) | ((s & 4) ? go[2] : 0u) |
                                       ((s & 8) ? go[3] : 0u);
                    pat |= ((g & mr) == mr ? 1u : 0u) << s;
                }
                // K_SIGNS ops of one sweep share the staging arrays: `t1` carries the op's slot offset (see host)
                s_sign_base[op.t1 + tid] = mk & ~regmask;
                s_sign_pat[op.t1 + tid] = pat;
            }
        } else {
            const C* m = prog.pool + op.off + ((N.branch >> op.sel_shift) & op.sel_mask) * op.stride;
            const uint32_t cnt = op.kind == K_M1 ? 4u : (op.kind == K_M2 ? 16u : 2u);
            if (tid < cnt) s_mats[o][tid] = m[tid];
        }
    }
    const bool is_prep = (D.flags & SW_PREP) != 0;
    if (is_prep && tid < 2u * (uint32_t)n_bits) s_prep[tid] = prog.pool[D.prep_off + tid];
    if (tid == 0) {
        int so = -1;
        for (int o = 0; o < nops && so < 0; o++)
            if (D.ops[o].kind == K_M1 || D.ops[o].kind == K_M2) so = o;
        s_scale_op = so;
    }
    __syncthreads();
    if ((int)tid < nops) {
        // classify the staged matrix; the first matrix op absorbs the node's 1/sqrt(p) rescale
        const int kind = D.ops[tid].kind;
        const int cnt = kind == K_M1 ? 4 : (kind == K_M2 ? 16 : (kind == K_D1 ? 2 : 0));
        const T f = ((int)tid == s_scale_op) ? (T)N.scale : (T)1;
        bool real = true;
        for (int e = 0; e < cnt; e++) {
            C v = s_mats[tid][e];
            v.x *= f;
            v.y *= f;
            s_mats[tid][e] = v;
            real = real && v.y == (T)0;
        }
        const bool diag = real && kind == K_M1 && s_mats[tid][1].x == (T)0 && s_mats[tid][2].x == (T)0;
        s_real[tid] = diag ? 2u : (real ? 1u : 0u);
    }
    if (is_prep && tid >= 32 && tid < 48) {
        const int sidx = (int)tid - 32;
        C v;
        v.x = 1;
        v.y = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) v = cmul(v, s_prep[2 * D.rbit[i] + ((sidx >> i) & 1)]);
        s_prep_reg[sidx] = v;
    }
    __syncthreads();

    const C* src = reinterpret_cast<const C*>(args.states) + (uint64_t)N.src_slot * args.slot_stride;
    C* dst = reinterpret_cast<C*>(args.states) + (uint64_t)N.dst_slot * args.slot_stride;
    const bool need_scale = s_scale_op < 0;
    const T sc = (T)N.scale;
    const uint32_t groups_per_cta = 1u << (n_bits - 4 - (int)args.tiles_log2);
    const uint32_t grp0 = part * groups_per_cta;
    const int rq0 = D.red_q[0], rq1 = D.red_q[1];
    const bool do_red = rq0 >= 0;
    const bool do_chunks = (D.flags & SW_FINAL) != 0 && args.chunk_sums != nullptr;
    int r_low = 0;  // register bits below the chunk size
#pragma unroll
    for (int i = 0; i < 4; i++) r_low += (D.rbit[i] < CHUNK_LOG) ? 1 : 0;
    double* csum = args.chunk_sums + (uint64_t)N.dst_slot * args.chunk_stride;
    T p0 = 0, p1 = 0, p2 = 0, p3 = 0;  // joint populations P[b(rq0) + 2 b(rq1)]
    // Fast reduction: when every reduced bit is a register bit or constant over this thread's iterations, the bin of
    // value s is fixed per thread -> accumulate per-s sums in the loop and bin them once at the end.
    bool red_fast = true;
    {
        const int gl = n_bits - 4 - (int)args.tiles_log2;     // log2(register groups per CTA)
        const int rq[2] = {rq0, rq1};
#pragma unroll
        for (int j = 0; j < 2; j++) {
            if (rq[j] < 0 || ((regmask >> rq[j]) & 1u)) continue;
            const int cp = rq[j] - __popc(regmask & ((1u << rq[j]) - 1u));   // bit position in the group index
            if (cp >= (DT == 256 ? 8 : 7) && cp < gl) red_fast = false;
        }
    }
    T acc[16];
#pragma unroll
    for (int s = 0; s < 16; s++) acc[s] = 0;
    uint32_t gb_last = 0;

    for (uint32_t it = tid; it < groups_per_cta; it += DT) {
        uint32_t gb = grp0 + it;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t b = D.rbit[i];
            gb = ((gb >> b) << (b + 1)) | (gb & ((1u << b) - 1u));
        }
        C a[16];
        if (is_prep) {
            C base;
            base.x = 1;
            base.y = 0;
            for (int b = 0; b < n_bits; b++)
                if (!((regmask >> b) & 1u)) base = cmul(base, s_prep[2 * b + ((gb >> b) & 1u)]);
#pragma unroll
            for (int s = 0; s < 16; s++) a[s] = cmul(base, s_prep_reg[s]);
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
        if (need_scale) {
#pragma unroll
            for (int s = 0; s < 16; s++) {
                a[s].x *= sc;
                a[s].y *= sc;
            }
        }
        for (int o = 0; o < nops; o++) {
            const KOp op = D.ops[o];
            switch (op.kind) {
                case K_M1:
                    if (s_real[o] == 2u) {
                        switch (op.t0) {
                            case 0: m1d<0>(a, s_mats[o]); break;
                            case 1: m1d<1>(a, s_mats[o]); break;
                            case 2: m1d<2>(a, s_mats[o]); break;
                            default: m1d<3>(a, s_mats[o]); break;
                        }
                    } else if (s_real[o]) {
                        switch (op.t0) {
                            case 0: m1r<0>(a, s_mats[o]); break;
                            case 1: m1r<1>(a, s_mats[o]); break;
                            case 2: m1r<2>(a, s_mats[o]); break;
                            default: m1r<3>(a, s_mats[o]); break;
                        }
                    } else {
                        switch (op.t0) {
                            case 0: m1<0>(a, s_mats[o]); break;
                            case 1: m1<1>(a, s_mats[o]); break;
                            case 2: m1<2>(a, s_mats[o]); break;
                            default: m1<3>(a, s_mats[o]); break;
                        }
                    }
                    break;
                case K_M2:
                    if (s_real[o]) {
                        switch (op.t0 * 4 + op.t1) {
                            case 1: m2r<0, 1>(a, s_mats[o]); break;
                            case 2: m2r<0, 2>(a, s_mats[o]); break;
                            case 3: m2r<0, 3>(a, s_mats[o]); break;
                            case 6: m2r<1, 2>(a, s_mats[o]); break;
                            case 7: m2r<1, 3>(a, s_mats[o]); break;
                            default: m2r<2, 3>(a, s_mats[o]); break;
                        }
                    } else {
                        switch (op.t0 * 4 + op.t1) {
                            case 1: m2<0, 1>(a, s_mats[o]); break;
                            case 2: m2<0, 2>(a, s_mats[o]); break;
                            case 3: m2<0, 3>(a, s_mats[o]); break;
                            case 6: m2<1, 2>(a, s_mats[o]); break;
                            case 7: m2<1, 3>(a, s_mats[o]); break;
                            default: m2<2, 3>(a, s_mats[o]); break;
                        }
                    }
                    break;
                case K_SIGNS: {
                    uint32_t par = 0;
                    for (uint32_t t = 0; t < op.count; t++) {
                        const uint32_t mb = s_sign_base[op.t1 + t];
                        par ^= ((gb & mb) == mb) ? s_sign_pat[op.t1 + t] : 0u;
                    }
#pragma unroll
                    for (int s = 0; s < 16; s++)
                        if ((par >> s) & 1u) {
                            a[s].x = -a[s].x;
                            a[s].y = -a[s].y;
                        }
                    break;
                }
                default: {  // K_D1
                    const C d0 = s_mats[o][0], d1 = s_mats[o][1];
#pragma unroll
                    for (int s = 0; s < 16; s++) {
                        const uint32_t g = gb | ((s & 1) ? go[0] : 0u) | ((s & 2) ? go[1] : 0u) | ((s & 4) ? go[2] : 0u) |
                                           ((s & 8) ? go[3] : 0u);
                        a[s] = cmul(a[s], ((g >> op.gbit) & 1u) ? d1 : d0);
                    }
                    break;
                }
            }
        }
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
        if (do_red || do_chunks) {
            T w[16];
#pragma unroll
            for (int s = 0; s < 16; s++) w[s] = a[s].x * a[s].x + a[s].y * a[s].y;
            gb_last = gb;
            if (do_red && red_fast) {
#pragma unroll
                for (int s = 0; s < 16; s++) acc[s] += w[s];
            } else if (do_red) {
#pragma unroll
                for (int s = 0; s < 16; s++) {
                    const uint32_t g = gb | ((s & 1) ? go[0] : 0u) | ((s & 2) ? go[1] : 0u) | ((s & 4) ? go[2] : 0u) |
                                       ((s & 8) ? go[3] : 0u);
                    const uint32_t bin = ((g >> rq0) & 1u) | (rq1 >= 0 ? (((g >> rq1) & 1u) << 1) : 0u);
                    p0 += bin == 0 ? w[s] : (T)0;
                    p1 += bin == 1 ? w[s] : (T)0;
                    p2 += bin == 2 ? w[s] : (T)0;
                    p3 += bin == 3 ? w[s] : (T)0;
                }
            }
            if (do_chunks) {
                // (1) fold the register bits that lie inside the 32-amplitude chunk (they are the lowest register
                //     indices: rbit is ascending); live values afterwards: s with the low r_low bits clear
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if (i < r_low) {
#pragma unroll
                        for (int s = 0; s < 16; s++)
                            if (!(s & (1 << i))) w[s] += w[s | (1 << i)];
                    }
                }
                // (2) the other chunk bits are the lowest (5 - r_low) lane bits
                const int lane_bits = CHUNK_LOG - r_low;
#pragma unroll
                for (int s = 0; s < 16; s++) {
                    if (s & ((1 << r_low) - 1)) continue;
                    T vt = w[s];   // <= 32 terms: accumulate in the state's own precision, store as double
                    for (int o = 1; o < (1 << lane_bits); o <<= 1) vt += __shfl_xor_sync(0xffffffffu, vt, o);
                    const double v = (double)vt;
                    if ((tid & ((1u << lane_bits) - 1u)) == 0) {
                        const uint32_t g = gb | ((s & 1) ? go[0] : 0u) | ((s & 2) ? go[1] : 0u) | ((s & 4) ? go[2] : 0u) |
                                           ((s & 8) ? go[3] : 0u);
                        csum[g >> CHUNK_LOG] = v;
                    }
                }
            }
        }
    }
    if (do_red) {
        if (red_fast) {
#pragma unroll
            for (int s = 0; s < 16; s++) {
                const uint32_t g = gb_last | ((s & 1) ? go[0] : 0u) | ((s & 2) ? go[1] : 0u) | ((s & 4) ? go[2] : 0u) |
                                   ((s & 8) ? go[3] : 0u);
                const uint32_t bin = ((g >> rq0) & 1u) | (rq1 >= 0 ? (((g >> rq1) & 1u) << 1) : 0u);
                p0 += bin == 0 ? acc[s] : (T)0;
                p1 += bin == 1 ? acc[s] : (T)0;
                p2 += bin == 2 ? acc[s] : (T)0;
                p3 += bin == 3 ? acc[s] : (T)0;
            }
        }
        double d[4] = {(double)p0, (double)p1, (double)p2, (double)p3};
#pragma unroll
        for (int b = 0; b < 4; b++)
            for (int o = 16; o > 0; o >>= 1) d[b] += __shfl_down_sync(0xffffffffu, d[b], o);
        if ((tid & 31u) == 0) {
#pragma unroll
            for (int b = 0; b < 4; b++) s_red[tid >> 5][b] = d[b];
        }
        __syncthreads();
        if (tid < 4) {
            double acc = 0;
            for (int w = 0; w < DT / 32; w++) acc += s_red[w][tid];
            args.partials[((uint64_t)N.dst_slot * args.part_stride + part) * 4u + tid] = acc;
        }
    }
}

}  // namespace

template <typename T>
void launch_sweep_direct(const DevProgram<T>& prog, const SweepArgs& a, cudaStream_t st) {
    // direct_vec: index bit 0 is a register bit of EVERY sweep in this launch (128-bit accesses); the scalar variant
    // is correct for any sweep
    if (a.direct_vec) sweep_direct_kernel<T, true><<<a.n_items, DT, 0, st>>>(prog, a);
    else sweep_direct_kernel<T, false><<<a.n_items, DT, 0, st>>>(prog, a);
}

template void launch_sweep_direct<float>(const DevProgram<float>&, const SweepArgs&, cudaStream_t);
template void launch_sweep_direct<double>(const DevProgram<double>&, const SweepArgs&, cudaStream_t);

}  // namespace qcd
