// sweep.cu -- the state-sweep kernel (the HBM-bound hot kernel of the engine).
//
// Work item = (node, tile).  A tile is 2^k amplitudes: the 2^L lowest index bits (contiguous runs, L >= 5 -> >= 256 B
// in complex64) x n_high scattered high bits chosen by the scheduler so that every non-diagonal op of the sweep acts
// on bits inside the tile.  Algorithmic traffic: one read + one write of the state (write only for SW_PREP).
//
//   phase 1  global -> registers (128-bit loads, all in flight before first use) -> * node.scale -> swizzled smem
//   phase 2  per pass: each thread gathers 2^4 amplitudes (its register bits = the pass's active bits), applies all
//            ops of the pass in registers (1q/2q/4q matrices with per-branch variants, sign layers, 1q phases),
//            scatters back.  One __syncthreads per pass.
//   phase 3  smem -> global (128-bit stores) fused with (a) joint populations of the next site's qubits, reduced
//            warp-shuffle -> smem -> one double4 per (slot, tile), deterministic order; (b) on the final sweep, the
//            sums of |amp|^2 over 32-amplitude chunks (half-warp shuffle) for inverse-CDF sampling.
//
// Shared-memory layout: element i of the tile lives at i ^ fold(i >> W) (W = log2(128 B / sizeof(amp))), which spreads
// any 5 varying index bits of a warp over all banks for both the strided pass accesses and the linear load/store.
#include "direct_ops.cuh"

namespace qcd {
namespace {

template <typename T> struct Tr;
template <> struct Tr<float> {
    typedef float2 C;
    static constexpr int W = 4;
    static constexpr int MINB = 3;
};
template <> struct Tr<double> {
    typedef double2 C;
    static constexpr int W = 3;
    static constexpr int MINB = 2;
};

template <int W>
__device__ __forceinline__ uint32_t swz(uint32_t i) {
    uint32_t h = i >> W;
    uint32_t f = h ^ (h >> W) ^ (h >> (2 * W)) ^ (h >> (3 * W));
    return i ^ (f & ((1u << W) - 1u));
}

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

// vector (2-amplitude) global access
__device__ __forceinline__ void ld2(const float2* p, float2& a, float2& b) {
    float4 v = *reinterpret_cast<const float4*>(p);
    a = make_float2(v.x, v.y);
    b = make_float2(v.z, v.w);
}
__device__ __forceinline__ void ld2(const double2* p, double2& a, double2& b) {
    a = p[0];
    b = p[1];
}
__device__ __forceinline__ void st2(float2* p, float2 a, float2 b) {
    *reinterpret_cast<float4*>(p) = make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void st2(double2* p, double2 a, double2 b) {
    p[0] = a;
    p[1] = b;
}

template <int TB, typename C>
__device__ __forceinline__ void apply_m1(C (&a)[16], const C* __restrict__ m) {
    C m00 = m[0], m01 = m[1], m10 = m[2], m11 = m[3];
#pragma unroll
    for (int s = 0; s < 16; s++) {
        if (s & (1 << TB)) continue;
        C x0 = a[s], x1 = a[s | (1 << TB)];
        C y0 = cmul(m00, x0);
        cfma(y0, m01, x1);
        C y1 = cmul(m10, x0);
        cfma(y1, m11, x1);
        a[s] = y0;
        a[s | (1 << TB)] = y1;
    }
}

template <int T0, int T1, typename C>
__device__ __forceinline__ void apply_m2(C (&a)[16], const C* __restrict__ m) {
#pragma unroll
    for (int s = 0; s < 16; s++) {
        if (s & ((1 << T0) | (1 << T1))) continue;
        C x[4] = {a[s], a[s | (1 << T0)], a[s | (1 << T1)], a[s | (1 << T0) | (1 << T1)]};
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

template <typename C>
__device__ __forceinline__ void apply_m4(C (&a)[16], const C* __restrict__ m) {
    C y[16];
#pragma unroll
    for (int r = 0; r < 16; r++) {
        C acc = cmul(m[r * 16], a[0]);
#pragma unroll
        for (int c = 1; c < 16; c++) cfma(acc, m[r * 16 + c], a[c]);
        y[r] = acc;
    }
#pragma unroll
    for (int r = 0; r < 16; r++) a[r] = y[r];
}

template <typename T>
__global__ void __launch_bounds__(SWEEP_THREADS, Tr<T>::MINB) sweep_kernel(DevProgram<T> prog, SweepArgs args) {
    typedef typename Tr<T>::C C;
    constexpr int W = Tr<T>::W;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* tile = reinterpret_cast<C*>(smem_raw);
    __shared__ Sweep S;
    __shared__ NodeRec N;
    __shared__ uint32_t htab[2][32];
    __shared__ uint32_t s_tile_base;
    __shared__ C s_prep[64];
    __shared__ C s_prep_base;
    __shared__ double s_red[SWEEP_THREADS / 32][4];
    __shared__ dops::Staged<T> St;   // the current pass, staged (matrices of this node's branch, decoded ops, sign patterns)

    const uint32_t tid = threadIdx.x;
    const uint32_t item = blockIdx.x;
    const uint32_t node_idx = item >> args.tiles_log2;
    const uint32_t tau = item & ((1u << args.tiles_log2) - 1u);

    if (tid == 0) {
        N = args.nodes[node_idx];
        S = prog.sweeps[N.sweep];
    }
    __syncthreads();
    const int k = S.k, L = S.L, n_high = S.n_high, n_bits = S.n_bits;
    if (tid < 64) {
        int half = tid >> 5, v = tid & 31;
        uint32_t g = 0;
        for (int j = 0; j < 5; j++) {
            int hj = half * 5 + j;
            if (hj < n_high && ((v >> j) & 1)) g |= 1u << S.high[hj];
        }
        htab[half][v] = g;
    } else if (tid == 64) {
        // deposit the tile number into the global bits >= L that are not tile bits
        uint32_t g = 0, t = tau;
        int hi = 0;
        for (int b = L; b < n_bits; b++) {
            if (hi < n_high && S.high[hi] == b) {
                hi++;
                continue;
            }
            if (t & 1u) g |= 1u << b;
            t >>= 1;
        }
        s_tile_base = g;
    }
    if ((S.flags & SW_PREP) && tid >= 128 && tid < 128 + 2 * n_bits) s_prep[tid - 128] = prog.pool[S.prep_off + tid - 128];
    __syncthreads();
    const uint32_t tile_base = s_tile_base;
    const uint32_t lowmask = (1u << L) - 1u;
    auto gidx = [&](uint32_t idx) -> uint32_t {
        uint32_t h = idx >> L;
        return tile_base | (idx & lowmask) | htab[0][h & 31u] | htab[1][(h >> 5) & 31u];
    };

    C* dst = reinterpret_cast<C*>(args.states) + (uint64_t)N.dst_slot * args.slot_stride;
    const uint32_t nvec = 1u << (k - 1);

    // ------------------------------------------------------------------ phase 1: load (or prepare)
    if (S.flags & SW_PREP) {
        if (tid == 0) {
            // factor of the product state that is constant over the tile
            C p;
            p.x = 1;
            p.y = 0;
            int hi = 0;
            for (int b = L; b < n_bits; b++) {
                if (hi < n_high && S.high[hi] == b) {
                    hi++;
                    continue;
                }
                p = cmul(p, s_prep[2 * b + ((tile_base >> b) & 1u)]);
            }
            s_prep_base = p;
        }
        __syncthreads();
        const C pb = s_prep_base;
        for (uint32_t v = tid; v < nvec; v += SWEEP_THREADS) {
            uint32_t idx = 2u * v;
            uint32_t g = gidx(idx);
            C p = pb;
            for (int b = 1; b < L; b++) p = cmul(p, s_prep[2 * b + ((g >> b) & 1u)]);
            for (int j = 0; j < n_high; j++) {
                int b = S.high[j];
                p = cmul(p, s_prep[2 * b + ((g >> b) & 1u)]);
            }
            tile[swz<W>(idx)] = cmul(p, s_prep[0]);
            tile[swz<W>(idx + 1)] = cmul(p, s_prep[1]);
        }
    } else {
        const C* src = reinterpret_cast<const C*>(args.states) + (uint64_t)N.src_slot * args.slot_stride;
        const T sc = (T)N.scale;
        constexpr int U = 8;
        for (uint32_t v0 = 0; v0 < nvec; v0 += SWEEP_THREADS * U) {
            C a0[U], a1[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                uint32_t v = v0 + u * SWEEP_THREADS + tid;
                if (v < nvec) ld2(src + gidx(2u * v), a0[u], a1[u]);
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                uint32_t v = v0 + u * SWEEP_THREADS + tid;
                if (v < nvec) {
                    a0[u].x *= sc;
                    a0[u].y *= sc;
                    a1[u].x *= sc;
                    a1[u].y *= sc;
                    tile[swz<W>(2u * v)] = a0[u];
                    tile[swz<W>(2u * v + 1u)] = a1[u];
                }
            }
        }
    }
    __syncthreads();

    // ------------------------------------------------------------------ phase 2: passes
    for (uint32_t pi = S.pass_begin; pi < S.pass_end; pi++) {
        const Pass ps = prog.passes[pi];
        const int nreg = ps.nreg;
        uint32_t lo[REG_BITS], go[REG_BITS];
#pragma unroll
        for (int i = 0; i < REG_BITS; i++) {
            if (i < nreg) {
                int b = ps.rbits[i];
                lo[i] = 1u << b;
                go[i] = b < L ? (1u << b) : (1u << S.high[b - L]);
            } else {
                lo[i] = 0;
                go[i] = 0;
            }
        }
        const uint32_t ngroups = 1u << (k - nreg);
        const uint32_t nval = 1u << nreg;
        // Fast path: the pass has a staged descriptor (<= 12 ops, <= 48 sign masks, no 16x16 op): ops come from shared
        // memory, pre-decoded, with real / real-diagonal / slotted paths -- the same code the register-only kernel runs.
        // The rescale factor was applied while loading the tile, so the staging gets scale = 1.
        dops::stage_sweep<T, SWEEP_THREADS, false>(St, prog, prog.n_sweeps + pi, N.branch, 1.0, tid);
        if (St.D.ok && !St.has_m4 && nreg == REG_BITS) {
            const int nops_s = St.D.nops;
            const uint32_t slots = St.slots, sign_code = St.sign_code;
            for (uint32_t grp = tid; grp < ngroups; grp += SWEEP_THREADS) {
                uint32_t base = grp;
#pragma unroll
                for (int i = 0; i < REG_BITS; i++) {
                    const uint32_t b = ps.rbits[i];
                    base = ((base >> b) << (b + 1)) | (base & ((1u << b) - 1u));
                }
                C a[16];
#pragma unroll
                for (int s = 0; s < 16; s++) {
                    const uint32_t off = ((s & 1) ? lo[0] : 0u) | ((s & 2) ? lo[1] : 0u) | ((s & 4) ? lo[2] : 0u) | ((s & 8) ? lo[3] : 0u);
                    a[s] = tile[swz<W>(base | off)];
                }
                const uint32_t gb = gidx(base);
                if (slots) dops::apply_slotted<T>(a, St, slots, sign_code, gb);
                else dops::apply_sweep<T>(a, St, nops_s, gb, go, nullptr, 0, 0);
#pragma unroll
                for (int s = 0; s < 16; s++) {
                    const uint32_t off = ((s & 1) ? lo[0] : 0u) | ((s & 2) ? lo[1] : 0u) | ((s & 4) ? lo[2] : 0u) | ((s & 8) ? lo[3] : 0u);
                    tile[swz<W>(base | off)] = a[s];
                }
            }
            __syncthreads();
            continue;
        }
        for (uint32_t grp = tid; grp < ngroups; grp += SWEEP_THREADS) {
            uint32_t base = grp;
#pragma unroll
            for (int i = 0; i < REG_BITS; i++) {
                if (i < nreg) {
                    uint32_t b = ps.rbits[i];
                    base = ((base >> b) << (b + 1)) | (base & ((1u << b) - 1u));
                }
            }
            C a[16];
#pragma unroll
            for (int s = 0; s < 16; s++) {
                uint32_t off = ((s & 1) ? lo[0] : 0u) | ((s & 2) ? lo[1] : 0u) | ((s & 4) ? lo[2] : 0u) | ((s & 8) ? lo[3] : 0u);
                if ((uint32_t)s < nval) {
                    a[s] = tile[swz<W>(base | off)];
                } else {
                    a[s].x = 0;
                    a[s].y = 0;
                }
            }
            const uint32_t gb = gidx(base);
            for (uint32_t oi = ps.op_begin; oi < ps.op_end; oi++) {
                const KOp op = prog.kops[oi];
                const C* m = prog.pool + op.off + ((N.branch >> op.sel_shift) & op.sel_mask) * op.stride;
                switch (op.kind) {
                    case K_M1:
                        switch (op.t0) {
                            case 0: apply_m1<0>(a, m); break;
                            case 1: apply_m1<1>(a, m); break;
                            case 2: apply_m1<2>(a, m); break;
                            default: apply_m1<3>(a, m); break;
                        }
                        break;
                    case K_M2:
                        switch (op.t0 * 4 + op.t1) {
                            case 1: apply_m2<0, 1>(a, m); break;
                            case 2: apply_m2<0, 2>(a, m); break;
                            case 3: apply_m2<0, 3>(a, m); break;
                            case 6: apply_m2<1, 2>(a, m); break;
                            case 7: apply_m2<1, 3>(a, m); break;
                            default: apply_m2<2, 3>(a, m); break;
                        }
                        break;
                    case K_M4:
                        apply_m4(a, m);
                        break;
                    case K_SIGNS: {
                        uint32_t par = 0;
                        for (uint32_t t = 0; t < op.count; t++) {
                            const uint32_t mk = prog.masks[op.off + t];
#pragma unroll
                            for (int s = 0; s < 16; s++) {
                                uint32_t g = gb | ((s & 1) ? go[0] : 0u) | ((s & 2) ? go[1] : 0u) | ((s & 4) ? go[2] : 0u) |
                                             ((s & 8) ? go[3] : 0u);
                                par ^= ((g & mk) == mk ? 1u : 0u) << s;
                            }
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
                        const C d0 = m[0], d1 = m[1];
#pragma unroll
                        for (int s = 0; s < 16; s++) {
                            uint32_t g = gb | ((s & 1) ? go[0] : 0u) | ((s & 2) ? go[1] : 0u) | ((s & 4) ? go[2] : 0u) |
                                         ((s & 8) ? go[3] : 0u);
                            a[s] = cmul(a[s], ((g >> op.gbit) & 1u) ? d1 : d0);
                        }
                        break;
                    }
                }
            }
#pragma unroll
            for (int s = 0; s < 16; s++) {
                uint32_t off = ((s & 1) ? lo[0] : 0u) | ((s & 2) ? lo[1] : 0u) | ((s & 4) ? lo[2] : 0u) | ((s & 8) ? lo[3] : 0u);
                if ((uint32_t)s < nval) tile[swz<W>(base | off)] = a[s];
            }
        }
        __syncthreads();
    }

    // ------------------------------------------------------------------ phase 3: store + reductions
    const int rq0 = S.red_q[0], rq1 = S.red_q[1];
    const bool do_red = rq0 >= 0;
    const bool do_chunks = (S.flags & SW_FINAL) != 0 && args.chunk_sums != nullptr;
    const int chunk_log = n_bits < CHUNK_LOG ? n_bits : CHUNK_LOG;
    double* csum = args.chunk_sums + (uint64_t)N.dst_slot * args.chunk_stride;
    T p0 = 0, p1 = 0, p2 = 0, p3 = 0;
    const uint32_t nvec_pad = (nvec + 31u) & ~31u;  // whole warps iterate together (shuffles below)
    for (uint32_t v = tid; v < nvec_pad; v += SWEEP_THREADS) {
        const bool valid = v < nvec;
        C a0, a1;
        a0.x = a0.y = a1.x = a1.y = 0;
        uint32_t g = 0;
        if (valid) {
            a0 = tile[swz<W>(2u * v)];
            a1 = tile[swz<W>(2u * v + 1u)];
            g = gidx(2u * v);
            st2(dst + g, a0, a1);
        }
        const T w0 = a0.x * a0.x + a0.y * a0.y;
        const T w1 = a1.x * a1.x + a1.y * a1.y;
        if (do_red) {
            // bit 0 differs between the two amplitudes of the vector
            uint32_t g1 = g | 1u;
            uint32_t b0 = ((g >> rq0) & 1u) | (rq1 >= 0 ? (((g >> rq1) & 1u) << 1) : 0u);
            uint32_t b1 = ((g1 >> rq0) & 1u) | (rq1 >= 0 ? (((g1 >> rq1) & 1u) << 1) : 0u);
            p0 += (b0 == 0 ? w0 : (T)0) + (b1 == 0 ? w1 : (T)0);
            p1 += (b0 == 1 ? w0 : (T)0) + (b1 == 1 ? w1 : (T)0);
            p2 += (b0 == 2 ? w0 : (T)0) + (b1 == 2 ? w1 : (T)0);
            p3 += (b0 == 3 ? w0 : (T)0) + (b1 == 3 ? w1 : (T)0);
        }
        if (do_chunks) {
            // chunk = 2^chunk_log consecutive indices = 2^(chunk_log-1) consecutive lanes
            double cs = (double)w0 + (double)w1;
            const int lanes = 1 << (chunk_log - 1);
            for (int o = lanes >> 1; o > 0; o >>= 1) cs += __shfl_xor_sync(0xffffffffu, cs, o);
            if (valid && ((tid & (lanes - 1)) == 0)) csum[g >> chunk_log] = cs;
        }
    }
    if (do_red) {
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
            for (int w = 0; w < SWEEP_THREADS / 32; w++) acc += s_red[w][tid];
            args.partials[((uint64_t)N.dst_slot * args.part_stride + tau) * 4u + tid] = acc;
        }
    }
}

}  // namespace

template <typename T>
cudaError_t configure_sweep(int k) {
    size_t bytes = ((size_t)1 << k) * sizeof(typename Tr<T>::C);
    return cudaFuncSetAttribute(sweep_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

template <typename T>
void launch_sweep(const DevProgram<T>& prog, const SweepArgs& a, int k, cudaStream_t st) {
    size_t bytes = ((size_t)1 << k) * sizeof(typename Tr<T>::C);
    sweep_kernel<T><<<a.n_items, SWEEP_THREADS, bytes, st>>>(prog, a);
}

template cudaError_t configure_sweep<float>(int);
template cudaError_t configure_sweep<double>(int);
template void launch_sweep<float>(const DevProgram<float>&, const SweepArgs&, int, cudaStream_t);
template void launch_sweep<double>(const DevProgram<double>&, const SweepArgs&, int, cudaStream_t);

}  // namespace qcd
