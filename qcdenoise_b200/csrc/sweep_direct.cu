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
#include "direct_ops.cuh"

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
#ifndef CROSS_MINB
#define CROSS_MINB 4      // leaf_cross_kernel, complex64: 128 registers per thread (5: 96 registers and 88 bytes of spills)
#endif
using namespace dops;

// The 16 amplitudes of register group gb BEFORE the sweep's ops: product state (prep sweeps) or the parent state.
// Prep sweeps: the factor of every non-register bit of gb.  A thread's iterations differ in a few group-index bits only
// (vmask); the factor of all other bits (base_rest) is formed once per thread, not once per group -- the write-only first
// density-matrix pass was 81 % issue-bound on this 24-step complex product (ncu, round 1).
template <typename T, bool VEC>
__device__ __forceinline__ void load_group(typename Cx<T>::C (&a)[16], const Staged<T>& S,
                                           const typename Cx<T>::C* src, uint32_t gb, const uint32_t (&go)[4],
                                           uint32_t regmask, T sc, bool is_prep, bool need_scale,
                                           typename Cx<T>::C base_rest, uint32_t vmask) {
    typedef typename Cx<T>::C C;
    if (is_prep) {
        C base = base_rest;
        for (uint32_t m = vmask; m; m &= m - 1u) {
            const int b = __ffs(m) - 1;
            base = cmul(base, S.prep[2 * b + ((gb >> b) & 1u)]);
        }
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

template <typename T>
__device__ __forceinline__ void chunk_sums_dyn(int r_low, T (&w)[16], double* csum, uint32_t gb, const uint32_t (&go)[4],
                                               uint32_t lane) {
    switch (r_low) {
        case 0: chunk_sums<0, T>(w, csum, gb, go, lane); break;
        case 1: chunk_sums<1, T>(w, csum, gb, go, lane); break;
        case 2: chunk_sums<2, T>(w, csum, gb, go, lane); break;
        case 3: chunk_sums<3, T>(w, csum, gb, go, lane); break;
        default: chunk_sums<4, T>(w, csum, gb, go, lane); break;
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
    const uint32_t diag_bits = slots ? slot_diag_bits<T>(S, slots) : 0u;
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
    // prep sweeps: bits of gb that change between this thread's iterations, and the product-state factor of all others
    uint32_t prep_vmask = 0;
    C prep_rest;
    prep_rest.x = 1;
    prep_rest.y = 0;
    if (is_prep) {
        auto spread4 = [&](uint32_t g) {
            g = ((g & ~(go[0] - 1u)) << 1) | (g & (go[0] - 1u));
            g = ((g & ~(go[1] - 1u)) << 1) | (g & (go[1] - 1u));
            g = ((g & ~(go[2] - 1u)) << 1) | (g & (go[2] - 1u));
            g = ((g & ~(go[3] - 1u)) << 1) | (g & (go[3] - 1u));
            return g;
        };
        prep_vmask = spread4((groups_per_cta - 1u) & ~(uint32_t)(DT - 1));      // grp0 is a multiple of groups_per_cta
        const uint32_t gb0 = spread4(grp0 + tid);
        for (int b = 0; b < n_bits; b++)
            if (!(((regmask | prep_vmask) >> b) & 1u)) prep_rest = cmul(prep_rest, S.prep[2 * b + ((gb0 >> b) & 1u)]);
    }

    for (uint32_t it = tid; it < groups_per_cta; it += DT) {
        uint32_t gb = grp0 + it;   // insert a zero at every register bit (ascending)
        gb = ((gb & ~(go[0] - 1u)) << 1) | (gb & (go[0] - 1u));
        gb = ((gb & ~(go[1] - 1u)) << 1) | (gb & (go[1] - 1u));
        gb = ((gb & ~(go[2] - 1u)) << 1) | (gb & (go[2] - 1u));
        gb = ((gb & ~(go[3] - 1u)) << 1) | (gb & (go[3] - 1u));
        C a[16];
        load_group<T, VEC>(a, S, src, gb, go, regmask, sc, is_prep, need_scale, prep_rest, prep_vmask);
        if (slots) apply_slotted<T>(a, S, slots, sign_code, gb, diag_bits);
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
                chunk_sums_dyn<T>(r_low, w, csum, gb, go, tid & 31u);
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
    const uint32_t* owner, const uint32_t* tids, uint32_t n_traj, uint64_t seed, TrajMap map, const double* flip,
    uint32_t* hist, uint32_t* outcomes, uint64_t shot_begin) {
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
    const uint32_t pcirc = map.first_circuit + nd.circuit;      // circuit index in the Philox counters
    if (live) {
        shot = traj_shot(map, tids[i], nd.circuit);
        const uint32_t sum_slot = (nd.flags & NODE_SHARED) ? nd.pad : nd.dst_slot;
        const double* cs = chunk_sums + (uint64_t)sum_slot * chunk_stride;
        const double* bs = block_sums + (uint64_t)sum_slot * block_stride;
        const uint32_t n_blocks = (n_chunks + (1u << SCAN_BLOCK_LOG) - 1u) >> SCAN_BLOCK_LOG;
        target = u53(draw(seed, pcirc, shot, QCD_STREAM_MEASURE)) * bs[n_blocks - 1];
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
            C one;
            one.x = 1;
            one.y = 0;
            load_group<T, false>(a, S, src, gb, go, regmask, (T)scale_j, (S.D.flags & SW_PREP) != 0, S.scale_op < 0, one,
                                 ~regmask & (S.D.n_bits >= 32 ? 0xffffffffu : (1u << S.D.n_bits) - 1u));
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
        // no hit: the target lies within rounding of the chunk's total (chunk sums formed from a sibling group's shared
        // sums are not the bit-identical total of these 32 values) -> the last element that has any weight
        const uint32_t has_w = __ballot_sync(0xffffffffu, s_w[wid][lane] > 0.0);
        const uint32_t pick = hit ? (uint32_t)(__ffs(hit) - 1) : (has_w ? 31u - (uint32_t)__clz(has_w) : 31u);
        if ((int)lane == srcl) e_sel = pick;
    }
    if (!live) return;
    uint32_t x = (c << CHUNK_LOG) + e_sel;
    if (flip) {
        const double* fc = flip + (uint64_t)nd.circuit * n_bits * 2u;
        for (int blk = 0; blk * 4 < n_bits; blk++) {
            const uint4 w = draw(seed, pcirc, shot, QCD_STREAM_READOUT + blk);
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
    if (outcomes) outcomes[(uint64_t)nd.circuit * map.spc + shot - shot_begin] = x;
}


// ---------------------------------------------------------------------------------------------------------------
// Shared passes of sibling groups (see SW_CROSS, CrossRec).  The leaves (parent, branch word, member) of a run of circuits
// whose final sweeps differ only in a trailing real 2x2 on one qubit k share everything up to that op; with phi = the
// state after the shared ops, member k's probability chunk sums are
//     u00^2 C0[c] + u01^2 C0[c ^ k] + 2 u00 u01 X_k[c]   (bit k of c clear; u10, u11 when set),
//     C0[c] = sum_{x in c} |phi(x)|^2,   X_k[c] = sum_{x in c} Re(conj(phi(x)) phi(x ^ 2^k)),
// so ONE pass over the parent serves C0, four lane-exchanged X_k and the X of the shared ops' own high qubits, instead of
// one sweep per member.  Nothing is written but those sums (the members are NODE_NOSTORE leaves: the sampler recomputes
// the one chunk a trajectory lands in from the parent with the member's own descriptor).
//
// Index bits of a pass:  register bits (4; 16 amplitudes per thread; NH of them >= CHUNK_LOG), lane bits (5: the cross bits
// + the lowest free bits), iteration bits (chunk bits that are neither: a thread loops over them and accumulates), outer
// bits (everything else: one "unit" per value; a warp owns a strided set of units).  Chunk sums are complete after the
// iteration loop and one xor-shuffle per lane bit that lies inside the chunk.
template <typename T, int NH>
__global__ void __launch_bounds__(DT, sizeof(T) == 4 ? CROSS_MINB : 2)
    leaf_cross_kernel(DevProgram<T> prog, const CrossRec* recs, const void* states, uint64_t slot_stride, double* c0_sums,
                      double* chunk_sums, uint64_t chunk_stride, int n_bits, uint32_t ctas_log2) {
    typedef typename Cx<T>::C C;
    constexpr int NC = 1 << NH;          // chunks a thread's 16 values fall into
    constexpr int LOWR = 4 - NH;         // register bits below the chunk size
    __shared__ Staged<T> S;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const CrossRec R = recs[blockIdx.x >> ctas_log2];
    const uint32_t part = blockIdx.x & ((1u << ctas_log2) - 1u);
    stage_sweep<T, DT, false>(S, prog, R.sweep, R.branch, R.scale, tid);
    const DirectSweep& D = S.D;
    uint32_t go[4];
#pragma unroll
    for (int i = 0; i < 4; i++) go[i] = 1u << D.rbit[i];
    const uint32_t regmask = go[0] | go[1] | go[2] | go[3];
    uint32_t lanemask = 0, lane_part = 0, red_lanes = 0;      // red_lanes: lane bits that lie inside the chunk
#pragma unroll
    for (int j = 0; j < 5; j++) {
        const uint32_t b = 1u << R.lane_bit[j];
        lanemask |= b;
        lane_part |= ((lane >> j) & 1u) ? b : 0u;
        if (R.lane_bit[j] < CHUNK_LOG) red_lanes |= 1u << j;
    }
    const uint32_t chunkmask = (1u << CHUNK_LOG) - 1u;
    const uint32_t itermask = chunkmask & ~regmask & ~lanemask;
    const uint32_t allmask = n_bits >= 32 ? 0xffffffffu : (1u << n_bits) - 1u;
    const uint32_t outermask = allmask & ~(chunkmask | regmask | lanemask);
    const uint32_t n_it = 1u << __popc(itermask), n_units = 1u << __popc(outermask);
    auto deposit = [](uint32_t v, uint32_t mask) {       // bits of v onto the set bits of mask, ascending
        uint32_t r = 0;
        for (uint32_t m = mask; m; m &= m - 1u, v >>= 1) r |= (v & 1u) ? (m & (0u - m)) : 0u;
        return r;
    };
    const C* src = reinterpret_cast<const C*>(states) + (uint64_t)R.src_slot * slot_stride;
    const T sc = (T)R.scale;
    const bool need_scale = S.scale_op < 0;
    const int nk = R.n_klanes;
    const uint32_t slots = S.slots, sign_code = S.sign_code;
    if (slots == 0) __trap();      // not the slotted shape: build_cross (engine.cu) must not have offered this run
    const uint32_t diag_bits = slot_diag_bits<T>(S, slots);
    C one;
    one.x = 1;
    one.y = 0;
    // units of this warp: u = first, first + stride, ...; the deposited index advances by a masked add (the carry runs
    // through the gaps of the mask when they are filled with ones)
    const uint32_t warps_total = (DT / 32) << ctas_log2;
    const uint32_t ustep = deposit(warps_total, outermask);
    const bool want_c0 = R.c0_slot != NO_SLOT;
    const bool want_r0 = NH > 0 && R.x_reg[0] != NO_SLOT, want_r1 = NH > 1 && R.x_reg[1] != NO_SLOT;
    uint32_t udep = deposit(part * (DT / 32) + warp, outermask);
    for (uint32_t u = part * (DT / 32) + warp; u < n_units;
         u += warps_total, udep = ((udep | ~outermask) + ustep) & outermask) {
        const uint32_t ubase = udep | lane_part;
        T c0[NC], xl[CROSS_KLANES][NC], xr[NH > 0 ? NH : 1][NC];
#pragma unroll
        for (int h = 0; h < NC; h++) {
            c0[h] = 0;
#pragma unroll
            for (int j = 0; j < CROSS_KLANES; j++) xl[j][h] = 0;
#pragma unroll
            for (int r = 0; r < (NH > 0 ? NH : 1); r++) xr[r][h] = 0;
        }
        uint32_t idep = 0;
        for (uint32_t it = 0; it < n_it; it++, idep = ((idep | ~itermask) + 1u) & itermask) {
            const uint32_t gb = ubase | idep;
            C a[16];
            load_group<T, true>(a, S, src, gb, go, regmask, sc, false, need_scale, one, 0u);
            apply_slotted<T>(a, S, slots, sign_code, gb, diag_bits);      // (the host only groups runs of this shape)
            if (want_c0) {
#pragma unroll
                for (int s = 0; s < 16; s++) c0[s >> LOWR] += a[s].x * a[s].x + a[s].y * a[s].y;
            }
            // Re(conj(a) a') with a' held by the lane across bit j: the two lanes split the product by component -- the
            // lower one forms a.x a'.x, the upper one a.y a'.y -- so that ONE exchange per amplitude serves both (the
            // SM moves 32 lanes of shuffle per clock: two exchanges per amplitude made this kernel shuffle bound);
            // the halves meet after the iteration loop
#pragma unroll 1
            for (int j = 0; j < nk; j++) {      // one copy of the exchange code (the kernel is instruction-fetch sensitive)
                const bool up = (lane >> j) & 1u;
                T part[NC];
#pragma unroll
                for (int h = 0; h < NC; h++) part[h] = 0;
#pragma unroll
                for (int s = 0; s < 16; s++) {
                    const T own = up ? a[s].y : a[s].x, send = up ? a[s].x : a[s].y;
                    part[s >> LOWR] += own * __shfl_xor_sync(0xffffffffu, send, 1 << j);
                }
#pragma unroll
                for (int jj = 0; jj < CROSS_KLANES; jj++) {
                    if (jj == j) {
#pragma unroll
                        for (int h = 0; h < NC; h++) xl[jj][h] += part[h];
                    }
                }
            }
            // the shared ops' own high qubits: the partner is in this thread; both partner chunks carry the sum
            if (want_r0) {
                constexpr int bit = 1 << (NH > 0 ? LOWR : 0);
#pragma unroll
                for (int s = 0; s < 16; s++) {
                    if (s & bit) continue;
                    const T x = a[s].x * a[s | bit].x + a[s].y * a[s | bit].y;
                    xr[0][s >> LOWR] += x;
                    xr[0][(s | bit) >> LOWR] += x;
                }
            }
            if (want_r1) {
                constexpr int bit = 1 << (NH > 1 ? LOWR + 1 : 0);
#pragma unroll
                for (int s = 0; s < 16; s++) {
                    if (s & bit) continue;
                    const T x = a[s].x * a[s | bit].x + a[s].y * a[s | bit].y;
                    xr[NH > 1 ? 1 : 0][s >> LOWR] += x;
                    xr[NH > 1 ? 1 : 0][(s | bit) >> LOWR] += x;
                }
            }
        }
        // the two component halves of every lane-exchanged sum
#pragma unroll
        for (int j = 0; j < CROSS_KLANES; j++) {
            if (j < nk) {
#pragma unroll
                for (int h = 0; h < NC; h++) xl[j][h] += __shfl_xor_sync(0xffffffffu, xl[j][h], 1 << j);
            }
        }
        // complete the chunk sums over the lane bits inside the chunk
#pragma unroll
        for (int j = 0; j < 5; j++) {
            if ((red_lanes >> j) & 1u) {     // uniform
#pragma unroll
                for (int h = 0; h < NC; h++) {
                    c0[h] += __shfl_xor_sync(0xffffffffu, c0[h], 1 << j);
#pragma unroll
                    for (int jj = 0; jj < CROSS_KLANES; jj++)
                        if (jj < nk) xl[jj][h] += __shfl_xor_sync(0xffffffffu, xl[jj][h], 1 << j);
#pragma unroll
                    for (int r = 0; r < NH; r++) xr[r][h] += __shfl_xor_sync(0xffffffffu, xr[r][h], 1 << j);
                }
            }
        }
        if ((lane & red_lanes) == 0) {
#pragma unroll
            for (int h = 0; h < NC; h++) {
                uint32_t g = ubase;
#pragma unroll
                for (int r = 0; r < NH; r++) g |= ((h >> r) & 1) ? go[LOWR + r] : 0u;
                const uint64_t c = g >> CHUNK_LOG;
                if (R.c0_slot != NO_SLOT) c0_sums[(uint64_t)R.c0_slot * chunk_stride + c] = (double)c0[h];
#pragma unroll
                for (int j = 0; j < CROSS_KLANES; j++)
                    if (j < nk && R.x_lane[j] != NO_SLOT)
                        chunk_sums[(uint64_t)R.x_lane[j] * chunk_stride + c] = (double)xl[j][h];
#pragma unroll
                for (int r = 0; r < NH; r++)
                    if (R.x_reg[r] != NO_SLOT) chunk_sums[(uint64_t)R.x_reg[r] * chunk_stride + c] = (double)xr[r][h];
            }
        }
    }
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
                           const uint32_t* tids, uint32_t n_traj, uint64_t seed, TrajMap map, const double* flip,
                           uint32_t* hist, uint32_t* outcomes, uint64_t shot_begin, cudaStream_t st) {
    if (!n_traj) return;
    const uint32_t bs = 32 * SN_WARPS;
    sample_nostore_kernel<T><<<(n_traj + bs - 1) / bs, bs, 0, st>>>(prog, leaves, states, slot_stride, chunk_sums,
                                                                    chunk_stride, block_sums, block_stride, n_chunks,
                                                                    n_bits, owner, tids, n_traj, seed, map, flip, hist,
                                                                    outcomes, shot_begin);
}


template <typename T>
void launch_leaf_cross(const DevProgram<T>& prog, const CrossRec* recs, uint32_t n_recs, const void* states,
                       uint64_t slot_stride, double* c0_sums, double* chunk_sums, uint64_t chunk_stride, int n_bits,
                       int n_high_reg, uint32_t ctas_log2, cudaStream_t st) {
    if (!n_recs) return;
    const uint32_t grid = n_recs << ctas_log2;
    switch (n_high_reg) {
        case 0: leaf_cross_kernel<T, 0><<<grid, DT, 0, st>>>(prog, recs, states, slot_stride, c0_sums, chunk_sums, chunk_stride, n_bits, ctas_log2); break;
        case 1: leaf_cross_kernel<T, 1><<<grid, DT, 0, st>>>(prog, recs, states, slot_stride, c0_sums, chunk_sums, chunk_stride, n_bits, ctas_log2); break;
        default: leaf_cross_kernel<T, 2><<<grid, DT, 0, st>>>(prog, recs, states, slot_stride, c0_sums, chunk_sums, chunk_stride, n_bits, ctas_log2); break;
    }
}

template void launch_leaf_cross<float>(const DevProgram<float>&, const CrossRec*, uint32_t, const void*, uint64_t, double*,
                                       double*, uint64_t, int, int, uint32_t, cudaStream_t);
template void launch_leaf_cross<double>(const DevProgram<double>&, const CrossRec*, uint32_t, const void*, uint64_t, double*,
                                        double*, uint64_t, int, int, uint32_t, cudaStream_t);
template void launch_sweep_direct<float>(const DevProgram<float>&, const SweepArgs&, cudaStream_t);
template void launch_sweep_direct<double>(const DevProgram<double>&, const SweepArgs&, cudaStream_t);
template void launch_sample_nostore<float>(const DevProgram<float>&, const NodeRec*, const void*, uint64_t, const double*,
                                           uint64_t, const double*, uint64_t, uint32_t, int, const uint32_t*,
                                           const uint32_t*, uint32_t, uint64_t, TrajMap, const double*, uint32_t*,
                                           uint32_t*, uint64_t, cudaStream_t);
template void launch_sample_nostore<double>(const DevProgram<double>&, const NodeRec*, const void*, uint64_t,
                                            const double*, uint64_t, const double*, uint64_t, uint32_t, int,
                                            const uint32_t*, const uint32_t*, uint32_t, uint64_t, TrajMap, const double*,
                                            uint32_t*, uint32_t*, uint64_t, cudaStream_t);

}  // namespace qcd
