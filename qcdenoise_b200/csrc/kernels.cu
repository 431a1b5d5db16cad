// kernels.cu -- decide (trajectory tree), prefix-sum sampling, density-matrix helpers and parity reductions.
#include "kernels.cuh"

#include "../../include/qcdsim.h"

namespace qcd {
namespace {

// ------------------------------------------------------------------------------------------------ Philox4x32-10
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
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
    uint64_t x = ((uint64_t)w.y << 32) | (uint64_t)w.x;
    return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}

// ------------------------------------------------------------------------------------------------ decide
__global__ void node_probs_kernel(const NodeRec* nodes, uint32_t n_nodes, const double* partials,
                                  uint32_t part_stride, uint32_t tiles, double* P) {
    // one CTA (128 threads) per node: P[b] = sum over tiles of partials[slot][tile][b], fixed order
    __shared__ double red[128][4];
    const uint32_t nd = blockIdx.x;
    const double* p = partials + (uint64_t)nodes[nd].dst_slot * part_stride * 4u;
    double acc[4] = {0, 0, 0, 0};
    for (uint32_t t = threadIdx.x; t < tiles; t += 128) {
#pragma unroll
        for (int b = 0; b < 4; b++) acc[b] += p[(uint64_t)t * 4u + b];
    }
#pragma unroll
    for (int b = 0; b < 4; b++) red[threadIdx.x][b] = acc[b];
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
#pragma unroll
            for (int b = 0; b < 4; b++) red[threadIdx.x][b] += red[threadIdx.x + o][b];
        }
        __syncthreads();
    }
    if (threadIdx.x < 4) P[(uint64_t)nd * 4u + threadIdx.x] = red[0][threadIdx.x];
}

// Resolve the decisions of one site group for one trajectory (force < 0) or replay a given branch word (force >= 0)
// to obtain the rescale factor.  Returns the branch word; *scale_out (if given) = 1/sqrt(prod p_j) over the
// non-unit decisions.
__device__ __forceinline__ uint32_t resolve_group(const Group& G, const SubSite* subs, const double* gdata,
                                                  const double* Pin, uint64_t seed, uint32_t circuit, uint64_t shot,
                                                  long long force, double* scale_out) {
    double P[4] = {1.0, 0.0, 0.0, 0.0};
    if (G.q[0] >= 0) {
        P[0] = Pin[0];
        P[1] = Pin[1];
        P[2] = Pin[2];
        P[3] = Pin[3];
    }
    uint32_t word = 0;
    double norm2 = 1.0;
    for (uint32_t u = G.sub_begin; u < G.sub_end; u++) {
        const SubSite ss = subs[u];
        const double* d = gdata + ss.data_off;
        double tot = 0.0;
        for (int j = 0; j < ss.nbranch; j++) {
            const double* w = d + j * 20;
            tot += w[0] * P[0] + w[1] * P[1] + w[2] * P[2] + w[3] * P[3];
        }
        int sel;
        if (force >= 0) {
            sel = (int)((force >> ss.shift) & ((1u << ss.nbits) - 1u));
        } else {
            const double target = u53(draw(seed, circuit, shot, ss.stream)) * tot;
            double cum = 0.0;
            sel = ss.nbranch - 1;
            for (int j = 0; j < ss.nbranch; j++) {
                const double* w = d + j * 20;
                cum += w[0] * P[0] + w[1] * P[1] + w[2] * P[2] + w[3] * P[3];
                if (cum > target) {
                    sel = j;
                    break;
                }
            }
        }
        const double* w = d + sel * 20;
        const double pj = w[0] * P[0] + w[1] * P[1] + w[2] * P[2] + w[3] * P[3];
        const double psum = P[0] + P[1] + P[2] + P[3];
        if (!ss.unit) norm2 *= pj / psum;
        const double* Tm = w + 4;
        double Q[4];
#pragma unroll
        for (int x = 0; x < 4; x++) Q[x] = Tm[x * 4] * P[0] + Tm[x * 4 + 1] * P[1] + Tm[x * 4 + 2] * P[2] + Tm[x * 4 + 3] * P[3];
        const double inv = pj > 0.0 ? 1.0 / pj : 0.0;
        // keep sum(P) (the squared norm of the stored state) unchanged: P' = T P * (psum / p_j)
#pragma unroll
        for (int x = 0; x < 4; x++) P[x] = Q[x] * inv * psum;
        word |= (uint32_t)sel << ss.shift;
    }
    if (scale_out) {
        // the stored state has squared norm sum(Pin) (~1); renormalise to exactly 1 as well
        double n0 = G.q[0] >= 0 ? (Pin[0] + Pin[1] + Pin[2] + Pin[3]) : 1.0;
        *scale_out = 1.0 / sqrt(norm2 * n0);
    }
    return word;
}

template <typename T>
__global__ void traj_branch_kernel(DevProgram<T> prog, const NodeRec* parents, const double* P, const uint32_t* owner,
                                   const uint32_t* tids, uint32_t n_traj, uint64_t seed, TrajMap map,
                                   const CircuitInfo* circuits, long long root_base, uint64_t* keys, int gbits, int mbits,
                                   uint32_t* table) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t key = ~0ull;                        // finished circuit (leaf owner) or past the end: no child
    gbits += mbits;                              // table index = parent << (gbits + mbits) | member << gbits | word
    if (i < n_traj) {
        const uint32_t o = owner[i];
        const NodeRec nd = parents[o];
        if (!(nd.flags & NODE_LEAF)) {
            uint32_t circuit;
            uint64_t shot;
            traj_locate(map, tids[i], circuit, shot);
            const Sweep& S = prog.sweeps[nd.sweep + 1];
            uint32_t word = 0;
            if (S.group >= 0)
                word = resolve_group(prog.groups[S.group], prog.subs, prog.gdata, P + (uint64_t)o * 4u, seed,
                                     map.first_circuit + circuit, shot, -1, nullptr);
            // the node may hold trajectories of several circuits that share everything up to their final sweeps:
            // children that execute a final sweep are per circuit
            if (mbits > 0 && nd.sweep + 2u == circuits[nd.circuit].sweep_end)
                word |= (circuit - circuits[circuit].lead) << (gbits - mbits);
            const uint32_t ko = root_base >= 0 ? (uint32_t)((long long)nd.parent - root_base) : o;
            key = ((uint64_t)ko << 32) | word;
        }
        keys[i] = key;
    }
    if (table == nullptr) return;
    // first step of the grouping, fused here (see count_keys_kernel): table[owner << gbits | word]++, warp-aggregated
    const bool live = key != ~0ull;
    const uint32_t active = __ballot_sync(0xffffffffu, live);
    if (!live) return;
    const uint32_t idx = (uint32_t)(((uint64_t)(key >> 32) << gbits) | (uint32_t)key);
    const uint32_t first = __shfl_sync(active, idx, __ffs(active) - 1);
    const uint32_t peers = __all_sync(active, idx == first) ? active : __match_any_sync(active, idx);
    if ((threadIdx.x & 31u) == (uint32_t)(__ffs(peers) - 1)) atomicAdd(table + idx, (uint32_t)__popc(peers));
}

template <typename T>
__global__ void child_scale_kernel(DevProgram<T> prog, NodeRec* children, const uint32_t* n_children, const double* P) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n_children) return;
    NodeRec& c = children[i];
    const Sweep& S = prog.sweeps[c.sweep];
    double sc = 1.0;
    if (S.group >= 0)
        resolve_group(prog.groups[S.group], prog.subs, prog.gdata, P + (uint64_t)c.parent * 4u, 0, 0, 0,
                      (long long)c.branch, &sc);
    c.scale = sc;
}

__global__ void fill_owner_kernel(const NodeRec* nodes, uint32_t n_nodes, uint32_t* owner, uint32_t n_traj) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_traj) return;
    // nodes are sorted by traj_off (relative to nodes[0].traj_off): last node with traj_off <= i + base
    const uint32_t base = nodes[0].traj_off;
    uint32_t lo = 0, hi = n_nodes - 1;
    while (lo < hi) {
        uint32_t mid = (lo + hi + 1) >> 1;
        if (nodes[mid].traj_off - base <= i) lo = mid;
        else hi = mid - 1;
    }
    owner[i] = lo;
}

__global__ void iota_kernel(uint32_t* p, uint32_t n, uint32_t start) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = start + i;
}


// ------------------------------------------------------------------------------------------------ grouping
// Trajectories -> child nodes without a sort: a counting table indexed (owner << gbits | branch word).
//   count_keys  : table[idx]++                                   (atomics; integer -> order independent)
//   group_l1/l2/l3 : exclusive scan of (count, count != 0) over the table; every non-empty entry becomes one child
//                 NodeRec (children are therefore ordered by parent, then branch word -- deterministic)
//   scatter     : shots_out[offs[idx] + (--table[idx])] = shot   (leaves the table all-zero for the next use)
__global__ void count_keys_kernel(const uint64_t* keys, uint32_t n_traj, int gbits, uint32_t* table) {
    // warp-aggregated: trajectories of one node mostly draw the same branch word, so the lanes of a warp collapse
    // to a handful of distinct table entries -- one atomic per distinct entry instead of one per trajectory
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t key = i < n_traj ? keys[i] : ~0ull;
    const bool live = key != ~0ull;
    const uint32_t active = __ballot_sync(0xffffffffu, live);
    if (!live) return;
    const uint32_t idx = (uint32_t)(((uint64_t)(key >> 32) << gbits) | (uint32_t)key);
    // common case: the whole warp sits in one node and drew the same word -> no match instruction needed
    const uint32_t first = __shfl_sync(active, idx, __ffs(active) - 1);
    const uint32_t peers = __all_sync(active, idx == first) ? active : __match_any_sync(active, idx);
    if ((threadIdx.x & 31u) == (uint32_t)(__ffs(peers) - 1)) atomicAdd(table + idx, (uint32_t)__popc(peers));
}

// exclusive scan of up to 1024 (count, flag) pairs by 256 threads (4 consecutive each); returns block totals
__device__ __forceinline__ uint2 block_excl_scan_pairs(uint2 (&v)[4], uint2* warp_tot /*[8]*/) {
    uint2 t = make_uint2(v[0].x + v[1].x + v[2].x + v[3].x, v[0].y + v[1].y + v[2].y + v[3].y);
    const uint2 mine = t;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t nx = __shfl_up_sync(0xffffffffu, t.x, o), ny = __shfl_up_sync(0xffffffffu, t.y, o);
        if (lane >= o) {
            t.x += nx;
            t.y += ny;
        }
    }
    if (lane == 31) warp_tot[wid] = t;
    __syncthreads();
    uint2 off = make_uint2(t.x - mine.x, t.y - mine.y), tot = make_uint2(0, 0);
    for (int w = 0; w < 8; w++) {
        if (w < wid) {
            off.x += warp_tot[w].x;
            off.y += warp_tot[w].y;
        }
        tot.x += warp_tot[w].x;
        tot.y += warp_tot[w].y;
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const uint2 x = v[e];
        v[e] = off;
        off.x += x.x;
        off.y += x.y;
    }
    return tot;
}

__global__ void group_l1_kernel(const uint32_t* table, uint32_t M, uint2* bsum) {
    __shared__ uint2 warp_tot[8];
    const uint32_t base = blockIdx.x * 1024u + threadIdx.x * 4u;
    uint2 v[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const uint32_t c = (base + e < M) ? table[base + e] : 0u;
        v[e] = make_uint2(c, c ? 1u : 0u);
    }
    const uint2 tot = block_excl_scan_pairs(v, warp_tot);
    if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

__global__ void group_l2_kernel(uint2* bsum, uint32_t n_blocks, uint32_t* totals /*[2]: trajectories, children*/) {
    __shared__ uint2 warp_tot[8];
    uint2 carry = make_uint2(0, 0);
    for (uint32_t b0 = 0; b0 < n_blocks; b0 += 1024u) {
        const uint32_t base = b0 + threadIdx.x * 4u;
        uint2 v[4];
#pragma unroll
        for (int e = 0; e < 4; e++) v[e] = (base + e < n_blocks) ? bsum[base + e] : make_uint2(0, 0);
        const uint2 tot = block_excl_scan_pairs(v, warp_tot);
#pragma unroll
        for (int e = 0; e < 4; e++)
            if (base + e < n_blocks) bsum[base + e] = make_uint2(v[e].x + carry.x, v[e].y + carry.y);
        carry.x += tot.x;
        carry.y += tot.y;
    }
    if (threadIdx.x == 0) {
        totals[0] = carry.x;
        totals[1] = carry.y;
    }
}

__global__ void group_l3_kernel(const uint32_t* table, uint32_t M, int gbits, int mbits, const uint2* bsum,
                                const NodeRec* parents, const CircuitInfo* circuits, uint32_t traj_base, uint32_t* offs,
                                NodeRec* children) {
    __shared__ uint2 warp_tot[8];
    const uint32_t base = blockIdx.x * 1024u + threadIdx.x * 4u;
    uint2 v[4];
    uint32_t cnt[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
        cnt[e] = (base + e < M) ? table[base + e] : 0u;
        v[e] = make_uint2(cnt[e], cnt[e] ? 1u : 0u);
    }
    block_excl_scan_pairs(v, warp_tot);
    const uint2 boff = bsum[blockIdx.x];
#pragma unroll
    for (int e = 0; e < 4; e++) {
        if (base + e >= M) continue;
        const uint32_t excl = v[e].x + boff.x;
        offs[base + e] = excl;
        if (cnt[e]) {
            const uint32_t idx = base + e;
            const uint32_t par = idx >> (gbits + mbits);
            const NodeRec pn = parents[par];
            NodeRec c;
            c.src_slot = pn.dst_slot;
            c.dst_slot = 0xFFFFFFFFu;
            c.sweep = pn.sweep + 1u;
            c.branch = idx & ((1u << gbits) - 1u);
            c.scale = 1.0;
            c.circuit = pn.circuit;
            const CircuitInfo pci = circuits[pn.circuit];
            if (mbits > 0 && pn.sweep + 2u == pci.sweep_end) {
                // the final sweep belongs to the trajectories' own circuit (member of the parent's run)
                c.circuit = pci.lead + ((idx >> gbits) & ((1u << mbits) - 1u));
                c.sweep = circuits[c.circuit].sweep_end - 1u;
            }
            c.traj_off = traj_base + excl;
            c.traj_cnt = cnt[e];
            c.parent = par;
            c.flags = (c.sweep + 1u == circuits[c.circuit].sweep_end) ? NODE_LEAF : 0u;
            c.pad = 0;
            children[v[e].y + boff.y] = c;
        }
    }
}

__global__ void scatter_kernel(const uint64_t* keys, const uint32_t* shots_in, uint32_t n_traj, int gbits,
                               const uint32_t* offs, uint32_t* table, uint32_t* shots_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t key = i < n_traj ? keys[i] : ~0ull;
    const bool live = key != ~0ull;
    const uint32_t active = __ballot_sync(0xffffffffu, live);
    if (!live) return;
    const uint32_t idx = (uint32_t)(((uint64_t)(key >> 32) << gbits) | (uint32_t)key);
    const uint32_t first = __shfl_sync(active, idx, __ffs(active) - 1);
    const uint32_t peers = __all_sync(active, idx == first) ? active : __match_any_sync(active, idx);
    const uint32_t lane = threadIdx.x & 31u;
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicSub(table + idx, (uint32_t)__popc(peers));   // old count
    base = __shfl_sync(peers, base, leader);
    // the group takes slots [old - popc, old); this lane's rank inside its group picks one
    const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    shots_out[offs[idx] + base - (uint32_t)__popc(peers) + rank] = shots_in[i];
}

// virtual root parents: node.pad = first shot of the node's range
__global__ void root_shots_kernel(const NodeRec* nodes, const uint32_t* owner, uint32_t* shots, uint32_t n_traj) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_traj) return;
    const NodeRec nd = nodes[owner[i]];
    shots[i] = nd.pad + (i - (nd.traj_off - nodes[0].traj_off));
}

__global__ void fill_root_probs_kernel(double* P, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) P[i] = (i & 3u) ? 0.0 : 1.0;
}

// ------------------------------------------------------------------------------------------------ scans
// inclusive scan of up to 1024 doubles by 256 threads (4 consecutive each), fixed association order
__device__ __forceinline__ void block_scan_1024(double (&v)[4], double* warp_tot /*[8]*/, double carry_in) {
    v[1] += v[0];
    v[2] += v[1];
    v[3] += v[2];
    double t = v[3];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double n = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += n;
    }
    if (lane == 31) warp_tot[wid] = t;
    __syncthreads();
    double off = carry_in;
    for (int w = 0; w < wid; w++) off += warp_tot[w];
    off += t - v[3];
#pragma unroll
    for (int e = 0; e < 4; e++) v[e] += off;
    __syncthreads();
}

// level 1: each CTA scans one block of 1024 chunk sums in place, writes the block total.  Members of a sibling group
// (NODE_CROSSED) first form their chunk sums from the group's shared sums: see ComposeRec.
__global__ void scan_l1_kernel(const NodeRec* leaves, const uint32_t* slots, double* chunk_sums, uint64_t chunk_stride,
                               double* block_sums, uint64_t block_stride, uint32_t n_chunks, const ComposeRec* comp,
                               const double* c0_sums) {
    __shared__ double warp_tot[8];
    const uint32_t item = blockIdx.y;
    uint32_t slot;
    const ComposeRec* cr = nullptr;
    if (leaves) {
        const uint32_t fl = leaves[item].flags;
        if (!(fl & NODE_LEAF) || (fl & NODE_SHARED)) return;
        slot = leaves[item].dst_slot;
        if (comp && (fl & NODE_CROSSED)) cr = comp + (leaves[item].pad - 1u);
    } else {
        slot = slots ? slots[item] : item;
    }
    double* cs = chunk_sums + (uint64_t)slot * chunk_stride;
    const uint32_t b = blockIdx.x;
    const uint32_t base = (b << SCAN_BLOCK_LOG) + threadIdx.x * 4u;
    double v[4];
    if (cr) {
        const double* c0 = c0_sums + (uint64_t)cr->c0_slot * chunk_stride;
        const uint32_t kb = cr->kb;
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const uint32_t c = base + e;
            double r = 0.0;
            if (c < n_chunks) {
                if (kb == 0) {
                    r = c0[c];
                } else {
                    const double A = c0[c & ~kb], B = c0[c | kb], X = cs[c];
                    const double* q = cr->q + ((c & kb) ? 3 : 0);
                    r = q[0] * A + q[1] * B + q[2] * X;
                    r = r > 0.0 ? r : 0.0;      // destructive interference: the exact value is >= 0, rounding may not be
                }
            }
            v[e] = r;
        }
    } else {
#pragma unroll
        for (int e = 0; e < 4; e++) v[e] = (base + e < n_chunks) ? cs[base + e] : 0.0;
    }
    block_scan_1024(v, warp_tot, 0.0);
#pragma unroll
    for (int e = 0; e < 4; e++)
        if (base + e < n_chunks) cs[base + e] = v[e];
    if (threadIdx.x == 255) block_sums[(uint64_t)slot * block_stride + b] = v[3];
}

// level 2: one CTA per state scans its block totals in place
__global__ void scan_l2_kernel(const NodeRec* leaves, const uint32_t* slots, double* block_sums, uint64_t block_stride,
                               uint32_t n_blocks) {
    __shared__ double warp_tot[8];
    __shared__ double s_carry;
    const uint32_t item = blockIdx.x;
    uint32_t slot;
    if (leaves) {
        if (!(leaves[item].flags & NODE_LEAF) || (leaves[item].flags & NODE_SHARED)) return;
        slot = leaves[item].dst_slot;
    } else {
        slot = slots ? slots[item] : item;
    }
    double* bs = block_sums + (uint64_t)slot * block_stride;
    double carry = 0.0;
    for (uint32_t b0 = 0; b0 < n_blocks; b0 += 1024) {
        const uint32_t base = b0 + threadIdx.x * 4u;
        double v[4];
#pragma unroll
        for (int e = 0; e < 4; e++) v[e] = (base + e < n_blocks) ? bs[base + e] : 0.0;
        block_scan_1024(v, warp_tot, carry);
#pragma unroll
        for (int e = 0; e < 4; e++)
            if (base + e < n_blocks) bs[base + e] = v[e];
        if (threadIdx.x == 255) s_carry = v[3];
        __syncthreads();
        carry = s_carry;
        __syncthreads();
    }
}

// first index in [lo, hi) with arr[i] > target, or hi - 1
__device__ __forceinline__ uint32_t upper_search(const double* arr, uint32_t lo, uint32_t hi, double target) {
    uint32_t a = lo, b = hi;  // invariant: answer in [a, b]
    while (a < b) {
        uint32_t mid = (a + b) >> 1;
        if (arr[mid] > target) b = mid;
        else a = mid + 1;
    }
    return a < hi ? a : hi - 1;
}

// flip_c = this circuit's row of the flip table: [n][2] = P(1|0), P(0|1)
__device__ __forceinline__ uint32_t readout_flips(uint32_t x, int n, const double* flip_c, uint64_t seed,
                                                  uint32_t circuit, uint64_t shot) {
    for (int blk = 0; blk * 4 < n; blk++) {
        const uint4 w = draw(seed, circuit, shot, QCD_STREAM_READOUT + blk);
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int q = blk * 4 + j;
            if (q < n) {
                const uint32_t bit = (x >> q) & 1u;
                const double pf = flip_c[q * 2 + bit];
                if ((double)ww[j] * (1.0 / 4294967296.0) < pf) x ^= 1u << q;
            }
        }
    }
    return x;
}

// Warp-cooperative inverse CDF inside one probability chunk: every lane holds one element's weight (0 beyond len);
// inclusive warp-shuffle prefix sum in double, ballot of (prefix > target), first set lane wins (last element if none).
__device__ __forceinline__ uint32_t warp_pick(double w, double target, uint32_t len) {
    const uint32_t lane = threadIdx.x & 31u;
    double acc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double nb = __shfl_up_sync(0xffffffffu, acc, o);
        if (lane >= (uint32_t)o) acc += nb;
    }
    const uint32_t hit = __ballot_sync(0xffffffffu, acc > target && lane < len);
    return hit ? (uint32_t)(__ffs(hit) - 1) : len - 1u;
}

template <typename T>
__global__ void sample_kernel(const NodeRec* leaves, const void* states, uint64_t slot_stride, const double* chunk_sums,
                              uint64_t chunk_stride, const double* block_sums, uint64_t block_stride, uint32_t n_chunks,
                              int n_bits, const uint32_t* owner, const uint32_t* tids, uint32_t n_traj, uint64_t seed,
                              TrajMap map, const double* flip, uint32_t* hist, uint32_t* outcomes, uint64_t shot_begin) {
    typedef typename Cplx<T>::type C;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    // phase 1 (per thread): locate the trajectory's chunk by two binary searches over the scanned sums
    bool live = i < n_traj;
    NodeRec nd;
    nd.circuit = 0;
    nd.dst_slot = 0;
    uint64_t shot = 0;
    uint32_t c = 0;
    double target = 0.0;
    if (live) {
        nd = leaves[owner[i]];
        live = (nd.flags & NODE_LEAF) != 0 && !(nd.flags & NODE_NOSTORE);
    }
    const uint32_t pcirc = map.first_circuit + nd.circuit;      // circuit index in the Philox counters
    if (live) {
        shot = traj_shot(map, tids[i], nd.circuit);
        const double* cs = chunk_sums + (uint64_t)nd.dst_slot * chunk_stride;
        const double* bs = block_sums + (uint64_t)nd.dst_slot * block_stride;
        const uint32_t n_blocks = (n_chunks + (1u << SCAN_BLOCK_LOG) - 1u) >> SCAN_BLOCK_LOG;
        target = u53(draw(seed, pcirc, shot, QCD_STREAM_MEASURE)) * bs[n_blocks - 1];
        const uint32_t b = upper_search(bs, 0, n_blocks, target);
        if (b > 0) target -= bs[b - 1];
        const uint32_t cbeg = b << SCAN_BLOCK_LOG;
        const uint32_t cend = min(cbeg + (1u << SCAN_BLOCK_LOG), n_chunks);
        c = upper_search(cs, cbeg, cend, target);
        if (c > cbeg) target -= cs[c - 1];
    }
    // phase 2 (per warp): for each live trajectory of the warp, all lanes read its chunk (one coalesced request) and
    // pick the element by a warp-shuffle prefix sum
    const int chunk_log = n_bits < CHUNK_LOG ? n_bits : CHUNK_LOG;
    const uint32_t len = 1u << chunk_log;
    uint32_t todo = __ballot_sync(0xffffffffu, live);
    uint32_t e_sel = 0;
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1u;
        const uint32_t slot_j = __shfl_sync(0xffffffffu, nd.dst_slot, src);
        const uint32_t c_j = __shfl_sync(0xffffffffu, c, src);
        const double t_j = __shfl_sync(0xffffffffu, target, src);
        double w = 0.0;
        if (lane < len) {
            const C a = reinterpret_cast<const C*>(states)[(uint64_t)slot_j * slot_stride + ((uint64_t)c_j << chunk_log) + lane];
            w = (double)(a.x * a.x + a.y * a.y);
        }
        const uint32_t pick = warp_pick(w, t_j, len);
        if ((int)lane == src) e_sel = pick;
    }
    if (!live) return;
    uint32_t x = (c << chunk_log) + e_sel;
    if (flip) x = readout_flips(x, n_bits, flip + (uint64_t)nd.circuit * n_bits * 2u, seed, pcirc, shot);
    if (hist) atomicAdd(hist + ((uint64_t)nd.circuit << n_bits) + x, 1u);
    if (outcomes) outcomes[(uint64_t)nd.circuit * map.spc + shot - shot_begin] = x;
}

// ------------------------------------------------------------------------------------------------ DM helpers
template <typename T>
__global__ void extract_diag_kernel(const void* states, uint64_t slot_stride, const uint32_t* slots, int n, T* probs) {
    typedef typename Cplx<T>::type C;
    const uint32_t v = blockIdx.y;
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= (1u << n)) return;
    const C* st = reinterpret_cast<const C*>(states) + (uint64_t)slots[v] * slot_stride;
    probs[((uint64_t)v << n) + x] = st[(uint64_t)x * ((1ull << n) + 1ull)].x;
}

// Outcome distributions of several measurement SETTINGS of one density matrix.  Setting s of vector v applies one
// final single-qubit step (a 4x4 superoperator on (bit q, bit q + n), index = ket + 2 * bra) and measures; its
// distribution only needs the 2x2 block of rho in qubit q at equal other bits:
//   p_s[x] = Re sum_{a,b} S[(x_q, x_q), (a, b)] rho[x with bit q = a, x with bit q = b].        q = 0xFF: plain diagonal.
template <typename T>
__global__ void setting_probs_kernel(const void* states, uint64_t slot_stride, uint32_t first_slot, int n, int n_settings,
                                     const uint8_t* qubits, const double2* sops, T* probs) {
    typedef typename Cplx<T>::type C;
    const uint32_t vs = blockIdx.y;                      // (vector, setting) flattened
    const uint32_t v = vs / (uint32_t)n_settings;
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= (1u << n)) return;
    const C* st = reinterpret_cast<const C*>(states) + (uint64_t)(first_slot + v) * slot_stride;
    const uint32_t q = qubits[vs];
    const uint64_t N = 1ull << n;
    double p;
    if (q == 0xFFu) {
        p = (double)st[(uint64_t)x * (N + 1ull)].x;
    } else {
        const double2* S = sops + (uint64_t)vs * 16u;
        const uint32_t xq = (x >> q) & 1u, x0 = x & ~(1u << q);
        const double2* row = S + (xq + 2u * xq) * 4u;
        p = 0.0;
#pragma unroll
        for (uint32_t ab = 0; ab < 4u; ab++) {
            const uint32_t a = ab & 1u, b = ab >> 1;
            const C r = st[(uint64_t)(x0 | (a << q)) + N * (uint64_t)(x0 | (b << q))];
            p += row[ab].x * (double)r.x - row[ab].y * (double)r.y;     // Re(S * rho)
        }
    }
    probs[((uint64_t)vs << n) + x] = (T)p;
}

template <typename T>
__global__ void readout_probs_kernel(T* probs, int n, int q, const double* conf) {
    // pairs (x with bit q = 0, x | 1<<q): p0' = c00 p0 + c10 p1 ; p1' = c01 p0 + c11 p1   (row = true, col = measured)
    const uint32_t v = blockIdx.y;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1u << (n - 1))) return;
    const uint32_t x0 = ((i >> q) << (q + 1)) | (i & ((1u << q) - 1u));
    T* p = probs + ((uint64_t)v << n);
    const double* c = conf + ((uint64_t)v * n + q) * 4u;
    const double p0 = (double)p[x0], p1 = (double)p[x0 | (1u << q)];
    p[x0] = (T)(c[0] * p0 + c[2] * p1);
    p[x0 | (1u << q)] = (T)(c[1] * p0 + c[3] * p1);
}

template <typename T>
__global__ void prob_chunks_kernel(const T* probs, int n, double* chunk_sums, uint64_t chunk_stride) {
    const uint32_t v = blockIdx.y;
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    const int chunk_log = n < CHUNK_LOG ? n : CHUNK_LOG;
    const uint32_t n_chunks = 1u << (n - chunk_log);
    if (c >= n_chunks) return;
    const T* p = probs + ((uint64_t)v << n) + ((uint64_t)c << chunk_log);
    double acc = 0.0;
    for (uint32_t e = 0; e < (1u << chunk_log); e++) acc += (double)p[e];
    chunk_sums[(uint64_t)v * chunk_stride + c] = acc;
}

template <typename T>
__global__ void sample_probs_kernel(const T* probs, int n, const double* chunk_sums, uint64_t chunk_stride,
                                    const double* block_sums, uint64_t block_stride, uint32_t n_chunks,
                                    uint64_t shot_begin, uint64_t n_shots, uint64_t seed, uint32_t first_circuit,
                                    const double* flip, uint32_t* hist) {
    const uint32_t v = blockIdx.y;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    const bool live = i < n_shots;
    const uint64_t shot = shot_begin + i;
    const uint32_t circuit = first_circuit + v;
    const int chunk_log = n < CHUNK_LOG ? n : CHUNK_LOG;
    const uint32_t len = 1u << chunk_log;
    uint32_t c = 0;
    double target = 0.0;
    if (live) {
        const double* cs = chunk_sums + (uint64_t)v * chunk_stride;
        const double* bs = block_sums + (uint64_t)v * block_stride;
        const uint32_t n_blocks = (n_chunks + (1u << SCAN_BLOCK_LOG) - 1u) >> SCAN_BLOCK_LOG;
        target = u53(draw(seed, circuit, shot, QCD_STREAM_MEASURE)) * bs[n_blocks - 1];
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
        const int src = __ffs(todo) - 1;
        todo &= todo - 1u;
        const uint32_t c_j = __shfl_sync(0xffffffffu, c, src);
        const double t_j = __shfl_sync(0xffffffffu, target, src);
        const double w = lane < len ? (double)probs[((uint64_t)v << n) + ((uint64_t)c_j << chunk_log) + lane] : 0.0;
        const uint32_t pick = warp_pick(w, t_j, len);
        if ((int)lane == src) e_sel = pick;
    }
    if (!live) return;
    uint32_t x = (c << chunk_log) + e_sel;
    if (flip) x = readout_flips(x, n, flip + (uint64_t)v * n * 2u, seed, circuit, shot);
    atomicAdd(hist + ((uint64_t)v << n) + x, 1u);
}

// ------------------------------------------------------------------------------------------------ parity
// signed_sums[v][m] = sum_x hist[v][x] * (-1)^popc(x & mask[m]) ; integer arithmetic -> order independent
constexpr int PAR_GROUP = 8;
__global__ void parity_counts_kernel(const uint32_t* hist, int n, const uint32_t* masks, int n_masks, int per_vec,
                                     long long* signed_sums, long long* totals) {
    const uint32_t v = blockIdx.y;
    const uint32_t* h = hist + ((uint64_t)v << n);
    const uint32_t* mk = masks + (per_vec ? (uint64_t)v * n_masks : 0);
    const uint64_t N = 1ull << n;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (int m0 = 0; m0 < n_masks; m0 += PAR_GROUP) {
        uint32_t mm[PAR_GROUP];
        long long acc[PAR_GROUP];
#pragma unroll
        for (int j = 0; j < PAR_GROUP; j++) {
            mm[j] = (m0 + j < n_masks) ? mk[m0 + j] : 0u;
            acc[j] = 0;
        }
        long long tot = 0;
        for (uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < N; x += stride) {
            const long long c = (long long)h[x];
            if (c == 0) continue;
            tot += c;
#pragma unroll
            for (int j = 0; j < PAR_GROUP; j++) acc[j] += (__popc((uint32_t)x & mm[j]) & 1) ? -c : c;
        }
#pragma unroll
        for (int j = 0; j < PAR_GROUP; j++) {
            long long a = acc[j];
            for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
            if ((threadIdx.x & 31) == 0 && m0 + j < n_masks && a != 0)
                atomicAdd((unsigned long long*)(signed_sums + (uint64_t)v * n_masks + m0 + j), (unsigned long long)a);
        }
        if (m0 == 0) {
            for (int o = 16; o > 0; o >>= 1) tot += __shfl_down_sync(0xffffffffu, tot, o);
            if ((threadIdx.x & 31) == 0 && tot != 0)
                atomicAdd((unsigned long long*)(totals + v), (unsigned long long)tot);
        }
    }
}

// signed_sums[v][m] = sum over the shots of vector v of (-1)^popc(outcome & mask[m])   (sparse form of parity_counts)
__global__ void parity_outcomes_kernel(const uint32_t* outcomes, uint64_t shots, const uint32_t* masks, int n_masks,
                                       int per_vec, long long* signed_sums) {
    const uint32_t v = blockIdx.y;
    const uint32_t* o = outcomes + (uint64_t)v * shots;
    const uint32_t* mk = masks + (per_vec ? (uint64_t)v * n_masks : 0);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (int m0 = 0; m0 < n_masks; m0 += PAR_GROUP) {
        uint32_t mm[PAR_GROUP];
        int acc[PAR_GROUP];
#pragma unroll
        for (int j = 0; j < PAR_GROUP; j++) {
            mm[j] = (m0 + j < n_masks) ? mk[m0 + j] : 0u;
            acc[j] = 0;
        }
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < shots; i += stride) {
            const uint32_t x = o[i];
#pragma unroll
            for (int j = 0; j < PAR_GROUP; j++) acc[j] += 1 - 2 * (__popc(x & mm[j]) & 1);
        }
#pragma unroll
        for (int j = 0; j < PAR_GROUP; j++) {
            int a = acc[j];
            for (int off = 16; off > 0; off >>= 1) a += __shfl_down_sync(0xffffffffu, a, off);
            if ((threadIdx.x & 31) == 0 && m0 + j < n_masks && a != 0)
                atomicAdd((unsigned long long*)(signed_sums + (uint64_t)v * n_masks + m0 + j), (unsigned long long)(long long)a);
        }
    }
}

// expvals from exact probabilities: per-CTA partial sums (fixed order), then a second kernel adds the partials in order
template <typename T>
__global__ void parity_probs_kernel(const T* probs, int n, const uint32_t* masks, int n_masks, int per_vec,
                                    double* partial /*[v][m][gridDim.x]*/) {
    __shared__ double red[8];
    const uint32_t v = blockIdx.y;
    const T* p = probs + ((uint64_t)v << n);
    const uint32_t* mk = masks + (per_vec ? (uint64_t)v * n_masks : 0);
    const uint64_t N = 1ull << n;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (int m = 0; m < n_masks; m++) {
        const uint32_t mm = mk[m];
        double acc = 0.0;
        for (uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < N; x += stride) {
            const double c = (double)p[x];
            acc += (__popc((uint32_t)x & mm) & 1) ? -c : c;
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[w];
            partial[((uint64_t)v * n_masks + m) * gridDim.x + blockIdx.x] = t;
        }
        __syncthreads();
    }
}
__global__ void parity_final_kernel(const double* partial, uint32_t n_part, uint32_t n_out, double* expvals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    double t = 0.0;
    for (uint32_t j = 0; j < n_part; j++) t += partial[(uint64_t)i * n_part + j];
    expvals[i] = t;
}

inline uint32_t cdiv(uint64_t a, uint32_t b) { return (uint32_t)((a + b - 1) / b); }

}  // namespace

// ==================================================================================================== launchers
void launch_node_probs(const NodeRec* nodes, uint32_t n_nodes, const double* partials, uint32_t part_stride,
                       uint32_t n_parts, double* P, cudaStream_t st) {
    if (n_nodes) node_probs_kernel<<<n_nodes, 128, 0, st>>>(nodes, n_nodes, partials, part_stride, n_parts, P);
}
template <typename T>
void launch_traj_branch(const DevProgram<T>& prog, const NodeRec* parents, const double* P, const uint32_t* owner,
                        const uint32_t* tids, uint32_t n_traj, uint64_t seed, TrajMap map, const CircuitInfo* circuits,
                        long long root_base, uint64_t* keys, int gbits, int mbits, uint32_t* table, cudaStream_t st) {
    if (n_traj)
        traj_branch_kernel<T><<<cdiv(n_traj, 256), 256, 0, st>>>(prog, parents, P, owner, tids, n_traj, seed, map, circuits,
                                                                 root_base, keys, gbits, mbits, table);
}
template <typename T>
void launch_child_scale(const DevProgram<T>& prog, NodeRec* children, uint32_t max_children, const uint32_t* n_children,
                        const double* P, cudaStream_t st) {
    if (max_children) child_scale_kernel<T><<<cdiv(max_children, 128), 128, 0, st>>>(prog, children, n_children, P);
}
void launch_group_children(const uint64_t* keys, const uint32_t* shots_in, uint32_t n_traj, int gbits, int mbits,
                           uint32_t* table, uint32_t M, uint2* bsum, uint32_t* totals, const NodeRec* parents,
                           const CircuitInfo* circuits, uint32_t traj_base, uint32_t* offs, NodeRec* children,
                           uint32_t* shots_out, bool counted, cudaStream_t st) {
    if (!n_traj || !M) return;
    const uint32_t nb = cdiv(M, 1024);
    const int gtot = gbits + mbits;
    if (!counted) count_keys_kernel<<<cdiv(n_traj, 256), 256, 0, st>>>(keys, n_traj, gtot, table);
    group_l1_kernel<<<nb, 256, 0, st>>>(table, M, bsum);
    group_l2_kernel<<<1, 256, 0, st>>>(bsum, nb, totals);
    group_l3_kernel<<<nb, 256, 0, st>>>(table, M, gbits, mbits, bsum, parents, circuits, traj_base, offs, children);
    scatter_kernel<<<cdiv(n_traj, 256), 256, 0, st>>>(keys, shots_in, n_traj, gtot, offs, table, shots_out);
}
void launch_root_shots(const NodeRec* nodes, const uint32_t* owner, uint32_t* shots, uint32_t n_traj, cudaStream_t st) {
    if (n_traj) root_shots_kernel<<<cdiv(n_traj, 256), 256, 0, st>>>(nodes, owner, shots, n_traj);
}
void launch_fill_root_probs(double* P, uint32_t n_nodes, cudaStream_t st) {
    if (n_nodes) fill_root_probs_kernel<<<cdiv((uint64_t)n_nodes * 4u, 256), 256, 0, st>>>(P, n_nodes * 4u);
}
void launch_fill_owner(const NodeRec* nodes, uint32_t n_nodes, uint32_t* owner, uint32_t n_traj, cudaStream_t st) {
    if (n_traj) fill_owner_kernel<<<cdiv(n_traj, 256), 256, 0, st>>>(nodes, n_nodes, owner, n_traj);
}
void launch_iota(uint32_t* p, uint32_t n, uint32_t start, cudaStream_t st) {
    if (n) iota_kernel<<<cdiv(n, 256), 256, 0, st>>>(p, n, start);
}
void launch_scan_chunks(const NodeRec* leaves, uint32_t n_leaves, double* chunk_sums, uint64_t chunk_stride,
                        double* block_sums, uint64_t block_stride, uint32_t n_chunks, cudaStream_t st,
                        const ComposeRec* comp, const double* c0_sums) {
    if (!n_leaves) return;
    const uint32_t n_blocks = (n_chunks + (1u << SCAN_BLOCK_LOG) - 1u) >> SCAN_BLOCK_LOG;
    for (uint32_t y0 = 0; y0 < n_leaves; y0 += 32768) {
        uint32_t ny = n_leaves - y0 < 32768 ? n_leaves - y0 : 32768;
        scan_l1_kernel<<<dim3(n_blocks, ny), 256, 0, st>>>(leaves + y0, nullptr, chunk_sums, chunk_stride, block_sums,
                                                           block_stride, n_chunks, comp, c0_sums);
    }
    scan_l2_kernel<<<n_leaves, 256, 0, st>>>(leaves, nullptr, block_sums, block_stride, n_blocks);
}
void launch_scan_vectors(uint32_t n_vec, double* chunk_sums, uint64_t chunk_stride, double* block_sums,
                         uint64_t block_stride, uint32_t n_chunks, cudaStream_t st) {
    if (!n_vec) return;
    const uint32_t n_blocks = (n_chunks + (1u << SCAN_BLOCK_LOG) - 1u) >> SCAN_BLOCK_LOG;
    for (uint32_t y0 = 0; y0 < n_vec; y0 += 32768) {
        uint32_t ny = n_vec - y0 < 32768 ? n_vec - y0 : 32768;
        scan_l1_kernel<<<dim3(n_blocks, ny), 256, 0, st>>>(nullptr, nullptr, chunk_sums + (uint64_t)y0 * chunk_stride,
                                                           chunk_stride, block_sums + (uint64_t)y0 * block_stride,
                                                           block_stride, n_chunks, nullptr, nullptr);
    }
    scan_l2_kernel<<<n_vec, 256, 0, st>>>(nullptr, nullptr, block_sums, block_stride, n_blocks);
}
template <typename T>
void launch_sample(const NodeRec* leaves, uint32_t n_leaves, const void* states, uint64_t slot_stride,
                   const double* chunk_sums, uint64_t chunk_stride, const double* block_sums, uint64_t block_stride,
                   uint32_t n_chunks, int n_bits, const uint32_t* owner, const uint32_t* tids, uint32_t n_traj,
                   uint64_t seed, TrajMap map, const double* flip, uint32_t* hist, uint32_t* outcomes,
                   uint64_t shot_begin, cudaStream_t st) {
    (void)n_leaves;
    if (n_traj)
        sample_kernel<T><<<cdiv(n_traj, 128), 128, 0, st>>>(leaves, states, slot_stride, chunk_sums, chunk_stride,
                                                            block_sums, block_stride, n_chunks, n_bits, owner, tids,
                                                            n_traj, seed, map, flip, hist, outcomes, shot_begin);
}
template <typename T>
void launch_extract_diag(const void* states, uint64_t slot_stride, const uint32_t* slots, uint32_t n_vec, int n,
                         T* probs, cudaStream_t st) {
    for (uint32_t y0 = 0; y0 < n_vec; y0 += 32768) {
        uint32_t ny = n_vec - y0 < 32768 ? n_vec - y0 : 32768;
        extract_diag_kernel<T><<<dim3(cdiv(1ull << n, 256), ny), 256, 0, st>>>(states, slot_stride, slots + y0, n,
                                                                                probs + ((uint64_t)y0 << n));
    }
}
template <typename T>
void launch_setting_probs(const void* states, uint64_t slot_stride, uint32_t n_vec, int n, int n_settings,
                          const uint8_t* qubits, const double2* sops, T* probs, cudaStream_t st) {
    const uint32_t per = 32768u / (uint32_t)n_settings > 0 ? 32768u / (uint32_t)n_settings : 1u;   // vectors per launch
    for (uint32_t v0 = 0; v0 < n_vec; v0 += per) {
        const uint32_t nv = n_vec - v0 < per ? n_vec - v0 : per;
        setting_probs_kernel<T><<<dim3(cdiv(1ull << n, 256), nv * n_settings), 256, 0, st>>>(
            states, slot_stride, v0, n, n_settings, qubits + (uint64_t)v0 * n_settings,
            sops + (uint64_t)v0 * n_settings * 16u, probs + (((uint64_t)v0 * n_settings) << n));
    }
}
template <typename T>
void launch_readout_probs(T* probs, uint32_t n_vec, int n, const double* conf, cudaStream_t st) {
    for (int q = 0; q < n; q++)
        for (uint32_t y0 = 0; y0 < n_vec; y0 += 32768) {
            uint32_t ny = n_vec - y0 < 32768 ? n_vec - y0 : 32768;
            readout_probs_kernel<T><<<dim3(cdiv(1ull << (n - 1), 256), ny), 256, 0, st>>>(
                probs + ((uint64_t)y0 << n), n, q, conf + (uint64_t)y0 * n * 4u);
        }
}
template <typename T>
void launch_prob_chunks(const T* probs, uint32_t n_vec, int n, double* chunk_sums, uint64_t chunk_stride,
                        cudaStream_t st) {
    const int chunk_log = n < CHUNK_LOG ? n : CHUNK_LOG;
    for (uint32_t y0 = 0; y0 < n_vec; y0 += 32768) {
        uint32_t ny = n_vec - y0 < 32768 ? n_vec - y0 : 32768;
        prob_chunks_kernel<T><<<dim3(cdiv(1ull << (n - chunk_log), 128), ny), 128, 0, st>>>(
            probs + ((uint64_t)y0 << n), n, chunk_sums + (uint64_t)y0 * chunk_stride, chunk_stride);
    }
}
template <typename T>
void launch_sample_probs(const T* probs, uint32_t n_vec, int n, const double* chunk_sums, uint64_t chunk_stride,
                         const double* block_sums, uint64_t block_stride, uint32_t n_chunks, uint64_t shot_begin,
                         uint64_t n_shots, uint64_t seed, uint32_t first_circuit, const double* flip, uint32_t* hist,
                         cudaStream_t st) {
    if (!n_shots) return;
    for (uint32_t y0 = 0; y0 < n_vec; y0 += 32768) {
        uint32_t ny = n_vec - y0 < 32768 ? n_vec - y0 : 32768;
        sample_probs_kernel<T><<<dim3(cdiv(n_shots, 128), ny), 128, 0, st>>>(
            probs + ((uint64_t)y0 << n), n, chunk_sums + (uint64_t)y0 * chunk_stride, chunk_stride,
            block_sums + (uint64_t)y0 * block_stride, block_stride, n_chunks, shot_begin, n_shots, seed,
            first_circuit + y0, flip ? flip + (uint64_t)y0 * n * 2u : nullptr, hist + ((uint64_t)y0 << n));
    }
}
void launch_parity_counts(const uint32_t* hist, uint32_t n_vec, int n, const uint32_t* masks, int n_masks,
                          int per_vec, long long* signed_sums, long long* totals, cudaStream_t st) {
    uint32_t gx = cdiv(1ull << n, 256 * 8);
    if (gx > 1184) gx = 1184;
    for (uint32_t y0 = 0; y0 < n_vec; y0 += 32768) {
        uint32_t ny = n_vec - y0 < 32768 ? n_vec - y0 : 32768;
        parity_counts_kernel<<<dim3(gx, ny), 256, 0, st>>>(hist + ((uint64_t)y0 << n), n,
                                                          masks + (per_vec ? (uint64_t)y0 * n_masks : 0), n_masks,
                                                          per_vec, signed_sums + (uint64_t)y0 * n_masks, totals + y0);
    }
}
void launch_parity_outcomes(const uint32_t* outcomes, uint32_t n_vec, uint64_t shots, const uint32_t* masks, int n_masks,
                            int per_vec, long long* signed_sums, cudaStream_t st) {
    uint32_t gx = cdiv(shots, 256 * 4);
    if (gx > 1184) gx = 1184;
    if (gx < 1) gx = 1;
    for (uint32_t y0 = 0; y0 < n_vec; y0 += 32768) {
        uint32_t ny = n_vec - y0 < 32768 ? n_vec - y0 : 32768;
        parity_outcomes_kernel<<<dim3(gx, ny), 256, 0, st>>>(outcomes + (uint64_t)y0 * shots, shots,
                                                            masks + (per_vec ? (uint64_t)y0 * n_masks : 0), n_masks,
                                                            per_vec, signed_sums + (uint64_t)y0 * n_masks);
    }
}
template <typename T>
void launch_parity_probs(const T* probs, uint32_t n_vec, int n, const uint32_t* masks, int n_masks, int per_vec,
                         double* partial, uint32_t n_part, double* expvals, cudaStream_t st) {
    for (uint32_t y0 = 0; y0 < n_vec; y0 += 32768) {
        uint32_t ny = n_vec - y0 < 32768 ? n_vec - y0 : 32768;
        parity_probs_kernel<T><<<dim3(n_part, ny), 256, 0, st>>>(probs + ((uint64_t)y0 << n), n,
                                                                 masks + (per_vec ? (uint64_t)y0 * n_masks : 0), n_masks,
                                                                 per_vec, partial + (uint64_t)y0 * n_masks * n_part);
    }
    parity_final_kernel<<<cdiv((uint64_t)n_vec * n_masks, 128), 128, 0, st>>>(partial, n_part, n_vec * n_masks, expvals);
}

#define INST(T)                                                                                                        \
    template void launch_traj_branch<T>(const DevProgram<T>&, const NodeRec*, const double*, const uint32_t*,         \
                                        const uint32_t*, uint32_t, uint64_t, TrajMap, const CircuitInfo*, long long,  \
                                        uint64_t*, int, int, uint32_t*, cudaStream_t);                                \
    template void launch_child_scale<T>(const DevProgram<T>&, NodeRec*, uint32_t, const uint32_t*, const double*,    \
                                        cudaStream_t);                                                                \
    template void launch_sample<T>(const NodeRec*, uint32_t, const void*, uint64_t, const double*, uint64_t,          \
                                   const double*, uint64_t, uint32_t, int, const uint32_t*, const uint32_t*, uint32_t, \
                                   uint64_t, TrajMap, const double*, uint32_t*, uint32_t*, uint64_t, cudaStream_t);   \
    template void launch_extract_diag<T>(const void*, uint64_t, const uint32_t*, uint32_t, int, T*, cudaStream_t);    \
    template void launch_setting_probs<T>(const void*, uint64_t, uint32_t, int, int, const uint8_t*, const double2*,  \
                                          T*, cudaStream_t);                                                          \
    template void launch_readout_probs<T>(T*, uint32_t, int, const double*, cudaStream_t);                            \
    template void launch_prob_chunks<T>(const T*, uint32_t, int, double*, uint64_t, cudaStream_t);                    \
    template void launch_sample_probs<T>(const T*, uint32_t, int, const double*, uint64_t, const double*, uint64_t,   \
                                         uint32_t, uint64_t, uint64_t, uint64_t, uint32_t, const double*, uint32_t*,  \
                                         cudaStream_t);                                                               \
    template void launch_parity_probs<T>(const T*, uint32_t, int, const uint32_t*, int, int, double*, uint32_t,       \
                                         double*, cudaStream_t);
INST(float)
INST(double)

}  // namespace qcd
