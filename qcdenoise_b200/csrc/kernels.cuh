// kernels.cuh -- hand-written sm_100a kernels of the sampling engine.
//
//  sweep_kernel      one full read+write pass over a batch of states: a CTA owns a 2^k-amplitude tile (contiguous
//                    2^L runs + scattered high bits) staged in XOR-swizzled shared memory; passes keep 2^4
//                    amplitudes per thread in registers and apply every fused op of the pass there; the store phase
//                    also reduces the joint populations the next stochastic site needs and, on the final sweep, the
//                    32-amplitude probability chunk sums used for measurement sampling.           (HBM bound)
//  node_probs / traj_branch / child_scale / fill_owner   the per-level "decide" step of the trajectory tree,
//                    Philox4x32-10 keyed (seed; shot, stream, circuit)
//  scan_* / sample_* hierarchical prefix sums + inverse-CDF measurement sampling into histograms
//  parity_*          bitmask-parity stabilizer reductions
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "program.h"

namespace qcd {

struct NodeRec {
    uint32_t src_slot;   // parent state (ignored for SW_PREP sweeps)
    uint32_t dst_slot;
    uint32_t sweep;      // absolute index into sweeps[] of the sweep that PRODUCES this node
    uint32_t branch;     // branch word resolved for that sweep's group
    double scale;        // 1/sqrt(prod p_j) applied while loading
    uint32_t circuit;
    uint32_t traj_off;   // trajectories of this node: [traj_off, traj_off + traj_cnt) of the level's array
    uint32_t traj_cnt;
    uint32_t parent;     // index of the parent in the previous level's node array (virtual roots: index of the root
                         // that collects this circuit's trajectories -- the first circuit of its run in the chunk)
    uint32_t flags;      // NODE_LEAF: `sweep` is the circuit's last sweep
    uint32_t pad;
};
// Trajectory ids: a tree works on one chunk [g0, g1) of the flattened [circuit][shot] index space; a trajectory is named
// by its position in the chunk (32 bits), from which its circuit and shot follow.  (A node shared by several circuits
// holds trajectories of all of them, so the circuit cannot be read off the node.)
struct TrajMap {
    uint64_t spc;            // shots per circuit
    uint64_t off0;           // shot index of trajectory id 0 inside its circuit
    double inv_spc;
    uint32_t c_first;        // circuit (index among the uploaded ones) of trajectory id 0
    uint32_t first_circuit;  // added to the circuit index in every Philox counter (a rank's slice of a larger batch)
};
#ifdef __CUDACC__
__device__ __forceinline__ void traj_locate(const TrajMap& m, uint32_t tid, uint32_t& circuit, uint64_t& shot) {
    const uint64_t local = (uint64_t)tid + m.off0;
    uint64_t q = (uint64_t)((double)local * m.inv_spc);      // exact up to +-1 for local < 2^52
    long long r = (long long)(local - q * m.spc);
    if (r < 0) {
        q--;
        r += (long long)m.spc;
    } else if (r >= (long long)m.spc) {
        q++;
        r -= (long long)m.spc;
    }
    circuit = m.c_first + (uint32_t)q;
    shot = (uint64_t)r;
}
__device__ __forceinline__ uint64_t traj_shot(const TrajMap& m, uint32_t tid, uint32_t circuit) {
    return (uint64_t)tid + m.off0 - (uint64_t)(circuit - m.c_first) * m.spc;
}
#endif
constexpr uint32_t NODE_LEAF = 1;
constexpr uint32_t NODE_NOSTORE = 2;   // leaf whose state is not written: sampled by recomputing chunks from the parent
constexpr uint32_t NODE_CROSSED = 4;   // NOSTORE leaf whose chunk sums come from its sibling group's shared passes (pad =
                                       // 1 + index of its ComposeRec); it takes no sweep of its own
constexpr uint32_t NODE_SHARED = 8;    // NODE_CROSSED leaf whose chunk sums ARE its group's C0 (no trailing op, or one inside the
                                       // chunks) and whose group has an earlier such member: it reads the scanned sums of that
                                       // member's slot (pad = that slot) instead of scanning a copy
constexpr uint32_t NO_SLOT = 0xFFFFFFFFu;

// Chunk sums of one member of a sibling group from the group's shared sums (see SW_CROSS): with A = C0[c & ~kb],
// B = C0[c | kb], X = X_k[c]:  sum(c) = q[0] A + q[1] B + q[2] X  (chunk bit kb clear),  q[3] A + q[4] B + q[5] X (set);
// kb == 0: sum(c) = C0[c] (no trailing op, or one that acts inside the 32-amplitude chunks).
// One shared pass over the parent state of a sibling group (leaf_cross_kernel).  A thread holds the 16 amplitudes of the
// SW_CROSS descriptor's register bits (the shared ops' qubits + the lowest other bits); the 5 lane bits of a warp are up to
// 5 cross bits k (their X_k come from lane exchanges) + the lowest free bits.
struct CrossRec {
    uint32_t src_slot;       // parent state
    uint32_t c0_slot;        // slot of c0_sums that receives C0 (NO_SLOT: not emitted by this pass)
    uint32_t sweep;          // SW_CROSS descriptor
    uint32_t branch;
    double scale;
    uint8_t lane_bit[5];     // global index bit carried by lane bit j
    uint8_t n_klanes;        // lane bits [0, n_klanes) are cross bits
    uint8_t pad[2];
    uint32_t x_lane[5];      // chunk_sums slot that receives X of lane_bit[j]
    uint32_t x_reg[2];       // ... of the descriptor's j-th register bit >= CHUNK_LOG (NO_SLOT: not wanted)
    uint32_t pad2;
};
constexpr int CROSS_KLANES = 5;

struct ComposeRec {
    uint32_t c0_slot;
    uint32_t kb;
    double q[6];
};

template <typename T> struct Cplx;
template <> struct Cplx<float> { typedef float2 type; };
template <> struct Cplx<double> { typedef double2 type; };

template <typename T>
struct DevProgram {
    const Sweep* sweeps;
    const Pass* passes;
    const KOp* kops;
    const typename Cplx<T>::type* pool;
    const uint32_t* masks;
    const Group* groups;
    const SubSite* subs;
    const double* gdata;
    const DirectSweep* direct;   // [n_sweeps] per single-pass sweep, then [n_passes] per pass
    uint32_t n_sweeps;
};

constexpr int SWEEP_THREADS = 256;
constexpr int CHUNK_LOG = 5;        // probability chunk = 32 amplitudes
constexpr int SCAN_BLOCK_LOG = 10;  // 1024 chunks per scan block

struct SweepArgs {
    const NodeRec* nodes;
    void* states;            // base of the slot arena
    uint64_t slot_stride;    // amplitudes per slot
    double* partials;        // [slot][tiles][4]
    double* chunk_sums;      // [slot][2^(n_bits - chunk_log)]
    uint64_t chunk_stride;   // doubles per slot in chunk_sums
    uint32_t tiles_log2;     // log2(tiles per state)      (direct kernel: log2(CTAs per state))
    uint32_t n_items;        // nodes * tiles
    uint32_t part_stride;    // partials: [slot][part_stride][4]; a launch writes tiles (or CTAs-per-state) of them
    uint32_t direct_vec;     // direct kernel: every sweep of the launch has index bit 0 as a register bit
    uint32_t desc_from_pad;  // direct kernel: descriptor index = node.pad (per-pass launches) instead of node.sweep
    uint32_t no_slots;       // direct kernel: disable the slotted fast path (tuning / A-B only)
    uint32_t direct_kind;    // direct kernel variant: 0 plain (+prep), 1 inner (reduction), 2 final (chunk sums), 3 generic
};

template <typename T>
void launch_sweep(const DevProgram<T>& prog, const SweepArgs& a, int k, cudaStream_t st);
template <typename T>
cudaError_t configure_sweep(int k);
// every node's sweep must have DirectSweep.ok; a.tiles_log2 = log2(CTAs per state), each CTA owns
// 2^(n_bits - 4 - tiles_log2) >= 256 register groups of 16 amplitudes
template <typename T>
void launch_sweep_direct(const DevProgram<T>& prog, const SweepArgs& a, cudaStream_t st);

// c0_sums / chunk_sums: [slot][chunk_stride]; ctas_log2: CTAs per pass
template <typename T>
void launch_leaf_cross(const DevProgram<T>& prog, const CrossRec* recs, uint32_t n_recs, const void* states,
                       uint64_t slot_stride, double* c0_sums, double* chunk_sums, uint64_t chunk_stride, int n_bits,
                       int n_high_reg, uint32_t ctas_log2, cudaStream_t st);

template <typename T>
void launch_sample_nostore(const DevProgram<T>& prog, const NodeRec* leaves, const void* states, uint64_t slot_stride,
                           const double* chunk_sums, uint64_t chunk_stride, const double* block_sums,
                           uint64_t block_stride, uint32_t n_chunks, int n_bits, const uint32_t* owner,
                           const uint32_t* tids, uint32_t n_traj, uint64_t seed, TrajMap map, const double* flip,
                           uint32_t* hist, uint32_t* outcomes, uint64_t shot_begin, cudaStream_t st);

// ---- decide
void launch_node_probs(const NodeRec* nodes, uint32_t n_nodes, const double* partials, uint32_t part_stride,
                       uint32_t n_parts, double* P, cudaStream_t st);
// gbits = branch-word bits, mbits = member bits (children that execute a circuit's FINAL sweep are keyed by the
// trajectory's own circuit as well: (parent, member, word)); root_base >= 0: the parents are virtual roots, a
// trajectory is keyed to parents[owner].parent - root_base instead of its owner.
template <typename T>
void launch_traj_branch(const DevProgram<T>& prog, const NodeRec* parents, const double* P, const uint32_t* owner,
                        const uint32_t* tids, uint32_t n_traj, uint64_t seed, TrajMap map, const CircuitInfo* circuits,
                        long long root_base, uint64_t* keys, int gbits, int mbits, uint32_t* table,
                        cudaStream_t st);   // table != NULL: also counts the keys into the grouping table
template <typename T>
void launch_child_scale(const DevProgram<T>& prog, NodeRec* children, uint32_t max_children, const uint32_t* n_children,
                        const double* P, cudaStream_t st);
// trajectories (keys = parent << 32 | member << gbits | branch word, ~0 = finished) -> child NodeRecs (dst_slot
// unassigned) ordered by (parent, member, word), totals[0] = trajectories placed, totals[1] = children; trajectory ids
// permuted so that each child's trajectories are contiguous.  table [M = n_parents << (gbits + mbits)] must be all-zero
// on entry and is all-zero on exit.
void launch_group_children(const uint64_t* keys, const uint32_t* shots_in, uint32_t n_traj, int gbits, int mbits,
                           uint32_t* table, uint32_t M, uint2* bsum, uint32_t* totals, const NodeRec* parents,
                           const CircuitInfo* circuits, uint32_t traj_base, uint32_t* offs, NodeRec* children,
                           uint32_t* shots_out, bool counted, cudaStream_t st);   // counted: traj_branch filled the table
void launch_root_shots(const NodeRec* nodes, const uint32_t* owner, uint32_t* shots, uint32_t n_traj, cudaStream_t st);
void launch_fill_root_probs(double* P, uint32_t n_nodes, cudaStream_t st);
void launch_fill_owner(const NodeRec* nodes, uint32_t n_nodes, uint32_t* owner, uint32_t n_traj, cudaStream_t st);
void launch_iota(uint32_t* p, uint32_t n, uint32_t start, cudaStream_t st);

// ---- sampling
// comp != NULL: leaves flagged NODE_CROSSED first form their chunk sums from their group's shared sums (ComposeRec)
void launch_scan_chunks(const NodeRec* leaves, uint32_t n_leaves, double* chunk_sums, uint64_t chunk_stride,
                        double* block_sums, uint64_t block_stride, uint32_t n_chunks, cudaStream_t st,
                        const ComposeRec* comp = nullptr, const double* c0_sums = nullptr);
template <typename T>
void launch_sample(const NodeRec* leaves, uint32_t n_leaves, const void* states, uint64_t slot_stride,
                   const double* chunk_sums, uint64_t chunk_stride, const double* block_sums, uint64_t block_stride,
                   uint32_t n_chunks, int n_bits, const uint32_t* owner, const uint32_t* tids, uint32_t n_traj,
                   uint64_t seed, TrajMap map, const double* flip /*dev [circuit][n][2] = P(1|0), P(0|1); or null*/,
                   uint32_t* hist, uint32_t* outcomes, uint64_t shot_begin, cudaStream_t st);

// ---- density-matrix helpers / probability vectors
template <typename T>
void launch_setting_probs(const void* states, uint64_t slot_stride, uint32_t n_vec, int n, int n_settings,
                          const uint8_t* qubits, const double2* sops, T* probs, cudaStream_t st);
template <typename T>
void launch_extract_diag(const void* states, uint64_t slot_stride, const uint32_t* slots, uint32_t n_vec, int n,
                         T* probs, cudaStream_t st);
template <typename T>
void launch_readout_probs(T* probs, uint32_t n_vec, int n, const double* conf /*dev [vec][n][4]*/, cudaStream_t st);
template <typename T>
void launch_prob_chunks(const T* probs, uint32_t n_vec, int n, double* chunk_sums, uint64_t chunk_stride,
                        cudaStream_t st);
template <typename T>
void launch_sample_probs(const T* probs, uint32_t n_vec, int n, const double* chunk_sums, uint64_t chunk_stride,
                         const double* block_sums, uint64_t block_stride, uint32_t n_chunks, uint64_t shot_begin,
                         uint64_t n_shots, uint64_t seed, uint32_t first_circuit, const double* flip, uint32_t* hist,
                         cudaStream_t st);
void launch_scan_vectors(uint32_t n_vec, double* chunk_sums, uint64_t chunk_stride, double* block_sums,
                         uint64_t block_stride, uint32_t n_chunks, cudaStream_t st);

// ---- parity
void launch_parity_counts(const uint32_t* hist, uint32_t n_vec, int n, const uint32_t* masks, int n_masks,
                          int per_vec, long long* signed_sums, long long* totals, cudaStream_t st);
void launch_parity_outcomes(const uint32_t* outcomes, uint32_t n_vec, uint64_t shots, const uint32_t* masks, int n_masks,
                            int per_vec, long long* signed_sums, cudaStream_t st);
template <typename T>
void launch_parity_probs(const T* probs, uint32_t n_vec, int n, const uint32_t* masks, int n_masks, int per_vec,
                         double* partial, uint32_t n_part, double* expvals, cudaStream_t st);

// ---- full stabilizer group (stab_group.cu)
void launch_stabilizer_group(const uint32_t* gen_x, const uint32_t* gen_z, int n_gens, uint64_t begin, uint64_t count,
                             uint32_t* x_out, uint32_t* z_out, uint8_t* sign_out, cudaStream_t st);
uint32_t group_expvals_chunks(int n);    // column chunks per density matrix: partial holds [vec][chunks][2^n] doubles
template <typename T>
void launch_group_expvals(const void* states, uint64_t slot_stride, uint32_t n_vec, int n, const uint32_t* adj_dev,
                          double* partial, double* out /*[vec][2^n], indexed by the element's x mask*/, cudaStream_t st);

}  // namespace qcd
