// dm_cluster.h -- on-chip density-matrix interpreter (dm_cluster.cu): host-side schedule + launch.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/qcdsim.h"

namespace qcd {

constexpr int DMI_MAX_PASS_OPS = 24;    // op records held in shared memory per pass

// One pass: the ops sched[first .. first + n_ops) (indices into the batch's op array, execution order) all act inside
// the qubit pair {q0, q1} (q1 = 0xFF: a single qubit).
struct DmPass {
    uint32_t first;
    uint16_t n_ops;
    uint8_t q0, q1;
};

struct DmInterpArgs {
    const qcd_op* ops;          // [dev] explicit ops of the whole batch (noise channels included)
    const double* params;       // [dev]
    const uint32_t* sched;      // [dev]
    const DmPass* passes;       // [dev]
    const uint32_t* pass_off;   // [dev] [n_batch + 1] pass range of every circuit of the batch
    int n_qubits;
    int n_circuits;             // circuits of this launch: circuit0 .. circuit0 + n_circuits of the batch
    int circuit0;
    int local_bits;             // index bits held by one CTA
    int rec_cap;                // op records that fit the dynamic shared memory behind the amplitudes (>= DMI_MAX_PASS_OPS)
    void* probs;                // [dev] T [n_circuits][2^n] or NULL
    void* rho_out;              // [dev] complex<T> [n_circuits][slot_stride] or NULL
    uint64_t slot_stride;
};

// CTAs per cluster that hold one density matrix of n_qubits (1, 2, 4, 8, 16), or 0 if it does not fit on chip.
int dm_interp_cluster_size(int n_qubits, int precision);

// Records the kernel can hold per CTA next to the amplitudes: enough for a whole circuit when that fits (max_circuit_ops =
// most scheduled ops of any circuit of the batch), so that one build phase serves all its passes.
int dm_interp_rec_cap(int n_qubits, int precision, uint32_t max_circuit_ops);
size_t dm_interp_record_bytes(int precision);

// Host scheduler, O(ops), no matrix arithmetic: validates the ops of one circuit (op_base = index of ops[0] in the batch
// array) and appends its passes / execution order.  Every two-qubit gate opens (or extends) a pass on its pair; single-
// qubit ops join the last pass that touched their qubit, or wait for the first one that will.
int dm_interp_schedule(const qcd_op* ops, uint32_t n_ops, uint32_t op_base, const double* params, size_t n_params,
                       int n_qubits, std::vector<DmPass>& passes, std::vector<uint32_t>& sched, std::string& err);

// max_clusters <= 0: as many clusters as the device holds at once (persistent: each walks over several circuits).
template <typename T>
cudaError_t launch_dm_interp(const DmInterpArgs& a, int cluster_size, int max_clusters, cudaStream_t st);

}  // namespace qcd
