// lower.h -- host-side compiler: gate-list IR (include/qcdsim.h) -> fused, tiled sweep program (program.h).
#pragma once
#include <complex>
#include <string>
#include <vector>

#include "../../include/qcdsim.h"
#include "program.h"

namespace qcd {

typedef std::complex<double> cd;

enum Mode { MODE_SV = 0, MODE_DM = 1 };

// Content hash of everything a circuit executes before its final sweep (sweeps, passes, ops, matrices, sign masks, site
// groups -- including the group the final sweep resolves), with offsets taken relative to the circuit's own arrays.
// Equal signatures <=> the trajectory tree of the two circuits is identical down to the parents of the leaves.
struct PrefixSig {
    uint64_t h0 = 0, h1 = 0;
    uint32_t n_sweeps = 0;
    bool operator==(const PrefixSig& o) const { return h0 == o.h0 && h1 == o.h1 && n_sweeps == o.n_sweeps; }
};

struct Program {
    int n_qubits = 0;
    int n_bits = 0;   // n_qubits (MODE_SV) or 2 n_qubits (MODE_DM)
    int mode = MODE_SV;
    int k_max = 13;   // tile bits for this precision
    int max_group_bits = 0;
    int max_sweeps = 0;
    std::vector<Sweep> sweeps;
    std::vector<Pass> passes;
    std::vector<KOp> kops;
    std::vector<cd> pool;
    std::vector<uint64_t> masks;
    std::vector<Group> groups;
    std::vector<SubSite> subs;
    std::vector<double> gdata;
    std::vector<CircuitInfo> circuits;
    std::vector<PrefixSig> prefix;   // one per circuit
};

// Appends one circuit to P. Returns QCD_OK or an error code with a message in err.
int compile_circuit(Program& P, const qcd_op* ops, uint32_t n_ops, const double* params, size_t n_params,
                    std::string& err);

// Appends every circuit of `src` (compiled separately with the same n_qubits / mode / k_max) to `dst`, re-basing all
// offsets.  Used to lower large batches on several host threads.
// pool_base >= 0: src's matrix pool is NOT copied; its offsets are re-based to pool_base (the caller places the pools).
void merge_program(Program& dst, const Program& src, long long pool_base = -1);

// JSON description of circuit `c` of P (for the CPU-side schedule interpreter in tests/).
std::string program_to_json(const Program& P, int c);

}  // namespace qcd
