// device_noise.h -- calibration-style ("device") noise made explicit on the host, natively and for whole batches:
// (T1, T2, gate_error, gate_length) -> depolarizing Pauli mixture + per-qubit thermal-relaxation Kraus operators
// written after every gate of a gate list.  This is what qiskit-aer's NoiseModel.from_backend + the noise-model
// attachment at execute() do for the reference (qcdenoise/samplers.py:577-611, :322-324).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/qcdsim.h"

namespace qcd {

// Kraus operators of T1/T2 relaxation over `duration` (same unit as t1, t2; decay towards |0>).  Writes
// out[0] = m and then m 2x2 complex matrices (8 doubles each, row-major (re, im)); returns 1 + 8 m, or -1 for
// parameters out of range.  out must hold 49 doubles.
int thermal_relaxation_kraus(double t1, double t2, double duration, double* out);

struct GateError {
    bool has_pauli = false;
    bool has_relax = false;
    double pauli[16];      // 4^k probabilities, index = P(q0) + 4 P(q1)
    double kraus[2][49];   // per gate qubit: m, m matrices
    int kraus_len[2] = {0, 0};
};

// Error of one k-qubit gate (k = 1, 2): depolarizing strength chosen so that depolarizing o relaxation has average
// gate infidelity `gate_error`.  NaN / <= 0 gate_error or gate_length = that part is absent.  Returns false for
// parameters out of range.
bool device_gate_error(int k, double gate_error, double gate_length_ns, const double* t1_us, const double* t2_us,
                       GateError& out);

// Expands circuits [c_lo, c_hi) of a noise-free batch: after every op whose op_gate_props row is not (NaN, NaN) come
// PAULI1Q/2Q and one KRAUS1Q per gate qubit.  Appends to ops_out / params_out (params of the source ops are copied
// over), n_ops_out[c - c_lo] = expanded op count of circuit c.
int expand_device_noise(const qcd_op* ops, const uint32_t* op_offsets, const double* params, size_t n_params, int c_lo,
                        int c_hi, int n_qubits, const double* qubit_t1_t2, const double* op_gate_props,
                        std::vector<qcd_op>& ops_out, std::vector<uint32_t>& n_ops_out, std::vector<double>& params_out,
                        std::string& err);

}  // namespace qcd
