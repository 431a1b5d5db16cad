// device_noise.cpp -- see device_noise.h.  Host-only, no CUDA calls.  The formulas restate qiskit-aer's published
// noise constructors (thermal_relaxation_error, depolarizing_error, basic_device_gate_errors), which is what the
// reference reaches through NoiseModel.from_backend (qcdenoise/samplers.py:588, :608).  Operator order and the
// drop thresholds are the ones of qcdenoise_b200/channels.py, so that the explicit (Python) and the native route
// produce the same program and therefore the same shots.
#include "device_noise.h"

#include <cmath>
#include <cstring>

namespace qcd {

namespace {

// m <- m + one matrix [[a, b], [c, d]] (real entries) if its largest entry exceeds thr
inline void push_real(double* out, int& m, double a, double b, double c, double d, double thr) {
    if (std::max(std::max(std::fabs(a), std::fabs(b)), std::max(std::fabs(c), std::fabs(d))) <= thr) return;
    double* k = out + 1 + 8 * m;
    k[0] = a; k[1] = 0; k[2] = b; k[3] = 0; k[4] = c; k[5] = 0; k[6] = d; k[7] = 0;
    m++;
}

inline size_t op_param_len(const qcd_op& op, const double* params, size_t n_params) {
    switch (op.opcode) {
        case QCD_OP_RX: case QCD_OP_RY: case QCD_OP_RZ: return 1;
        case QCD_OP_U1Q: return 8;
        case QCD_OP_U2Q: return 32;
        case QCD_OP_PAULI1Q: return 4;
        case QCD_OP_PAULI2Q: return 16;
        case QCD_OP_KRAUS1Q: case QCD_OP_KRAUS2Q: {
            if (op.param_off >= n_params) return (size_t)-1;
            const double m = params[op.param_off];
            if (!(m >= 1 && m <= 64)) return (size_t)-1;
            return 1 + (size_t)m * (op.opcode == QCD_OP_KRAUS1Q ? 8 : 32);
        }
        default: return 0;
    }
}

}  // namespace

int thermal_relaxation_kraus(double t1, double t2, double duration, double* out) {
    if (!(t1 > 0) || !(t2 > 0) || !(duration >= 0)) return -1;
    if (t2 > 2 * t1 * (1 + 1e-12)) return -1;
    const double p_reset = std::isfinite(t1) ? 1.0 - std::exp(-duration / t1) : 0.0;
    const double coh = std::isfinite(t2) ? std::exp(-duration / t2) : 1.0;
    const double to0 = p_reset;   // decay towards |0> only (no thermal population)
    int m = 0;
    if (t2 <= t1) {
        // mixture {identity, Z, reset}
        const double surv = 1.0 - p_reset;
        const double pz = surv > 0 ? 0.5 * surv * (1.0 - coh / surv) : 0.0;
        const double pid = 1.0 - pz - to0 - 0.0;
        const double sid = std::sqrt(std::max(pid, 0.0)), sz = std::sqrt(std::max(pz, 0.0)), s0 = std::sqrt(to0);
        push_real(out, m, sid, 0, 0, sid, 1e-15);
        push_real(out, m, sz, 0, 0, -sz, 1e-15);
        push_real(out, m, s0, 0, 0, 0, 1e-15);
        push_real(out, m, 0, s0, 0, 0, 1e-15);
    } else {
        // T1 < T2 <= 2 T1: Kraus operators from the Choi matrix; its coherent 2x2 block couples |00> and |11>
        const double a = 1.0, d = 1.0 - to0;
        const double tr = a + d, det = a * d - coh * coh;
        const double disc = std::sqrt(std::max(tr * tr / 4 - det, 0.0));
        const double lams[2] = {tr / 2 + disc, tr / 2 - disc};
        for (double lam : lams) {
            if (lam <= 1e-14) continue;
            double v0, v1;
            if (std::fabs(coh) > 1e-300) {
                v0 = coh;
                v1 = lam - a;
            } else if (std::fabs(lam - a) < std::fabs(lam - d)) {
                v0 = 1.0; v1 = 0.0;
            } else {
                v0 = 0.0; v1 = 1.0;
            }
            const double nrm = std::sqrt(v0 * v0 + v1 * v1);
            const double s = std::sqrt(lam);
            double* k = out + 1 + 8 * m;
            std::memset(k, 0, 8 * sizeof(double));
            k[0] = s * (v0 / nrm);
            k[6] = s * (v1 / nrm);
            m++;
        }
        if (to0 > 1e-15) push_real(out, m, 0, std::sqrt(to0), 0, 0, 0.0);
    }
    out[0] = (double)m;
    return 1 + 8 * m;
}

bool device_gate_error(int k, double gate_error, double gate_length_ns, const double* t1_us, const double* t2_us,
                       GateError& out) {
    out.has_pauli = out.has_relax = false;
    out.kraus_len[0] = out.kraus_len[1] = 0;
    if (k != 1 && k != 2) return false;
    double f_relax = 1.0;
    if (gate_length_ns > 0) {   // false for NaN
        double f_pro = 1.0;
        for (int i = 0; i < k; i++) {
            const double t1n = t1_us[i] * 1e3;
            const int len = thermal_relaxation_kraus(t1n, std::min(t2_us[i] * 1e3, 2 * t1n), gate_length_ns, out.kraus[i]);
            if (len < 0) return false;
            out.kraus_len[i] = len;
            // process fidelity of one qubit's channel: sum_k |tr K_k|^2 / 4 (all entries here are real)
            const int m = (int)out.kraus[i][0];
            double f = 0;
            for (int j = 0; j < m; j++) {
                const double* kk = out.kraus[i] + 1 + 8 * j;
                const double t = kk[0] + kk[6];
                f += t * t;
            }
            f_pro *= f / 4.0;
        }
        const double dim_all = (double)(1 << k);
        f_relax = (dim_all * f_pro + 1.0) / (dim_all + 1.0);
        out.has_relax = true;
    }
    if (gate_error > 0) {
        const double dim = (double)(1 << k);
        const double target = std::min(gate_error, dim / (dim + 1));
        if (1.0 - f_relax <= target) {
            const int terms = 1 << (2 * k);
            const double p = std::min(dim * (target - (1.0 - f_relax)) / (dim * f_relax - 1.0), (double)terms / (terms - 1));
            if (p > 0) {
                for (int i = 0; i < terms; i++) out.pauli[i] = p / terms;
                out.pauli[0] = 1.0 - p * (terms - 1) / terms;
                out.has_pauli = true;
            }
        }
    }
    return true;
}

int expand_device_noise(const qcd_op* ops, const uint32_t* op_offsets, const double* params, size_t n_params, int c_lo,
                        int c_hi, int n_qubits, const double* qubit_t1_t2, const double* op_gate_props,
                        std::vector<qcd_op>& ops_out, std::vector<uint32_t>& n_ops_out, std::vector<double>& params_out,
                        std::string& err) {
    struct Memo {
        uint16_t opcode;
        uint8_t q0, q1;
        double ge, gl;
        bool has_pauli;
        uint32_t pauli_off, kraus_off[2];
        int n_kraus;
    };
    std::vector<Memo> memo;
    GateError ge;
    for (int c = c_lo; c < c_hi; c++) {
        memo.clear();   // one (gate, qubits) entry is shared by all its occurrences in the circuit
        const size_t first = ops_out.size();
        const double* tq = qubit_t1_t2 + (size_t)c * n_qubits * 2;
        for (uint32_t i = op_offsets[c]; i < op_offsets[c + 1]; i++) {
            qcd_op op = ops[i];
            struct Where {   // built only on error
                int c;
                uint32_t i;
                std::string operator+(const char* m) const {
                    return "circuit " + std::to_string(c) + ", op " + std::to_string(i) + ": " + m;
                }
            } where{c, i - op_offsets[c]};
            if (op.opcode >= QCD_OP_COUNT) { err = where + "unknown opcode"; return QCD_E_INVALID; }
            const bool two = op.opcode == QCD_OP_CX || op.opcode == QCD_OP_CZ || op.opcode == QCD_OP_SWAP ||
                             op.opcode == QCD_OP_U2Q || op.opcode == QCD_OP_KRAUS2Q || op.opcode == QCD_OP_PAULI2Q;
            if (op.q0 >= n_qubits || (two && (op.q1 >= n_qubits || op.q1 == op.q0))) {
                err = where + "qubit index out of range";
                return QCD_E_INVALID;
            }
            const size_t plen = op_param_len(op, params, n_params);
            if (plen == (size_t)-1 || (plen && (size_t)op.param_off + plen > n_params)) {
                err = where + "op parameters run past params[]";
                return QCD_E_INVALID;
            }
            if (plen) {
                const uint32_t off = (uint32_t)params_out.size();
                params_out.insert(params_out.end(), params + op.param_off, params + op.param_off + plen);
                op.param_off = off;
            }
            ops_out.push_back(op);
            const double g_err = op_gate_props[2 * (size_t)i], g_len = op_gate_props[2 * (size_t)i + 1];
            if (std::isnan(g_err) && std::isnan(g_len)) continue;
            const bool stochastic = op.opcode >= QCD_OP_KRAUS1Q;
            if (stochastic) { err = where + "a noise op cannot carry a gate error"; return QCD_E_INVALID; }
            const int k = two ? 2 : 1;
            const Memo* hit = nullptr;
            for (const Memo& mm : memo)
                if (mm.opcode == op.opcode && mm.q0 == op.q0 && mm.q1 == op.q1 &&
                    std::memcmp(&mm.ge, &g_err, sizeof(double)) == 0 && std::memcmp(&mm.gl, &g_len, sizeof(double)) == 0) {
                    hit = &mm;
                    break;
                }
            if (!hit) {
                const int qs[2] = {op.q0, two ? op.q1 : 0};
                double t1[2], t2[2];
                for (int j = 0; j < k; j++) {
                    t1[j] = std::max(tq[2 * qs[j]], 1e-9);
                    t2[j] = std::max(tq[2 * qs[j] + 1], 1e-9);
                }
                if (!device_gate_error(k, g_err, g_len, t1, t2, ge)) {
                    err = where + "relaxation parameters out of range";
                    return QCD_E_INVALID;
                }
                Memo mm;
                mm.opcode = op.opcode; mm.q0 = op.q0; mm.q1 = op.q1; mm.ge = g_err; mm.gl = g_len;
                mm.has_pauli = ge.has_pauli;
                mm.pauli_off = mm.kraus_off[0] = mm.kraus_off[1] = 0;
                mm.n_kraus = ge.has_relax ? k : 0;
                if (ge.has_pauli) {
                    mm.pauli_off = (uint32_t)params_out.size();
                    params_out.insert(params_out.end(), ge.pauli, ge.pauli + (k == 1 ? 4 : 16));
                }
                for (int j = 0; j < mm.n_kraus; j++) {
                    mm.kraus_off[j] = (uint32_t)params_out.size();
                    params_out.insert(params_out.end(), ge.kraus[j], ge.kraus[j] + ge.kraus_len[j]);
                }
                memo.push_back(mm);
                hit = &memo.back();
            }
            if (hit->has_pauli) {
                qcd_op e;
                e.opcode = two ? QCD_OP_PAULI2Q : QCD_OP_PAULI1Q;
                e.q0 = op.q0;
                e.q1 = two ? op.q1 : 0xFF;
                e.param_off = hit->pauli_off;
                ops_out.push_back(e);
            }
            for (int j = 0; j < hit->n_kraus; j++) {
                qcd_op e;
                e.opcode = QCD_OP_KRAUS1Q;
                e.q0 = j == 0 ? op.q0 : op.q1;
                e.q1 = 0xFF;
                e.param_off = hit->kraus_off[j];
                ops_out.push_back(e);
            }
        }
        n_ops_out.push_back((uint32_t)(ops_out.size() - first));
    }
    return QCD_OK;
}

}  // namespace qcd
