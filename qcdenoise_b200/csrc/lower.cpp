// lower.cpp -- gate-list IR -> sweep program.
//
// Pipeline (per circuit):
//   1. lower     : IR ops -> LOps on state bits.  CX = H.CZ.H, Z/CZ = sign masks, phases = D1, stochastic ops =
//                  SITE marker + variant matrices selected by the site's branch field.  In density-matrix mode
//                  every gate acts on row bit q and (conjugated) on column bit q + n; channels become
//                  superoperators on (q, q+n) / (q0, q1, q0+n, q1+n); there are no sites.
//   2. fuse      : peephole: products of 1q matrices on the same bit, 1q absorbed into neighbouring 2q, identities
//                  dropped, diagonals demoted to D1 / sign masks, duplicate sign masks cancelled.  State-dependent
//                  sites are barriers (their populations are measured at that point).
//   3. segment   : cut into sweeps.  A sweep starts with the branch-dependent matrices of one site group (whose
//                  populations the previous sweep reduced) and runs until the next state-dependent site or until
//                  the non-diagonal target bits no longer fit one tile.
//   4. tile/pass : choose contiguous low bits + scattered high bits; group ops into passes of <= 4 register bits.
#include "lower.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <set>
#include <sstream>

namespace qcd {
namespace {

const double EPS = 1e-14;
const double SQ2 = 0.70710678118654752440;

struct LOp {
    enum Kind { M1, M2, M4, SIGN, D1, SITE } kind = M1;
    int nq = 0;
    int q[4] = {0, 0, 0, 0};
    int nvar = 1;
    std::vector<cd> mats;  // nvar * dim * dim (row-major), dim = 1 << nq ; D1: 2 entries
    int site = -1;         // variant selector / SITE index
    int fshift = 0, fmask = 0;
    uint64_t mask = 0;     // SIGN
    bool dead = false;
};

struct LSite {
    bool dependent = false;
    bool monomial = true;
    bool unit = false;  // variants are norm-preserving (bare Paulis): no 1/sqrt(p) rescale
    int nq = 1;
    int q[2] = {-1, -1};
    int nbranch = 1;
    uint32_t stream = 0;
    std::vector<double> w;  // nbranch * dim       diag(K^dag K)
    std::vector<double> T;  // nbranch * dim * dim |K[r][c]|^2
};

bool touches(const LOp& o, int b) {
    if (o.kind == LOp::SIGN) return (o.mask >> b) & 1;
    for (int i = 0; i < o.nq; i++)
        if (o.q[i] == b) return true;
    return false;
}

std::vector<cd> matmul(const std::vector<cd>& a, const std::vector<cd>& b, int d) {  // a * b
    std::vector<cd> c(d * d, cd(0, 0));
    for (int i = 0; i < d; i++)
        for (int k = 0; k < d; k++) {
            cd aik = a[i * d + k];
            if (aik == cd(0, 0)) continue;
            for (int j = 0; j < d; j++) c[i * d + j] += aik * b[k * d + j];
        }
    return c;
}

// embed a 1q matrix acting on slot `slot` of an nq-bit operator
std::vector<cd> embed1(const cd* m, int slot, int nq) {
    int d = 1 << nq;
    std::vector<cd> r(d * d, cd(0, 0));
    for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++) {
            if ((i & ~(1 << slot)) != (j & ~(1 << slot))) continue;
            r[i * d + j] = m[((i >> slot) & 1) * 2 + ((j >> slot) & 1)];
        }
    return r;
}

// re-index an operator on bits (q[0..nq)) to the order given by perm: new slot s holds old slot perm[s]
std::vector<cd> permute_bits(const std::vector<cd>& m, const int* perm, int nq) {
    int d = 1 << nq;
    int map[16];
    for (int idx = 0; idx < d; idx++) {
        int o = 0;
        for (int s = 0; s < nq; s++)
            if ((idx >> s) & 1) o |= 1 << perm[s];
        map[idx] = o;
    }
    std::vector<cd> r(d * d);
    for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++) r[i * d + j] = m[map[i] * d + map[j]];
    return r;
}

bool is_identity(const std::vector<cd>& m, int d) {
    for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++)
            if (std::norm(m[i * d + j] - (i == j ? cd(1, 0) : cd(0, 0))) > EPS * EPS) return false;   // |z| > EPS, no hypot
    return true;
}
bool is_diag(const cd* m, int d) {
    for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++)
            if (i != j && std::norm(m[i * d + j]) > EPS * EPS) return false;
    return true;
}

struct Lowerer {
    int n, nb, mode;
    const double* params;
    size_t n_params;
    std::vector<LOp> ops;
    std::vector<LSite> sites;
    std::string err;
    int code = QCD_OK;

    bool fail(int c, const std::string& m) {
        code = c;
        err = m;
        return false;
    }
    bool need(uint32_t off, size_t cnt) {
        if ((size_t)off + cnt > n_params) return fail(QCD_E_INVALID, "op parameters run past params[]");
        return true;
    }
    void push_m(int kind, std::initializer_list<int> qs, std::vector<cd> m) {
        LOp o;
        o.kind = (LOp::Kind)kind;
        o.nq = (int)qs.size();
        int i = 0;
        for (int q : qs) o.q[i++] = q;
        o.mats = std::move(m);
        ops.push_back(std::move(o));
    }
    void push_sign(uint64_t mask) {
        LOp o;
        o.kind = LOp::SIGN;
        o.mask = mask;
        ops.push_back(o);
    }
    static std::vector<cd> conj(const std::vector<cd>& m) {
        std::vector<cd> r(m.size());
        for (size_t i = 0; i < m.size(); i++) r[i] = std::conj(m[i]);
        return r;
    }
    // unitary 1q gate on qubit q (both modes)
    void gate1(int q, const std::vector<cd>& u) {
        push_m(LOp::M1, {q}, u);
        if (mode == MODE_DM) push_m(LOp::M1, {q + n}, conj(u));
    }
    void gate2(int q0, int q1, const std::vector<cd>& u) {
        push_m(LOp::M2, {q0, q1}, u);
        if (mode == MODE_DM) push_m(LOp::M2, {q0 + n, q1 + n}, conj(u));
    }
    void cz(int a, int b) {
        push_sign((1ull << a) | (1ull << b));
        if (mode == MODE_DM) push_sign((1ull << (a + n)) | (1ull << (b + n)));
    }
    void zgate(int a) {
        push_sign(1ull << a);
        if (mode == MODE_DM) push_sign(1ull << (a + n));
    }

    // channel given as Kraus list on nq qubits
    // K = true Kraus operators (define the channel, the branch weights and the population maps);
    // pauli_probs: K_j = sqrt(p_j) Pauli_j -- the trajectory variants are then the bare Paulis (unit norm).
    bool channel(const int* q, int nq, const std::vector<std::vector<cd>>& K, uint32_t stream, bool pauli_probs,
                 const double* probs) {
        int d = 1 << nq;
        if (mode == MODE_DM) {
            // superoperator on (q..., q+n...): index = x + d*y ; S[(x,y),(x',y')] = sum_j K[x,x'] conj(K[y,y'])
            int D = d * d;
            std::vector<cd> S(D * D, cd(0, 0));
            for (const auto& k : K) {
                // Kraus operators are sparse (Paulis, damping operators: <= d non-zeros): visit non-zero pairs only
                int nz_idx[16], nz = 0;
                for (int e = 0; e < d * d; e++)
                    if (k[e] != cd(0, 0)) nz_idx[nz++] = e;
                for (int i = 0; i < nz; i++) {
                    const int x = nz_idx[i] / d, xp = nz_idx[i] % d;
                    const cd a = k[nz_idx[i]];
                    for (int j = 0; j < nz; j++) {
                        const int y = nz_idx[j] / d, yp = nz_idx[j] % d;
                        S[(x + d * y) * D + (xp + d * yp)] += a * std::conj(k[nz_idx[j]]);
                    }
                }
            }
            if (nq == 1)
                push_m(LOp::M2, {q[0], q[0] + n}, std::move(S));
            else
                push_m(LOp::M4, {q[0], q[1], q[0] + n, q[1] + n}, std::move(S));
            return true;
        }
        LSite s;
        s.nq = nq;
        s.q[0] = q[0];
        s.q[1] = nq > 1 ? q[1] : -1;
        s.nbranch = (int)K.size();
        s.stream = stream;
        s.w.assign(s.nbranch * d, 0.0);
        s.T.assign(s.nbranch * d * d, 0.0);
        bool dep = false;
        for (int j = 0; j < s.nbranch; j++) {
            const auto& k = K[j];
            for (int c = 0; c < d; c++)
                for (int c2 = 0; c2 < d; c2++) {
                    cd acc(0, 0);
                    for (int r = 0; r < d; r++) acc += std::conj(k[r * d + c]) * k[r * d + c2];
                    if (c == c2)
                        s.w[j * d + c] = acc.real();
                    else if (std::abs(acc) > 1e-12)
                        return fail(QCD_E_UNSUPPORTED,
                                    "Kraus operator with non-diagonal K^dag K in trajectory mode (use the "
                                    "density-matrix mode for this channel)");
                }
            if (pauli_probs)
                for (int c = 0; c < d; c++) s.w[j * d + c] = probs[j];  // exact, state-independent
            for (int c = 1; c < d; c++)
                if (std::abs(s.w[j * d + c] - s.w[j * d]) > 1e-15) dep = true;
            for (int r = 0; r < d; r++) {
                int nz = 0;
                for (int c = 0; c < d; c++) {
                    double a = std::norm(k[r * d + c]);
                    s.T[(j * d + r) * d + c] = a;
                    if (a > 1e-30) nz++;
                }
                if (nz > 1) s.monomial = false;
            }
            for (int c = 0; c < d; c++) {
                int nz = 0;
                for (int r = 0; r < d; r++)
                    if (std::norm(k[r * d + c]) > 1e-30) nz++;
                if (nz > 1) s.monomial = false;
            }
        }
        s.dependent = dep;
        s.unit = pauli_probs;
        int sidx = (int)sites.size();
        sites.push_back(s);
        LOp mark;
        mark.kind = LOp::SITE;
        mark.site = sidx;
        ops.push_back(mark);
        int bits = 0;
        while ((1 << bits) < s.nbranch) bits++;
        if (pauli_probs) {
            // index = P(q0) [+ 4 P(q1)]: one bare-Pauli variant op per qubit, each reading its own 2-bit field
            static const cd PA[4][4] = {{1, 0, 0, 1}, {0, 1, 1, 0}, {0, cd(0, -1), cd(0, 1), 0}, {1, 0, 0, -1}};
            for (int h = 0; h < nq; h++) {
                LOp o;
                o.kind = LOp::M1;
                o.nq = 1;
                o.q[0] = q[h];
                o.nvar = 4;
                for (int p = 0; p < 4; p++)
                    for (int e = 0; e < 4; e++) o.mats.push_back(PA[p][e]);
                o.site = sidx;
                o.fshift = 2 * h;
                o.fmask = 3;
                ops.push_back(std::move(o));
            }
            return true;
        }
        LOp o;
        o.kind = nq == 1 ? LOp::M1 : LOp::M2;
        o.nq = nq;
        o.q[0] = q[0];
        o.q[1] = nq > 1 ? q[1] : 0;
        o.nvar = s.nbranch;
        for (const auto& k : K) o.mats.insert(o.mats.end(), k.begin(), k.end());
        o.site = sidx;
        o.fshift = 0;
        o.fmask = (1 << bits) - 1;
        ops.push_back(std::move(o));
        return true;
    }

    bool lower(const qcd_op* in, uint32_t n_ops) {
        const std::vector<cd> H = {SQ2, SQ2, SQ2, -SQ2};
        uint32_t stream = 0;
        for (uint32_t i = 0; i < n_ops; i++) {
            const qcd_op& op = in[i];
            int a = op.q0, b = op.q1;
            bool two = false;
            switch (op.opcode) {
                case QCD_OP_CX: case QCD_OP_CZ: case QCD_OP_SWAP: case QCD_OP_U2Q: case QCD_OP_KRAUS2Q:
                case QCD_OP_PAULI2Q:
                    two = true;
            }
            if (a >= n || (two && (b >= n || b == a))) return fail(QCD_E_INVALID, "qubit index out of range");
            const double* p = params + op.param_off;
            switch (op.opcode) {
                case QCD_OP_ID: break;
                case QCD_OP_H: gate1(a, H); break;
                case QCD_OP_X: gate1(a, {0, 1, 1, 0}); break;
                case QCD_OP_Y: gate1(a, {0, cd(0, -1), cd(0, 1), 0}); break;
                case QCD_OP_Z: zgate(a); break;
                case QCD_OP_S: gate1(a, {1, 0, 0, cd(0, 1)}); break;
                case QCD_OP_SDG: gate1(a, {1, 0, 0, cd(0, -1)}); break;
                case QCD_OP_T: gate1(a, {1, 0, 0, cd(SQ2, SQ2)}); break;
                case QCD_OP_TDG: gate1(a, {1, 0, 0, cd(SQ2, -SQ2)}); break;
                case QCD_OP_RX: case QCD_OP_RY: case QCD_OP_RZ: {
                    if (!need(op.param_off, 1)) return false;
                    double c = std::cos(p[0] / 2), s = std::sin(p[0] / 2);
                    if (op.opcode == QCD_OP_RX) gate1(a, {c, cd(0, -s), cd(0, -s), c});
                    else if (op.opcode == QCD_OP_RY) gate1(a, {c, -s, s, c});
                    else gate1(a, {cd(c, -s), 0, 0, cd(c, s)});
                    break;
                }
                case QCD_OP_U1Q: {
                    if (!need(op.param_off, 8)) return false;
                    std::vector<cd> u(4);
                    for (int e = 0; e < 4; e++) u[e] = cd(p[2 * e], p[2 * e + 1]);
                    gate1(a, u);
                    break;
                }
                case QCD_OP_CX: gate1(b, H); cz(a, b); gate1(b, H); break;
                case QCD_OP_CZ: cz(a, b); break;
                case QCD_OP_SWAP: gate2(a, b, {1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1}); break;
                case QCD_OP_U2Q: {
                    if (!need(op.param_off, 32)) return false;
                    std::vector<cd> u(16);
                    for (int e = 0; e < 16; e++) u[e] = cd(p[2 * e], p[2 * e + 1]);
                    gate2(a, b, u);
                    break;
                }
                case QCD_OP_KRAUS1Q: case QCD_OP_KRAUS2Q: {
                    if (!need(op.param_off, 1)) return false;
                    int m = (int)p[0];
                    int nq = op.opcode == QCD_OP_KRAUS1Q ? 1 : 2, d2 = nq == 1 ? 4 : 16;
                    if (m < 1 || m > 64 || !need(op.param_off + 1, (size_t)m * d2 * 2))
                        return fail(QCD_E_INVALID, "bad Kraus operator count");
                    if (nq == 2 && mode == MODE_SV && m > 16)
                        return fail(QCD_E_UNSUPPORTED, "more than 16 two-qubit Kraus operators");
                    std::vector<std::vector<cd>> K(m, std::vector<cd>(d2));
                    for (int j = 0; j < m; j++)
                        for (int e = 0; e < d2; e++) K[j][e] = cd(p[1 + (j * d2 + e) * 2], p[2 + (j * d2 + e) * 2]);
                    int qq[2] = {a, b};
                    if (!channel(qq, nq, K, stream++, false, nullptr)) return false;
                    break;
                }
                case QCD_OP_PAULI1Q: case QCD_OP_PAULI2Q: {
                    int nq = op.opcode == QCD_OP_PAULI1Q ? 1 : 2, np = nq == 1 ? 4 : 16;
                    if (!need(op.param_off, np)) return false;
                    static const cd PA[4][4] = {{1, 0, 0, 1}, {0, 1, 1, 0}, {0, cd(0, -1), cd(0, 1), 0}, {1, 0, 0, -1}};
                    if (mode == MODE_DM) {
                        // Pauli mixture as a superoperator, straight from the monomial form of the Paulis:
                        // P_j[r, r ^ f_j] = ph_j(r)  =>  S[(x, y), (x ^ f, y ^ f)] += p_j ph_j(x) conj(ph_j(y))
                        static const int FL[4] = {0, 1, 1, 0};
                        static const cd PH[4][2] = {{1, 1}, {1, 1}, {cd(0, -1), cd(0, 1)}, {1, -1}};   // P[r][r ^ f]
                        const int d = 1 << nq, D = d * d;
                        std::vector<cd> S(D * D, cd(0, 0));
                        double tot = 0;
                        for (int j = 0; j < np; j++) {
                            if (p[j] < 0) return fail(QCD_E_INVALID, "negative Pauli probability");
                            tot += p[j];
                            if (p[j] == 0) continue;
                            const int j0 = j & 3, j1 = j >> 2;
                            const int f = nq == 1 ? FL[j0] : (FL[j0] | (FL[j1] << 1));
                            for (int x = 0; x < d; x++) {
                                const cd px = nq == 1 ? PH[j0][x] : PH[j0][x & 1] * PH[j1][x >> 1];
                                for (int y = 0; y < d; y++) {
                                    const cd py = nq == 1 ? PH[j0][y] : PH[j0][y & 1] * PH[j1][y >> 1];
                                    S[(x + d * y) * D + ((x ^ f) + d * (y ^ f))] += p[j] * px * std::conj(py);
                                }
                            }
                        }
                        if (std::abs(tot - 1.0) > 1e-9) return fail(QCD_E_INVALID, "Pauli probabilities do not sum to 1");
                        if (nq == 1)
                            push_m(LOp::M2, {a, a + n}, std::move(S));
                        else
                            push_m(LOp::M4, {a, b, a + n, b + n}, std::move(S));
                        stream++;
                        break;
                    }
                    std::vector<std::vector<cd>> K;
                    double tot = 0;
                    for (int j = 0; j < np; j++) {
                        if (p[j] < 0) return fail(QCD_E_INVALID, "negative Pauli probability");
                        tot += p[j];
                        double s = std::sqrt(p[j]);
                        std::vector<cd> k;
                        if (nq == 1) {
                            for (int e = 0; e < 4; e++) k.push_back(s * PA[j][e]);
                        } else {
                            k.assign(16, cd(0, 0));
                            const cd* m0 = PA[j & 3];
                            const cd* m1 = PA[j >> 2];
                            for (int r = 0; r < 4; r++)
                                for (int c = 0; c < 4; c++)
                                    k[r * 4 + c] = s * m0[(r & 1) * 2 + (c & 1)] * m1[(r >> 1) * 2 + (c >> 1)];
                        }
                        K.push_back(k);
                    }
                    if (std::abs(tot - 1.0) > 1e-9) return fail(QCD_E_INVALID, "Pauli probabilities do not sum to 1");
                    int qq[2] = {a, b};
                    if (!channel(qq, nq, K, stream++, mode == MODE_SV, p)) return false;
                    break;
                }
                case QCD_OP_RESET: {
                    std::vector<std::vector<cd>> K = {{1, 0, 0, 0}, {0, 1, 0, 0}};
                    int qq[2] = {a, -1};
                    if (!channel(qq, 1, K, stream++, false, nullptr)) return false;
                    break;
                }
                default:
                    return fail(QCD_E_INVALID, "unknown opcode");
            }
        }
        return true;
    }

    // ---------------------------------------------------------------- fusion
    bool is_barrier(const LOp& o) const { return o.kind == LOp::SITE && sites[o.site].dependent; }

    // index of the last live op touching bit b in out[0..end), stopping at barriers; -1 if none / barrier first
    int last_toucher(const std::vector<LOp>& out, int b) const {
        for (int i = (int)out.size() - 1; i >= 0; i--) {
            if (out[i].dead) continue;
            if (is_barrier(out[i])) return -1;
            if (out[i].kind == LOp::SITE) continue;
            if (touches(out[i], b)) return i;
        }
        return -1;
    }

    void fuse() {
        std::vector<LOp> out;
        for (LOp& x : ops) {
            if (x.kind == LOp::M1) {
                // diagonal single-variant matrices become D1 / SIGN early (they commute with other diagonals)
                int j = last_toucher(out, x.q[0]);
                // A per-branch (variant) op may only absorb DIAGONAL neighbours: the decide step tracks the joint
                // populations through the site's own Kraus operators, so nothing that moves population (H, X, ...)
                // may hide inside a variant matrix.
                const bool x_diag = x.nvar == 1 && is_diag(x.mats.data(), 2);
                const bool p_diag = j >= 0 && out[j].kind == LOp::M1 && out[j].nvar == 1 && is_diag(out[j].mats.data(), 2);
                const bool both_plain = j >= 0 && out[j].kind == LOp::M1 && x.nvar == 1 && out[j].nvar == 1;
                const bool can_fuse = j >= 0 && out[j].kind == LOp::M1 &&
                                      (both_plain || (x.nvar > 1 && out[j].nvar == 1 && p_diag) ||
                                       (x.nvar == 1 && out[j].nvar > 1 && x_diag));
                if (can_fuse) {
                    LOp& p = out[j];  // x after p : new = x * p
                    int nv = std::max(x.nvar, p.nvar);
                    std::vector<cd> m;
                    for (int v = 0; v < nv; v++) {
                        std::vector<cd> xv(x.mats.begin() + (x.nvar == 1 ? 0 : v * 4), x.mats.begin() + (x.nvar == 1 ? 0 : v * 4) + 4);
                        std::vector<cd> pv(p.mats.begin() + (p.nvar == 1 ? 0 : v * 4), p.mats.begin() + (p.nvar == 1 ? 0 : v * 4) + 4);
                        std::vector<cd> r = matmul(xv, pv, 2);
                        m.insert(m.end(), r.begin(), r.end());
                    }
                    if (x.nvar > 1) {
                        // keep the variant op AFTER its SITE marker: absorb p into x and emit x here
                        p.dead = true;
                        x.mats = std::move(m);
                        out.push_back(std::move(x));
                        continue;
                    }
                    p.nvar = nv;
                    p.mats = std::move(m);
                    // the merged op sits at position j: valid because nothing between j and the end touches this bit
                    if (p.nvar == 1 && is_identity(p.mats, 2)) p.dead = true;
                    continue;
                }
                if (j >= 0 && (out[j].kind == LOp::M2 || out[j].kind == LOp::M4) && x.nvar == 1 && out[j].nvar == 1) {
                    LOp& p = out[j];
                    int slot = 0;
                    for (int s = 0; s < p.nq; s++)
                        if (p.q[s] == x.q[0]) slot = s;
                    p.mats = matmul(embed1(x.mats.data(), slot, p.nq), p.mats, 1 << p.nq);
                    continue;
                }
                if (x.nvar == 1 && is_identity(x.mats, 2)) continue;
                out.push_back(std::move(x));
                continue;
            }
            if ((x.kind == LOp::M2 || x.kind == LOp::M4) && x.nvar == 1) {
                int d = 1 << x.nq;
                if (is_identity(x.mats, d)) continue;  // e.g. the reference's labelled identity `unitary` markers
                // same-support predecessor: multiply
                bool merged = false;
                {
                    int j = last_toucher(out, x.q[0]);
                    if (j >= 0 && out[j].kind == x.kind && out[j].nvar == 1) {
                        bool same = true;
                        for (int s = 0; s < x.nq; s++) same = same && out[j].q[s] == x.q[s];
                        for (int s = 1; s < x.nq && same; s++) same = last_toucher(out, x.q[s]) == j;
                        if (same) {
                            out[j].mats = matmul(x.mats, out[j].mats, d);
                            merged = true;
                        }
                    }
                }
                if (merged) continue;
                // absorb single 1q predecessors
                for (int s = 0; s < x.nq; s++) {
                    int j = last_toucher(out, x.q[s]);
                    if (j >= 0 && out[j].kind == LOp::M1 && out[j].nvar == 1) {
                        x.mats = matmul(x.mats, embed1(out[j].mats.data(), s, x.nq), d);
                        out[j].dead = true;
                    }
                }
                out.push_back(std::move(x));
                continue;
            }
            out.push_back(std::move(x));
        }
        // demote diagonal 1q matrices; drop dead ops
        std::vector<LOp> res;
        for (LOp& o : out) {
            if (o.dead) continue;
            if (o.kind == LOp::M1 && o.nvar == 1 && is_diag(o.mats.data(), 2)) {
                cd d0 = o.mats[0], d1 = o.mats[3];
                if (std::abs(d0 - cd(1, 0)) < EPS && std::abs(d1 - cd(1, 0)) < EPS) continue;
                if (std::abs(d0 - cd(1, 0)) < EPS && std::abs(d1 + cd(1, 0)) < EPS) {
                    LOp s;
                    s.kind = LOp::SIGN;
                    s.mask = 1ull << o.q[0];
                    res.push_back(s);
                    continue;
                }
                LOp d;
                d.kind = LOp::D1;
                d.nq = 1;
                d.q[0] = o.q[0];
                d.mats = {d0, d1};
                res.push_back(std::move(d));
                continue;
            }
            res.push_back(std::move(o));
        }
        ops.swap(res);
    }
};

// ------------------------------------------------------------------------ tiling
struct Tile {
    int k, L, n_high;
    int high[MAX_HIGH];
};

bool choose_tile(const std::set<int>& active, int nb, int k_max, Tile& t) {
    int k = std::min(k_max, nb);
    std::vector<int> a(active.rbegin(), active.rend());  // descending
    int min_low = std::min(MIN_LOW, k);
    for (int h = 0; h <= k - min_low && h <= MAX_HIGH; h++) {
        int L = k - h;
        int cnt = 0;
        for (int b : a)
            if (b >= L) cnt++;
        if (cnt > h) continue;
        t.k = k;
        t.L = L;
        t.n_high = h;
        std::vector<int> hs;
        for (int b : a)
            if (b >= L) hs.push_back(b);
        for (int b = L; (int)hs.size() < h && b < nb; b++)
            if (!active.count(b)) hs.push_back(b);
        if ((int)hs.size() < h) continue;
        std::sort(hs.begin(), hs.end());
        for (int i = 0; i < h; i++) t.high[i] = hs[i];
        return true;
    }
    return false;
}

int local_bit(const Tile& t, int g) {
    if (g < t.L) return g;
    for (int i = 0; i < t.n_high; i++)
        if (t.high[i] == g) return t.L + i;
    return -1;
}

void active_bits(const LOp& o, std::set<int>& s) {
    if (o.kind == LOp::M1 || o.kind == LOp::M2 || o.kind == LOp::M4)
        for (int i = 0; i < o.nq; i++) s.insert(o.q[i]);
}

}  // namespace

int compile_circuit(Program& P, const qcd_op* in, uint32_t n_ops, const double* params, size_t n_params,
                    std::string& err) {
    Lowerer lw;
    lw.n = P.n_qubits;
    lw.mode = P.mode;
    lw.nb = P.n_bits;
    lw.params = params;
    lw.n_params = n_params;
    if (!lw.lower(in, n_ops)) {
        err = lw.err;
        return lw.code;
    }
    lw.fuse();
    std::vector<LOp>& ops = lw.ops;
    const int nb = P.n_bits;

    // ---- product-state prep: leading single-variant M1 on untouched bits fold into |psi0> = (x) (alpha, beta)
    std::vector<cd> prep(2 * nb);
    for (int b = 0; b < nb; b++) {
        prep[2 * b] = 1;
        prep[2 * b + 1] = 0;
    }
    {
        std::vector<char> touched(nb, 0);
        for (LOp& o : ops) {
            if (o.kind == LOp::SITE) {
                if (lw.sites[o.site].dependent) break;
                continue;
            }
            if (o.kind == LOp::M1 && o.nvar == 1 && !touched[o.q[0]]) {
                prep[2 * o.q[0]] = o.mats[0];
                prep[2 * o.q[0] + 1] = o.mats[2];
                touched[o.q[0]] = 1;
                o.dead = true;
                continue;
            }
            if (o.kind == LOp::SIGN) {
                for (int b = 0; b < nb; b++)
                    if ((o.mask >> b) & 1) touched[b] = 1;
            } else {
                for (int i = 0; i < o.nq; i++) touched[o.q[i]] = 1;
            }
        }
    }

    // ---- segmentation into sweeps
    struct SweepB {
        std::vector<int> ops;  // indices into `ops`
        std::vector<int> sites;
        std::vector<int> site_shift;
        int q[2] = {-1, -1};   // qubits whose joint populations the sweep's decisions read
        int bits = 0;          // branch-word bits used
        bool run_open = true;  // nothing but site matrices / diagonals so far: a state-dependent site may join
        bool all_mono = true;  // every site matrix so far maps populations to populations
        std::set<int> active;
        int red_q[2] = {-1, -1};
    };
    std::vector<SweepB> sw(1);
    auto fits_bits = [&](const SweepB& s, const int* qs, int nq) {
        // bits that are already active were accepted together before: nothing new to place in the tile
        bool all_known = !s.active.empty();
        for (int i = 0; i < nq && all_known; i++) all_known = s.active.count(qs[i]) != 0;
        if (all_known) return true;
        std::set<int> a = s.active;
        for (int i = 0; i < nq; i++) a.insert(qs[i]);
        Tile t;
        return choose_tile(a, nb, P.k_max, t);
    };
    std::map<int, int> shift_of_site;
    for (size_t i = 0; i < ops.size(); i++) {
        LOp& o = ops[i];
        if (o.dead) continue;
        SweepB* cur = &sw.back();
        if (o.kind == LOp::SITE) {
            LSite& s = lw.sites[o.site];
            int b = 0;
            while ((1 << b) < s.nbranch) b++;
            bool join = cur->bits + b <= MAX_BRANCH_BITS && fits_bits(*cur, s.q, s.nq);
            if (s.dependent) {
                // its branch probabilities are functions of the joint populations of (q0, q1) at the START of the
                // sweep (reduced by the previous sweep), propagated through the monomial site matrices before it
                int nq2[2] = {cur->q[0], cur->q[1]};
                bool okq = true;
                for (int h = 0; h < s.nq; h++) {
                    int qq = s.q[h];
                    if (nq2[0] == qq || nq2[1] == qq) continue;
                    if (nq2[0] < 0) nq2[0] = qq;
                    else if (nq2[1] < 0) nq2[1] = qq;
                    else okq = false;
                }
                join = join && okq && cur->run_open && cur->all_mono && sw.size() > 1;
                if (!join) {
                    sw.emplace_back();
                    cur = &sw.back();
                    nq2[0] = s.q[0];
                    nq2[1] = s.nq > 1 ? s.q[1] : -1;
                }
                cur->q[0] = nq2[0];
                cur->q[1] = nq2[1];
                SweepB& prev = sw[sw.size() - 2];
                prev.red_q[0] = cur->q[0];
                prev.red_q[1] = cur->q[1];
            } else if (!join) {
                sw.emplace_back();
                cur = &sw.back();
            }
            cur->sites.push_back(o.site);
            cur->site_shift.push_back(cur->bits);
            shift_of_site[o.site] = cur->bits;
            cur->bits += b;
            cur->all_mono = cur->all_mono && s.monomial;
            continue;
        }
        bool variant_of_group = o.nvar > 1;
        bool diagonal = o.kind == LOp::SIGN || o.kind == LOp::D1;
        if (!diagonal && !fits_bits(*cur, o.q, o.nq)) {
            if (variant_of_group) {  // cannot happen: the SITE marker checked the fit
                err = "internal: site matrices do not fit the tile of their sweep";
                return QCD_E_UNSUPPORTED;
            }
            sw.emplace_back();
            cur = &sw.back();
        }
        if (!variant_of_group && !diagonal) cur->run_open = false;
        cur->ops.push_back((int)i);
        active_bits(o, cur->active);
    }
    {
        std::map<int, int> sweep_of_site;
        for (size_t s = 0; s < sw.size(); s++)
            for (int si : sw[s].sites) sweep_of_site[si] = (int)s;
        for (size_t s = 0; s < sw.size(); s++)
            for (int oi : sw[s].ops)
                if (ops[oi].nvar > 1 && sweep_of_site[ops[oi].site] != (int)s) {
                    err = "internal: variant op separated from its site";
                    return QCD_E_UNSUPPORTED;
                }
    }

    // ---- emit
    CircuitInfo ci;
    ci.sweep_begin = (uint32_t)P.sweeps.size();
    ci.n_sites = (uint32_t)lw.sites.size();
    ci.lead = (uint32_t)P.circuits.size();
    // array sizes at the start of this circuit and where its final sweep's own passes begin (prefix signature)
    struct Mark {
        size_t sweeps, passes, kops, pool, masks, groups, subs, gdata;
    };
    auto mark = [&]() {
        return Mark{P.sweeps.size(), P.passes.size(), P.kops.size(), P.pool.size(), P.masks.size(), P.groups.size(),
                    P.subs.size(), P.gdata.size()};
    };
    const Mark m0 = mark();
    Mark m1 = m0;
    int32_t final_group_rel = -1;
    uint32_t prep_off = (uint32_t)P.pool.size();
    P.pool.insert(P.pool.end(), prep.begin(), prep.end());
    // ---- late peephole inside each sweep: a per-branch 1q matrix and a plain 1q matrix on the same bit, with nothing
    // touching that bit in between, become one per-branch matrix (e.g. Kraus variant followed by the H that closes the
    // CX).  Safe here and not in fuse(): segmentation is done and the decide step never looks at these matrices.
    for (SweepB& s : sw) {
        for (size_t j = 0; j < s.ops.size(); j++) {
            LOp& x = ops[s.ops[j]];
            if (x.dead || x.kind != LOp::M1) continue;
            int prev = -1;
            for (int i = (int)j - 1; i >= 0; i--) {
                const LOp& o = ops[s.ops[i]];
                if (!o.dead && touches(o, x.q[0])) {
                    prev = i;
                    break;
                }
            }
            if (prev < 0) continue;
            LOp& p = ops[s.ops[prev]];
            if (p.kind != LOp::M1 || (p.nvar > 1 && x.nvar > 1)) continue;
            const int nv = std::max(p.nvar, x.nvar);
            std::vector<cd> m;
            for (int v = 0; v < nv; v++) {
                std::vector<cd> xv(x.mats.begin() + (x.nvar == 1 ? 0 : v * 4), x.mats.begin() + (x.nvar == 1 ? 0 : v * 4) + 4);
                std::vector<cd> pv(p.mats.begin() + (p.nvar == 1 ? 0 : v * 4), p.mats.begin() + (p.nvar == 1 ? 0 : v * 4) + 4);
                std::vector<cd> r = matmul(xv, pv, 2);   // x after p
                m.insert(m.end(), r.begin(), r.end());
            }
            // keep the merged op at the LATER position (everything between commutes with it: it does not touch the bit)
            if (p.nvar > 1) {
                x.site = p.site;
                x.fshift = p.fshift;
                x.fmask = p.fmask;
            }
            x.nvar = nv;
            x.mats = m;
            p.dead = true;
            if (nv == 1 && is_identity(x.mats, 2)) x.dead = true;
        }
        std::vector<int> live;
        for (int oi : s.ops)
            if (!ops[oi].dead) live.push_back(oi);
        s.ops.swap(live);
    }

    for (size_t si = 0; si < sw.size(); si++) {
        SweepB& s = sw[si];
        Tile t;
        // the qubits whose populations this sweep reduces are welcome in the tile (they can then become register
        // bits of a single-pass sweep: cheapest reduction); never at the price of not fitting
        std::set<int> wanted = s.active;
        for (int h = 0; h < 2; h++)
            if (s.red_q[h] >= 0) wanted.insert(s.red_q[h]);
        if (!choose_tile(wanted, nb, P.k_max, t) && !choose_tile(s.active, nb, P.k_max, t)) {
            err = "an operator does not fit a " + std::to_string(P.k_max) + "-bit tile (tile_bits too small)";
            return QCD_E_UNSUPPORTED;
        }
        Sweep S;
        memset(&S, 0, sizeof(S));
        S.n_bits = (uint8_t)nb;
        S.k = (uint8_t)t.k;
        S.L = (uint8_t)t.L;
        S.n_high = (uint8_t)t.n_high;
        for (int i = 0; i < t.n_high; i++) S.high[i] = (uint8_t)t.high[i];
        S.flags = (si == 0 ? SW_PREP : 0) | (si + 1 == sw.size() ? SW_FINAL : 0);
        S.red_q[0] = (int8_t)s.red_q[0];
        S.red_q[1] = (int8_t)s.red_q[1];
        S.prep_off = prep_off;
        S.group = -1;
        if (!s.sites.empty()) {
            Group G;
            memset(&G, 0, sizeof(G));
            G.q[0] = (int8_t)s.q[0];
            G.q[1] = (int8_t)s.q[1];
            G.total_bits = (uint8_t)s.bits;
            G.sub_begin = (uint32_t)P.subs.size();
            for (size_t u = 0; u < s.sites.size(); u++) {
                LSite& ls = lw.sites[s.sites[u]];
                SubSite ss;
                memset(&ss, 0, sizeof(ss));
                ss.nbranch = (uint8_t)ls.nbranch;
                ss.shift = (uint8_t)s.site_shift[u];
                int b = 0;
                while ((1 << b) < ls.nbranch) b++;
                ss.nbits = (uint8_t)b;
                ss.unit = ls.unit ? 1 : 0;
                ss.stream = ls.stream;
                ss.data_off = (uint32_t)P.gdata.size();
                // express w (over the site's own qubits) and T on the group's joint populations P[b(q0)+2 b(q1)]
                int slot[2] = {-1, -1};
                for (int h = 0; h < ls.nq; h++) {
                    if (ls.q[h] == s.q[0]) slot[h] = 0;
                    else if (ls.q[h] == s.q[1] && s.q[1] >= 0) slot[h] = 1;
                }
                int d = 1 << ls.nq;
                for (int j = 0; j < ls.nbranch; j++) {
                    double w4[4], T16[16];
                    for (int x = 0; x < 4; x++) {
                        // site-local index of joint population x (bits outside Q: weights must not depend on them)
                        int loc = 0;
                        for (int h = 0; h < ls.nq; h++)
                            if (slot[h] >= 0 && ((x >> slot[h]) & 1)) loc |= 1 << h;
                        w4[x] = ls.w[j * d + loc];
                        for (int y = 0; y < 4; y++) {
                            // T[x][y]: population flow y -> x ; bits of Q not acted on must match
                            int locy = 0;
                            bool same_rest = true;
                            for (int sb = 0; sb < 2; sb++) {
                                bool acted = (slot[0] == sb) || (ls.nq > 1 && slot[1] == sb);
                                if (!acted && (((x >> sb) & 1) != ((y >> sb) & 1))) same_rest = false;
                            }
                            for (int h = 0; h < ls.nq; h++)
                                if (slot[h] >= 0 && ((y >> slot[h]) & 1)) locy |= 1 << h;
                            double v = 0;
                            if (same_rest) {
                                // marginalise site qubits outside Q (only valid when the flow does not depend on
                                // them: true for independent sites; dependent sites have all qubits in Q)
                                int free_mask = 0;
                                for (int h = 0; h < ls.nq; h++)
                                    if (slot[h] < 0) free_mask |= 1 << h;
                                if (free_mask == 0) {
                                    v = ls.T[(j * d + loc) * d + locy];
                                } else {
                                    // sum over outputs of the free qubits for a fixed free input (0): column sums
                                    for (int fo = 0; fo < d; fo++) {
                                        if ((fo & ~free_mask) != 0) continue;
                                        v += ls.T[(j * d + (loc | fo)) * d + locy];
                                    }
                                }
                            }
                            T16[x * 4 + y] = v;
                        }
                    }
                    P.gdata.insert(P.gdata.end(), w4, w4 + 4);
                    P.gdata.insert(P.gdata.end(), T16, T16 + 16);
                }
                P.subs.push_back(ss);
            }
            G.sub_end = (uint32_t)P.subs.size();
            S.group = (int32_t)P.groups.size();
            P.groups.push_back(G);
            P.max_group_bits = std::max(P.max_group_bits, s.bits);
        }
        if (si + 1 == sw.size()) {   // everything emitted so far (incl. the final sweep's site group) is shared prefix
            m1 = mark();
            final_group_rel = S.group >= 0 ? S.group - (int32_t)m0.groups : -1;
        }
        // passes
        S.pass_begin = (uint32_t)P.passes.size();
        struct PB {
            std::set<int> R;
            std::vector<int> ops;
        };
        std::vector<PB> pbs;
        int nreg = std::min(REG_BITS, t.k);
        for (int oi : s.ops) {
            LOp& o = ops[oi];
            std::set<int> B;
            if (o.kind == LOp::M1 || o.kind == LOp::M2 || o.kind == LOp::M4)
                for (int i = 0; i < o.nq; i++) B.insert(local_bit(t, o.q[i]));
            if (B.count(-1)) {
                err = "internal: op bit not in tile";
                return QCD_E_UNSUPPORTED;
            }
            if ((int)B.size() > nreg) {
                err = "state too small for a " + std::to_string(B.size()) + "-bit operator";
                return QCD_E_UNSUPPORTED;
            }
            bool open_new = pbs.empty();
            if (!open_new) {
                std::set<int> U = pbs.back().R;
                U.insert(B.begin(), B.end());
                if ((int)U.size() > nreg) open_new = true;
            }
            if (open_new) pbs.emplace_back();
            pbs.back().R.insert(B.begin(), B.end());
            pbs.back().ops.push_back(oi);
        }
        for (PB& pb : pbs) {
            if (pbs.size() == 1 && si != 0) {
                // single-pass sweep -> direct (register-only) kernel: pad with the LOWEST free bits, so that a thread's
                // 16 amplitudes contain contiguous runs (128-bit accesses) and whole probability chunks
                if ((int)pb.R.size() < nreg) pb.R.insert(0);   // index bit 0 first: 128-bit accesses
                for (int h = 0; h < 2; h++) {
                    const int lb = s.red_q[h] >= 0 ? local_bit(t, s.red_q[h]) : -1;
                    if (lb >= 0 && (int)pb.R.size() < nreg) pb.R.insert(lb);
                }
                for (int b = 0; b < t.k && (int)pb.R.size() < nreg; b++)
                    if (!pb.R.count(b)) pb.R.insert(b);
            }
            // otherwise pad with the highest free local bits (keeps lane-contiguous smem accesses)
            for (int b = t.k - 1; b >= 0 && (int)pb.R.size() < nreg; b--)
                if (!pb.R.count(b)) pb.R.insert(b);
            Pass ps;
            memset(&ps, 0, sizeof(ps));
            ps.nreg = (uint8_t)nreg;
            std::vector<int> R(pb.R.begin(), pb.R.end());
            for (int i = 0; i < nreg; i++) ps.rbits[i] = (uint8_t)R[i];
            auto reg_of = [&](int g) {
                int lb = local_bit(t, g);
                return (int)(std::find(R.begin(), R.end(), lb) - R.begin());
            };
            ps.op_begin = (uint32_t)P.kops.size();
            // merge runs of SIGN ops into one K_SIGNS with duplicate masks cancelled
            for (size_t x = 0; x < pb.ops.size(); x++) {
                LOp& o = ops[pb.ops[x]];
                KOp ko;
                memset(&ko, 0, sizeof(ko));
                if (o.kind == LOp::SIGN) {
                    std::map<uint64_t, int> cnt;
                    size_t y = x;
                    while (y < pb.ops.size() && ops[pb.ops[y]].kind == LOp::SIGN) cnt[ops[pb.ops[y++]].mask] ^= 1;
                    x = y - 1;
                    ko.kind = K_SIGNS;
                    ko.off = (uint32_t)P.masks.size();
                    for (auto& kv : cnt)
                        if (kv.second) P.masks.push_back(kv.first);
                    ko.count = (uint16_t)(P.masks.size() - ko.off);
                    if (ko.count == 0) continue;
                    P.kops.push_back(ko);
                    continue;
                }
                if (o.kind == LOp::D1) {
                    ko.kind = K_D1;
                    ko.gbit = (uint8_t)o.q[0];
                    ko.off = (uint32_t)P.pool.size();
                    P.pool.insert(P.pool.end(), o.mats.begin(), o.mats.end());
                    P.kops.push_back(ko);
                    continue;
                }
                int dim = 1 << o.nq;
                int regs[4];
                for (int i = 0; i < o.nq; i++) regs[i] = reg_of(o.q[i]);
                // order operator bits by ascending register index
                int order[4] = {0, 1, 2, 3};
                std::sort(order, order + o.nq, [&](int a, int b) { return regs[a] < regs[b]; });
                // new slot s holds old slot order[s]
                ko.kind = o.kind == LOp::M1 ? K_M1 : (o.kind == LOp::M2 ? K_M2 : K_M4);
                ko.t0 = (uint8_t)regs[order[0]];
                ko.t1 = o.nq > 1 ? (uint8_t)regs[order[1]] : 0;
                if (o.kind == LOp::M4) {
                    // M4 indexes by register index directly: requires its 4 bits == R
                    for (int i = 0; i < 4; i++)
                        if (regs[order[i]] != i) {
                            err = "internal: M4 bits do not match the pass registers";
                            return QCD_E_UNSUPPORTED;
                        }
                }
                ko.off = (uint32_t)P.pool.size();
                ko.stride = (uint32_t)(dim * dim);
                if (o.nvar > 1) {
                    ko.sel_shift = (uint8_t)(shift_of_site[o.site] + o.fshift);
                    ko.sel_mask = (uint8_t)o.fmask;
                }
                for (int v = 0; v < o.nvar; v++) {
                    std::vector<cd> m(o.mats.begin() + v * dim * dim, o.mats.begin() + (v + 1) * dim * dim);
                    if (o.nq > 1) m = permute_bits(m, order, o.nq);
                    P.pool.insert(P.pool.end(), m.begin(), m.end());
                }
                // pad variants up to the field size so that any field value addresses valid memory
                if (o.nvar > 1)
                    for (int v = o.nvar; v <= o.fmask; v++)
                        P.pool.insert(P.pool.end(), dim * dim, cd(0, 0));
                P.kops.push_back(ko);
            }
            ps.op_end = (uint32_t)P.kops.size();
            P.passes.push_back(ps);
        }
        S.pass_end = (uint32_t)P.passes.size();
        P.sweeps.push_back(S);
    }
    ci.sweep_end = (uint32_t)P.sweeps.size();
    P.max_sweeps = std::max(P.max_sweeps, (int)(ci.sweep_end - ci.sweep_begin));
    P.circuits.push_back(ci);
    // ---- prefix signature: two independent 64-bit hashes over the prefix content with circuit-relative offsets
    {
        struct Hasher {
            uint64_t a = 0xcbf29ce484222325ull, b = 0x9e3779b97f4a7c15ull;
            void bytes(const void* p, size_t n) {
                const unsigned char* c = static_cast<const unsigned char*>(p);
                for (size_t i = 0; i < n; i++) {
                    a = (a ^ c[i]) * 0x100000001b3ull;
                    b = (b + c[i] + 0x632be59bd9b4e019ull) * 0xff51afd7ed558ccdull;
                    b ^= b >> 29;
                }
            }
        } H;
        auto pod = [&H](const auto& v) { H.bytes(&v, sizeof(v)); };
        for (size_t i = m0.sweeps; i < m1.sweeps; i++) {
            Sweep S = P.sweeps[i];
            S.pass_begin -= (uint32_t)m0.passes;
            S.pass_end -= (uint32_t)m0.passes;
            S.prep_off -= (uint32_t)m0.pool;
            if (S.group >= 0) S.group -= (int32_t)m0.groups;
            pod(S);
        }
        for (size_t i = m0.passes; i < m1.passes; i++) {
            Pass ps = P.passes[i];
            ps.op_begin -= (uint32_t)m0.kops;
            ps.op_end -= (uint32_t)m0.kops;
            pod(ps);
        }
        for (size_t i = m0.kops; i < m1.kops; i++) {
            KOp k = P.kops[i];
            k.off -= (uint32_t)(k.kind == K_SIGNS ? m0.masks : m0.pool);
            pod(k);
        }
        if (m1.pool > m0.pool) H.bytes(&P.pool[m0.pool], (m1.pool - m0.pool) * sizeof(cd));
        if (m1.masks > m0.masks) H.bytes(&P.masks[m0.masks], (m1.masks - m0.masks) * sizeof(uint64_t));
        for (size_t i = m0.groups; i < m1.groups; i++) {
            Group g = P.groups[i];
            g.sub_begin -= (uint32_t)m0.subs;
            g.sub_end -= (uint32_t)m0.subs;
            pod(g);
        }
        for (size_t i = m0.subs; i < m1.subs; i++) {
            SubSite ss = P.subs[i];
            ss.data_off -= (uint32_t)m0.gdata;
            pod(ss);
        }
        if (m1.gdata > m0.gdata) H.bytes(&P.gdata[m0.gdata], (m1.gdata - m0.gdata) * sizeof(double));
        pod(final_group_rel);
        PrefixSig sig;
        sig.h0 = H.a;
        sig.h1 = H.b;
        sig.n_sweeps = (uint32_t)(m1.sweeps - m0.sweeps);
        P.prefix.push_back(sig);
    }
    return QCD_OK;
}

void merge_program(Program& dst, const Program& src, long long pool_base) {
    const uint32_t sweep0 = (uint32_t)dst.sweeps.size(), pass0 = (uint32_t)dst.passes.size();
    const uint32_t kop0 = (uint32_t)dst.kops.size();
    const uint32_t pool0 = pool_base >= 0 ? (uint32_t)pool_base : (uint32_t)dst.pool.size();
    const uint32_t mask0 = (uint32_t)dst.masks.size(), group0 = (uint32_t)dst.groups.size();
    const uint32_t sub0 = (uint32_t)dst.subs.size(), gdata0 = (uint32_t)dst.gdata.size();
    for (Sweep s : src.sweeps) {
        s.pass_begin += pass0;
        s.pass_end += pass0;
        s.prep_off += pool0;
        if (s.group >= 0) s.group += (int32_t)group0;
        dst.sweeps.push_back(s);
    }
    for (Pass p : src.passes) {
        p.op_begin += kop0;
        p.op_end += kop0;
        dst.passes.push_back(p);
    }
    for (KOp k : src.kops) {
        k.off += k.kind == K_SIGNS ? mask0 : pool0;
        dst.kops.push_back(k);
    }
    if (pool_base < 0) dst.pool.insert(dst.pool.end(), src.pool.begin(), src.pool.end());
    dst.masks.insert(dst.masks.end(), src.masks.begin(), src.masks.end());
    for (Group g : src.groups) {
        g.sub_begin += sub0;
        g.sub_end += sub0;
        dst.groups.push_back(g);
    }
    for (SubSite ss : src.subs) {
        ss.data_off += gdata0;
        dst.subs.push_back(ss);
    }
    dst.gdata.insert(dst.gdata.end(), src.gdata.begin(), src.gdata.end());
    const uint32_t circ0 = (uint32_t)dst.circuits.size();
    for (CircuitInfo ci : src.circuits) {
        ci.sweep_begin += sweep0;
        ci.sweep_end += sweep0;
        ci.lead += circ0;
        dst.circuits.push_back(ci);
    }
    dst.prefix.insert(dst.prefix.end(), src.prefix.begin(), src.prefix.end());
    dst.max_group_bits = std::max(dst.max_group_bits, src.max_group_bits);
    dst.max_sweeps = std::max(dst.max_sweeps, src.max_sweeps);
}

std::string program_to_json(const Program& P, int c) {
    std::ostringstream o;
    o.precision(17);
    const CircuitInfo& ci = P.circuits[c];
    o << "{\"n_qubits\":" << P.n_qubits << ",\"n_bits\":" << P.n_bits << ",\"mode\":" << P.mode
      << ",\"n_sites\":" << ci.n_sites;
    if ((size_t)c < P.prefix.size())
        o << ",\"prefix_sig\":[\"" << std::hex << P.prefix[c].h0 << "\",\"" << P.prefix[c].h1 << "\"," << std::dec
          << P.prefix[c].n_sweeps << "]";
    o << ",\"sweeps\":[";
    auto cplx = [&](cd v) { o << "[" << v.real() << "," << v.imag() << "]"; };
    for (uint32_t si = ci.sweep_begin; si < ci.sweep_end; si++) {
        const Sweep& S = P.sweeps[si];
        if (si > ci.sweep_begin) o << ",";
        o << "{\"k\":" << (int)S.k << ",\"L\":" << (int)S.L << ",\"high\":[";
        for (int i = 0; i < S.n_high; i++) o << (i ? "," : "") << (int)S.high[i];
        o << "],\"flags\":" << (int)S.flags << ",\"red_q\":[" << (int)S.red_q[0] << "," << (int)S.red_q[1] << "]";
        if (S.flags & SW_PREP) {
            o << ",\"prep\":[";
            for (int b = 0; b < 2 * P.n_bits; b++) {
                if (b) o << ",";
                cplx(P.pool[S.prep_off + b]);
            }
            o << "]";
        }
        if (S.group >= 0) {
            const Group& G = P.groups[S.group];
            o << ",\"group\":{\"q\":[" << (int)G.q[0] << "," << (int)G.q[1] << "],\"bits\":" << (int)G.total_bits
              << ",\"subs\":[";
            for (uint32_t u = G.sub_begin; u < G.sub_end; u++) {
                const SubSite& ss = P.subs[u];
                if (u > G.sub_begin) o << ",";
                o << "{\"nbranch\":" << (int)ss.nbranch << ",\"shift\":" << (int)ss.shift << ",\"stream\":" << ss.stream
                  << ",\"unit\":" << (int)ss.unit << ",\"data\":[";
                for (int j = 0; j < ss.nbranch * 20; j++) o << (j ? "," : "") << P.gdata[ss.data_off + j];
                o << "]}";
            }
            o << "]}";
        }
        o << ",\"passes\":[";
        for (uint32_t pi = S.pass_begin; pi < S.pass_end; pi++) {
            const Pass& ps = P.passes[pi];
            if (pi > S.pass_begin) o << ",";
            o << "{\"rbits\":[";
            for (int i = 0; i < ps.nreg; i++) o << (i ? "," : "") << (int)ps.rbits[i];
            o << "],\"ops\":[";
            for (uint32_t oi = ps.op_begin; oi < ps.op_end; oi++) {
                const KOp& k = P.kops[oi];
                if (oi > ps.op_begin) o << ",";
                o << "{\"kind\":" << (int)k.kind << ",\"t0\":" << (int)k.t0 << ",\"t1\":" << (int)k.t1
                  << ",\"gbit\":" << (int)k.gbit << ",\"sel_shift\":" << (int)k.sel_shift << ",\"sel_mask\":"
                  << (int)k.sel_mask;
                if (k.kind == K_SIGNS) {
                    o << ",\"masks\":[";
                    for (int m = 0; m < k.count; m++) o << (m ? "," : "") << P.masks[k.off + m];
                    o << "]";
                } else {
                    int dim = k.kind == K_M1 ? 2 : (k.kind == K_M2 ? 4 : (k.kind == K_M4 ? 16 : 1));
                    int per = k.kind == K_D1 ? 2 : dim * dim;
                    int nv = k.kind == K_D1 ? 1 : (int)k.sel_mask + 1;
                    o << ",\"nvar\":" << nv << ",\"mats\":[";
                    for (int e = 0; e < per * nv; e++) {
                        if (e) o << ",";
                        cplx(P.pool[k.off + e]);
                    }
                    o << "]";
                }
                o << "}";
            }
            o << "]}";
        }
        o << "]}";
    }
    o << "]}";
    return o.str();
}

}  // namespace qcd
