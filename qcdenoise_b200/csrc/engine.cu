// engine.cu -- the C ABI of libqcdsim.so (include/qcdsim.h): contexts, program upload, and the host-side drivers
// that turn a compiled sweep program (lower.cpp) into kernel launches (sweep.cu, kernels.cu).
//
// Statevector trajectories are executed as a TREE: level d holds the distinct states reached after d sweeps; the
// trajectories of a node that draw the same branch word at the next site group share one child state.  Results are
// identical to one-state-per-shot execution (every draw is a pure function of (seed, circuit, shot, stream)), the
// number of state sweeps is not.  The tree is walked breadth-first while state slots last and depth-first beyond
// that (children are swept in chunks, each chunk is finished before the next), so any circuit depth runs in
// max_sweeps + 1 slots.
#include <cuda_runtime.h>
#include <sys/mman.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/qcdsim.h"
#include "dm_cluster.h"
#include "kernels.cuh"
#include "device_noise.h"
#include "lower.h"

using namespace qcd;

namespace {

thread_local std::string g_create_error;

// Large host arrays that are written once (the lowering threads' matrix pools): ask for transparent huge pages, the
// first-touch page faults of 4 KB pages otherwise cost more than computing the matrices.
void advise_huge(const void* p, size_t bytes) {
#ifdef MADV_HUGEPAGE
    const uintptr_t two_mb = (uintptr_t)2 << 20;
    const uintptr_t a = ((uintptr_t)p + two_mb - 1) & ~(two_mb - 1);
    const uintptr_t e = ((uintptr_t)p + bytes) & ~(two_mb - 1);
    if (e > a) madvise(reinterpret_cast<void*>(a), e - a, MADV_HUGEPAGE);   // best effort
#else
    (void)p;
    (void)bytes;
#endif
}

struct Trace {
    bool on = getenv("QCD_TRACE") != nullptr;
    double t_sync = 0, t_h2d = 0, t_total = 0, t_prep = 0, t_launch = 0, t_leaves = 0, t_run0 = 0;
    int n_sync = 0;
    static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
} g_trace;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes, bool zero = false) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = std::max(bytes, (size_t)256);
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            p = nullptr;
            return e;
        }
        cap = want;
        if (zero) e = cudaMemset(p, 0, want);
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename U> U* as() const { return reinterpret_cast<U*>(p); }
};

// page-locked host staging (node records travel host <-> device once per tree level)
struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = std::max(bytes + bytes / 2, (size_t)4096);
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e != cudaSuccess) {
            p = nullptr;
            return e;
        }
        cap = want;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <typename U> U* as() const { return reinterpret_cast<U*>(p); }
};

struct DevProg {
    bool valid = false;
    bool host_ready = false;   // P was lowered at upload time (device copies still to be made)
    Program P;
    DevBuf sweeps, passes, kops, pool, masks, groups, subs, gdata, circuits, direct;
    std::vector<DirectSweep> direct_host;
    std::vector<uint8_t> direct_flags;   // per descriptor: DF_* (what a launch holding this sweep must provide)
    int member_bits = 0;                 // ceil(log2(longest run of circuits sharing a prefix)), statevector mode
    // Sibling groups (statevector mode).  The final sweeps of a run of circuits that share a prefix usually differ only in a
    // trailing real 2x2 on ONE qubit k (stabilizers.py:136-156: the basis change of a Toth generator is one Hadamard).  The
    // leaves (parent, branch word, member) of such a run then share everything but that last op: see SW_CROSS.
    struct CrossRun {
        uint32_t lead = 0, n_members = 0;
        std::vector<int16_t> kbit;       // per member: >= CHUNK_LOG: the trailing op's qubit; -1: chunk sums = C0; -2: not eligible
        std::vector<double> u;           // per member: the trailing real 2x2 (row-major), 4 doubles
        uint32_t desc = 0;               // SW_CROSS descriptor: the shared ops; register bits = their qubits + the lowest others
        uint32_t reg_mask = 0;           // its register bits
        int n_high = 0;                  // ... of which this many are >= CHUNK_LOG (the top register indices)
        uint8_t high_bit[2] = {0, 0};    // ... namely these
    };
    std::vector<CrossRun> cross_runs;
    std::vector<int32_t> cross_of_circuit;   // per circuit: index into cross_runs, -1: none
    // A new upload invalidates the compiled program but keeps the device buffers: the next program of similar size
    // reuses them (cudaFree synchronises the whole device and cudaMalloc of hundreds of MB is not free either).
    void invalidate() {
        cross_runs.clear();
        cross_of_circuit.clear();
        direct_host.clear();
        direct_flags.clear();
        valid = false;
        host_ready = false;
        P = Program();
    }
    void release() {
        direct.release();
        cross_runs.clear();
        cross_of_circuit.clear();
        direct_host.clear();
        direct_flags.clear();
        sweeps.release(); passes.release(); kops.release(); pool.release(); masks.release();
        groups.release(); subs.release(); gdata.release(); circuits.release();
        valid = false;
        host_ready = false;
        P = Program();
    }
};

enum { DF_OK = 1, DF_VEC = 2, DF_RED = 4, DF_FINAL = 8, DF_PREP = 16 };

// Descriptor of pass `pi` of sweep S for the register-only kernel (D.ok stays 0 when the pass does not qualify).
void describe_pass(const Program& P, DirectSweep& D, const Sweep& S, uint32_t pi, uint8_t flags, bool reduce) {
    // (states below 13 index bits still get descriptors: the tile kernel stages its passes from them)
    const Pass& ps = P.passes[pi];
    const uint32_t nops = ps.op_end - ps.op_begin;
    if (ps.nreg != REG_BITS || nops > (uint32_t)MAX_DIRECT_OPS) return;
    uint32_t n_signs = 0, n_m4 = 0;
    for (uint32_t o = 0; o < nops; o++) {
        KOp op = P.kops[ps.op_begin + o];
        if (op.kind == K_SIGNS) {
            if (n_signs + op.count > 48) return;
            op.t1 = (uint8_t)n_signs;
            n_signs += op.count;
        }
        if (op.kind == K_M4 && (++n_m4 > 1 || op.sel_mask != 0)) return;  // one plain 16x16 superoperator per pass
        D.ops[o] = op;
    }
    D.ok = 1;
    D.n_bits = S.n_bits;
    D.nreg = ps.nreg;
    D.nops = (uint8_t)nops;
    for (int i = 0; i < REG_BITS; i++) {
        const int b = ps.rbits[i];
        D.rbit[i] = (uint8_t)(b < S.L ? b : S.high[b - S.L]);
    }
    D.red_q[0] = reduce ? S.red_q[0] : -1;
    D.red_q[1] = reduce ? S.red_q[1] : -1;
    D.flags = flags;
    D.prep_off = S.prep_off;
}

// Runs of consecutive circuits whose sweeps are identical up to the final one (a graph circuit followed by its stabilizer
// settings): one trajectory tree serves the whole run down to the parents of the leaves.  -> member bits.
int assign_runs(Program& P, int mode) {
    size_t longest = 1;
    for (size_t c = 0; c < P.circuits.size(); c++) {
        uint32_t lead = (uint32_t)c;
        if (mode == MODE_SV && c > 0 && P.prefix[c] == P.prefix[c - 1] &&
            c - P.circuits[c - 1].lead < ((size_t)1 << MAX_MEMBER_BITS))
            lead = P.circuits[c - 1].lead;
        P.circuits[c].lead = lead;
        longest = std::max(longest, c - lead + 1);
    }
    int bits = 0;
    while (((size_t)1 << bits) < longest) bits++;
    return bits;
}

// Sibling-group descriptors of every run of circuits that share a prefix (see DevProg::CrossRun): appended to
// dp.direct_host behind the per-sweep and per-pass descriptors.
template <typename PoolAt>
void build_cross(DevProg& dp, PoolAt pool_at /* pool offset -> std::complex<double>, in device precision */) {
    const Program& P = dp.P;
    dp.cross_runs.clear();
    dp.cross_of_circuit.assign(P.circuits.size(), -1);
    if (P.mode != MODE_SV || dp.member_bits == 0 || P.n_bits < 10) return;
    const int chunk_log = CHUNK_LOG;
    auto final_desc = [&](size_t c) -> const DirectSweep* {
        const CircuitInfo& ci = P.circuits[c];
        if (ci.sweep_end <= ci.sweep_begin) return nullptr;
        const DirectSweep& D = dp.direct_host[ci.sweep_end - 1];
        if (!D.ok || !(D.flags & SW_FINAL) || (D.flags & SW_PREP)) return nullptr;
        return &D;
    };
    auto n_elems = [](int kind) { return kind == K_M1 ? 4u : (kind == K_M2 ? 16u : (kind == K_D1 ? 2u : 0u)); };
    // the same op (kind, global qubits, variant selection, matrix / mask contents) in two descriptors
    auto same_op = [&](const DirectSweep& A, const KOp& a, const DirectSweep& B, const KOp& b) {
        if (a.kind != b.kind || a.sel_shift != b.sel_shift || a.sel_mask != b.sel_mask) return false;
        if (a.kind == K_SIGNS) {
            if (a.count != b.count) return false;
            for (uint32_t e = 0; e < a.count; e++)
                if (P.masks[a.off + e] != P.masks[b.off + e]) return false;
            return true;
        }
        if (a.kind == K_M4) return false;
        if (a.kind == K_D1) {
            if (a.gbit != b.gbit) return false;
        } else {
            if (A.rbit[a.t0] != B.rbit[b.t0]) return false;
            if (a.kind == K_M2 && A.rbit[a.t1] != B.rbit[b.t1]) return false;
        }
        const uint32_t ne = n_elems(a.kind);
        for (uint32_t v = 0; v <= a.sel_mask; v++)
            for (uint32_t e = 0; e < ne; e++)
                if (pool_at(a.off + v * a.stride + e) != pool_at(b.off + v * b.stride + e)) return false;
        return true;
    };
    for (size_t c0 = 0; c0 < P.circuits.size();) {
        size_t c1 = c0 + 1;
        while (c1 < P.circuits.size() && P.circuits[c1].lead == (uint32_t)c0) c1++;
        const size_t m = c1 - c0;
        const size_t run_begin = c0;
        c0 = c1;
        if (m < 2) continue;
        // base = a member with the fewest ops (the dataset circuit / the identity setting: no basis change).  A setting
        // whose Hadamard was fused into the site's own matrices has as few ops but other contents: of the candidates,
        // take the one most members extend.
        const DirectSweep* base = nullptr;
        {
            int min_ops = 1 << 30;
            for (size_t c = run_begin; c < c1; c++)
                if (const DirectSweep* D = final_desc(c)) min_ops = std::min<int>(min_ops, D->nops);
            size_t best = 0;
            for (size_t c = run_begin; c < c1; c++) {
                const DirectSweep* B = final_desc(c);
                if (!B || B->nops != min_ops) continue;
                size_t cnt = 0;
                for (size_t c2 = run_begin; c2 < c1; c2++) {
                    const DirectSweep* D = final_desc(c2);
                    if (!D || D->nops < B->nops || D->nops > B->nops + 1) continue;
                    bool same = true;
                    for (int o = 0; o < B->nops && same; o++) same = same_op(*D, D->ops[o], *B, B->ops[o]);
                    cnt += same ? 1 : 0;
                }
                if (cnt > best) {
                    best = cnt;
                    base = B;
                }
                if (best * 2 > m) break;      // a majority: no other candidate can beat it
            }
        }
        if (g_trace.on) fprintf(stderr, "[qcd trace] sibling run candidate: circuits %zu..%zu, base %s\n", run_begin, c1 - 1, base ? "found" : "none");
        if (!base) continue;
        bool ok = true;
        uint32_t r0_mask = 0;
        for (int o = 0; o < base->nops && ok; o++) {
            const KOp& op = base->ops[o];
            if (op.kind == K_M4) ok = false;
            if (op.kind == K_M1 || op.kind == K_M2) r0_mask |= 1u << base->rbit[op.t0];
            if (op.kind == K_M2) r0_mask |= 1u << base->rbit[op.t1];
        }
        if (!ok || __builtin_popcount(r0_mask) > 2) {
            if (g_trace.on) fprintf(stderr, "[qcd trace] sibling run rejected: shared-op qubits 0x%x\n", r0_mask);
            continue;
        }
        {
            // the shared passes only implement the slotted shape of stage_sweep (direct_ops.cuh): real 2x2 ops on distinct
            // qubits, at most one sign layer, real 2x2 ops on distinct qubits -- what a noise site of real Kraus operators
            // followed by CZ gates lowers to
            uint32_t seen[2] = {0, 0};
            int phase = 0;
            for (int o = 0; o < base->nops && ok; o++) {
                const KOp& op = base->ops[o];
                if (op.kind == K_SIGNS) {
                    if (phase == 1) ok = false;
                    phase = 1;
                } else if (op.kind == K_M1) {
                    const uint32_t b = 1u << base->rbit[op.t0];
                    if (seen[phase] & b) ok = false;
                    seen[phase] |= b;
                    for (uint32_t v = 0; v <= op.sel_mask && ok; v++)
                        for (uint32_t e = 0; e < 4; e++)
                            if (pool_at(op.off + v * op.stride + e).imag() != 0.0) ok = false;
                } else {
                    ok = false;
                }
            }
            if (!ok) {
                if (g_trace.on) fprintf(stderr, "[qcd trace] sibling run rejected: shared ops are not of the slotted shape\n");
                continue;
            }
        }
        DevProg::CrossRun R;
        R.lead = (uint32_t)run_begin;
        R.n_members = (uint32_t)m;
        R.kbit.assign(m, -2);
        R.u.assign(4 * m, 0.0);
        uint32_t k_mask = 0;
        size_t n_ok = 0;
        // member j ends in the real 2x2 `u` on qubit k after the shared ops
        auto accept = [&](size_t j, int k, const double* u) {
            if (k < chunk_log) {
                // acts inside the 32-amplitude chunks: their sums are unchanged iff the 2x2 is norm preserving
                const double e0 = u[0] * u[0] + u[2] * u[2] - 1.0, e1 = u[1] * u[1] + u[3] * u[3] - 1.0,
                             e2 = u[0] * u[1] + u[2] * u[3];
                if (std::fabs(e0) > 1e-6 || std::fabs(e1) > 1e-6 || std::fabs(e2) > 1e-6) return false;   // device precision
                R.kbit[j] = -1;
            } else {
                if ((k_mask >> k) & 1u) return false;      // a second member on the same qubit: it keeps its own sweep
                k_mask |= 1u << k;
                R.kbit[j] = (int16_t)k;
                for (int e = 0; e < 4; e++) R.u[4 * j + e] = u[e];
            }
            return true;
        };
        for (size_t c = run_begin; c < c1; c++) {
            const DirectSweep* D = final_desc(c);
            if (!D || D->nops < base->nops || D->nops > base->nops + 1) continue;
            const size_t j = c - run_begin;
            int n_diff = 0, od = -1, om = -1;      // od / om: the differing op in the base / in the member
            for (int o = 0; o < base->nops; o++)
                if (!same_op(*D, D->ops[o], *base, base->ops[o])) {
                    n_diff++;
                    od = om = o;
                }
            if (n_diff > 1 && D->nops == base->nops) {
                // the same ops in another order (a fused op is emitted where its last factor stood): match them one to one
                uint32_t used = 0;
                int un_b = -1, un_m = -1, n_un = 0;
                for (int o = 0; o < base->nops; o++) {
                    int hit = -1;
                    for (int q = 0; q < D->nops && hit < 0; q++)
                        if (!((used >> q) & 1u) && same_op(*D, D->ops[q], *base, base->ops[o])) hit = q;
                    if (hit >= 0) used |= 1u << hit;
                    else {
                        un_b = o;
                        n_un++;
                    }
                }
                for (int q = 0; q < D->nops; q++)
                    if (!((used >> q) & 1u)) un_m = q;
                if (n_un == 1 && un_m >= 0) {      // (whether the odd op commutes past the others is checked below)
                    n_diff = 1;
                    od = un_b;
                    om = un_m;
                }
            }
            if (n_diff == 0 && D->nops == base->nops) {
                R.kbit[j] = -1;
                n_ok++;
                continue;
            }
            if (n_diff == 0) {      // one trailing op
                const KOp& t = D->ops[base->nops];
                if (t.kind != K_M1 || t.sel_mask != 0) continue;
                double u[4];
                bool real = true;
                for (int e = 0; e < 4; e++) {
                    u[e] = pool_at(t.off + e).real();
                    real = real && pool_at(t.off + e).imag() == 0.0;
                }
                if (real && accept(j, D->rbit[t.t0], u)) n_ok++;
                continue;
            }
            // The trailing 2x2 may have been fused into the shared op on its own qubit (no sign mask touches that qubit,
            // so it commuted there): the one differing op is then U x (the shared op), variant by variant.
            if (n_diff != 1 || D->nops != base->nops) continue;
            const KOp &a = D->ops[om], &b = base->ops[od];
            if (a.kind != K_M1 || b.kind != K_M1 || a.sel_shift != b.sel_shift || a.sel_mask != b.sel_mask ||
                D->rbit[a.t0] != base->rbit[b.t0])
                continue;
            const int k = D->rbit[a.t0];
            bool commutes = true;
            for (int o = 0; o < base->nops && commutes; o++) {
                const KOp& op = base->ops[o];
                if (op.kind == K_SIGNS)
                    for (uint32_t e = 0; e < op.count; e++)
                        if ((P.masks[op.off + e] >> k) & 1u) commutes = false;
                if (op.kind == K_D1 && op.gbit == k) commutes = false;
                if (o != od && op.kind == K_M1 && base->rbit[op.t0] == k) commutes = false;
            }
            if (!commutes) continue;
            double u[4] = {0, 0, 0, 0};
            bool have_u = false, fits = true;
            for (uint32_t v = 0; v <= a.sel_mask && fits; v++) {
                double M[4], N[4];
                for (int e = 0; e < 4; e++) {
                    M[e] = pool_at(b.off + v * b.stride + e).real();
                    N[e] = pool_at(a.off + v * a.stride + e).real();
                    if (pool_at(a.off + v * a.stride + e).imag() != 0.0) fits = false;
                }
                const double det = M[0] * M[3] - M[1] * M[2];
                if (!have_u && std::fabs(det) > 1e-3) {       // U = N M^-1 from an invertible variant (the no-jump operator)
                    u[0] = (N[0] * M[3] - N[1] * M[2]) / det;
                    u[1] = (N[1] * M[0] - N[0] * M[1]) / det;
                    u[2] = (N[2] * M[3] - N[3] * M[2]) / det;
                    u[3] = (N[3] * M[0] - N[2] * M[1]) / det;
                    have_u = true;
                }
            }
            if (!have_u || !fits) continue;
            for (uint32_t v = 0; v <= a.sel_mask && fits; v++) {
                double M[4], N[4];
                for (int e = 0; e < 4; e++) {
                    M[e] = pool_at(b.off + v * b.stride + e).real();
                    N[e] = pool_at(a.off + v * a.stride + e).real();
                }
                const double r[4] = {u[0] * M[0] + u[1] * M[2], u[0] * M[1] + u[1] * M[3], u[2] * M[0] + u[3] * M[2],
                                     u[2] * M[1] + u[3] * M[3]};
                for (int e = 0; e < 4; e++) fits = fits && std::fabs(r[e] - N[e]) < 1e-5;
            }
            if (fits && accept(j, k, u)) n_ok++;
        }
        if (n_ok < 2) {
            if (g_trace.on) fprintf(stderr, "[qcd trace] sibling run rejected: %zu eligible member(s)\n", n_ok);
            continue;
        }
        // the descriptor: register bits = base's matrix qubits + the lowest other bits (so that index bit 0 is always a
        // register bit: 128-bit loads)
        {
            DirectSweep D = *base;               // (copy first: direct_host grows below)
            uint32_t bits = r0_mask;
            for (int b = 0; __builtin_popcount(bits) < REG_BITS; b++) bits |= 1u << b;
            int idx_of[32];
            int ri = 0;
            for (int b = 0; b < 32; b++)
                if ((bits >> b) & 1u) {
                    idx_of[b] = ri;
                    D.rbit[ri++] = (uint8_t)b;
                    if (b >= chunk_log) R.high_bit[R.n_high++] = (uint8_t)b;
                }
            for (int o = 0; o < D.nops; o++) {
                KOp& op = D.ops[o];
                if (op.kind == K_M1 || op.kind == K_M2) {
                    const int g0 = base->rbit[op.t0], g1 = base->rbit[op.t1 & 3];
                    op.t0 = (uint8_t)idx_of[g0];
                    if (op.kind == K_M2) op.t1 = (uint8_t)idx_of[g1];
                }
            }
            D.red_q[0] = D.red_q[1] = -1;
            D.flags = SW_FINAL | SW_CROSS;
            R.reg_mask = bits;
            dp.direct_host.push_back(D);
            R.desc = (uint32_t)(dp.direct_host.size() - 1);
        }
        if (g_trace.on)
            fprintf(stderr, "[qcd trace] sibling run: circuits %zu..%zu, %zu eligible, shared-op qubits 0x%x, %d high\n",
                    run_begin, c1 - 1, n_ok, r0_mask, R.n_high);
        for (size_t c = run_begin; c < c1; c++) dp.cross_of_circuit[c] = (int32_t)dp.cross_runs.size();
        dp.cross_runs.push_back(std::move(R));
    }
}

}  // namespace

struct qcd_ctx {
    int device = 0;
    size_t budget = 0;
    std::string err;
    // options
    bool time_sweeps = false;
    int tile_bits = 0;           // 0 = default for the precision
    uint64_t max_chunk = 1u << 25;
    int table_cap_log2 = 22;
    uint64_t max_slots = 0;      // 0 = as many as the budget allows
    bool use_direct = true;      // register-only kernel for single-pass sweeps
    int direct_min_bits = 10;    // ... of states with at least this many index bits (smaller states: tile kernel).  Was 13;
                                 // on resident programs the register-only kernel is 27 / 25 / 45 % faster than the tile
                                 // kernel at n = 10 / 11 / 12 (2 000 reference-shaped circuits x 1024 shots: 12.4 -> 9.8,
                                 // 15.3 -> 12.2, 22.0 -> 15.2 ms), same complex128 outcomes
    bool no_slots = false;
    uint32_t nostore_max = 1024; // leaves with at most this many trajectories are sampled without writing their state
                                 // (graph 0 before sibling groups: 8 -> 56.6, 64 -> 54.2, 512 -> 54.0, 4096 -> 55.3, 32768 -> 65.3
                                 // ms per step; with sibling groups, the 8-graph job: 256 -> 2640, 1024 -> 2603, 4096 -> 2600 ms)
    bool share_prefix = true;    // consecutive circuits with identical sweeps up to the final one share their tree
    bool leaf_cross = true;      // unstored sibling leaves (same parent, same branch word, members of one run) take their
                                 // chunk sums from shared passes over the parent (SW_CROSS) instead of one sweep each
    uint32_t hist_row0 = 0;      // uploaded circuit that row 0 of the caller's histogram buffer stands for
    int dm_on_chip = 1;          // density matrices that fit a thread-block cluster's shared memory: interpret the gate list
                                 // there (dm_cluster.cu) instead of lowering it to sweeps.  0 never, 1 when the cluster
                                 // has at most 4 CTAs (measured end to end, circuits/s on chip vs lowered: n = 7 226 k vs
                                 // 136 k / 167 k vs 44 k, n = 8 156 k vs 112 k / 151 k vs 40 k for unitary / device noise;
                                 // the 16-CTA clusters of n = 9 lose: 33 k vs 101 k / 35 k vs 33 k), 2 whenever it fits
    int dm_max_clusters = 0;     // 0 = as many clusters as are resident at once
    bool dm_cluster_refused = false;   // the device refused the cluster launch once: use the lowered route from then on
    bool dm_test_refuse = false;       // test hook: behave as if the next cluster launch were refused
    size_t cached_budget = 0;    // last usable_budget() answer and how many runs have reused it
    int budget_age = 0;
    // programs (host copy of the IR; compiled lazily per (mode, precision))
    int n_qubits = 0, n_circuits = 0;
    std::vector<qcd_op> ops;
    std::vector<uint32_t> offsets;
    std::vector<double> params;
    // device-noise batches keep the noise-free gate lists + calibration numbers; every lowering thread writes a
    // circuit's channels into a small private buffer right before compiling it (no expanded copy of the batch)
    bool device_noise = false;
    std::vector<double> qubit_t1_t2, op_gate_props;
    // on-chip interpreter: explicit ops + pass schedule of the batch on the device (built at the first run)
    struct InterpProg {
        bool ready = false;
        DevBuf ops, params, sched, passes, pass_off;
        size_t n_passes = 0;
        uint32_t max_circuit_ops = 0;     // most scheduled ops of one circuit
        cudaEvent_t done = nullptr;       // recorded after the last launch that reads these buffers
        cudaStream_t copy = nullptr;      // non-blocking: uploads do not queue behind kernels on the legacy stream
    } interp;
    DevProg prog[2][2];  // [mode][precision == 64]
    int prog_tile_bits[2][2] = {{0, 0}, {0, 0}};
    // arenas / scratch
    DevBuf states, partials, chunk_sums, block_sums;
    DevBuf owner, keys, probs4, table, offs, children, counter, bsum, flip, masks, scratch, slots, compose, cross_recs;
    std::vector<DevBuf> level_nodes, level_shots;
    PinBuf flip_host;                 // staging of the readout flip table (asynchronous upload)
    cudaEvent_t flip_staged = nullptr;
    std::vector<PinBuf> host_nodes;   // children read back at depth d (= the nodes of depth d + 1), page-locked
    std::vector<cudaEvent_t> ev_pool;
    std::vector<uint8_t> ev_kind;     // kind of the launch bracketed by events (2 i, 2 i + 1)
    size_t ev_used = 0;
    qcd_stats stats;
};

namespace {

int fail(qcd_ctx* c, int code, const std::string& m) {
    c->err = m;
    return code;
}
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(ctx, QCD_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));    \
    } while (0)

template <typename U>
int upload_vec(qcd_ctx* ctx, DevBuf& b, const std::vector<U>& v) {
    CU(b.ensure(std::max<size_t>(v.size(), 1) * sizeof(U)));
    if (!v.empty()) CU(cudaMemcpy(b.p, v.data(), v.size() * sizeof(U), cudaMemcpyHostToDevice));
    return QCD_OK;
}

int default_tile_bits(int precision) { return precision == 64 ? 12 : 13; }

// compile + upload the program for (mode, precision) if not done yet
int get_program(qcd_ctx* ctx, int mode, int precision, DevProg** out) {
    if (ctx->n_circuits <= 0) return fail(ctx, QCD_E_STATE, "no programs uploaded (call qcd_upload_programs first)");
    int pi = precision == 64 ? 1 : 0;
    int kb = ctx->tile_bits > 0 ? ctx->tile_bits : default_tile_bits(precision);
    DevProg& dp = ctx->prog[mode][pi];
    if (dp.valid && ctx->prog_tile_bits[mode][pi] == kb) {
        *out = &dp;
        return QCD_OK;
    }
    const bool reuse = dp.host_ready && dp.P.k_max == kb && dp.P.mode == mode;
    if (!reuse) dp.invalidate();
    Program& P = dp.P;
    // matrix pools stay per lowering thread, already in device precision: no merged host copy of the (large) pool
    std::vector<std::vector<char>> dev_pools;
    std::vector<size_t> pool_elems;
    const double t_begin = Trace::now();
    if (!reuse) {
        P.n_qubits = ctx->n_qubits;
        P.mode = mode;
        P.n_bits = mode == MODE_DM ? 2 * ctx->n_qubits : ctx->n_qubits;
        P.k_max = kb;
        if (P.n_bits > 30) return fail(ctx, QCD_E_UNSUPPORTED, "state of more than 30 index bits");
        // lower on several host threads: contiguous blocks of circuits into private Programs, merged in order
        const int nc = ctx->n_circuits;
        int n_thr = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u);
        n_thr = std::max(1, std::min(n_thr, nc / 2));   // a 20-qubit circuit lowers in ~1 ms: threads pay off early
        std::vector<Program> parts(n_thr);
        dev_pools.assign(n_thr, std::vector<char>());
        pool_elems.assign(n_thr, 0);
        std::vector<int> part_rc(n_thr, QCD_OK);
        std::vector<std::string> part_err(n_thr);
        auto work = [&](int t) {
            Program& Q = parts[t];
            Q.n_qubits = P.n_qubits;
            Q.mode = P.mode;
            Q.n_bits = P.n_bits;
            Q.k_max = P.k_max;
            const int c_lo = (int)((long long)nc * t / n_thr), c_hi = (int)((long long)nc * (t + 1) / n_thr);
            std::vector<qcd_op> x_ops;
            std::vector<uint32_t> x_cnt;
            std::vector<double> x_par;
            for (int c = c_lo; c < c_hi; c++) {
                std::string err;
                uint32_t b = ctx->offsets[c], e = ctx->offsets[c + 1];
                int rc;
                if (ctx->device_noise) {
                    x_ops.clear();
                    x_cnt.clear();
                    x_par.clear();
                    rc = expand_device_noise(ctx->ops.data(), ctx->offsets.data(), ctx->params.data(), ctx->params.size(), c,
                                             c + 1, ctx->n_qubits, ctx->qubit_t1_t2.data(), ctx->op_gate_props.data(), x_ops,
                                             x_cnt, x_par, err);
                    if (rc == QCD_OK) rc = compile_circuit(Q, x_ops.data(), (uint32_t)x_ops.size(), x_par.data(), x_par.size(), err);
                } else {
                    rc = compile_circuit(Q, ctx->ops.data() + b, e - b, ctx->params.data(), ctx->params.size(), err);
                }
                if (rc != QCD_OK) {
                    part_rc[t] = rc;
                    part_err[t] = "circuit " + std::to_string(c) + ": " + err;
                    return;
                }
                if (c == c_lo && c_hi - c_lo > 8) {
                    // circuits of one batch are alike: size the block's arrays from the first one (no re-growth of
                    // 100 MB vectors, no page faults on the copies)
                    const size_t cnt = (size_t)(c_hi - c_lo) + 1;
                    auto grow = [&](auto& v) { v.reserve(v.size() * cnt + v.size() * cnt / 4); };
                    grow(Q.pool); grow(Q.kops); grow(Q.passes); grow(Q.sweeps); grow(Q.masks); grow(Q.groups);
                    grow(Q.subs); grow(Q.gdata);
                    Q.circuits.reserve(cnt);
                    advise_huge(Q.pool.data(), Q.pool.capacity() * sizeof(cd));
                }
            }
            const size_t np = Q.pool.size();
            pool_elems[t] = np;
            dev_pools[t].reserve(np * (precision == 64 ? sizeof(double2) : sizeof(float2)));
            advise_huge(dev_pools[t].data(), dev_pools[t].capacity());
            if (precision == 64) {
                dev_pools[t].resize(np * sizeof(double2));
                double2* d = reinterpret_cast<double2*>(dev_pools[t].data());
                for (size_t i = 0; i < np; i++) d[i] = make_double2(Q.pool[i].real(), Q.pool[i].imag());
            } else {
                dev_pools[t].resize(np * sizeof(float2));
                float2* d = reinterpret_cast<float2*>(dev_pools[t].data());
                for (size_t i = 0; i < np; i++) d[i] = make_float2((float)Q.pool[i].real(), (float)Q.pool[i].imag());
            }
            std::vector<cd>().swap(Q.pool);
        };
        if (n_thr == 1) {
            work(0);
        } else {
            std::vector<std::thread> pool;
            for (int t = 0; t < n_thr; t++) pool.emplace_back(work, t);
            for (std::thread& th : pool) th.join();
        }
        for (int t = 0; t < n_thr; t++)
            if (part_rc[t] != QCD_OK) return fail(ctx, part_rc[t], part_err[t]);
        {
            size_t n_sw = 0, n_pa = 0, n_ko = 0, n_ma = 0, n_gr = 0, n_su = 0, n_gd = 0;
            for (const Program& q : parts) {
                n_sw += q.sweeps.size(); n_pa += q.passes.size(); n_ko += q.kops.size(); n_ma += q.masks.size();
                n_gr += q.groups.size(); n_su += q.subs.size(); n_gd += q.gdata.size();
            }
            P.sweeps.reserve(n_sw); P.passes.reserve(n_pa); P.kops.reserve(n_ko); P.masks.reserve(n_ma);
            P.groups.reserve(n_gr); P.subs.reserve(n_su); P.gdata.reserve(n_gd); P.circuits.reserve(nc);
        }
        size_t pool_total = 0;
        for (int t = 0; t < n_thr; t++) {
            merge_program(P, parts[t], (long long)pool_total);
            pool_total += pool_elems[t];
            parts[t] = Program();
        }
        if (pool_total > 0xffffffffull) return fail(ctx, QCD_E_UNSUPPORTED, "matrix pool exceeds 2^32 entries: split the upload");
        dp.member_bits = assign_runs(P, mode);
    }
    const double t_lowered = Trace::now();
    dp.host_ready = false;
    // The device buffers of the previous program are about to be overwritten: kernels of an earlier asynchronous run
    // (on any stream) may still read them.  Lowering above ran concurrently with them; wait only now.
    CU(cudaDeviceSynchronize());
    int rc;
    if ((rc = upload_vec(ctx, dp.sweeps, P.sweeps))) return rc;
    if ((rc = upload_vec(ctx, dp.passes, P.passes))) return rc;
    if ((rc = upload_vec(ctx, dp.kops, P.kops))) return rc;
    {
        const size_t elem = precision == 64 ? sizeof(double2) : sizeof(float2);
        size_t total = 0;
        for (size_t n_el : pool_elems) total += n_el;
        CU(dp.pool.ensure(std::max<size_t>(total, 1) * elem));
        // one copy per lowering thread's pool, issued from parallel host threads (pageable copies are CPU-bound)
        std::vector<cudaError_t> cp_rc(dev_pools.size(), cudaSuccess);
        auto copy_part = [&](size_t t, size_t base) {
            cudaSetDevice(ctx->device);
            if (pool_elems[t])
                cp_rc[t] = cudaMemcpy(dp.pool.as<char>() + base * elem, dev_pools[t].data(), pool_elems[t] * elem,
                                      cudaMemcpyHostToDevice);
        };
        std::vector<std::thread> cps;
        size_t base = 0;
        for (size_t t = 0; t < dev_pools.size(); t++) {
            if (dev_pools.size() > 1 && pool_elems[t] * elem > (1u << 20)) cps.emplace_back(copy_part, t, base);
            else copy_part(t, base);
            base += pool_elems[t];
        }
        for (std::thread& th : cps) th.join();
        for (cudaError_t e : cp_rc)
            if (e != cudaSuccess) return fail(ctx, QCD_E_CUDA, std::string("pool upload: ") + cudaGetErrorString(e));
    }
    std::vector<uint32_t> masks(P.masks.size());
    for (size_t i = 0; i < masks.size(); i++) masks[i] = (uint32_t)P.masks[i];
    if ((rc = upload_vec(ctx, dp.masks, masks))) return rc;
    if ((rc = upload_vec(ctx, dp.groups, P.groups))) return rc;
    if ((rc = upload_vec(ctx, dp.subs, P.subs))) return rc;
    if ((rc = upload_vec(ctx, dp.gdata, P.gdata))) return rc;
    if ((rc = upload_vec(ctx, dp.circuits, P.circuits))) return rc;
    // direct (register-only) descriptors: entry si for single-pass sweep si, entry n_sweeps + pi for pass pi run as a
    // sweep of its own (density-matrix mode launches multi-pass sweeps pass by pass on the register-only kernel)
    dp.direct_host.clear();
    dp.direct_host.reserve(P.sweeps.size() + P.passes.size());
    advise_huge(dp.direct_host.data(), dp.direct_host.capacity() * sizeof(DirectSweep));
    {
        DirectSweep zero;
        memset(&zero, 0, sizeof(zero));
        dp.direct_host.assign(P.sweeps.size() + P.passes.size(), zero);
    }
    auto describe = [&](DirectSweep& D, const Sweep& S, uint32_t pi, uint8_t flags, bool reduce) {
        describe_pass(P, D, S, pi, flags, reduce);
    };
    dp.direct_flags.assign(dp.direct_host.size(), 0);
    auto flags_of = [&](size_t i) {
        const DirectSweep& D = dp.direct_host[i];
        uint8_t f = 0;
        if (D.ok) f |= DF_OK;
        if (D.rbit[0] == 0) f |= DF_VEC;
        if (D.red_q[0] >= 0) f |= DF_RED;
        if (D.flags & SW_FINAL) f |= DF_FINAL;
        if (D.flags & SW_PREP) f |= DF_PREP;
        for (int o = 0; o < D.nops; o++)
            if (D.ops[o].kind == K_M4) f |= DF_PREP;   // a 16x16 superoperator forces the generic launch kind
        dp.direct_flags[i] = f;
    };
    // sweeps own disjoint pass ranges: the descriptors of a block of sweeps are written by one host thread
    auto describe_block = [&](size_t s_lo, size_t s_hi) {
        for (size_t si = s_lo; si < s_hi; si++) {
            const Sweep& S = P.sweeps[si];
            if (S.pass_end - S.pass_begin == 1) describe(dp.direct_host[si], S, S.pass_begin, S.flags, true);
            flags_of(si);
            for (uint32_t pi = S.pass_begin; pi < S.pass_end; pi++) {
                uint8_t fl = 0;
                if (pi == S.pass_begin) fl |= S.flags & SW_PREP;
                if (pi + 1 == S.pass_end) fl |= S.flags & SW_FINAL;
                describe(dp.direct_host[P.sweeps.size() + pi], S, pi, fl, pi + 1 == S.pass_end);
                flags_of(P.sweeps.size() + pi);
            }
        }
    };
    {
        const size_t ns = P.sweeps.size();
        int n_thr = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
        n_thr = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_thr, ns / 2048));
        if (n_thr == 1) {
            describe_block(0, ns);
        } else {
            std::vector<std::thread> pool;
            for (int t = 0; t < n_thr; t++) pool.emplace_back(describe_block, ns * t / n_thr, ns * (t + 1) / n_thr);
            for (std::thread& th : pool) th.join();
        }
    }
    if (mode == MODE_SV) {
        const size_t n0 = dp.direct_host.size();
        // the matrix pool only exists per lowering thread, already in device precision
        build_cross(dp, [&](size_t off) -> std::complex<double> {
            size_t t = 0;
            while (t + 1 < pool_elems.size() && off >= pool_elems[t]) off -= pool_elems[t++];
            if (precision == 64) {
                const double2 v = reinterpret_cast<const double2*>(dev_pools[t].data())[off];
                return std::complex<double>(v.x, v.y);
            }
            const float2 v = reinterpret_cast<const float2*>(dev_pools[t].data())[off];
            return std::complex<double>(v.x, v.y);
        });
        dp.direct_flags.resize(dp.direct_host.size(), 0);
        for (size_t i = n0; i < dp.direct_host.size(); i++) flags_of(i);
    }
    if ((rc = upload_vec(ctx, dp.direct, dp.direct_host))) return rc;
    int k = std::min(P.k_max, P.n_bits);
    if (precision == 64) CU(configure_sweep<double>(k));
    else CU(configure_sweep<float>(k));
    dp.valid = true;
    ctx->prog_tile_bits[mode][pi] = kb;
    *out = &dp;
    if (g_trace.on)
        fprintf(stderr, "[qcd] program: %d circuits lowered in %.1f ms, uploaded in %.1f ms (%zu sweeps, %zu passes)\n",
                ctx->n_circuits, (t_lowered - t_begin) * 1e3, (Trace::now() - t_lowered) * 1e3, P.sweeps.size(),
                P.passes.size());
    return QCD_OK;
}

template <typename T>
DevProgram<T> dev_view(const DevProg& dp) {
    DevProgram<T> v;
    v.sweeps = dp.sweeps.as<Sweep>();
    v.passes = dp.passes.as<Pass>();
    v.kops = dp.kops.as<KOp>();
    v.pool = dp.pool.as<typename Cplx<T>::type>();
    v.masks = dp.masks.as<uint32_t>();
    v.groups = dp.groups.as<Group>();
    v.subs = dp.subs.as<SubSite>();
    v.gdata = dp.gdata.as<double>();
    v.direct = dp.direct.as<DirectSweep>();
    v.n_sweeps = (uint32_t)dp.P.sweeps.size();
    return v;
}

size_t usable_budget(qcd_ctx* ctx) {
    if (ctx->budget) return ctx->budget;
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) return (size_t)1 << 30;
    // memory we already hold for states is reusable
    fr += ctx->states.cap + ctx->partials.cap + ctx->chunk_sums.cap + ctx->block_sums.cap;
    return (size_t)(0.8 * (double)fr);
}

cudaEvent_t next_event(qcd_ctx* ctx) {
    if (ctx->ev_used == ctx->ev_pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        ctx->ev_pool.push_back(e);
    }
    return ctx->ev_pool[ctx->ev_used++];
}

enum { SK_PREP = 0, SK_INNER = 1, SK_LEAF = 2, SK_MIXED = 3 };

// bracket one sweep launch: launch/byte counters always, CUDA events only when "time_sweeps" is on
void sweep_begin(qcd_ctx* ctx, int kind, uint64_t bytes, cudaStream_t st) {
    ctx->stats.sweep_launches++;
    ctx->stats.sweep_bytes += bytes;
    ctx->stats.kind_launches[kind]++;
    ctx->stats.kind_bytes[kind] += bytes;
    if (ctx->time_sweeps) {
        if (ctx->ev_kind.size() <= ctx->ev_used / 2) ctx->ev_kind.resize(ctx->ev_used / 2 + 1);
        ctx->ev_kind[ctx->ev_used / 2] = (uint8_t)kind;
        cudaEventRecord(next_event(ctx), st);
    }
}
void sweep_end(qcd_ctx* ctx, cudaStream_t st) {
    if (ctx->time_sweeps) cudaEventRecord(next_event(ctx), st);
}

int finish_timing(qcd_ctx* ctx) {
    if (!ctx->time_sweeps) return QCD_OK;
    for (size_t i = 0; i + 1 < ctx->ev_used; i += 2) {
        float ms = 0;
        CU(cudaEventSynchronize(ctx->ev_pool[i + 1]));
        CU(cudaEventElapsedTime(&ms, ctx->ev_pool[i], ctx->ev_pool[i + 1]));
        ctx->stats.sweep_ms += ms;
        ctx->stats.kind_ms[ctx->ev_kind[i / 2]] += ms;
    }
    ctx->ev_used = 0;
    return QCD_OK;
}

// ============================================================================================ statevector trajectories
template <typename T>
struct SvRunner {
    typedef typename Cplx<T>::type C;
    qcd_ctx* ctx;
    DevProg* dp;
    DevProgram<T> prog;
    cudaStream_t st;
    int n = 0, k = 0, gbits = 0, max_sweeps = 0;
    uint32_t tiles_log2 = 0, tiles = 1, n_chunks = 1, n_blocks = 1, part_stride = 1;
    uint64_t slot_stride = 0;
    uint64_t seed = 0, spc = 0, shot_begin = 0;
    int mbits = 0;           // member bits of the grouping keys (0: no tree sharing between circuits)
    TrajMap map;             // of the chunk being processed
    const double* flip_dev = nullptr;
    uint32_t* hist = nullptr;
    uint32_t* outcomes = nullptr;
    std::vector<uint32_t> free_slots;
    uint64_t table_cap = 0;
    uint64_t live = 0;
    uint64_t n_slots = 0;
    bool cross = false;      // sibling groups enabled for this run (chunk_sums then holds a second region: C0 per slot)
    std::vector<NodeRec> cross_own;           // per leaf launch: the nodes that keep their own sweep
    std::vector<CrossRec> cross_passes;       // ... and the shared passes of the sibling groups
    std::vector<ComposeRec> cross_comp;
    std::vector<uint32_t> cross_idx;
    uint64_t grp_members[66] = {0}, grp_passes[66] = {0}, own_leaves = 0;   // QCD_TRACE: sibling groups by size
    double* c0_sums() const { return cross ? ctx->chunk_sums.as<double>() + n_slots * (uint64_t)n_chunks : nullptr; }

    uint32_t alloc_slot() {
        uint32_t s = free_slots.back();
        free_slots.pop_back();
        live++;
        if (live > ctx->stats.max_live_states) ctx->stats.max_live_states = live;
        return s;
    }
    void free_slot(uint32_t s) {
        free_slots.push_back(s);
        live--;
    }

    SweepArgs sweep_args(const NodeRec* nodes_dev, uint32_t n_nodes) {
        SweepArgs a;
        a.nodes = nodes_dev;
        a.states = ctx->states.p;
        a.slot_stride = slot_stride;
        a.partials = ctx->partials.as<double>();
        a.chunk_sums = ctx->chunk_sums.as<double>();
        a.chunk_stride = n_chunks;
        a.tiles_log2 = tiles_log2;
        a.n_items = n_nodes * tiles;
        a.part_stride = part_stride;
        a.direct_vec = 0;
        a.desc_from_pad = 0;
        a.direct_kind = 3;
        a.no_slots = ctx->no_slots ? 1 : 0;
        return a;
    }


    // Sibling groups of a leaf launch: unstored leaves with the same parent and branch word whose circuits belong to one
    // CrossRun.  Members are flagged NODE_CROSSED (pad = 1 + index of their ComposeRec) and take no sweep of their own;
    // cross_passes receives the groups' shared passes (4 lane-exchanged cross bits each), cross_own the nodes that keep
    // their own sweep.  Returns the number of shared passes (0: nothing grouped, the launch is unchanged).
    size_t plan_cross(NodeRec* chunk, size_t c, size_t* n_crossed, int* n_high) {
        cross_own.clear();
        cross_passes.clear();
        cross_comp.clear();
        *n_crossed = 0;
        *n_high = -1;
        const int chunk_log = std::min(n, CHUNK_LOG);
        for (size_t i = 0; i < c;) {
            size_t j = i;
            while (j < c && chunk[j].parent == chunk[i].parent) j++;
            cross_idx.clear();
            for (size_t t = i; t < j; t++) {
                const NodeRec& r = chunk[t];
                if ((r.flags & (NODE_LEAF | NODE_NOSTORE)) != (NODE_LEAF | NODE_NOSTORE)) continue;
                const int32_t cr = dp->cross_of_circuit[r.circuit];
                if (cr < 0) continue;
                const DevProg::CrossRun& CR = dp->cross_runs[cr];
                if (*n_high >= 0 && CR.n_high != *n_high) continue;      // one kernel variant per launch
                if (CR.kbit[r.circuit - CR.lead] != -2) cross_idx.push_back((uint32_t)t);
            }
            i = j;
            if (cross_idx.size() < 2) continue;
            std::stable_sort(cross_idx.begin(), cross_idx.end(),
                             [&](uint32_t a, uint32_t b) { return chunk[a].branch < chunk[b].branch; });
            for (size_t a = 0; a < cross_idx.size();) {
                size_t b = a;
                while (b < cross_idx.size() && chunk[cross_idx[b]].branch == chunk[cross_idx[a]].branch) b++;
                const size_t members = b - a;
                const NodeRec first = chunk[cross_idx[a]];
                const DevProg::CrossRun& CR = dp->cross_runs[dp->cross_of_circuit[first.circuit]];
                // cross bits of the members: the shared ops' own high qubits come for free with the first pass, the
                // others ride on lane bits, five per pass
                uint32_t x_reg[2] = {NO_SLOT, NO_SLOT};
                std::pair<uint8_t, uint32_t> kl[64];      // (bit, slot of the member that owns it)
                size_t nk = 0;
                for (size_t t = a; t < b; t++) {
                    const NodeRec& r = chunk[cross_idx[t]];
                    const int k = CR.kbit[r.circuit - CR.lead];
                    if (k < 0) continue;
                    if ((CR.reg_mask >> k) & 1u) x_reg[CR.high_bit[0] == k ? 0 : 1] = r.dst_slot;
                    else kl[nk++] = std::make_pair((uint8_t)k, r.dst_slot);
                }
                static const size_t KL = getenv("QCD_CROSS_KLANES")       // tuning only (measured: 5 > 4 > 3 > 2)
                                             ? std::min(CROSS_KLANES, std::max(1, atoi(getenv("QCD_CROSS_KLANES"))))
                                             : CROSS_KLANES;
                const size_t n_pass = std::max<size_t>(1, (nk + KL - 1) / KL);
                // a shared pass costs about 2.5 sweeps of one unstored leaf (measured: 3.1 us against 1.2 us at n = 20)
                if (5 * n_pass >= 2 * members) {      // nothing saved
                    a = b;
                    continue;
                }
                *n_high = CR.n_high;
                for (size_t pi = 0; pi < n_pass; pi++) {
                    CrossRec pr;
                    memset(&pr, 0, sizeof(pr));
                    pr.src_slot = first.src_slot;
                    pr.c0_slot = pi == 0 ? first.dst_slot : NO_SLOT;
                    pr.sweep = CR.desc;
                    pr.branch = first.branch;
                    pr.scale = first.scale;
                    uint32_t used = CR.reg_mask;
                    int nl = 0;
                    for (size_t q = KL * pi; q < std::min(nk, KL * pi + KL); q++) {
                        pr.lane_bit[nl] = kl[q].first;
                        pr.x_lane[nl] = kl[q].second;
                        used |= 1u << kl[q].first;
                        nl++;
                    }
                    pr.n_klanes = (uint8_t)nl;
                    for (int bb = 0; nl < 5; bb++)      // the lowest free bits fill the warp
                        if (!((used >> bb) & 1u)) pr.lane_bit[nl++] = (uint8_t)bb;
                    for (int q = pr.n_klanes; q < CROSS_KLANES; q++) pr.x_lane[q] = NO_SLOT;
                    pr.x_reg[0] = pi == 0 ? x_reg[0] : NO_SLOT;
                    pr.x_reg[1] = pi == 0 ? x_reg[1] : NO_SLOT;
                    cross_passes.push_back(pr);
                }
                uint32_t c0_owner = NO_SLOT;      // first member whose chunk sums are C0 itself: the others share its scan
                for (size_t t = a; t < b; t++) {
                    NodeRec& r = chunk[cross_idx[t]];
                    const size_t m = r.circuit - CR.lead;
                    const int k = CR.kbit[m];
                    if (k < 0) {
                        if (c0_owner != NO_SLOT) {
                            r.flags |= NODE_CROSSED | NODE_SHARED;
                            r.pad = c0_owner;
                            continue;
                        }
                        c0_owner = r.dst_slot;
                    }
                    ComposeRec cr;
                    cr.c0_slot = first.dst_slot;
                    cr.kb = k >= 0 ? 1u << (k - chunk_log) : 0u;
                    const double* u = &CR.u[4 * m];
                    cr.q[0] = u[0] * u[0]; cr.q[1] = u[1] * u[1]; cr.q[2] = 2.0 * u[0] * u[1];
                    cr.q[3] = u[2] * u[2]; cr.q[4] = u[3] * u[3]; cr.q[5] = 2.0 * u[2] * u[3];
                    cross_comp.push_back(cr);
                    r.flags |= NODE_CROSSED;
                    r.pad = (uint32_t)cross_comp.size();
                }
                *n_crossed += members;
                if (g_trace.on) {
                    grp_members[std::min<size_t>(members, 65)]++;
                    grp_passes[std::min<size_t>(members, 65)] += n_pass;
                }
                a = b;
            }
        }
        if (cross_passes.empty()) return 0;
        for (size_t t = 0; t < c; t++)
            if (!(chunk[t].flags & NODE_CROSSED)) cross_own.push_back(chunk[t]);
        return cross_passes.size();
    }

    // Measurement sampling of every leaf of nodes[begin, end) (the WHOLE batch, before anything below it is swept:
    // NODE_NOSTORE leaves read their parent's slot, which the host may already have returned to the free list).
    int sample_leaves(int d, const NodeRec* nodes, size_t begin, size_t end, bool* owner_filled = nullptr) {
        bool any_leaf = false, any_nostore = false, any_stored = false;
        for (size_t i = begin; i < end; i++)
            if (nodes[i].flags & NODE_LEAF) {
                any_leaf = true;
                if (nodes[i].flags & NODE_NOSTORE) any_nostore = true;
                else any_stored = true;
            }
        if (!any_leaf) return QCD_OK;
        const size_t n_nodes = end - begin;
        const uint32_t t0 = nodes[begin].traj_off;
        const uint32_t tc = nodes[end - 1].traj_off + nodes[end - 1].traj_cnt - t0;
        const NodeRec* ndev = ctx->level_nodes[d].as<NodeRec>() + begin;
        uint32_t* shots_d = ctx->level_shots[d].as<uint32_t>() + t0;
        uint32_t* owner = ctx->owner.as<uint32_t>();
        launch_fill_owner(ndev, (uint32_t)n_nodes, owner, tc, st);
        if (owner_filled) *owner_filled = true;   // same nodes, same trajectory range: the decide step reuses it
        launch_scan_chunks(ndev, (uint32_t)n_nodes, ctx->chunk_sums.as<double>(), n_chunks,
                           ctx->block_sums.as<double>(), n_blocks, n_chunks, st,
                           cross ? ctx->compose.as<ComposeRec>() : nullptr, c0_sums());
        ctx->stats.aux_launches += 3;
        if (any_stored) {
            launch_sample<T>(ndev, (uint32_t)n_nodes, ctx->states.p, slot_stride, ctx->chunk_sums.as<double>(), n_chunks,
                             ctx->block_sums.as<double>(), n_blocks, n_chunks, n, owner, shots_d, tc, seed, map, flip_dev,
                             hist, outcomes, shot_begin, st);
            ctx->stats.aux_launches++;
        }
        if (any_nostore) {
            launch_sample_nostore<T>(prog, ndev, ctx->states.p, slot_stride, ctx->chunk_sums.as<double>(), n_chunks,
                                     ctx->block_sums.as<double>(), n_blocks, n_chunks, n, owner, shots_d, tc, seed, map,
                                     flip_dev, hist, outcomes, shot_begin, st);
            ctx->stats.aux_launches++;
        }
        CU(cudaGetLastError());
        for (size_t i = begin; i < end; i++)
            if (nodes[i].flags & NODE_LEAF) {
                free_slot(nodes[i].dst_slot);
                ctx->stats.leaves++;
            }
        return QCD_OK;
    }

    // nodes[begin, end) at depth d have executed their sweep (d == 0: virtual roots); their device copy is
    // level_nodes[d] + begin and their trajectories are level_shots[d][traj_off ...).
    int process(int d, const NodeRec* nodes, size_t begin, size_t end, bool is_root, uint32_t n_parts,
                bool leaves_done = false) {
        const size_t n_nodes = end - begin;
        if (n_nodes == 0) return QCD_OK;
        double tp0 = g_trace.on ? Trace::now() : 0;
        bool owner_filled = false;
        if (!is_root && !leaves_done) {
            int rc = sample_leaves(d, nodes, begin, end, &owner_filled);
            if (rc) return rc;
        }
        if (g_trace.on) { g_trace.t_leaves += Trace::now() - tp0; tp0 = Trace::now(); }
        const int gtot = gbits + mbits;
        if (((uint64_t)n_nodes << gtot) > table_cap && n_nodes > 1) {
            const size_t piece = std::max<uint64_t>(1, table_cap >> gtot);
            bool split = false;
            for (size_t b = begin; b < end;) {
                size_t e = std::min(end, b + piece);
                // virtual roots hand their trajectories to the first root of their run: never cut a run
                while (is_root && e < end && nodes[e].parent != e) e++;
                if (b == begin && e == end) break;      // one run larger than the table cap: the table grows instead
                split = true;
                int rc = process(d, nodes, b, e, is_root, n_parts, true);
                if (rc) return rc;
                b = e;
            }
            if (split) return QCD_OK;
        }
        const uint32_t t0 = nodes[begin].traj_off;
        const uint32_t tc = nodes[end - 1].traj_off + nodes[end - 1].traj_cnt - t0;
        const NodeRec* ndev = ctx->level_nodes[d].as<NodeRec>() + begin;
        uint32_t* shots_d = ctx->level_shots[d].as<uint32_t>() + t0;
        uint32_t* owner = ctx->owner.as<uint32_t>();
        double* P = ctx->probs4.as<double>();
        CU(ctx->probs4.ensure(n_nodes * 4 * sizeof(double)));
        P = ctx->probs4.as<double>();

        if (!owner_filled) {
            launch_fill_owner(ndev, (uint32_t)n_nodes, owner, tc, st);
            ctx->stats.aux_launches++;
        }
        if (is_root) {
            launch_root_shots(ndev, owner, shots_d, tc, st);
            launch_fill_root_probs(P, (uint32_t)n_nodes, st);
            ctx->stats.aux_launches += 2;
        } else {
            bool any_inner = false;
            for (size_t i = begin; i < end; i++)
                if (!(nodes[i].flags & NODE_LEAF)) any_inner = true;
            if (!any_inner) return QCD_OK;
            launch_node_probs(ndev, (uint32_t)n_nodes, ctx->partials.as<double>(), part_stride, n_parts, P, st);
            ctx->stats.aux_launches++;
        }
        // ---- decide: branch word per trajectory, then group trajectories into children
        uint64_t* keys = ctx->keys.as<uint64_t>();
        if (((uint64_t)n_nodes << gtot) > 0xffffffffull)
            return fail(ctx, QCD_E_UNSUPPORTED, "trajectory-grouping table exceeds 2^32 entries");
        const uint32_t M = (uint32_t)(n_nodes << gtot);
        const uint32_t max_children = (uint32_t)std::min<uint64_t>(tc, M);
        CU(ctx->table.ensure((size_t)M * 4, true));
        // branch draw + first step of the grouping (key counting) in one pass over the trajectories
        launch_traj_branch<T>(prog, ndev, P, owner, shots_d, tc, seed, map, dp->circuits.as<CircuitInfo>(),
                              is_root ? (long long)begin : -1ll, keys, gbits, mbits, ctx->table.as<uint32_t>(), st);
        CU(ctx->offs.ensure((size_t)M * 4));
        CU(ctx->bsum.ensure(((size_t)M / 1024 + 2) * sizeof(uint2)));
        CU(ctx->children.ensure((size_t)max_children * sizeof(NodeRec)));
        CU(ctx->counter.ensure(2 * sizeof(uint32_t)));
        uint32_t* shots_next = ctx->level_shots[d + 1].as<uint32_t>() + t0;
        launch_group_children(keys, shots_d, tc, gbits, mbits, ctx->table.as<uint32_t>(), M, ctx->bsum.as<uint2>(),
                              ctx->counter.as<uint32_t>(), ndev, dp->circuits.as<CircuitInfo>(), t0,
                              ctx->offs.as<uint32_t>(), ctx->children.as<NodeRec>(), shots_next, true, st);
        launch_child_scale<T>(prog, ctx->children.as<NodeRec>(), max_children, ctx->counter.as<uint32_t>() + 1, P, st);
        ctx->stats.aux_launches += 6;
        uint32_t totals[2] = {0, 0};
        double tt0 = g_trace.on ? Trace::now() : 0;
        if (g_trace.on) g_trace.t_launch += tt0 - tp0;
        CU(cudaMemcpyAsync(totals, ctx->counter.p, sizeof(totals), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (g_trace.on) { g_trace.t_sync += Trace::now() - tt0; g_trace.n_sync++; }
        const uint32_t n_children = totals[1];
        if (n_children == 0) return fail(ctx, QCD_E_CUDA, "internal: inner nodes produced no children");
        if ((int)ctx->host_nodes.size() <= d) ctx->host_nodes.resize(d + 1);
        CU(ctx->host_nodes[d].ensure((size_t)n_children * sizeof(NodeRec)));
        NodeRec* ch = ctx->host_nodes[d].as<NodeRec>();   // stable while this call's subtree runs: deeper levels use their own
        tt0 = g_trace.on ? Trace::now() : 0;
        CU(cudaMemcpyAsync(ch, ctx->children.p, (size_t)n_children * sizeof(NodeRec), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (g_trace.on) { g_trace.t_h2d += Trace::now() - tt0; tt0 = Trace::now(); }

        // ---- sweep the children in chunks; finish each chunk's subtree before the next
        size_t pos = 0;
        size_t next_parent_to_free = 0;  // relative to `begin`
        const int reserve = std::max(0, max_sweeps - (d + 1));
        while (pos < n_children) {
            const long long avail = (long long)free_slots.size() - reserve;
            if (avail < 1)
                return fail(ctx, QCD_E_NOMEM, "not enough state slots for this circuit depth (raise the memory budget)");
            const size_t c = (size_t)std::min<long long>((long long)(n_children - pos), avail);
            // the chunk is edited in place in the page-locked readback buffer and goes back to the device from there
            NodeRec* chunk = ch + pos;
            NodeRec* const chunk_end = chunk + c;
            CU(ctx->level_nodes[d + 1].ensure(c * sizeof(NodeRec)));
            NodeRec* cdev = ctx->level_nodes[d + 1].as<NodeRec>();
            // single-pass sweeps run on the register-only kernel; anything else on the shared-memory tile kernel
            uint8_t f_and = 0xFF, f_or = 0;
            const uint8_t* dflags = dp->direct_flags.data();
            for (NodeRec* r = chunk; r != chunk_end; ++r) {
                r->dst_slot = alloc_slot();
                const uint8_t f = dflags[r->sweep];
                f_and &= f;
                f_or |= f;
            }
            const bool direct = ctx->use_direct && n >= ctx->direct_min_bits && (f_and & DF_OK);
            const bool vec = (f_and & DF_VEC) != 0, need_red = (f_or & DF_RED) != 0, need_chunks = (f_or & DF_FINAL) != 0,
                       need_prep = (f_or & DF_PREP) != 0;
            size_t n_leaf_nodes = 0;
            for (const NodeRec* r = chunk; r != chunk_end; ++r) n_leaf_nodes += (r->flags & NODE_LEAF) ? 1 : 0;
            const int launch_kind_of = n_leaf_nodes == 0 ? SK_INNER : (n_leaf_nodes == c ? SK_LEAF : SK_MIXED);
            uint64_t n_nostore = 0;
            if (direct && !is_root && ctx->nostore_max > 0) {
                const uint32_t lim = ctx->nostore_max;
                for (NodeRec* r = chunk; r != chunk_end; ++r)
                    if ((r->flags & NODE_LEAF) && r->traj_cnt <= lim) {
                        r->flags |= NODE_NOSTORE;
                        n_nostore++;
                    }
            }
            // sibling groups: their members take no sweep of their own
            size_t n_crossed = 0, n_cross_passes = 0;
            int cross_nh = 0;
            if (cross && direct && !is_root && n_leaf_nodes == c && n_nostore >= 2)
                n_cross_passes = plan_cross(chunk, c, &n_crossed, &cross_nh);
            const size_t n_own = n_cross_passes ? c - n_crossed : c;     // nodes swept by the ordinary launch
            CU(cudaMemcpyAsync(cdev, chunk, c * sizeof(NodeRec), cudaMemcpyHostToDevice, st));
            const NodeRec* own_dev = cdev;
            auto ctas_log2 = [&](uint64_t cnt) {
                // CTAs per state: 8 register groups per thread when there are plenty of states, fewer otherwise
                int pl = std::max(0, n - 4 - 8 - 3);
                while (pl < n - 4 - 8 && (cnt << pl) < 1184) pl++;
                while ((cnt << pl) > 0x7fffffffull) pl--;
                return pl;
            };
            if (n_cross_passes) {
                CU(ctx->children.ensure(std::max<size_t>(1, cross_own.size()) * sizeof(NodeRec)));
                CU(cudaMemcpyAsync(ctx->children.p, cross_own.data(), cross_own.size() * sizeof(NodeRec),
                                   cudaMemcpyHostToDevice, st));
                CU(ctx->compose.ensure(cross_comp.size() * sizeof(ComposeRec)));
                CU(cudaMemcpyAsync(ctx->compose.p, cross_comp.data(), cross_comp.size() * sizeof(ComposeRec),
                                   cudaMemcpyHostToDevice, st));
                CU(ctx->cross_recs.ensure(cross_passes.size() * sizeof(CrossRec)));
                CU(cudaMemcpyAsync(ctx->cross_recs.p, cross_passes.data(), cross_passes.size() * sizeof(CrossRec),
                                   cudaMemcpyHostToDevice, st));
                own_dev = ctx->children.as<NodeRec>();
                n_nostore -= n_crossed;
                ctx->stats.cross_passes += n_cross_passes;
                ctx->stats.cross_leaves += n_crossed;
                // a pass has >= 2^(n - 11) warp units (16 amplitudes x 32 lanes x the iteration bits): about 4 per warp,
                // i.e. as many CTAs per state as the sweep kernel uses -- the CTAs resident at one time then cover a few
                // parents only, and the later passes of a group find theirs in L2 (one CTA per pass streamed a whole
                // state each: every pass came from DRAM, 3.9 us per pass measured)
                int pl = std::max(0, n - 11 - 2 - 2);
                static const int pl_bias = getenv("QCD_CROSS_PL") ? atoi(getenv("QCD_CROSS_PL")) : 0;      // tuning only
                pl = std::max(0, pl + pl_bias);
                while (pl < n - 11 - 2 && ((uint64_t)n_cross_passes << pl) < 1184) pl++;
                sweep_begin(ctx, SK_LEAF, (uint64_t)n_cross_passes * slot_stride * sizeof(C), st);
                launch_leaf_cross<T>(prog, ctx->cross_recs.as<CrossRec>(), (uint32_t)n_cross_passes, ctx->states.p,
                                     slot_stride, c0_sums(), ctx->chunk_sums.as<double>(), n_chunks, n, cross_nh,
                                     (uint32_t)std::max(pl, 0), st);
                sweep_end(ctx, st);
                CU(cudaGetLastError());
                ctx->stats.node_sweeps += n_cross_passes;
            }
            SweepArgs sa = sweep_args(own_dev, (uint32_t)n_own);
            uint32_t parts_used = tiles;
            if (direct) {
                const int pl = ctas_log2(std::max<size_t>(n_own, 1));
                sa.tiles_log2 = (uint32_t)pl;
                sa.n_items = (uint32_t)(n_own << pl);
                sa.direct_vec = vec ? 1 : 0;
                sa.direct_kind = (!need_red && !need_chunks) ? 0u : ((need_prep || (need_red && need_chunks)) ? 3u
                                                                     : (need_red ? 1u : 2u));
                parts_used = 1u << pl;
            }
            // Sweeps of up to 3 passes run pass by pass on the register-only kernel (each pass a streaming read + write:
            // cheaper than one trip through the shared-memory tile kernel); longer ones stay on the tile kernel.
            size_t max_passes = 0;
            bool by_pass = !direct && ctx->use_direct && n >= ctx->direct_min_bits;
            if (by_pass) {
                const size_t n_sw = dp->P.sweeps.size();
                for (const NodeRec* rp = chunk; rp != chunk_end; ++rp) {
                    const NodeRec& r = *rp;
                    const Sweep& S = dp->P.sweeps[r.sweep];
                    const size_t np = S.pass_end - S.pass_begin;
                    if (np == 0 || np > 3) by_pass = false;
                    max_passes = std::max(max_passes, np);
                    for (uint32_t pi = S.pass_begin; pi < S.pass_end && by_pass; pi++)
                        by_pass = by_pass && dp->direct_host[n_sw + pi].ok;
                    if (!by_pass) break;
                }
            }
            if (by_pass) {
                const size_t n_sw = dp->P.sweeps.size();
                int pl = std::max(0, n - 4 - 8 - 3);
                while (pl < n - 4 - 8 && ((uint64_t)c << pl) < 1184) pl++;
                while (((uint64_t)c << pl) > 0x7fffffffull) pl--;
                parts_used = 1u << pl;     // the same split at every pass level: node_probs reads parts_used partials
                std::vector<NodeRec> recs;
                std::vector<size_t> lvl_off;
                std::vector<char> lvl_vec;
                for (size_t l = 0; l < max_passes; l++) {
                    lvl_off.push_back(recs.size());
                    bool v = true;
                    for (const NodeRec* rp = chunk; rp != chunk_end; ++rp) {
                        const NodeRec& r0 = *rp;
                        const Sweep& S = dp->P.sweeps[r0.sweep];
                        if (S.pass_end - S.pass_begin <= l) continue;
                        NodeRec r = r0;
                        r.pad = (uint32_t)(n_sw + S.pass_begin + l);
                        if (l > 0) {
                            r.src_slot = r.dst_slot;   // later passes work in place
                            r.scale = 1.0;
                        }
                        v = v && dp->direct_host[r.pad].rbit[0] == 0;
                        recs.push_back(r);
                    }
                    lvl_vec.push_back(v);
                }
                lvl_off.push_back(recs.size());
                CU(ctx->children.ensure(recs.size() * sizeof(NodeRec)));
                CU(cudaMemcpyAsync(ctx->children.p, recs.data(), recs.size() * sizeof(NodeRec), cudaMemcpyHostToDevice, st));
                for (size_t l = 0; l < max_passes; l++) {
                    const uint64_t cnt = lvl_off[l + 1] - lvl_off[l];
                    SweepArgs pa = sweep_args(ctx->children.as<NodeRec>() + lvl_off[l], (uint32_t)cnt);
                    pa.tiles_log2 = (uint32_t)pl;
                    pa.n_items = (uint32_t)(cnt << pl);
                    pa.direct_vec = lvl_vec[l] ? 1 : 0;
                    pa.direct_kind = 3;
                    pa.desc_from_pad = 1;
                    sweep_begin(ctx, (is_root && l == 0) ? SK_PREP : launch_kind_of,
                                cnt * ((is_root && l == 0) ? 1 : 2) * slot_stride * sizeof(C), st);
                    launch_sweep_direct<T>(prog, pa, st);
                    sweep_end(ctx, st);
                    CU(cudaGetLastError());
                }
                ctx->stats.node_sweeps += c;
            } else if (n_own) {
                sweep_begin(ctx, is_root ? SK_PREP : launch_kind_of,
                            ((uint64_t)n_own * (is_root ? 1 : 2) - n_nostore) * slot_stride * sizeof(C), st);
                if (direct) launch_sweep_direct<T>(prog, sa, st);
                else launch_sweep<T>(prog, sa, k, st);
                sweep_end(ctx, st);
                CU(cudaGetLastError());
                ctx->stats.node_sweeps += n_own;
            }
            pos += c;
            if (!is_root) {
                const size_t done_to = pos == n_children ? n_nodes : ch[pos].parent;
                for (size_t p = next_parent_to_free; p < done_to; p++)
                    if (!(nodes[begin + p].flags & NODE_LEAF)) free_slot(nodes[begin + p].dst_slot);
                next_parent_to_free = done_to;
            }
            if (g_trace.on) g_trace.t_prep += Trace::now() - tt0;
            int rc = process(d + 1, chunk, 0, c, false, parts_used);
            if (rc) return rc;
            tt0 = g_trace.on ? Trace::now() : 0;
        }
        return QCD_OK;
    }
};

int upload_flip(qcd_ctx* ctx, const double* readout, int n_vec, int n, const double** out, cudaStream_t st);

template <typename T>
int run_sv(qcd_ctx* ctx, DevProg* dp, uint64_t spc, uint64_t seed, uint64_t shot_begin, uint64_t shot_end,
           uint32_t first_circuit, const double* readout, uint32_t* hist, uint32_t* outcomes, uint32_t flags,
           cudaStream_t st) {
    typedef typename Cplx<T>::type C;
    if (g_trace.on) g_trace.t_run0 = Trace::now();
    SvRunner<T> R;
    R.ctx = ctx;
    R.dp = dp;
    R.prog = dev_view<T>(*dp);
    R.st = st;
    const Program& P = dp->P;
    const int n = P.n_bits;
    R.n = n;
    R.k = std::min(P.k_max, n);
    R.gbits = P.max_group_bits;
    R.max_sweeps = P.max_sweeps;
    R.tiles_log2 = (uint32_t)(n - R.k);
    R.tiles = 1u << R.tiles_log2;
    const int chunk_log = std::min(n, CHUNK_LOG);
    R.n_chunks = 1u << (n - chunk_log);
    R.n_blocks = (R.n_chunks + (1u << SCAN_BLOCK_LOG) - 1) >> SCAN_BLOCK_LOG;
    R.slot_stride = (uint64_t)1 << n;
    R.seed = seed;
    R.spc = spc;
    R.shot_begin = shot_begin;
    // histogram rows are only ever addressed through a trajectory's circuit index, so a caller that runs a slice of the
    // uploaded circuits may pass a buffer whose row 0 is circuit hist_row0 (checked against the slice by the caller below)
    R.hist = hist ? hist - ((size_t)ctx->hist_row0 << n) : nullptr;
    R.outcomes = outcomes;
    const bool no_dedup = (flags & QCD_RUN_NO_DEDUP) != 0;
    const bool share = ctx->share_prefix && !no_dedup && dp->member_bits > 0;
    R.mbits = share ? dp->member_bits : 0;
    R.table_cap = (uint64_t)1 << ctx->table_cap_log2;
    if (((uint64_t)1 << (R.gbits + R.mbits)) > R.table_cap) R.table_cap = (uint64_t)1 << (R.gbits + R.mbits);

    const uint64_t total = shot_end - shot_begin;
    // trajectories per tree: as many WHOLE circuits as fit max_chunk (a circuit split over two trees loses the
    // sharing between its halves); a single circuit larger than max_chunk is split
    uint64_t T_chunk = std::min<uint64_t>(total, ctx->max_chunk);
    if (spc <= T_chunk) T_chunk = (T_chunk / spc) * spc;

    // ---- state slots
    const size_t slot_bytes = R.slot_stride * sizeof(C);
    R.part_stride = std::max<uint32_t>(R.tiles, n >= 12 ? 1u << (n - 12) : 1u);
    // sibling groups: a second chunk-sum region (C0 of the group whose first member owns the slot)
    // (the region is sized from the options alone: the arena must not change shape between a program with sibling
    // groups and one without -- the reference-facing API alternates between the dataset circuit and its settings)
    const bool cross_region = ctx->leaf_cross && ctx->use_direct && n >= ctx->direct_min_bits && ctx->nostore_max > 0;
    R.cross = cross_region && share && !dp->cross_runs.empty();
    const size_t cs_per_slot = (size_t)R.n_chunks * 8 * (cross_region ? 2 : 1);
    const size_t per_slot = slot_bytes + (size_t)R.part_stride * 32 + cs_per_slot + (size_t)R.n_blocks * 8;
    // Keep the arena we already hold when it is adequate: the budget follows the device's free memory, which moves
    // with the caller's own allocations, and re-allocating a ~100 GB arena costs tens of milliseconds per call.  The
    // free-memory query itself is not free either: while the arena is adequate it is repeated every 16th run only.
    const uint64_t have = std::min(std::min<uint64_t>(ctx->states.cap / slot_bytes,
                                                      ctx->partials.cap / ((size_t)R.part_stride * 32)),
                                   std::min<uint64_t>(ctx->chunk_sums.cap / cs_per_slot,
                                                      ctx->block_sums.cap / ((size_t)R.n_blocks * 8)));
    const bool adequate = have >= (uint64_t)R.max_sweeps + 1 && !ctx->max_slots;
    size_t budget;
    if (adequate && ctx->cached_budget && ctx->budget_age < 16) {
        budget = ctx->cached_budget;
        ctx->budget_age++;
    } else {
        budget = usable_budget(ctx);
        ctx->cached_budget = budget;
        ctx->budget_age = 0;
    }
    uint64_t n_slots = budget / per_slot;
    const uint64_t wanted = 2 * T_chunk + 8 + (uint64_t)R.max_sweeps;
    n_slots = std::min<uint64_t>(n_slots, wanted);
    if (ctx->max_slots) n_slots = std::min<uint64_t>(n_slots, ctx->max_slots);
    if (adequate && have >= n_slots / 2) {
        n_slots = std::min(have, wanted);
    } else if (ctx->budget_age > 0) {
        // the arena has to grow: size it from the device's CURRENT free memory, not from a remembered figure
        budget = usable_budget(ctx);
        ctx->cached_budget = budget;
        ctx->budget_age = 0;
        n_slots = std::min<uint64_t>(budget / per_slot, wanted);
    }
    if (n_slots < (uint64_t)R.max_sweeps + 1)
        return fail(ctx, QCD_E_NOMEM,
                    "memory budget holds " + std::to_string(n_slots) + " states; this program needs at least " +
                        std::to_string(R.max_sweeps + 1));
    CU(ctx->states.ensure(n_slots * slot_bytes));
    CU(ctx->partials.ensure(n_slots * (size_t)R.part_stride * 32));
    CU(ctx->chunk_sums.ensure(n_slots * cs_per_slot));
    CU(ctx->block_sums.ensure(n_slots * (size_t)R.n_blocks * 8));
    R.n_slots = n_slots;
    R.free_slots.resize(n_slots);
    for (uint64_t i = 0; i < n_slots; i++) R.free_slots[i] = (uint32_t)(n_slots - 1 - i);

    // ---- per-level trajectory arrays
    const int n_levels = R.max_sweeps + 2;
    if ((int)ctx->level_nodes.size() < n_levels) ctx->level_nodes.resize(n_levels);
    if ((int)ctx->level_shots.size() < n_levels) ctx->level_shots.resize(n_levels);
    for (int l = 0; l < n_levels; l++) CU(ctx->level_shots[l].ensure(T_chunk * 4));
    CU(ctx->owner.ensure(T_chunk * 4));
    CU(ctx->keys.ensure(T_chunk * 8));
    if (ctx->table.p) CU(cudaMemsetAsync(ctx->table.p, 0, ctx->table.cap, st));  // a failed run may have left counts
    const double t_setup = g_trace.on ? Trace::now() - g_trace.t_run0 : 0;

    // ---- readout flip table [circuit][n][2] = P(1|0), P(0|1)
    if (readout) {      // stream-ordered upload: never overwrites the table under a kernel still reading it
        const double* fd = nullptr;
        int rc = upload_flip(ctx, readout, ctx->n_circuits, n, &fd, st);
        if (rc) return rc;
        R.flip_dev = fd;
    }

    // ---- outer loop over trajectory chunks of the flattened [circuit][shot] space
    for (uint64_t g0 = shot_begin; g0 < shot_end;) {
        // chunk boundaries on circuit boundaries whenever a chunk holds whole circuits
        uint64_t g1 = std::min(shot_end, g0 + T_chunk);
        if (spc <= T_chunk && g1 < shot_end) {
            const uint64_t aligned = (g1 / spc) * spc;
            if (aligned > g0) g1 = aligned;
            // ... and on the boundaries of runs of circuits that share a tree, when a run starts inside the chunk
            if (share && g1 % spc == 0 && g1 / spc < (uint64_t)ctx->n_circuits) {
                const uint64_t lead = P.circuits[g1 / spc].lead;
                if (lead * spc > g0) g1 = std::min(g1, lead * spc);
            }
        }
        // trajectory ids of this tree = positions in [g0, g1); virtual roots: one per circuit (one per shot without
        // de-duplication); node.pad = id of the root's first trajectory
        const uint64_t c_first = g0 / spc;
        R.map.spc = spc;
        R.map.off0 = g0 - c_first * spc;
        R.map.inv_spc = 1.0 / (double)spc;
        R.map.c_first = (uint32_t)c_first;
        R.map.first_circuit = first_circuit;
        std::vector<NodeRec> roots;
        uint32_t off = 0;
        for (uint64_t c = c_first; c <= (g1 - 1) / spc; c++) {
            const uint64_t s0 = std::max(g0, c * spc) - c * spc, s1 = std::min(g1, (c + 1) * spc) - c * spc;
            NodeRec r;
            memset(&r, 0, sizeof(r));
            r.sweep = P.circuits[c].sweep_begin - 1u;
            r.circuit = (uint32_t)c;
            r.scale = 1.0;
            if (no_dedup) {
                for (uint64_t s = s0; s < s1; s++) {
                    r.traj_off = off;
                    r.traj_cnt = 1;
                    r.pad = off;
                    r.parent = (uint32_t)roots.size();
                    off++;
                    roots.push_back(r);
                }
            } else {
                r.traj_off = off;
                r.traj_cnt = (uint32_t)(s1 - s0);
                r.pad = off;
                // the root that collects this circuit's trajectories: the first circuit of its run inside this chunk
                const uint64_t lead = share ? std::max<uint64_t>(P.circuits[c].lead, c_first) : c;
                r.parent = (uint32_t)(lead - c_first);
                off += r.traj_cnt;
                roots.push_back(r);
            }
        }
        CU(ctx->level_nodes[0].ensure(roots.size() * sizeof(NodeRec)));
        CU(cudaMemcpyAsync(ctx->level_nodes[0].p, roots.data(), roots.size() * sizeof(NodeRec), cudaMemcpyHostToDevice, st));
        int rc = R.process(0, roots.data(), 0, roots.size(), true, 1);
        if (rc) return rc;
        ctx->stats.trajectories += g1 - g0;
        g0 = g1;
    }
    double tt1 = g_trace.on ? Trace::now() : 0;
    CU(cudaStreamSynchronize(st));
    if (g_trace.on) {
        fprintf(stderr, "[qcd trace] run_sv: wait-for-decide %.2f ms in %d syncs, child readback %.2f ms, final drain %.2f ms; "
                "host: setup %.2f ms, leaf sampling launches %.2f ms, decide launches %.2f ms, child prep + sweep launch %.2f ms; "
                "total %.2f ms\n",
                g_trace.t_sync * 1e3, g_trace.n_sync, g_trace.t_h2d * 1e3, (Trace::now() - tt1) * 1e3, t_setup * 1e3,
                g_trace.t_leaves * 1e3, g_trace.t_launch * 1e3, g_trace.t_prep * 1e3, (Trace::now() - g_trace.t_run0) * 1e3);
        g_trace.t_sync = g_trace.t_h2d = g_trace.t_prep = g_trace.t_launch = g_trace.t_leaves = 0;
        g_trace.n_sync = 0;
        if (R.cross) {
            fprintf(stderr, "[qcd trace] sibling groups (members: groups / passes):");
            for (int m = 2; m < 66; m++)
                if (R.grp_members[m]) fprintf(stderr, " %d: %llu / %llu;", m, (unsigned long long)R.grp_members[m],
                                              (unsigned long long)R.grp_passes[m]);
            fprintf(stderr, "  leaves %llu, crossed %llu, passes %llu\n", (unsigned long long)ctx->stats.leaves,
                    (unsigned long long)ctx->stats.cross_leaves, (unsigned long long)ctx->stats.cross_passes);
        }
    }
    return finish_timing(ctx);
}

// ============================================================================================ density matrix

// ---------------------------------------------------------------------------------------------- on-chip density matrices
bool interp_eligible(const qcd_ctx* ctx, int precision) {
    const int cs = dm_interp_cluster_size(ctx->n_qubits, precision);
    return ctx->dm_on_chip && !ctx->dm_cluster_refused && cs > 0 && (ctx->dm_on_chip >= 2 || cs <= 4);
}

// Explicit ops (device-noise batches: the channels written natively, circuit block by circuit block) + the pass schedule
// of every circuit, built on host threads and copied to the device.  O(ops) per circuit, no matrix arithmetic.
int prepare_interp(qcd_ctx* ctx) {
    if (ctx->interp.ready) return QCD_OK;
    const double t_begin = Trace::now();
    const int nc = ctx->n_circuits, n = ctx->n_qubits;
    int n_thr = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u);
    n_thr = std::max(1, std::min(n_thr, nc / 64));
    struct Part {
        std::vector<qcd_op> ops;
        std::vector<double> par;
        std::vector<DmPass> passes;
        std::vector<uint32_t> sched, pass_cnt;
        int rc = QCD_OK;
        std::string err;
    };
    std::vector<Part> parts(n_thr);
    const bool expand = ctx->device_noise;
    auto build = [&](int t) {
        Part& Q = parts[t];
        const int c_lo = (int)((long long)nc * t / n_thr), c_hi = (int)((long long)nc * (t + 1) / n_thr);
        std::string err;
        if (expand) {
            std::vector<uint32_t> cnt;
            Q.rc = expand_device_noise(ctx->ops.data(), ctx->offsets.data(), ctx->params.data(), ctx->params.size(), c_lo, c_hi,
                                       n, ctx->qubit_t1_t2.data(), ctx->op_gate_props.data(), Q.ops, cnt, Q.par, err);
            if (Q.rc != QCD_OK) { Q.err = err; return; }
            uint32_t base = 0;
            for (int c = c_lo; c < c_hi; c++) {
                const size_t p0 = Q.passes.size();
                Q.rc = dm_interp_schedule(Q.ops.data() + base, cnt[c - c_lo], base, Q.par.data(), Q.par.size(), n, Q.passes,
                                          Q.sched, err);
                if (Q.rc != QCD_OK) { Q.err = "circuit " + std::to_string(c) + ": " + err; return; }
                Q.pass_cnt.push_back((uint32_t)(Q.passes.size() - p0));
                base += cnt[c - c_lo];
            }
        } else {
            for (int c = c_lo; c < c_hi; c++) {
                const uint32_t b = ctx->offsets[c], e = ctx->offsets[c + 1];
                const size_t p0 = Q.passes.size();
                Q.rc = dm_interp_schedule(ctx->ops.data() + b, e - b, b, ctx->params.data(), ctx->params.size(), n, Q.passes,
                                          Q.sched, err);
                if (Q.rc != QCD_OK) { Q.err = "circuit " + std::to_string(c) + ": " + err; return; }
                Q.pass_cnt.push_back((uint32_t)(Q.passes.size() - p0));
            }
        }
    };
    auto run_all = [&](auto&& fn) {
        if (n_thr == 1) { fn(0); return; }
        std::vector<std::thread> pool;
        for (int t = 0; t < n_thr; t++) pool.emplace_back(fn, t);
        for (std::thread& th : pool) th.join();
    };
    run_all(build);
    for (int t = 0; t < n_thr; t++)
        if (parts[t].rc != QCD_OK) return fail(ctx, parts[t].rc, parts[t].err);
    std::vector<size_t> ops_base(n_thr + 1, 0), par_base(n_thr + 1, 0), sched_base(n_thr + 1, 0), pass_base(n_thr + 1, 0);
    for (int t = 0; t < n_thr; t++) {
        ops_base[t + 1] = ops_base[t] + parts[t].ops.size();
        par_base[t + 1] = par_base[t] + parts[t].par.size();
        sched_base[t + 1] = sched_base[t] + parts[t].sched.size();
        pass_base[t + 1] = pass_base[t] + parts[t].passes.size();
    }
    if (ops_base[n_thr] > 0xffffffffull || par_base[n_thr] > 0xffffffffull || sched_base[n_thr] > 0xffffffffull)
        return fail(ctx, QCD_E_UNSUPPORTED, "batch exceeds 2^32 ops / parameters: split the upload");
    std::vector<uint32_t> pass_off;
    pass_off.reserve((size_t)nc + 1);
    pass_off.push_back(0);
    uint32_t max_circuit_ops = 0;
    for (int t = 0; t < n_thr; t++) {
        size_t pi = 0;
        for (uint32_t k : parts[t].pass_cnt) {
            pass_off.push_back(pass_off.back() + k);
            uint32_t ops_c = 0;
            for (uint32_t j = 0; j < k; j++) ops_c += parts[t].passes[pi + j].n_ops;
            pi += k;
            max_circuit_ops = std::max(max_circuit_ops, ops_c);
        }
    }
    qcd_ctx::InterpProg& I = ctx->interp;
    // the previous batch's buffers may still be read by THIS context's last interpreter launch (and by nothing else):
    // wait for that launch only -- another context's kernels keep running while this one prepares its next chunk
    if (I.done) CU(cudaEventSynchronize(I.done));
    if (!I.copy) CU(cudaStreamCreateWithFlags(&I.copy, cudaStreamNonBlocking));
    // Chunks of one batch differ by a few percent in size: grow with headroom, a re-allocation (cudaFree waits for every
    // kernel in flight, the other context's included) cost 13 - 74 ms when a chunk was slightly larger than the previous one.
    auto roomy = [](DevBuf& b, size_t bytes) -> cudaError_t { return bytes <= b.cap ? cudaSuccess : b.ensure(bytes + bytes / 4 + 4096); };
    auto upload_on_copy = [&](DevBuf& b, const void* src, size_t bytes) -> cudaError_t {
        cudaError_t ce = roomy(b, std::max<size_t>(bytes, 1));
        if (ce == cudaSuccess && bytes) ce = cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, I.copy);
        return ce;
    };
    if (expand) {
        CU(roomy(I.ops, std::max<size_t>(ops_base[n_thr], 1) * sizeof(qcd_op)));
        CU(roomy(I.params, std::max<size_t>(par_base[n_thr], 1) * sizeof(double)));
    } else {
        CU(upload_on_copy(I.ops, ctx->ops.data(), ctx->ops.size() * sizeof(qcd_op)));
        CU(upload_on_copy(I.params, ctx->params.data(), ctx->params.size() * sizeof(double)));
    }
    CU(roomy(I.sched, std::max<size_t>(sched_base[n_thr], 1) * sizeof(uint32_t)));
    CU(roomy(I.passes, std::max<size_t>(pass_base[n_thr], 1) * sizeof(DmPass)));
    std::vector<cudaError_t> cp_rc(n_thr, cudaSuccess);
    auto fix_and_copy = [&](int t) {     // re-base the block's offsets, then copy it from its own host thread
        Part& Q = parts[t];
        cudaSetDevice(ctx->device);
        auto cp = [&](void* dst, const void* src, size_t bytes) {
            if (bytes && cp_rc[t] == cudaSuccess) cp_rc[t] = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, I.copy);
        };
        if (expand) {
            for (qcd_op& o : Q.ops) o.param_off += (uint32_t)par_base[t];
            for (uint32_t& s : Q.sched) s += (uint32_t)ops_base[t];
            cp(I.ops.as<qcd_op>() + ops_base[t], Q.ops.data(), Q.ops.size() * sizeof(qcd_op));
            cp(I.params.as<double>() + par_base[t], Q.par.data(), Q.par.size() * sizeof(double));
        }
        for (DmPass& p : Q.passes) p.first += (uint32_t)sched_base[t];
        cp(I.sched.as<uint32_t>() + sched_base[t], Q.sched.data(), Q.sched.size() * sizeof(uint32_t));
        cp(I.passes.as<DmPass>() + pass_base[t], Q.passes.data(), Q.passes.size() * sizeof(DmPass));
    };
    run_all(fix_and_copy);
    for (cudaError_t ce : cp_rc)
        if (ce != cudaSuccess) return fail(ctx, QCD_E_CUDA, std::string("interpreter program upload: ") + cudaGetErrorString(ce));
    CU(upload_on_copy(I.pass_off, pass_off.data(), pass_off.size() * sizeof(uint32_t)));
    CU(cudaStreamSynchronize(I.copy));      // the host vectors go out of scope; the launch stream sees complete buffers
    I.n_passes = pass_base[n_thr];
    I.max_circuit_ops = max_circuit_ops;
    I.ready = true;
    if (g_trace.on)
        fprintf(stderr, "[qcd] on-chip DM program: %d circuits, %zu passes, %zu ops, %.2f ms on %d host threads\n", nc,
                I.n_passes, expand ? ops_base[n_thr] : ctx->ops.size(), 1e3 * (Trace::now() - t_begin), n_thr);
    return QCD_OK;
}

template <typename T>
int launch_interp(qcd_ctx* ctx, int precision, int c0, int c1, void* probs, void* rho, uint64_t slot_stride, cudaStream_t st) {
    const qcd_ctx::InterpProg& I = ctx->interp;
    DmInterpArgs a;
    a.ops = I.ops.as<qcd_op>();
    a.params = I.params.as<double>();
    a.sched = I.sched.as<uint32_t>();
    a.passes = I.passes.as<DmPass>();
    a.pass_off = I.pass_off.as<uint32_t>();
    a.n_qubits = ctx->n_qubits;
    a.n_circuits = c1 - c0;
    a.circuit0 = c0;
    a.local_bits = std::min(2 * ctx->n_qubits, precision == 64 ? 13 : 14);
    a.rec_cap = dm_interp_rec_cap(ctx->n_qubits, precision, I.max_circuit_ops);
    a.probs = probs;
    a.rho_out = rho;
    a.slot_stride = slot_stride;
    {
        // a device (or a partition of one) may not be able to co-schedule the cluster -- 16 CTAs is a non-portable size:
        // that is not an error of the request, the caller falls back to the lowered sweep route
        const cudaError_t le = ctx->dm_test_refuse ? cudaErrorLaunchOutOfResources
                                                   : launch_dm_interp<T>(a, dm_interp_cluster_size(ctx->n_qubits, precision),
                                                                         ctx->dm_max_clusters, st);
        if (le == cudaErrorLaunchOutOfResources || le == cudaErrorInvalidConfiguration || le == cudaErrorInvalidValue ||
            le == cudaErrorNotSupported || le == cudaErrorInvalidClusterSize) {
            cudaGetLastError();
            ctx->dm_cluster_refused = true;
            if (g_trace.on)
                fprintf(stderr, "[qcd] cluster launch refused (%s): n = %d, precision %d, %d CTAs per cluster, %d records -> lowered sweeps\n",
                        cudaGetErrorString(le), ctx->n_qubits, precision, dm_interp_cluster_size(ctx->n_qubits, precision), a.rec_cap);
            return fail(ctx, QCD_E_UNSUPPORTED, std::string("cluster launch refused: ") + cudaGetErrorString(le));
        }
        CU(le);
    }
    if (!ctx->interp.done) CU(cudaEventCreateWithFlags(&ctx->interp.done, cudaEventDisableTiming));
    CU(cudaEventRecord(ctx->interp.done, st));
    ctx->stats.aux_launches++;
    return QCD_OK;
}

template <typename T>
int run_dm(qcd_ctx* ctx, DevProg* dp, int cb, int ce, void* probs_out, cudaStream_t st, int n_settings = 0,
           const uint8_t* setting_qubits = nullptr, const double* setting_superops = nullptr,
           const uint32_t* group_adj = nullptr /*host [n]: read <S> of the whole stabilizer group into probs_out (double)*/) {
    typedef typename Cplx<T>::type C;
    // dp == nullptr: the circuits are interpreted on chip (dm_cluster.cu); the state arena is only needed when a
    // read-out kernel wants the whole density matrix (settings, stabilizer group)
    const bool on_chip = dp == nullptr;
    const int prec = sizeof(T) == 8 ? 64 : 32;
    if (on_chip && n_settings == 0 && !group_adj) {
        int rc = launch_interp<T>(ctx, prec, cb, ce, probs_out, nullptr, 0, st);
        if (rc) return rc;
        ctx->stats.leaves += ce - cb;
        return QCD_OK;
    }
    static const Program no_program;
    const Program& P = on_chip ? no_program : dp->P;
    const int nb = on_chip ? 2 * ctx->n_qubits : P.n_bits, n = on_chip ? ctx->n_qubits : P.n_qubits;
    const int k = std::min(on_chip ? 13 : P.k_max, nb);
    const uint32_t tiles_log2 = (uint32_t)(nb - k);
    const uint32_t tiles = 1u << tiles_log2;
    const uint64_t slot_stride = (uint64_t)1 << nb;
    const size_t slot_bytes = slot_stride * sizeof(C);
    // The free-memory query behind usable_budget() costs 0.1 - 86 ms depending on the box (it walks the allocations): only
    // ask when the arena we already hold cannot take the whole batch.
    uint64_t n_slots = (uint64_t)(ce - cb);
    if (ctx->states.cap / slot_bytes < n_slots) n_slots = std::min<uint64_t>(n_slots, usable_budget(ctx) / slot_bytes);
    if (ctx->max_slots) n_slots = std::min<uint64_t>(n_slots, ctx->max_slots);
    if (n_slots < 1) return fail(ctx, QCD_E_NOMEM, "memory budget too small for one density matrix");
    if (n_slots * tiles > 0x7fffffffull) n_slots = 0x7fffffffull / tiles;
    CU(ctx->states.ensure(n_slots * slot_bytes));
    DevProgram<T> prog;
    if (!on_chip) prog = dev_view<T>(*dp);
    T* out = reinterpret_cast<T*>(probs_out);
    for (int c0 = cb; c0 < ce; c0 += (int)n_slots) {
        const int c1 = (int)std::min<uint64_t>((uint64_t)ce, (uint64_t)c0 + n_slots);
        if (on_chip) {
            int rc = launch_interp<T>(ctx, prec, c0, c1, nullptr, ctx->states.p, slot_stride, st);
            if (rc) return rc;
        } else {
        // Register-only path: every pass of every circuit of the chunk becomes a launch level of its own on the direct
        // kernel (more state passes than the tiled kernel's one-per-sweep, each at streaming speed).
        bool all_direct = ctx->use_direct && nb >= ctx->direct_min_bits;
        for (int c = c0; c < c1 && all_direct; c++) {
            const CircuitInfo& ci = P.circuits[c];
            for (uint32_t si = ci.sweep_begin; si < ci.sweep_end && all_direct; si++) {
                if (P.sweeps[si].pass_begin == P.sweeps[si].pass_end) all_direct = false;   // op-less sweep
                for (uint32_t pi = P.sweeps[si].pass_begin; pi < P.sweeps[si].pass_end; pi++)
                    all_direct = all_direct && dp->direct_host[P.sweeps.size() + pi].ok;
            }
        }
        if (all_direct) {
            std::vector<std::vector<uint32_t>> pass_list(c1 - c0);
            size_t max_p = 0;
            for (int c = c0; c < c1; c++) {
                const CircuitInfo& ci = P.circuits[c];
                for (uint32_t si = ci.sweep_begin; si < ci.sweep_end; si++)
                    for (uint32_t pi = P.sweeps[si].pass_begin; pi < P.sweeps[si].pass_end; pi++)
                        pass_list[c - c0].push_back(pi);
                max_p = std::max(max_p, pass_list[c - c0].size());
            }
            std::vector<NodeRec> recs;
            std::vector<size_t> lvl_off;
            std::vector<char> lvl_vec, lvl_prep;
            for (size_t l = 0; l < max_p; l++) {
                lvl_off.push_back(recs.size());
                bool vec = true, prep = true;
                for (int c = c0; c < c1; c++) {
                    if (pass_list[c - c0].size() <= l) continue;
                    NodeRec r;
                    memset(&r, 0, sizeof(r));
                    r.src_slot = r.dst_slot = (uint32_t)(c - c0);
                    r.scale = 1.0;
                    r.circuit = (uint32_t)c;
                    r.pad = (uint32_t)(P.sweeps.size() + pass_list[c - c0][l]);
                    const DirectSweep& D = dp->direct_host[r.pad];
                    vec = vec && D.rbit[0] == 0;
                    prep = prep && (D.flags & SW_PREP);
                    recs.push_back(r);
                }
                lvl_vec.push_back(vec);
                lvl_prep.push_back(prep);
            }
            lvl_off.push_back(recs.size());
            CU(ctx->children.ensure(recs.size() * sizeof(NodeRec)));
            CU(cudaMemcpyAsync(ctx->children.p, recs.data(), recs.size() * sizeof(NodeRec), cudaMemcpyHostToDevice, st));
            for (size_t l = 0; l < max_p; l++) {
                const uint64_t cnt = lvl_off[l + 1] - lvl_off[l];
                int pl = std::max(0, nb - 4 - 8 - 3);
                while (pl < nb - 4 - 8 && (cnt << pl) < 1184) pl++;
                while ((cnt << pl) > 0x7fffffffull) pl--;
                SweepArgs a;
                a.nodes = ctx->children.as<NodeRec>() + lvl_off[l];
                a.states = ctx->states.p;
                a.slot_stride = slot_stride;
                a.partials = nullptr;
                a.chunk_sums = nullptr;
                a.chunk_stride = 0;
                a.tiles_log2 = (uint32_t)pl;
                a.n_items = (uint32_t)(cnt << pl);
                a.part_stride = 1;
                a.direct_vec = lvl_vec[l] ? 1 : 0;
                a.desc_from_pad = 1;
                a.direct_kind = 0;
                a.no_slots = ctx->no_slots ? 1 : 0;
                sweep_begin(ctx, lvl_prep[l] ? SK_PREP : SK_INNER, cnt * slot_bytes * (lvl_prep[l] ? 1 : 2), st);
                launch_sweep_direct<T>(prog, a, st);
                sweep_end(ctx, st);
                CU(cudaGetLastError());
                ctx->stats.node_sweeps += cnt;
            }
        } else {
        // one NodeRec per (circuit, sweep), grouped by sweep ordinal; the sweep works in place (src == dst)
        std::vector<NodeRec> recs;
        std::vector<size_t> lvl_off;
        for (int l = 0;; l++) {
            size_t before = recs.size();
            for (int c = c0; c < c1; c++) {
                const CircuitInfo& ci = P.circuits[c];
                if (ci.sweep_begin + l >= ci.sweep_end) continue;
                NodeRec r;
                memset(&r, 0, sizeof(r));
                r.src_slot = r.dst_slot = (uint32_t)(c - c0);
                r.sweep = ci.sweep_begin + l;
                r.scale = 1.0;
                r.circuit = (uint32_t)c;
                recs.push_back(r);
            }
            if (recs.size() == before) break;
            lvl_off.push_back(before);
        }
        lvl_off.push_back(recs.size());
        CU(ctx->children.ensure(recs.size() * sizeof(NodeRec)));
        CU(cudaMemcpyAsync(ctx->children.p, recs.data(), recs.size() * sizeof(NodeRec), cudaMemcpyHostToDevice, st));
        for (size_t l = 0; l + 1 < lvl_off.size(); l++) {
            const uint32_t cnt = (uint32_t)(lvl_off[l + 1] - lvl_off[l]);
            SweepArgs a;
            a.nodes = ctx->children.as<NodeRec>() + lvl_off[l];
            a.states = ctx->states.p;
            a.slot_stride = slot_stride;
            a.partials = nullptr;
            a.chunk_sums = nullptr;
            a.chunk_stride = 0;
            a.tiles_log2 = tiles_log2;
            a.n_items = cnt * tiles;
            a.part_stride = tiles;
            a.direct_vec = 0;
            a.desc_from_pad = 0;
            a.direct_kind = 3;
            a.no_slots = 0;
            sweep_begin(ctx, l == 0 ? SK_PREP : SK_INNER, (uint64_t)cnt * slot_bytes * (l == 0 ? 1 : 2), st);
            launch_sweep<T>(prog, a, k, st);
            sweep_end(ctx, st);
            CU(cudaGetLastError());
            ctx->stats.node_sweeps += cnt;
        }
        }
        }   // lowered route
        std::vector<uint32_t> slots(c1 - c0);
        for (int i = 0; i < c1 - c0; i++) slots[i] = (uint32_t)i;
        CU(ctx->slots.ensure(slots.size() * 4));
        CU(cudaMemcpyAsync(ctx->slots.p, slots.data(), slots.size() * 4, cudaMemcpyHostToDevice, st));
        if (group_adj) {
            // <S> of all 2^n elements of the graph state's stabilizer group, one pass over each density matrix
            const uint32_t chunks = group_expvals_chunks(n);
            const size_t part_bytes = (size_t)(c1 - c0) * chunks * ((size_t)1 << n) * sizeof(double);
            CU(ctx->scratch.ensure(part_bytes + 32 * sizeof(uint32_t)));
            uint32_t* adj_dev = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(ctx->scratch.p) + part_bytes);
            CU(cudaMemcpyAsync(adj_dev, group_adj, (size_t)n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
            launch_group_expvals<T>(ctx->states.p, slot_stride, (uint32_t)(c1 - c0), n, adj_dev, ctx->scratch.as<double>(),
                                    reinterpret_cast<double*>(probs_out) + ((size_t)(c0 - cb) << n), st);
            ctx->stats.aux_launches++;
        } else if (n_settings > 0) {
            // measurement settings of this batch's density matrices (slot i holds circuit c0 + i)
            const size_t cnt = (size_t)(c1 - c0) * n_settings;
            CU(ctx->scratch.ensure(cnt * (16 * sizeof(double2) + 1)));
            double2* sops_dev = ctx->scratch.as<double2>();
            uint8_t* q_dev = reinterpret_cast<uint8_t*>(sops_dev + cnt * 16);
            CU(cudaMemcpyAsync(sops_dev, setting_superops + (size_t)(c0 - cb) * n_settings * 32, cnt * 16 * sizeof(double2),
                               cudaMemcpyHostToDevice, st));
            CU(cudaMemcpyAsync(q_dev, setting_qubits + (size_t)(c0 - cb) * n_settings, cnt, cudaMemcpyHostToDevice, st));
            launch_setting_probs<T>(ctx->states.p, slot_stride, (uint32_t)(c1 - c0), n, n_settings, q_dev, sops_dev,
                                    out + (((size_t)(c0 - cb) * n_settings) << n), st);
        } else {
            launch_extract_diag<T>(ctx->states.p, slot_stride, ctx->slots.as<uint32_t>(), (uint32_t)(c1 - c0), n,
                                   out + ((size_t)(c0 - cb) << n), st);
        }
        CU(cudaGetLastError());
        ctx->stats.aux_launches++;
        ctx->stats.leaves += c1 - c0;
    }
    if (ctx->time_sweeps) {
        CU(cudaStreamSynchronize(st));
        return finish_timing(ctx);
    }
    return QCD_OK;
}

// Readout flip table [vector][n][2] = P(1|0), P(0|1) -> ctx->flip, ASYNCHRONOUSLY on the caller's stream from a page-locked
// staging buffer: a blocking cudaMemcpy here would make the host wait for every kernel enqueued before it (in the chunked
// array route: for the whole simulation of the chunk, so that no chunk ever overlapped the next one's preparation).
int upload_flip(qcd_ctx* ctx, const double* readout, int n_vec, int n, const double** out, cudaStream_t st) {
    *out = nullptr;
    if (!readout) return QCD_OK;
    const size_t cnt = (size_t)n_vec * n * 2;
    if (ctx->flip_staged) CU(cudaEventSynchronize(ctx->flip_staged));      // the previous copy has read the staging buffer
    CU(ctx->flip_host.ensure(cnt * sizeof(double)));
    double* flip = ctx->flip_host.as<double>();
    for (size_t cq = 0; cq < (size_t)n_vec * n; cq++) {
        flip[cq * 2] = readout[cq * 4 + 1];
        flip[cq * 2 + 1] = readout[cq * 4 + 2];
    }
    if (ctx->flip.cap < cnt * sizeof(double)) {
        CU(cudaStreamSynchronize(st));                 // growing frees the old buffer: kernels in flight may still read it
        CU(ctx->flip.ensure(cnt * sizeof(double)));
    }
    CU(cudaMemcpyAsync(ctx->flip.p, flip, cnt * sizeof(double), cudaMemcpyHostToDevice, st));
    if (!ctx->flip_staged) CU(cudaEventCreateWithFlags(&ctx->flip_staged, cudaEventDisableTiming));
    CU(cudaEventRecord(ctx->flip_staged, st));
    *out = ctx->flip.as<double>();
    return QCD_OK;
}

}  // namespace

// ==================================================================================================== C ABI
extern "C" {

int qcd_abi_version(void) { return QCD_ABI_VERSION; }

int qcd_create(int cuda_device, size_t mem_budget_bytes, qcd_ctx** out) {
    if (!out) {
        g_create_error = "qcd_create: out is NULL";
        return QCD_E_INVALID;
    }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("qcd_create: no CUDA device (") + cudaGetErrorString(e) +
                         "); this engine has no CPU fallback";
        return QCD_E_CUDA;
    }
    if (cuda_device < 0 || cuda_device >= count) {
        g_create_error = "qcd_create: device index out of range";
        return QCD_E_INVALID;
    }
    e = cudaSetDevice(cuda_device);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        return QCD_E_CUDA;
    }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, cuda_device);
    if (prop.major < 10) {
        g_create_error = "qcd_create: device is sm_" + std::to_string(prop.major * 10 + prop.minor) +
                         "; libqcdsim is built for sm_100a (B200) only";
        return QCD_E_UNSUPPORTED;
    }
    qcd_ctx* c = new qcd_ctx();
    c->device = cuda_device;
    c->budget = mem_budget_bytes;
    if (const char* v = getenv("QCD_DIRECT_MIN_BITS")) c->direct_min_bits = std::max(9, atoi(v));
    if (const char* v = getenv("QCD_DM_ON_CHIP")) c->dm_on_chip = atoi(v);          // A/B measurements of whole programs
    if (const char* v = getenv("QCD_LEAF_CROSS")) c->leaf_cross = atoi(v) != 0;
    memset(&c->stats, 0, sizeof(c->stats));
    *out = c;
    return QCD_OK;
}

void qcd_destroy(qcd_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (int m = 0; m < 2; m++)
        for (int p = 0; p < 2; p++) ctx->prog[m][p].release();
    DevBuf* bufs[] = {&ctx->states, &ctx->partials, &ctx->chunk_sums, &ctx->block_sums, &ctx->owner, &ctx->keys,
                      &ctx->probs4, &ctx->table, &ctx->offs, &ctx->children, &ctx->counter, &ctx->bsum, &ctx->flip,
                      &ctx->masks, &ctx->scratch, &ctx->slots, &ctx->compose, &ctx->cross_recs};
    for (DevBuf* b : bufs) b->release();
    for (DevBuf* b : {&ctx->interp.ops, &ctx->interp.params, &ctx->interp.sched, &ctx->interp.passes, &ctx->interp.pass_off})
        b->release();
    if (ctx->interp.done) cudaEventDestroy(ctx->interp.done);
    if (ctx->interp.copy) cudaStreamDestroy(ctx->interp.copy);
    for (DevBuf& b : ctx->level_nodes) b.release();
    for (DevBuf& b : ctx->level_shots) b.release();
    for (PinBuf& b : ctx->host_nodes) b.release();
    ctx->flip_host.release();
    if (ctx->flip_staged) cudaEventDestroy(ctx->flip_staged);
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    delete ctx;
}

const char* qcd_last_error(qcd_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int qcd_set_option(qcd_ctx* ctx, const char* key, double value) {
    if (!ctx || !key) return QCD_E_INVALID;
    std::string k(key);
    if (k == "time_sweeps") ctx->time_sweeps = value != 0;
    else if (k == "tile_bits") {
        if (value != 0 && (value < 5 || value > 13)) return fail(ctx, QCD_E_INVALID, "tile_bits must be 0 or in [5, 13]");
        ctx->tile_bits = (int)value;
    } else if (k == "max_chunk") {
        if (value < 1) return fail(ctx, QCD_E_INVALID, "max_chunk must be >= 1");
        ctx->max_chunk = (uint64_t)value;
    } else if (k == "table_cap_log2") {
        if (value < 4 || value > 28) return fail(ctx, QCD_E_INVALID, "table_cap_log2 must be in [4, 28]");
        ctx->table_cap_log2 = (int)value;
    } else if (k == "max_slots") ctx->max_slots = (uint64_t)value;
    else if (k == "use_direct") ctx->use_direct = value != 0;
    else if (k == "direct_min_bits") ctx->direct_min_bits = std::max(9, (int)value);
    else if (k == "nostore_max") ctx->nostore_max = (uint32_t)value;
    else if (k == "leaf_cross") ctx->leaf_cross = value != 0;
    else if (k == "no_slots") ctx->no_slots = value != 0;
    else if (k == "share_prefix") ctx->share_prefix = value != 0;
    else if (k == "hist_row0") ctx->hist_row0 = (uint32_t)value;
    else if (k == "release_memory") {
        // hand the state arena and the per-level scratch back to the device (they are sized from the free memory of the
        // first large run and otherwise kept for the life of the context); the next run allocates what it needs
        cudaSetDevice(ctx->device);
        cudaDeviceSynchronize();
        for (DevBuf* b : {&ctx->states, &ctx->partials, &ctx->chunk_sums, &ctx->block_sums, &ctx->owner, &ctx->keys,
                          &ctx->scratch})
            b->release();
        for (DevBuf& b : ctx->level_nodes) b.release();
        for (DevBuf& b : ctx->level_shots) b.release();
        ctx->cached_budget = 0;
        ctx->budget_age = 0;
    } else if (k == "dm_on_chip") ctx->dm_on_chip = (int)value;
    else if (k == "dm_max_clusters") ctx->dm_max_clusters = (int)value;
    else if (k == "dm_test_refuse") {          // 1: refuse the next cluster launch; 0: forget an earlier refusal
        ctx->dm_test_refuse = value != 0;
        if (value == 0) ctx->dm_cluster_refused = false;
    }
    else return fail(ctx, QCD_E_INVALID, "unknown option " + k);
    return QCD_OK;
}

int qcd_get_stats(qcd_ctx* ctx, qcd_stats* out) {
    if (!ctx || !out) return QCD_E_INVALID;
    *out = ctx->stats;
    return QCD_OK;
}

namespace {

// structural validation (cheap, O(ops)), so that malformed programs fail at upload time; lowering happens once per
// (mode, precision) at the first run
int validate_programs(qcd_ctx* ctx, const qcd_op* ops, const uint32_t* op_offsets, const double* params, size_t n_params,
                      int n_circuits, int n_qubits) {
    for (int c = 0; c < n_circuits; c++) {
        for (uint32_t i = op_offsets[c]; i < op_offsets[c + 1]; i++) {
            const qcd_op& op = ops[i];
            // the position string is only built when something is wrong (this loop sees millions of ops)
            struct Where {
                int c;
                uint32_t i;
                std::string operator+(const char* m) const {
                    return "circuit " + std::to_string(c) + ", op " + std::to_string(i) + ": " + m;
                }
            } where{c, i - op_offsets[c]};
            if (op.opcode >= QCD_OP_COUNT) return fail(ctx, QCD_E_INVALID, where + "unknown opcode");
            bool two = false;
            size_t need = 0;
            switch (op.opcode) {
                case QCD_OP_CX: case QCD_OP_CZ: case QCD_OP_SWAP: two = true; break;
                case QCD_OP_U2Q: two = true; need = 32; break;
                case QCD_OP_KRAUS2Q: two = true; need = 1; break;
                case QCD_OP_PAULI2Q: two = true; need = 16; break;
                case QCD_OP_RX: case QCD_OP_RY: case QCD_OP_RZ: need = 1; break;
                case QCD_OP_U1Q: need = 8; break;
                case QCD_OP_KRAUS1Q: need = 1; break;
                case QCD_OP_PAULI1Q: need = 4; break;
                default: break;
            }
            if (op.q0 >= n_qubits || (two && (op.q1 >= n_qubits || op.q1 == op.q0)))
                return fail(ctx, QCD_E_INVALID, where + "qubit index out of range");
            if (need && (size_t)op.param_off + need > n_params)
                return fail(ctx, QCD_E_INVALID, where + "op parameters run past params[]");
            if (op.opcode == QCD_OP_KRAUS1Q || op.opcode == QCD_OP_KRAUS2Q) {
                const double m = params[op.param_off];
                const size_t per = op.opcode == QCD_OP_KRAUS1Q ? 8 : 32;
                if (!(m >= 1 && m <= 64) || (size_t)op.param_off + 1 + (size_t)m * per > n_params)
                    return fail(ctx, QCD_E_INVALID, where + "bad Kraus operator count");
            }
            if (op.opcode == QCD_OP_PAULI1Q || op.opcode == QCD_OP_PAULI2Q) {
                double tot = 0;
                for (size_t j = 0; j < need; j++) {
                    if (params[op.param_off + j] < 0) return fail(ctx, QCD_E_INVALID, where + "negative Pauli probability");
                    tot += params[op.param_off + j];
                }
                if (std::abs(tot - 1.0) > 1e-9) return fail(ctx, QCD_E_INVALID, where + "Pauli probabilities do not sum to 1");
            }
        }
    }
    return QCD_OK;
}

int check_batch_args(qcd_ctx* ctx, const qcd_op* ops, const uint32_t* op_offsets, const double* params, size_t n_params,
                     int n_circuits, int n_qubits) {
    if (!op_offsets || n_circuits < 1) return fail(ctx, QCD_E_INVALID, "need at least one circuit");
    if (n_qubits < 1 || n_qubits > 30) return fail(ctx, QCD_E_INVALID, "n_qubits must be in [1, 30]");
    const uint32_t n_ops = op_offsets[n_circuits];
    if (n_ops && !ops) return fail(ctx, QCD_E_INVALID, "ops is NULL");
    if (n_params && !params) return fail(ctx, QCD_E_INVALID, "params is NULL");
    for (int c = 0; c < n_circuits; c++)
        if (op_offsets[c] > op_offsets[c + 1]) return fail(ctx, QCD_E_INVALID, "op_offsets must be non-decreasing");
    return QCD_OK;
}

void adopt_programs(qcd_ctx* ctx, int n_circuits, int n_qubits) {
    ctx->n_circuits = n_circuits;
    ctx->n_qubits = n_qubits;
    for (int m = 0; m < 2; m++)
        for (int p = 0; p < 2; p++) ctx->prog[m][p].invalidate();
}

// Expands a noise-free batch on several host threads (contiguous blocks of circuits, merged in order).
int expand_batch(const qcd_op* ops, const uint32_t* op_offsets, const double* params, size_t n_params, int n_circuits,
                 int n_qubits, const double* qubit_t1_t2, const double* op_gate_props, std::vector<qcd_op>& ops_out,
                 std::vector<uint32_t>& offsets_out, std::vector<double>& params_out, std::string& err) {
    int n_thr = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u);
    n_thr = std::max(1, std::min(n_thr, n_circuits / 64));
    struct Part {
        std::vector<qcd_op> ops;
        std::vector<uint32_t> counts;
        std::vector<double> params;
        std::string err;
        int rc = QCD_OK;
    };
    std::vector<Part> parts(n_thr);
    auto work = [&](int t) {
        Part& p = parts[t];
        const int c_lo = (int)((long long)n_circuits * t / n_thr), c_hi = (int)((long long)n_circuits * (t + 1) / n_thr);
        p.rc = expand_device_noise(ops, op_offsets, params, n_params, c_lo, c_hi, n_qubits, qubit_t1_t2, op_gate_props,
                                   p.ops, p.counts, p.params, p.err);
    };
    if (n_thr == 1) {
        work(0);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < n_thr; t++) pool.emplace_back(work, t);
        for (std::thread& th : pool) th.join();
    }
    size_t tot_ops = 0, tot_params = 0;
    for (const Part& p : parts) {
        if (p.rc != QCD_OK) {
            err = p.err;
            return p.rc;
        }
        tot_ops += p.ops.size();
        tot_params += p.params.size();
    }
    if (tot_ops > 0xffffffffull || tot_params > 0xffffffffull) {
        err = "expanded batch exceeds 2^32 ops or parameters: split the upload";
        return QCD_E_UNSUPPORTED;
    }
    ops_out.clear();
    params_out.clear();
    offsets_out.assign(1, 0u);
    ops_out.reserve(tot_ops);
    params_out.reserve(tot_params);
    for (const Part& p : parts) {
        const uint32_t base = (uint32_t)params_out.size();
        for (qcd_op op : p.ops) {
            const bool has_params = (op.opcode >= QCD_OP_RX && op.opcode <= QCD_OP_U1Q) || op.opcode == QCD_OP_U2Q ||
                                    (op.opcode >= QCD_OP_KRAUS1Q && op.opcode <= QCD_OP_PAULI2Q);
            if (has_params) op.param_off += base;
            ops_out.push_back(op);
        }
        params_out.insert(params_out.end(), p.params.begin(), p.params.end());
        for (uint32_t cnt : p.counts) offsets_out.push_back(offsets_out.back() + cnt);
    }
    return QCD_OK;
}

}  // namespace

int qcd_upload_programs(qcd_ctx* ctx, const qcd_op* ops, const uint32_t* op_offsets, const double* params,
                        size_t n_params, int n_circuits, int n_qubits) {
    if (!ctx) return QCD_E_INVALID;
    int rc = check_batch_args(ctx, ops, op_offsets, params, n_params, n_circuits, n_qubits);
    if (rc) return rc;
    if ((rc = validate_programs(ctx, ops, op_offsets, params, n_params, n_circuits, n_qubits))) return rc;
    cudaSetDevice(ctx->device);
    const uint32_t n_ops = op_offsets[n_circuits];
    ctx->ops.assign(ops, ops + n_ops);
    ctx->offsets.assign(op_offsets, op_offsets + n_circuits + 1);
    ctx->params.assign(params, params + n_params);
    ctx->device_noise = false;
    ctx->interp.ready = false;
    ctx->qubit_t1_t2.clear();
    ctx->op_gate_props.clear();
    adopt_programs(ctx, n_circuits, n_qubits);
    return QCD_OK;
}

int qcd_upload_device_noise_programs(qcd_ctx* ctx, const qcd_op* ops, const uint32_t* op_offsets, const double* params,
                                     size_t n_params, int n_circuits, int n_qubits, const double* qubit_t1_t2,
                                     const double* op_gate_props) {
    if (!ctx) return QCD_E_INVALID;
    int rc = check_batch_args(ctx, ops, op_offsets, params, n_params, n_circuits, n_qubits);
    if (rc) return rc;
    if (!qubit_t1_t2 || !op_gate_props) return fail(ctx, QCD_E_INVALID, "qubit_t1_t2 / op_gate_props is NULL");
    if ((rc = validate_programs(ctx, ops, op_offsets, params, n_params, n_circuits, n_qubits))) return rc;
    const uint32_t n_ops = op_offsets[n_circuits];
    // calibration numbers: T1, T2 > 0 (inf allowed); an op with an entry must be a gate
    for (size_t i = 0; i < (size_t)n_circuits * n_qubits * 2; i++)
        if (!(qubit_t1_t2[i] == qubit_t1_t2[i]))
            return fail(ctx, QCD_E_INVALID, "circuit " + std::to_string(i / (2 * n_qubits)) + ": T1 / T2 is NaN");
    for (uint32_t i = 0; i < n_ops; i++) {
        const double ge = op_gate_props[2 * (size_t)i], gl = op_gate_props[2 * (size_t)i + 1];
        if ((ge == ge || gl == gl) && ops[i].opcode >= QCD_OP_KRAUS1Q)
            return fail(ctx, QCD_E_INVALID, "op " + std::to_string(i) + ": a noise op cannot carry a gate error");
    }
    cudaSetDevice(ctx->device);
    ctx->ops.assign(ops, ops + n_ops);
    ctx->offsets.assign(op_offsets, op_offsets + n_circuits + 1);
    ctx->params.assign(params, params + n_params);
    ctx->device_noise = true;
    ctx->interp.ready = false;
    ctx->qubit_t1_t2.assign(qubit_t1_t2, qubit_t1_t2 + (size_t)n_circuits * n_qubits * 2);
    ctx->op_gate_props.assign(op_gate_props, op_gate_props + (size_t)n_ops * 2);
    adopt_programs(ctx, n_circuits, n_qubits);
    return QCD_OK;
}

int qcd_device_gate_superops_1q(int count, const double* unitary, const double* gate_error, const double* gate_length_ns,
                                const double* t1_us, const double* t2_us, double* superops_out) {
    if (count < 0 || !unitary || !gate_error || !gate_length_ns || !t1_us || !t2_us || !superops_out) return QCD_E_INVALID;
    typedef std::complex<double> cdd;
    cdd U[4];
    for (int e = 0; e < 4; e++) U[e] = cdd(unitary[2 * e], unitary[2 * e + 1]);
    // S = sum_j conj(K_j) (x) K_j, index = ket + 2 * bra:  S[(x + 2 y), (x' + 2 y')] = K[x][x'] conj(K[y][y'])
    auto accumulate = [](cdd* S, const cdd* K, double w) {
        for (int y = 0; y < 2; y++)
            for (int x = 0; x < 2; x++)
                for (int yp = 0; yp < 2; yp++)
                    for (int xp = 0; xp < 2; xp++)
                        S[(x + 2 * y) * 4 + (xp + 2 * yp)] += w * K[x * 2 + xp] * std::conj(K[y * 2 + yp]);
    };
    auto matmul4 = [](const cdd* A, const cdd* B, cdd* C) {
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) {
                cdd acc(0, 0);
                for (int k = 0; k < 4; k++) acc += A[i * 4 + k] * B[k * 4 + j];
                C[i * 4 + j] = acc;
            }
    };
    static const cdd PA[4][4] = {{1, 0, 0, 1}, {0, 1, 1, 0}, {0, cdd(0, -1), cdd(0, 1), 0}, {1, 0, 0, -1}};
    cdd SU[16];
    for (cdd& v : SU) v = 0;
    accumulate(SU, U, 1.0);
    GateError ge;
    for (int i = 0; i < count; i++) {
        cdd cur[16], tmp[16], T[16];
        std::copy(SU, SU + 16, cur);
        const double t1 = std::max(t1_us[i], 1e-9), t2 = std::max(t2_us[i], 1e-9);
        if (!(gate_error[i] != gate_error[i] && gate_length_ns[i] != gate_length_ns[i])) {   // an entry exists
            if (!device_gate_error(1, gate_error[i], gate_length_ns[i], &t1, &t2, ge)) return QCD_E_INVALID;
            if (ge.has_pauli) {
                for (cdd& v : T) v = 0;
                for (int p = 0; p < 4; p++) accumulate(T, PA[p], ge.pauli[p]);
                matmul4(T, cur, tmp);
                std::copy(tmp, tmp + 16, cur);
            }
            if (ge.has_relax) {
                for (cdd& v : T) v = 0;
                const int m = (int)ge.kraus[0][0];
                for (int j = 0; j < m; j++) {
                    cdd K[4];
                    for (int e = 0; e < 4; e++) K[e] = cdd(ge.kraus[0][1 + 8 * j + 2 * e], ge.kraus[0][2 + 8 * j + 2 * e]);
                    accumulate(T, K, 1.0);
                }
                matmul4(T, cur, tmp);
                std::copy(tmp, tmp + 16, cur);
            }
        }
        for (int e = 0; e < 16; e++) {
            superops_out[(size_t)i * 32 + 2 * e] = cur[e].real();
            superops_out[(size_t)i * 32 + 2 * e + 1] = cur[e].imag();
        }
    }
    return QCD_OK;
}

int qcd_expand_device_noise(const qcd_op* ops, const uint32_t* op_offsets, const double* params, size_t n_params,
                            int n_circuits, int n_qubits, const double* qubit_t1_t2, const double* op_gate_props,
                            qcd_op* ops_out, size_t ops_cap, uint32_t* offsets_out, double* params_out, size_t params_cap,
                            size_t* ops_needed, size_t* params_needed) {
    if (!op_offsets || n_circuits < 1 || n_qubits < 1 || n_qubits > 30 || !qubit_t1_t2 || !op_gate_props) return QCD_E_INVALID;
    if (op_offsets[n_circuits] && !ops) return QCD_E_INVALID;
    std::vector<qcd_op> e_ops;
    std::vector<uint32_t> e_off;
    std::vector<double> e_par;
    std::string err;
    int rc = expand_batch(ops, op_offsets, params, n_params, n_circuits, n_qubits, qubit_t1_t2, op_gate_props, e_ops, e_off,
                          e_par, err);
    if (rc) {
        g_create_error = err;   // readable through qcd_last_error(NULL)
        return rc;
    }
    if (ops_needed) *ops_needed = e_ops.size();
    if (params_needed) *params_needed = e_par.size();
    if (e_ops.size() > ops_cap || e_par.size() > params_cap || !offsets_out || (!ops_out && !e_ops.empty()) ||
        (!params_out && !e_par.empty()))
        return QCD_E_NOMEM;
    std::copy(e_ops.begin(), e_ops.end(), ops_out);
    std::copy(e_off.begin(), e_off.end(), offsets_out);
    std::copy(e_par.begin(), e_par.end(), params_out);
    return QCD_OK;
}

int qcd_run_sv_trajectories(qcd_ctx* ctx, int precision, uint64_t shots_per_circuit, uint64_t seed,
                            uint64_t shot_begin, uint64_t shot_end, uint32_t first_circuit, const double* readout,
                            uint32_t* hist, uint32_t* outcomes, uint32_t flags, void* stream) {
    if (!ctx) return QCD_E_INVALID;
    if (precision != 32 && precision != 64) return fail(ctx, QCD_E_INVALID, "precision must be 32 or 64");
    if (ctx->n_circuits <= 0) return fail(ctx, QCD_E_STATE, "no programs uploaded (call qcd_upload_programs first)");
    if (shots_per_circuit < 1 || shots_per_circuit > 0xffffffffull)
        return fail(ctx, QCD_E_INVALID, "shots_per_circuit must be in [1, 2^32)");
    if (shot_begin > shot_end || shot_end > shots_per_circuit * (uint64_t)ctx->n_circuits)
        return fail(ctx, QCD_E_INVALID, "shot range outside [0, n_circuits * shots_per_circuit]");
    if (!hist && !outcomes) return fail(ctx, QCD_E_INVALID, "hist and outcomes are both NULL");
    if (hist && shot_begin < shot_end && (uint64_t)ctx->hist_row0 > shot_begin / shots_per_circuit)
        return fail(ctx, QCD_E_INVALID, "option hist_row0 lies beyond the first circuit of the shot range");
    cudaSetDevice(ctx->device);
    memset(&ctx->stats, 0, sizeof(ctx->stats));
    ctx->ev_used = 0;
    if (shot_begin == shot_end) return QCD_OK;
    DevProg* dp = nullptr;
    int rc = get_program(ctx, MODE_SV, precision, &dp);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == 64)
        return run_sv<double>(ctx, dp, shots_per_circuit, seed, shot_begin, shot_end, first_circuit, readout, hist,
                              outcomes, flags, st);
    return run_sv<float>(ctx, dp, shots_per_circuit, seed, shot_begin, shot_end, first_circuit, readout, hist, outcomes,
                         flags, st);
}

int qcd_run_density_matrix(qcd_ctx* ctx, int precision, int circuit_begin, int circuit_end, void* probs_out,
                           void* stream) {
    if (!ctx) return QCD_E_INVALID;
    if (precision != 32 && precision != 64) return fail(ctx, QCD_E_INVALID, "precision must be 32 or 64");
    if (ctx->n_circuits <= 0) return fail(ctx, QCD_E_STATE, "no programs uploaded (call qcd_upload_programs first)");
    if (circuit_begin < 0 || circuit_begin > circuit_end || circuit_end > ctx->n_circuits)
        return fail(ctx, QCD_E_INVALID, "circuit range outside [0, n_circuits]");
    if (!probs_out) return fail(ctx, QCD_E_INVALID, "probs_out is NULL");
    if (ctx->n_qubits > 15) return fail(ctx, QCD_E_UNSUPPORTED, "density-matrix mode supports at most 15 qubits");
    cudaSetDevice(ctx->device);
    memset(&ctx->stats, 0, sizeof(ctx->stats));
    ctx->ev_used = 0;
    if (circuit_begin == circuit_end) return QCD_OK;
    cudaStream_t st = (cudaStream_t)stream;
    for (int attempt = 0; attempt < 2; attempt++) {
        DevProg* dp = nullptr;
        int rc = interp_eligible(ctx, precision) ? prepare_interp(ctx) : get_program(ctx, MODE_DM, precision, &dp);
        if (rc) return rc;
        rc = precision == 64 ? run_dm<double>(ctx, dp, circuit_begin, circuit_end, probs_out, st)
                             : run_dm<float>(ctx, dp, circuit_begin, circuit_end, probs_out, st);
        if (!(rc == QCD_E_UNSUPPORTED && ctx->dm_cluster_refused && dp == nullptr)) return rc;
    }
    return QCD_E_UNSUPPORTED;
}

int qcd_run_density_matrix_settings(qcd_ctx* ctx, int precision, int circuit_begin, int circuit_end, int n_settings,
                                    const uint8_t* setting_qubits, const double* setting_superops, void* probs_out,
                                    void* stream) {
    if (!ctx) return QCD_E_INVALID;
    if (precision != 32 && precision != 64) return fail(ctx, QCD_E_INVALID, "precision must be 32 or 64");
    if (ctx->n_circuits <= 0) return fail(ctx, QCD_E_STATE, "no programs uploaded (call qcd_upload_programs first)");
    if (circuit_begin < 0 || circuit_begin > circuit_end || circuit_end > ctx->n_circuits)
        return fail(ctx, QCD_E_INVALID, "circuit range outside [0, n_circuits]");
    if (n_settings < 1 || n_settings > 4096) return fail(ctx, QCD_E_INVALID, "n_settings must be in [1, 4096]");
    if (!probs_out || !setting_qubits || !setting_superops) return fail(ctx, QCD_E_INVALID, "NULL argument");
    if (ctx->n_qubits > 15) return fail(ctx, QCD_E_UNSUPPORTED, "density-matrix mode supports at most 15 qubits");
    const size_t cnt = (size_t)(circuit_end - circuit_begin) * n_settings;
    for (size_t i = 0; i < cnt; i++)
        if (setting_qubits[i] != 0xFF && setting_qubits[i] >= ctx->n_qubits)
            return fail(ctx, QCD_E_INVALID, "setting " + std::to_string(i) + ": qubit index out of range");
    cudaSetDevice(ctx->device);
    memset(&ctx->stats, 0, sizeof(ctx->stats));
    ctx->ev_used = 0;
    if (circuit_begin == circuit_end) return QCD_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = QCD_E_UNSUPPORTED;
    for (int attempt = 0; attempt < 2; attempt++) {      // second attempt: the lowered route, if the cluster launch was refused
        DevProg* dp = nullptr;
        rc = interp_eligible(ctx, precision) ? prepare_interp(ctx) : get_program(ctx, MODE_DM, precision, &dp);
        if (rc) return rc;
        // the per-batch host -> device copies of the setting tables read caller memory asynchronously: drain before returning
        if (precision == 64)
            rc = run_dm<double>(ctx, dp, circuit_begin, circuit_end, probs_out, st, n_settings, setting_qubits, setting_superops);
        else
            rc = run_dm<float>(ctx, dp, circuit_begin, circuit_end, probs_out, st, n_settings, setting_qubits, setting_superops);
        if (!(rc == QCD_E_UNSUPPORTED && ctx->dm_cluster_refused && dp == nullptr)) break;
    }
    if (rc) return rc;
    CU(cudaStreamSynchronize(st));
    return QCD_OK;
}

int qcd_run_density_matrix_group(qcd_ctx* ctx, int precision, int circuit_begin, int circuit_end,
                                 const uint32_t* adjacency, double* expvals_out, void* stream) {
    if (!ctx) return QCD_E_INVALID;
    if (precision != 32 && precision != 64) return fail(ctx, QCD_E_INVALID, "precision must be 32 or 64");
    if (ctx->n_circuits <= 0) return fail(ctx, QCD_E_STATE, "no programs uploaded (call qcd_upload_programs first)");
    if (circuit_begin < 0 || circuit_begin > circuit_end || circuit_end > ctx->n_circuits)
        return fail(ctx, QCD_E_INVALID, "circuit range outside [0, n_circuits]");
    if (!adjacency || !expvals_out) return fail(ctx, QCD_E_INVALID, "NULL argument");
    if (ctx->n_qubits > 15) return fail(ctx, QCD_E_UNSUPPORTED, "density-matrix mode supports at most 15 qubits");
    for (int k = 0; k < ctx->n_qubits; k++) {
        if (adjacency[k] >> ctx->n_qubits || (adjacency[k] >> k) & 1u)
            return fail(ctx, QCD_E_INVALID, "adjacency row " + std::to_string(k) + " names a vertex >= n_qubits or itself");
        for (int j = 0; j < ctx->n_qubits; j++)
            if (((adjacency[k] >> j) & 1u) != ((adjacency[j] >> k) & 1u))
                return fail(ctx, QCD_E_INVALID, "adjacency is not symmetric (the generators would not commute)");
    }
    cudaSetDevice(ctx->device);
    memset(&ctx->stats, 0, sizeof(ctx->stats));
    ctx->ev_used = 0;
    if (circuit_begin == circuit_end) return QCD_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = QCD_E_UNSUPPORTED;
    for (int attempt = 0; attempt < 2; attempt++) {
        DevProg* dp = nullptr;
        rc = interp_eligible(ctx, precision) ? prepare_interp(ctx) : get_program(ctx, MODE_DM, precision, &dp);
        if (rc) return rc;
        if (precision == 64)
            rc = run_dm<double>(ctx, dp, circuit_begin, circuit_end, expvals_out, st, 0, nullptr, nullptr, adjacency);
        else
            rc = run_dm<float>(ctx, dp, circuit_begin, circuit_end, expvals_out, st, 0, nullptr, nullptr, adjacency);
        if (!(rc == QCD_E_UNSUPPORTED && ctx->dm_cluster_refused && dp == nullptr)) break;
    }
    if (rc) return rc;
    CU(cudaStreamSynchronize(st));       // the adjacency rows were read from caller memory asynchronously
    return QCD_OK;
}

int qcd_stabilizer_group(qcd_ctx* ctx, const uint32_t* gen_x, const uint32_t* gen_z, int n_gens, uint64_t begin,
                         uint64_t end, uint32_t* x_out, uint32_t* z_out, uint8_t* sign_out, void* stream) {
    if (!ctx) return QCD_E_INVALID;
    if (!gen_x || !gen_z || !x_out || !z_out || !sign_out) return fail(ctx, QCD_E_INVALID, "NULL argument");
    if (n_gens < 1 || n_gens > 32) return fail(ctx, QCD_E_INVALID, "n_gens must be in [1, 32]");
    if (begin > end || end > ((uint64_t)1 << n_gens)) return fail(ctx, QCD_E_INVALID, "element range outside [0, 2^n_gens]");
    if (begin == end) return QCD_OK;
    cudaSetDevice(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    CU(ctx->masks.ensure(64 * sizeof(uint32_t)));
    uint32_t* g = ctx->masks.as<uint32_t>();
    CU(cudaMemcpyAsync(g, gen_x, (size_t)n_gens * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(g + 32, gen_z, (size_t)n_gens * 4, cudaMemcpyHostToDevice, st));
    launch_stabilizer_group(g, g + 32, n_gens, begin, end - begin, x_out, z_out, sign_out, st);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));       // generator masks were read from caller memory asynchronously
    return QCD_OK;
}

int qcd_apply_readout_probs(qcd_ctx* ctx, int precision, void* probs, int n_vectors, int n_qubits,
                            const double* readout, void* stream) {
    if (!ctx) return QCD_E_INVALID;
    if (precision != 32 && precision != 64) return fail(ctx, QCD_E_INVALID, "precision must be 32 or 64");
    if (!probs || !readout || n_vectors < 0 || n_qubits < 1 || n_qubits > 30)
        return fail(ctx, QCD_E_INVALID, "bad arguments to qcd_apply_readout_probs");
    if (n_vectors == 0) return QCD_OK;
    cudaSetDevice(ctx->device);
    std::vector<double> conf(readout, readout + (size_t)n_vectors * n_qubits * 4);
    int rc = upload_vec(ctx, ctx->flip, conf);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == 64) launch_readout_probs<double>((double*)probs, n_vectors, n_qubits, ctx->flip.as<double>(), st);
    else launch_readout_probs<float>((float*)probs, n_vectors, n_qubits, ctx->flip.as<double>(), st);
    CU(cudaGetLastError());
    return QCD_OK;
}

int qcd_sample_from_probs(qcd_ctx* ctx, int precision, const void* probs, int n_vectors, int n_qubits,
                          uint64_t shot_begin, uint64_t shot_end, uint64_t seed, uint32_t first_circuit,
                          const double* readout, uint32_t* hist, void* stream) {
    if (!ctx) return QCD_E_INVALID;
    if (precision != 32 && precision != 64) return fail(ctx, QCD_E_INVALID, "precision must be 32 or 64");
    if (!probs || !hist || n_vectors < 0 || n_qubits < 1 || n_qubits > 30 || shot_begin > shot_end)
        return fail(ctx, QCD_E_INVALID, "bad arguments to qcd_sample_from_probs");
    if (n_vectors == 0 || shot_begin == shot_end) return QCD_OK;
    cudaSetDevice(ctx->device);
    const int n = n_qubits;
    const int chunk_log = std::min(n, CHUNK_LOG);
    const uint32_t n_chunks = 1u << (n - chunk_log);
    const uint32_t n_blocks = (n_chunks + (1u << SCAN_BLOCK_LOG) - 1) >> SCAN_BLOCK_LOG;
    CU(ctx->chunk_sums.ensure((size_t)n_vectors * n_chunks * 8));
    CU(ctx->block_sums.ensure((size_t)n_vectors * n_blocks * 8));
    const double* flip = nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = upload_flip(ctx, readout, n_vectors, n, &flip, st);
    if (rc) return rc;
    double* cs = ctx->chunk_sums.as<double>();
    double* bs = ctx->block_sums.as<double>();
    if (precision == 64) {
        launch_prob_chunks<double>((const double*)probs, n_vectors, n, cs, n_chunks, st);
        launch_scan_vectors(n_vectors, cs, n_chunks, bs, n_blocks, n_chunks, st);
        launch_sample_probs<double>((const double*)probs, n_vectors, n, cs, n_chunks, bs, n_blocks, n_chunks, shot_begin,
                                    shot_end - shot_begin, seed, first_circuit, flip, hist, st);
    } else {
        launch_prob_chunks<float>((const float*)probs, n_vectors, n, cs, n_chunks, st);
        launch_scan_vectors(n_vectors, cs, n_chunks, bs, n_blocks, n_chunks, st);
        launch_sample_probs<float>((const float*)probs, n_vectors, n, cs, n_chunks, bs, n_blocks, n_chunks, shot_begin,
                                   shot_end - shot_begin, seed, first_circuit, flip, hist, st);
    }
    CU(cudaGetLastError());
    return QCD_OK;
}

int qcd_parity_counts(qcd_ctx* ctx, const uint32_t* hist, int n_vectors, int n_qubits, const uint32_t* masks,
                      int n_masks, int per_vector_masks, int64_t* signed_sums, int64_t* totals, void* stream) {
    if (!ctx) return QCD_E_INVALID;
    if (!hist || !masks || !signed_sums || !totals || n_vectors < 0 || n_masks < 1 || n_qubits < 1 || n_qubits > 30)
        return fail(ctx, QCD_E_INVALID, "bad arguments to qcd_parity_counts");
    if (n_vectors == 0) return QCD_OK;
    cudaSetDevice(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nm = (size_t)n_masks * (per_vector_masks ? n_vectors : 1);
    std::vector<uint32_t> mk(masks, masks + nm);
    int rc = upload_vec(ctx, ctx->masks, mk);
    if (rc) return rc;
    CU(cudaMemsetAsync(signed_sums, 0, (size_t)n_vectors * n_masks * 8, st));
    CU(cudaMemsetAsync(totals, 0, (size_t)n_vectors * 8, st));
    launch_parity_counts(hist, n_vectors, n_qubits, ctx->masks.as<uint32_t>(), n_masks, per_vector_masks,
                         (long long*)signed_sums, (long long*)totals, st);
    CU(cudaGetLastError());
    return QCD_OK;
}

int qcd_parity_outcomes(qcd_ctx* ctx, const uint32_t* outcomes, int n_vectors, uint64_t shots_per_vector,
                        const uint32_t* masks, int n_masks, int per_vector_masks, int64_t* signed_sums, void* stream) {
    if (!ctx) return QCD_E_INVALID;
    if (!outcomes || !masks || !signed_sums || n_vectors < 0 || n_masks < 1 || shots_per_vector > 0x7fffffffull)
        return fail(ctx, QCD_E_INVALID, "bad arguments to qcd_parity_outcomes");
    if (n_vectors == 0) return QCD_OK;
    cudaSetDevice(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nm = (size_t)n_masks * (per_vector_masks ? n_vectors : 1);
    std::vector<uint32_t> mk(masks, masks + nm);
    int rc = upload_vec(ctx, ctx->masks, mk);
    if (rc) return rc;
    CU(cudaMemsetAsync(signed_sums, 0, (size_t)n_vectors * n_masks * 8, st));
    if (shots_per_vector)
        launch_parity_outcomes(outcomes, n_vectors, shots_per_vector, ctx->masks.as<uint32_t>(), n_masks,
                               per_vector_masks, (long long*)signed_sums, st);
    CU(cudaGetLastError());
    return QCD_OK;
}

int qcd_parity_probs(qcd_ctx* ctx, int precision, const void* probs, int n_vectors, int n_qubits,
                     const uint32_t* masks, int n_masks, int per_vector_masks, double* expvals, void* stream) {
    if (!ctx) return QCD_E_INVALID;
    if (precision != 32 && precision != 64) return fail(ctx, QCD_E_INVALID, "precision must be 32 or 64");
    if (!probs || !masks || !expvals || n_vectors < 0 || n_masks < 1 || n_qubits < 1 || n_qubits > 30)
        return fail(ctx, QCD_E_INVALID, "bad arguments to qcd_parity_probs");
    if (n_vectors == 0) return QCD_OK;
    cudaSetDevice(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nm = (size_t)n_masks * (per_vector_masks ? n_vectors : 1);
    std::vector<uint32_t> mk(masks, masks + nm);
    int rc = upload_vec(ctx, ctx->masks, mk);
    if (rc) return rc;
    uint32_t n_part = (uint32_t)std::min<uint64_t>(((uint64_t)1 << n_qubits) / 256 + 1, 64);
    CU(ctx->scratch.ensure((size_t)n_vectors * n_masks * n_part * 8));
    if (precision == 64)
        launch_parity_probs<double>((const double*)probs, n_vectors, n_qubits, ctx->masks.as<uint32_t>(), n_masks,
                                    per_vector_masks, ctx->scratch.as<double>(), n_part, expvals, st);
    else
        launch_parity_probs<float>((const float*)probs, n_vectors, n_qubits, ctx->masks.as<uint32_t>(), n_masks,
                                   per_vector_masks, ctx->scratch.as<double>(), n_part, expvals, st);
    CU(cudaGetLastError());
    return QCD_OK;
}

int qcd_debug_dm_schedule(const qcd_op* ops, uint32_t n_ops, const double* params, size_t n_params, int n_qubits,
                          uint8_t* pass_qubits, uint32_t* pass_n_ops, size_t pass_cap, uint32_t* order, size_t order_cap,
                          size_t* n_passes, size_t* n_order) {
    if ((!ops && n_ops) || !n_passes || !n_order || n_qubits < 1 || n_qubits > 15) return QCD_E_INVALID;
    std::vector<DmPass> passes;
    std::vector<uint32_t> sched;
    std::string err;
    int rc = dm_interp_schedule(ops, n_ops, 0, params, n_params, n_qubits, passes, sched, err);
    if (rc != QCD_OK) {
        g_create_error = "qcd_debug_dm_schedule: " + err;
        return rc;
    }
    *n_passes = passes.size();
    *n_order = sched.size();
    if (passes.size() > pass_cap || sched.size() > order_cap || (passes.size() && (!pass_qubits || !pass_n_ops)) ||
        (sched.size() && !order))
        return QCD_E_NOMEM;
    for (size_t p = 0; p < passes.size(); p++) {
        pass_qubits[2 * p] = passes[p].q0;
        pass_qubits[2 * p + 1] = passes[p].q1;
        pass_n_ops[p] = passes[p].n_ops;
    }
    std::copy(sched.begin(), sched.end(), order);
    return QCD_OK;
}

int qcd_debug_schedule_json(const qcd_op* ops, uint32_t n_ops, const double* params, size_t n_params, int n_qubits,
                            int mode, int precision, int tile_bits, char* buf, size_t cap, size_t* needed) {
    if (n_qubits < 1 || n_qubits > 30 || (mode != 0 && mode != 1)) return QCD_E_INVALID;
    Program P;
    P.n_qubits = n_qubits;
    P.mode = mode;
    P.n_bits = mode == MODE_DM ? 2 * n_qubits : n_qubits;
    P.k_max = tile_bits > 0 ? tile_bits : default_tile_bits(precision);
    std::string err;
    int rc = compile_circuit(P, ops, n_ops, params, n_params, err);
    std::string js = rc == QCD_OK ? program_to_json(P, 0) : ("{\"error\":\"" + err + "\"}");
    if (needed) *needed = js.size() + 1;
    if (!buf || cap < js.size() + 1) return QCD_E_NOMEM;
    memcpy(buf, js.c_str(), js.size() + 1);
    return rc;
}

int qcd_debug_sibling_runs_json(const qcd_op* ops, const uint32_t* op_offsets, const double* params, size_t n_params,
                                int n_circuits, int n_qubits, int precision, char* buf, size_t cap, size_t* needed) {
    if (n_qubits < 1 || n_qubits > 30 || n_circuits < 1 || !ops || !op_offsets) return QCD_E_INVALID;
    DevProg dp;
    Program& P = dp.P;
    P.n_qubits = n_qubits;
    P.mode = MODE_SV;
    P.n_bits = n_qubits;
    P.k_max = default_tile_bits(precision);
    std::string err;
    int rc = QCD_OK;
    for (int c = 0; c < n_circuits && rc == QCD_OK; c++)
        rc = compile_circuit(P, ops + op_offsets[c], op_offsets[c + 1] - op_offsets[c], params, n_params, err);
    std::string js;
    if (rc != QCD_OK) {
        js = "{\"error\":\"" + err + "\"}";
    } else {
        dp.member_bits = assign_runs(P, MODE_SV);
        DirectSweep zero;
        memset(&zero, 0, sizeof(zero));
        dp.direct_host.assign(P.sweeps.size() + P.passes.size(), zero);
        for (size_t si = 0; si < P.sweeps.size(); si++) {
            const Sweep& S = P.sweeps[si];
            if (S.pass_end - S.pass_begin == 1) describe_pass(P, dp.direct_host[si], S, S.pass_begin, S.flags, true);
        }
        build_cross(dp, [&](size_t off) -> std::complex<double> {       // the pool in device precision, as the engine sees it
            const cd v = P.pool[off];
            return precision == 64 ? v : cd((double)(float)v.real(), (double)(float)v.imag());
        });
        js = "{\"runs\":[";
        char tmp[128];
        for (size_t r = 0; r < dp.cross_runs.size(); r++) {
            const DevProg::CrossRun& R = dp.cross_runs[r];
            const DirectSweep& D = dp.direct_host[R.desc];
            snprintf(tmp, sizeof(tmp), "%s{\"lead\":%u,\"members\":%u,\"n_high\":%d,\"reg_bits\":[%d,%d,%d,%d],\"nops\":%d,\"kbit\":[",
                     r ? "," : "", R.lead, R.n_members, R.n_high, D.rbit[0], D.rbit[1], D.rbit[2], D.rbit[3], D.nops);
            js += tmp;
            for (size_t j = 0; j < R.kbit.size(); j++) js += (j ? "," : "") + std::to_string((int)R.kbit[j]);
            js += "],\"u\":[";
            for (size_t j = 0; j < R.u.size(); j++) {
                snprintf(tmp, sizeof(tmp), "%s%.17g", j ? "," : "", R.u[j]);
                js += tmp;
            }
            js += "]}";
        }
        js += "]}";
    }
    if (needed) *needed = js.size() + 1;
    if (!buf || cap < js.size() + 1) return QCD_E_NOMEM;
    memcpy(buf, js.c_str(), js.size() + 1);
    return rc;
}

}  // extern "C"

