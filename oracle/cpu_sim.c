/* cpu_sim.c -- compiled CPU restatement of the simulation the reference delegates to qiskit-aer's `statevector`
 * method (samplers.py:308-326 -> qk.execute -> Aer C++): complex128, gate-at-a-time sweeps over the 2^n amplitudes,
 * ONE Kraus trajectory per shot (no sharing between shots), inverse-CDF `sample_measure`, OpenMP over shots the way
 * Aer parallelises shots.  TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py): it is the timed
 * "cpu_baseline" / `--impl reference` arm of bench.py and a cross-check of the numpy oracle (oracle/sim.py), never
 * part of the product path.  qiskit-aer itself is not vendored in the reference and not installable here, so this
 * is a port of its published algorithm, not the reference's own binary ("kind": "port").
 *
 * Circuit encoding: the qcd_op array of include/qcdsim.h.  RNG: Philox4x32-10 with the stream convention of
 * oracle/philox.py, so outcomes are comparable shot for shot with oracle/sim.py and with the CUDA engine. */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/qcdsim.h"

typedef double complex cd;

static void philox(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1,
                 n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
static double u53(uint64_t seed, uint32_t circuit, uint64_t shot, uint32_t stream) {
    uint32_t c[4] = {(uint32_t)shot, (uint32_t)(shot >> 32), stream, circuit};
    philox(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    uint64_t x = ((uint64_t)c[1] << 32) | c[0];
    return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}

static void apply1(cd* psi, uint64_t N, int q, const cd m[4]) {
    const uint64_t s = 1ull << q;
    for (uint64_t base = 0; base < N; base += 2 * s)
        for (uint64_t i = base; i < base + s; i++) {
            cd a0 = psi[i], a1 = psi[i + s];
            psi[i] = m[0] * a0 + m[1] * a1;
            psi[i + s] = m[2] * a0 + m[3] * a1;
        }
}
/* 4x4 on (q0, q1), basis index = bit(q0) + 2 bit(q1) */
static void apply2(cd* psi, uint64_t N, int q0, int q1, const cd m[16]) {
    const uint64_t s0 = 1ull << q0, s1 = 1ull << q1;
    for (uint64_t i = 0; i < N; i++) {
        if (i & (s0 | s1)) continue;
        cd a[4] = {psi[i], psi[i | s0], psi[i | s1], psi[i | s0 | s1]}, b[4];
        for (int r = 0; r < 4; r++) b[r] = m[4 * r] * a[0] + m[4 * r + 1] * a[1] + m[4 * r + 2] * a[2] + m[4 * r + 3] * a[3];
        psi[i] = b[0]; psi[i | s0] = b[1]; psi[i | s1] = b[2]; psi[i | s0 | s1] = b[3];
    }
}
static double norm_after1(const cd* psi, uint64_t N, int q, const cd m[4]) {
    const uint64_t s = 1ull << q;
    double acc = 0;
    for (uint64_t base = 0; base < N; base += 2 * s)
        for (uint64_t i = base; i < base + s; i++) {
            cd a0 = psi[i], a1 = psi[i + s];
            cd b0 = m[0] * a0 + m[1] * a1, b1 = m[2] * a0 + m[3] * a1;
            acc += creal(b0) * creal(b0) + cimag(b0) * cimag(b0) + creal(b1) * creal(b1) + cimag(b1) * cimag(b1);
        }
    return acc;
}
static double norm_after2(const cd* psi, uint64_t N, int q0, int q1, const cd m[16]) {
    const uint64_t s0 = 1ull << q0, s1 = 1ull << q1;
    double acc = 0;
    for (uint64_t i = 0; i < N; i++) {
        if (i & (s0 | s1)) continue;
        cd a[4] = {psi[i], psi[i | s0], psi[i | s1], psi[i | s0 | s1]};
        for (int r = 0; r < 4; r++) {
            cd b = m[4 * r] * a[0] + m[4 * r + 1] * a[1] + m[4 * r + 2] * a[2] + m[4 * r + 3] * a[3];
            acc += creal(b) * creal(b) + cimag(b) * cimag(b);
        }
    }
    return acc;
}
static void scale(cd* psi, uint64_t N, double f) {
    for (uint64_t i = 0; i < N; i++) psi[i] *= f;
}

/* ---- "wide" mode: ONE trajectory at a time, the loops over the amplitudes split over the OpenMP threads.  Used by the
 * parity tests at n >= 24, where a handful of shots must finish in seconds (the per-shot mode above keeps one state
 * per thread and is what the timed CPU baseline runs).  Same arithmetic per amplitude; only the association order of
 * the norm reductions differs. */
static int g_wide = 0;
static inline uint64_t spread1(uint64_t k, int q) { return ((k >> q) << (q + 1)) | (k & ((1ull << q) - 1)); }
static void apply1_wide(cd* psi, uint64_t N, int q, const cd m[4]) {
    const uint64_t s = 1ull << q;
#pragma omp parallel for schedule(static)
    for (long long k = 0; k < (long long)(N >> 1); k++) {
        const uint64_t i = spread1((uint64_t)k, q);
        cd a0 = psi[i], a1 = psi[i + s];
        psi[i] = m[0] * a0 + m[1] * a1;
        psi[i + s] = m[2] * a0 + m[3] * a1;
    }
}
static void apply2_wide(cd* psi, uint64_t N, int q0, int q1, const cd m[16]) {
    const uint64_t s0 = 1ull << q0, s1 = 1ull << q1;
    const int lo = q0 < q1 ? q0 : q1, hi = q0 < q1 ? q1 : q0;
#pragma omp parallel for schedule(static)
    for (long long k = 0; k < (long long)(N >> 2); k++) {
        const uint64_t i = spread1(spread1((uint64_t)k, lo), hi);
        cd a[4] = {psi[i], psi[i | s0], psi[i | s1], psi[i | s0 | s1]}, b[4];
        for (int r = 0; r < 4; r++) b[r] = m[4 * r] * a[0] + m[4 * r + 1] * a[1] + m[4 * r + 2] * a[2] + m[4 * r + 3] * a[3];
        psi[i] = b[0]; psi[i | s0] = b[1]; psi[i | s1] = b[2]; psi[i | s0 | s1] = b[3];
    }
}
static double norm_after1_wide(const cd* psi, uint64_t N, int q, const cd m[4]) {
    const uint64_t s = 1ull << q;
    double acc = 0;
#pragma omp parallel for schedule(static) reduction(+ : acc)
    for (long long k = 0; k < (long long)(N >> 1); k++) {
        const uint64_t i = spread1((uint64_t)k, q);
        cd a0 = psi[i], a1 = psi[i + s];
        cd b0 = m[0] * a0 + m[1] * a1, b1 = m[2] * a0 + m[3] * a1;
        acc += creal(b0) * creal(b0) + cimag(b0) * cimag(b0) + creal(b1) * creal(b1) + cimag(b1) * cimag(b1);
    }
    return acc;
}
static double norm_after2_wide(const cd* psi, uint64_t N, int q0, int q1, const cd m[16]) {
    const uint64_t s0 = 1ull << q0, s1 = 1ull << q1;
    const int lo = q0 < q1 ? q0 : q1, hi = q0 < q1 ? q1 : q0;
    double acc = 0;
#pragma omp parallel for schedule(static) reduction(+ : acc)
    for (long long k = 0; k < (long long)(N >> 2); k++) {
        const uint64_t i = spread1(spread1((uint64_t)k, lo), hi);
        cd a[4] = {psi[i], psi[i | s0], psi[i | s1], psi[i | s0 | s1]};
        for (int r = 0; r < 4; r++) {
            cd b = m[4 * r] * a[0] + m[4 * r + 1] * a[1] + m[4 * r + 2] * a[2] + m[4 * r + 3] * a[3];
            acc += creal(b) * creal(b) + cimag(b) * cimag(b);
        }
    }
    return acc;
}
static void scale_wide(cd* psi, uint64_t N, double f) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)N; i++) psi[i] *= f;
}
#define APPLY1(psi, N, q, m) (g_wide ? apply1_wide(psi, N, q, m) : apply1(psi, N, q, m))
#define APPLY2(psi, N, a, b, m) (g_wide ? apply2_wide(psi, N, a, b, m) : apply2(psi, N, a, b, m))
#define NORM1(psi, N, q, m) (g_wide ? norm_after1_wide(psi, N, q, m) : norm_after1(psi, N, q, m))
#define NORM2(psi, N, a, b, m) (g_wide ? norm_after2_wide(psi, N, a, b, m) : norm_after2(psi, N, a, b, m))
#define SCALE(psi, N, f) (g_wide ? scale_wide(psi, N, f) : scale(psi, N, f))
static int pick(const double* p, int m, double u) {
    double tot = 0, cum = 0;
    for (int j = 0; j < m; j++) tot += p[j];
    const double target = u * tot;
    for (int j = 0; j < m; j++) {
        cum += p[j];
        if (cum > target) return j;
    }
    return m - 1;
}
static void mat_from(const double* p, int n, cd* out) {
    for (int e = 0; e < n; e++) out[e] = p[2 * e] + I * p[2 * e + 1];
}

static const double SQ2 = 0.70710678118654752440;

static uint64_t run_one(cd* psi, const qcd_op* ops, uint32_t n_ops, const double* params, int n, uint64_t seed,
                        uint32_t circuit, uint64_t shot) {
    const uint64_t N = 1ull << n;
    memset(psi, 0, N * sizeof(cd));
    psi[0] = 1.0;
    uint32_t stream = 0;
    static const cd PAULI[4][4] = {{1, 0, 0, 1}, {0, 1, 1, 0}, {0, -I, I, 0}, {1, 0, 0, -1}};
    for (uint32_t i = 0; i < n_ops; i++) {
        const qcd_op op = ops[i];
        const double* p = params + op.param_off;
        const int a = op.q0, b = op.q1;
        cd m[16];
        switch (op.opcode) {
            case QCD_OP_ID: break;
            case QCD_OP_H: { cd h[4] = {SQ2, SQ2, SQ2, -SQ2}; APPLY1(psi, N, a, h); break; }
            case QCD_OP_X: APPLY1(psi, N, a, PAULI[1]); break;
            case QCD_OP_Y: APPLY1(psi, N, a, PAULI[2]); break;
            case QCD_OP_Z: APPLY1(psi, N, a, PAULI[3]); break;
            case QCD_OP_S: { cd g[4] = {1, 0, 0, I}; APPLY1(psi, N, a, g); break; }
            case QCD_OP_SDG: { cd g[4] = {1, 0, 0, -I}; APPLY1(psi, N, a, g); break; }
            case QCD_OP_T: { cd g[4] = {1, 0, 0, SQ2 + I * SQ2}; APPLY1(psi, N, a, g); break; }
            case QCD_OP_TDG: { cd g[4] = {1, 0, 0, SQ2 - I * SQ2}; APPLY1(psi, N, a, g); break; }
            case QCD_OP_RX: { double c = cos(p[0] / 2), s = sin(p[0] / 2); cd g[4] = {c, -I * s, -I * s, c}; APPLY1(psi, N, a, g); break; }
            case QCD_OP_RY: { double c = cos(p[0] / 2), s = sin(p[0] / 2); cd g[4] = {c, -s, s, c}; APPLY1(psi, N, a, g); break; }
            case QCD_OP_RZ: { double c = cos(p[0] / 2), s = sin(p[0] / 2); cd g[4] = {c - I * s, 0, 0, c + I * s}; APPLY1(psi, N, a, g); break; }
            case QCD_OP_U1Q: mat_from(p, 4, m); APPLY1(psi, N, a, m); break;
            case QCD_OP_CX: { cd g[16] = {1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0}; APPLY2(psi, N, a, b, g); break; }
            case QCD_OP_CZ: { cd g[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1}; APPLY2(psi, N, a, b, g); break; }
            case QCD_OP_SWAP: { cd g[16] = {1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1}; APPLY2(psi, N, a, b, g); break; }
            case QCD_OP_U2Q: mat_from(p, 16, m); APPLY2(psi, N, a, b, m); break;
            case QCD_OP_KRAUS1Q: case QCD_OP_RESET: {
                double reset_params[17] = {2, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0};
                if (op.opcode == QCD_OP_RESET) p = reset_params;
                const int cnt = (int)p[0];
                double pr[64];
                for (int j = 0; j < cnt; j++) { mat_from(p + 1 + 8 * j, 4, m); pr[j] = NORM1(psi, N, a, m); }
                const int j = pick(pr, cnt, u53(seed, circuit, shot, stream++));
                mat_from(p + 1 + 8 * j, 4, m);
                APPLY1(psi, N, a, m);
                SCALE(psi, N, 1.0 / sqrt(pr[j]));
                break;
            }
            case QCD_OP_KRAUS2Q: {
                const int cnt = (int)p[0];
                double pr[64];
                for (int j = 0; j < cnt; j++) { mat_from(p + 1 + 32 * j, 16, m); pr[j] = NORM2(psi, N, a, b, m); }
                const int j = pick(pr, cnt, u53(seed, circuit, shot, stream++));
                mat_from(p + 1 + 32 * j, 16, m);
                APPLY2(psi, N, a, b, m);
                SCALE(psi, N, 1.0 / sqrt(pr[j]));
                break;
            }
            case QCD_OP_PAULI1Q: {
                const int j = pick(p, 4, u53(seed, circuit, shot, stream++));
                if (j) APPLY1(psi, N, a, PAULI[j]);
                break;
            }
            case QCD_OP_PAULI2Q: {
                const int j = pick(p, 16, u53(seed, circuit, shot, stream++));
                if (j & 3) APPLY1(psi, N, a, PAULI[j & 3]);
                if (j >> 2) APPLY1(psi, N, b, PAULI[j >> 2]);
                break;
            }
            default: return ~0ull;
        }
    }
    double tot = 0;
    for (uint64_t x = 0; x < N; x++) tot += creal(psi[x]) * creal(psi[x]) + cimag(psi[x]) * cimag(psi[x]);
    const double target = u53(seed, circuit, shot, QCD_STREAM_MEASURE) * tot;
    double cum = 0;
    for (uint64_t x = 0; x < N; x++) {
        cum += creal(psi[x]) * creal(psi[x]) + cimag(psi[x]) * cimag(psi[x]);
        if (cum > target) return x;
    }
    return N - 1;
}

/* shots [shot_begin, shot_end) of one circuit -> outcomes[shot - shot_begin].  Returns 0, or -1 on a bad opcode /
 * allocation failure.  n_threads <= 0: all OpenMP threads. */
int cpu_run_trajectories(const qcd_op* ops, uint32_t n_ops, const double* params, int n_qubits, uint64_t shot_begin,
                         uint64_t shot_end, uint64_t seed, uint32_t circuit, uint64_t* outcomes, int n_threads) {
    int bad = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
    {
        cd* psi = (cd*)malloc(sizeof(cd) << n_qubits);
        if (!psi) {
#pragma omp atomic write
            bad = 1;
        } else {
#pragma omp for schedule(dynamic, 1)
            for (long long s = (long long)shot_begin; s < (long long)shot_end; s++) {
                uint64_t x = run_one(psi, ops, n_ops, params, n_qubits, seed, circuit, (uint64_t)s);
                if (x == ~0ull) {
#pragma omp atomic write
                    bad = 1;
                }
                outcomes[s - (long long)shot_begin] = x;
            }
            free(psi);
        }
    }
    return bad ? -1 : 0;
}

/* Same results, shots one after the other with the amplitude loops spread over the threads (n >= ~22, few shots). */
int cpu_run_trajectories_wide(const qcd_op* ops, uint32_t n_ops, const double* params, int n_qubits, uint64_t shot_begin,
                              uint64_t shot_end, uint64_t seed, uint32_t circuit, uint64_t* outcomes, int n_threads) {
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    cd* psi = (cd*)malloc(sizeof(cd) << n_qubits);
    if (!psi) return -1;
    int bad = 0;
    g_wide = 1;
    for (uint64_t s = shot_begin; s < shot_end; s++) {
        uint64_t x = run_one(psi, ops, n_ops, params, n_qubits, seed, circuit, s);
        if (x == ~0ull) bad = 1;
        outcomes[s - shot_begin] = x;
    }
    g_wide = 0;
    free(psi);
    return bad ? -1 : 0;
}

int cpu_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
