/* qcdsim.h -- C ABI of the B200-native noisy-circuit sampling engine (libqcdsim.so).
 *
 * The reference (nlaanait/qcdenoise) has NO plugin / FFI interface on this path: its only seam is the Python
 * call into qiskit-aer,
 *     qk.execute(circuit, backend=AerSimulator, noise_model=..., shots=n_shots).result().get_counts()
 *                                                                   qcdenoise/samplers.py:308-326
 * followed by pure-Python reductions (samplers.py:195-202, witnesses.py:40-67).  Each entry point below cites the
 * reference interface it replaces; INTEGRATION.md shows the ctypes stub a qcdenoise maintainer would add.
 *
 * Conventions
 *   - plain C: pointers + sizes only, no C++/torch types; every function returns 0 on success, a QCD_E_* code
 *     otherwise; qcd_last_error() gives the message of the last failure on that context.  Nothing throws.
 *   - pointers marked [dev] are CUDA device pointers owned by the caller (e.g. torch tensors' data_ptr());
 *     pointers marked [host] are ordinary host memory.  Work is enqueued on `stream` (a cudaStream_t passed as
 *     void*; NULL = legacy default stream).  Functions that must size device work from device results
 *     synchronise the stream internally (documented per function); outputs are complete on return only when
 *     stated, otherwise after the caller synchronises `stream`.
 *   - one qcd_ctx per (process, device); not thread-safe (one host thread per context, one process per GPU).
 *   - bit order: state index bit q <-> qubit q; histogram bin x <-> qiskit Counts key format(x, '0{n}b').
 *   - RNG: Philox4x32-10, key = (seed lo, seed hi), counter = (shot lo, shot hi, stream, circuit); `shot` is the
 *     shot index inside its circuit, `stream` = ordinal of the stochastic op in the circuit's op list,
 *     QCD_STREAM_MEASURE for the measurement draw, QCD_STREAM_READOUT + q/4 (32-bit word q%4) for the readout
 *     flip of qubit q.  Results are therefore a pure function of (seed, circuit, shot): identical for any
 *     partition of the shot range over GPUs and with trajectory de-duplication on or off.
 */
#ifndef QCDSIM_H
#define QCDSIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QCD_ABI_VERSION 4

#if defined(__GNUC__)
#define QCD_API __attribute__((visibility("default")))
#else
#define QCD_API
#endif

/* error codes */
#define QCD_OK 0
#define QCD_E_INVALID 1     /* bad argument / malformed program */
#define QCD_E_CUDA 2        /* CUDA runtime error (message has the cudaError string) */
#define QCD_E_NOMEM 3       /* device memory budget too small for the request */
#define QCD_E_UNSUPPORTED 4 /* valid request the engine does not implement (message says what) */
#define QCD_E_STATE 5       /* call order (e.g. run before upload) */

/* Gate-list IR: one qcd_op per instruction of the circuit, in program order.  This is the content of the qiskit
 * QuantumCircuit the reference builds in graph_circuit.py:81-99,114-130,151-173 and stabilizers.py:136-156, plus the
 * quantum errors its NoiseModel attaches (samplers.py:404-412, :588-608), made explicit.
 * param_off indexes the shared `params` array (doubles).  Matrices are row-major, complex entries as (re, im)
 * pairs, little-endian basis: index s = bit(q0) + 2*bit(q1). */
typedef struct qcd_op {
    uint16_t opcode;    /* QCD_OP_* */
    uint8_t q0;         /* first qubit (control for CX) */
    uint8_t q1;         /* second qubit or 0xFF */
    uint32_t param_off; /* offset of this op's parameters in params[] (ignored when it has none) */
} qcd_op;

enum {
    QCD_OP_ID = 0, QCD_OP_H, QCD_OP_X, QCD_OP_Y, QCD_OP_Z, QCD_OP_S, QCD_OP_SDG, QCD_OP_T, QCD_OP_TDG,
    QCD_OP_RX, QCD_OP_RY, QCD_OP_RZ, /* params: theta */
    QCD_OP_U1Q,                      /* params: 8 doubles (2x2 complex) */
    QCD_OP_CX, QCD_OP_CZ, QCD_OP_SWAP,
    QCD_OP_U2Q,                      /* params: 32 doubles (4x4 complex) */
    /* stochastic ops ("sites"); each consumes one Philox stream, numbered in program order */
    QCD_OP_KRAUS1Q,  /* params: m, then m * 8 doubles  (m Kraus matrices 2x2) */
    QCD_OP_KRAUS2Q,  /* params: m, then m * 32 doubles (m Kraus matrices 4x4) -- density-matrix mode only */
    QCD_OP_PAULI1Q,  /* params: 4 probabilities  (I, X, Y, Z) */
    QCD_OP_PAULI2Q,  /* params: 16 probabilities, index = P(q0) + 4 P(q1) */
    QCD_OP_RESET,    /* reset q0 to |0> */
    QCD_OP_COUNT
};

#define QCD_STREAM_MEASURE 0xFFFFFFFFu
#define QCD_STREAM_READOUT 0xFFFF0000u

/* run flags */
#define QCD_RUN_NO_DEDUP 1u /* one state per trajectory (literal Aer behaviour); default shares states between
                               trajectories with identical branch histories -- same results, fewer sweeps */

typedef struct qcd_ctx qcd_ctx;

/* run statistics, filled by the last qcd_run_* call */
typedef struct qcd_stats {
    uint64_t sweep_launches;   /* state-sweep kernel launches */
    uint64_t aux_launches;     /* all other kernels of ours (decide, scan, sample, parity, ...) */
    uint64_t node_sweeps;      /* (state, sweep) pairs executed = full read+write passes over one state */
    uint64_t sweep_bytes;      /* algorithmic bytes moved by sweeps: reads + writes of amplitudes */
    uint64_t leaves;           /* distinct final states sampled */
    uint64_t trajectories;     /* trajectories (shots) simulated */
    uint64_t max_live_states;  /* high-water mark of resident states */
    double sweep_ms;           /* CUDA-event time spent inside sweep kernels (only when timing is enabled) */
    /* the same three figures per KIND of sweep launch: [0] prep (write-only: the first sweep of a circuit / the first
     * density-matrix pass), [1] inner (every node of the launch has children), [2] leaf (every node is a final state:
     * probability chunk sums are reduced, small leaves are not written), [3] mixed */
    uint64_t kind_launches[4];
    uint64_t kind_bytes[4];
    double kind_ms[4];
    /* sibling groups (option "leaf_cross"): shared passes launched, and leaves whose chunk sums came from them instead of
     * a sweep of their own (both are included in node_sweeps / kind_*[2] as the passes they cost) */
    uint64_t cross_passes;
    uint64_t cross_leaves;
} qcd_stats;

QCD_API int qcd_abi_version(void);

/* Context.  mem_budget_bytes = device memory the engine may allocate for states (0 = 80% of free memory). */
QCD_API int qcd_create(int cuda_device, size_t mem_budget_bytes, qcd_ctx** out);
QCD_API void qcd_destroy(qcd_ctx* ctx);
QCD_API const char* qcd_last_error(qcd_ctx* ctx); /* ctx may be NULL: message of the last failed qcd_create */
/* Options (all optional; defaults in parentheses):
 *   "time_sweeps" 0/1 (0)   record CUDA events around every sweep launch -> qcd_stats.sweep_ms
 *   "tile_bits" 0|5..13 (0 = 13 complex64 / 12 complex128)   index bits of a shared-memory tile
 *   "max_chunk" (2^25)      trajectories per tree; whole circuits are kept together when they fit
 *   "max_slots" (0 = memory budget)   cap on resident states (exercises the depth-first schedule)
 *   "table_cap_log2" (22)   size of the trajectory-grouping table
 *   "use_direct" 0/1 (1)    register-only kernel for single-pass sweeps / density-matrix passes
 *   "direct_min_bits" (10)  smallest state (index bits) the register-only kernel takes; smaller states and sweeps of more
 *                           than 3 passes run on the shared-memory tile kernel
 *   "nostore_max" (1024)      leaves sampled by at most this many trajectories are not written to memory (0 = off)
 *   "leaf_cross" 0/1 (1)    unstored leaves of one parent and branch word whose circuits differ only in a trailing real 2x2
 *                           on one qubit k (a graph circuit and its Toth settings, stabilizers.py:136-156) share passes over
 *                           the parent that reduce |amp|^2 and Re(conj(amp[x]) amp[x ^ 2^k]) per 32-amplitude chunk; every
 *                           member's chunk sums follow from those (exact in exact arithmetic; rounding differs from a sweep
 *                           per leaf in the last bits)
 *   "no_slots" 0/1 (0)      disable the slotted op path of the register-only kernel (A/B measurements)
 *   "share_prefix" 0/1 (1)  consecutive circuits whose sweeps are identical up to the final one (a graph circuit and
 *                           its stabilizer settings, stabilizers.py:194-199) share one trajectory tree down to the
 *                           parents of the leaves; every draw stays keyed by the trajectory's own (circuit, shot), so
 *                           results are bit-identical with the option off
 *   "release_memory" (any value)   free the state arena and per-level scratch now (a context keeps the arena it sized from
 *                           the device's free memory at its first large run); the next run allocates what it needs
 *   "dm_on_chip" 0/1/2 (1)  density-matrix mode: the gate list is interpreted with the whole density matrix in the shared memory
 *                           of a thread-block cluster (no host lowering).  1 = when the matrix fits a cluster of at most 4
 *                           CTAs (n <= 8 qubits in complex64, n <= 7 in complex128: where it measured faster end to end),
 *                           2 = whenever it fits on chip (n <= 9 / n <= 8: clusters of up to 16 CTAs), 0 = always lower to
 *                           state sweeps.  If the device refuses the cluster launch (16 CTAs per cluster is a non-portable
 *                           size) the context falls back to lowered sweeps for good
 *   "dm_test_refuse" 0/1 (0)  test hook: 1 = treat the next cluster launch as refused by the device (exercises the fall
 *                           back to lowered sweeps), 0 = forget an earlier refusal
 *   "dm_max_clusters" (0)   cap on concurrently resident clusters of that kernel (0 = what the device holds)
 *   "hist_row0" (0)         qcd_run_sv_trajectories: the uploaded circuit that row 0 of `hist` stands for.  A rank that
 *                           holds the whole batch's programs and runs one slice of circuits at a time (self-scheduled
 *                           partition, qcdenoise_b200/distributed.py WorkQueue) passes a buffer with rows for that slice
 *                           only; must not exceed the first circuit of the shot range */
QCD_API int qcd_set_option(qcd_ctx* ctx, const char* key, double value);
QCD_API int qcd_get_stats(qcd_ctx* ctx, qcd_stats* out);

/* Replaces: building `circuit` + `noise_model` and handing them to qk.execute (samplers.py:322-324).
 * Uploads B circuits of n_qubits each: circuit c owns ops[op_offsets[c] .. op_offsets[c+1]).  [host] pointers;
 * data is copied.  Lowering (gate fusion, CX -> H.CZ.H, sweep scheduling) happens here, on the host. */
QCD_API int qcd_upload_programs(qcd_ctx* ctx, const qcd_op* ops, const uint32_t* op_offsets, const double* params,
                        size_t n_params, int n_circuits, int n_qubits);

/* Replaces: NoiseModel.from_backend(backend) (samplers.py:588, :608 -- qiskit-aer's basic_device_gate_errors) and the
 * attachment of that noise model to the circuit's gates at qk.execute (samplers.py:322-324), for a whole batch and on
 * native host threads.  ops / op_offsets / params describe NOISE-FREE gate lists as in qcd_upload_programs;
 *   qubit_t1_t2   [host] [n_circuits][n_qubits][2] doubles: T1, T2 in microseconds (inf = no relaxation);
 *   op_gate_props [host] [op_offsets[n_circuits]][2] doubles: gate_error, gate_length (ns) of the calibration entry
 *                 (gate, qubits) the op belongs to; (NaN, NaN) = no entry, the op stays noise-free.
 * After every op with an entry the library writes PAULI1Q/2Q (depolarizing strength such that depolarizing o
 * relaxation has average gate infidelity gate_error) and one KRAUS1Q per gate qubit (thermal relaxation over
 * gate_length; T2 clipped to 2 T1), exactly the program qcdenoise_b200.samplers.DeviceNoiseSampler builds op by op
 * in Python -- same operator order, hence the same shots.  The gate lists and the numbers are validated and copied
 * here; the channels are written circuit by circuit inside the lowering threads at the first run (no expanded copy
 * of the batch is ever held), so out-of-range T1/T2 surface as an error of that run. */
QCD_API int qcd_upload_device_noise_programs(qcd_ctx* ctx, const qcd_op* ops, const uint32_t* op_offsets,
                                             const double* params, size_t n_params, int n_circuits, int n_qubits,
                                             const double* qubit_t1_t2, const double* op_gate_props);

/* Host-only (no GPU needed): the expansion qcd_upload_device_noise_programs performs, written to caller buffers
 * (offsets_out holds n_circuits + 1 entries).  Returns QCD_E_NOMEM when a capacity is too small; the required sizes
 * are always stored in *ops_needed / *params_needed.  Used by the CPU tests to check the native channel
 * construction against the Python one and the oracle. */
QCD_API int qcd_expand_device_noise(const qcd_op* ops, const uint32_t* op_offsets, const double* params, size_t n_params,
                                    int n_circuits, int n_qubits, const double* qubit_t1_t2, const double* op_gate_props,
                                    qcd_op* ops_out, size_t ops_cap, uint32_t* offsets_out, double* params_out,
                                    size_t params_cap, size_t* ops_needed, size_t* params_needed);

/* Replaces: AerSimulator `statevector` method under a noise model, i.e. execute(...).result().get_counts() for
 * shots <= 2^n or n > ~14 (samplers.py:308-326; stabilizers.py:214).  Global shot index g in
 * [shot_begin, shot_end) of the flattened [n_circuits][shots_per_circuit] space: circuit = g / shots_per_circuit,
 * shot = g % shots_per_circuit  (the multi-GPU partition: rank r takes a contiguous slice).
 *   first_circuit: index of uploaded circuit 0 in the caller's whole batch; the Philox counters use
 *                  first_circuit + circuit, so a rank that uploaded only ITS slice of the circuits reproduces the draws
 *                  of a single-GPU run of the whole batch (mirrors qcd_sample_from_probs).  0 for a whole batch.
 *   precision: 32 (complex64) or 64 (complex128).
 *   readout [host, may be NULL]: [n_circuits][n_qubits][2][2] doubles, row = true bit, col = measured bit.
 *   hist [dev, may be NULL]: [n_circuits][2^n] uint32, ADDED to (caller zeroes).
 *   outcomes [dev, may be NULL]: [shot_end - shot_begin] uint32 basis indices (requires n_qubits <= 32).
 * Synchronises `stream` internally (tree scheduling needs device counts); results are complete on return. */
QCD_API int qcd_run_sv_trajectories(qcd_ctx* ctx, int precision, uint64_t shots_per_circuit, uint64_t seed,
                            uint64_t shot_begin, uint64_t shot_end, uint32_t first_circuit, const double* readout,
                            uint32_t* hist, uint32_t* outcomes, uint32_t flags, void* stream);

/* Replaces: AerSimulator `density_matrix` method (what Aer selects for noisy circuits with shots > 2^n), up to
 * the exact outcome distribution.  probs_out [dev]: [n_circuits][2^n] float (precision 32) or double (64): the
 * diagonal of rho after the last op (before classical readout error).  circuit_begin/end select a slice.
 * Asynchronous on `stream`. */
QCD_API int qcd_run_density_matrix(qcd_ctx* ctx, int precision, int circuit_begin, int circuit_end, void* probs_out,
                           void* stream);

/* Replaces: StabilizerSampler.sample re-simulating the whole graph circuit once per measurement setting
 * (stabilizers.py:183-215: `graph_circuit + stab_gate + measure_all` for each of the n + 1 Toth settings).  In
 * density-matrix mode a setting that differs from the uploaded circuit by ONE final single-qubit step (the basis
 * change of the generator's vertex, X -> H, Y -> H.Sdg, followed by that gate's own noise channel if any) is read
 * off the circuit's density matrix: p_s[x] = Re sum_{a,b} S_s[(x_q, x_q), (a, b)] rho[x|q=a, x|q=b].
 *   setting_qubits   [host] [count][n_settings] : qubit of the final step, 0xFF = measure as is;
 *   setting_superops [host] [count][n_settings][16] complex (re, im): 4x4 superoperator of the step on (ket, bra) of
 *                    that qubit, row-major, index = ket + 2 * bra  (S = sum_j conj(K_j) (x) K_j; ignored for 0xFF);
 *   probs_out [dev]  [count][n_settings][2^n] float / double, count = circuit_end - circuit_begin.
 * One density-matrix simulation per circuit instead of n_settings.  Complete on return (synchronises `stream`). */
QCD_API int qcd_run_density_matrix_settings(qcd_ctx* ctx, int precision, int circuit_begin, int circuit_end,
                                            int n_settings, const uint8_t* setting_qubits,
                                            const double* setting_superops, void* probs_out, void* stream);

/* Replaces: one simulation + one sampled measurement setting PER ELEMENT of the full stabilizer group of a graph state
 * (the "Jung" stabilizers of circuit_constructors.py:554-590; JungStabilizer, stabilizers.py:171-172, fed to
 * StabilizerSampler.sample, stabilizers.py:183-215).  In density-matrix mode the exact <S> = Tr(rho S) of ALL 2^n
 * elements is read off the circuit's density matrix in one pass: S = (+-) X^sx Z^sz has one non-zero per column, so
 * every entry of rho belongs to exactly one element (sx = row ^ column).
 *   adjacency   [host] [n_qubits] uint32: bit j of row k set <=> edge (k, j); generators g_k = X_k prod_j Z_j;
 *   expvals_out [dev]  [count][2^n] double, index = the element's x mask sx (element = product of g_k over the bits of
 *               sx), signed: the ideal graph state gives +1 everywhere; index 0 is the identity (= Tr rho).
 * Graph-state fidelity <G|rho|G> = mean of a row.  Complete on return (synchronises `stream`). */
QCD_API int qcd_run_density_matrix_group(qcd_ctx* ctx, int precision, int circuit_begin, int circuit_end,
                                         const uint32_t* adjacency, double* expvals_out, void* stream);

/* Replaces: the enumeration of the full signed stabilizer group through sympy Pauli algebra (sigma_prod,
 * stabilizers.py:39-79, driven by circuit_constructors.py:554-590) by bit arithmetic on the device: element `sel` in
 * [begin, end) is the product of the generators j whose bit (n_gens - 1 - j) of sel is set (the reference's
 * np.binary_repr order), generators and elements as (x, z) masks (bit q: X on qubit q / Z on qubit q / both: Y).
 *   gen_x, gen_z [host] [n_gens] uint32;  x_out, z_out [dev] [end - begin] uint32;
 *   sign_out [dev] [end - begin] uint8: 0 = '+', 1 = '-' (sign of the LABEL, Y = i X Z), 2 = the product is not
 *   Hermitian (the generators do not commute).  Complete on return (synchronises `stream`). */
QCD_API int qcd_stabilizer_group(qcd_ctx* ctx, const uint32_t* gen_x, const uint32_t* gen_z, int n_gens, uint64_t begin,
                                 uint64_t end, uint32_t* x_out, uint32_t* z_out, uint8_t* sign_out, void* stream);

/* Classical readout error applied to probability vectors, in place: probs [dev] [n_vectors][2^n] (float/double by
 * precision), readout [host] [n_vectors][n][2][2].  Replaces Aer's ReadoutError on `measure` in distribution. */
QCD_API int qcd_apply_readout_probs(qcd_ctx* ctx, int precision, void* probs, int n_vectors, int n_qubits,
                            const double* readout, void* stream);

/* Replaces: Aer `sample_measure` on an exact distribution.  Draws shots_per_vector outcomes from each of the
 * n_vectors distributions probs [dev] by inverse CDF with the QCD_STREAM_MEASURE draw of (seed, first_circuit + v,
 * shot) for shot in [shot_begin, shot_end); optional readout flips as above ([host], may be NULL).
 * hist [dev] [n_vectors][2^n] uint32 is ADDED to.  Asynchronous on `stream`. */
QCD_API int qcd_sample_from_probs(qcd_ctx* ctx, int precision, const void* probs, int n_vectors, int n_qubits,
                          uint64_t shot_begin, uint64_t shot_end, uint64_t seed, uint32_t first_circuit,
                          const double* readout, uint32_t* hist, void* stream);

/* Replaces: Witness._construct_diagonals + igmit.expectation_value (witnesses.py:40-67) without ever building the
 * 2^n diagonals:  D[x] = (-1)^popcount(x & mask).
 *   hist [dev] [n_vectors][2^n] uint32 counts ->
 *   signed_sums [dev] [n_vectors][n_masks] int64 = sum_x hist[x] * D_mask[x]     (bit-exact integer arithmetic)
 *   totals      [dev] [n_vectors] int64          = sum_x hist[x]
 * masks [host] [n_vectors][n_masks] uint32 when per_vector_masks != 0, else [n_masks] shared.
 * mean = signed/total, std = sqrt((1 - mean^2)/total) are formed by the caller in double.  Asynchronous. */
QCD_API int qcd_parity_counts(qcd_ctx* ctx, const uint32_t* hist, int n_vectors, int n_qubits, const uint32_t* masks,
                      int n_masks, int per_vector_masks, int64_t* signed_sums, int64_t* totals, void* stream);

/* Same reduction on per-shot outcome lists (the sparse form used when 2^n histograms are too large, n > 26):
 *   outcomes [dev] [n_vectors][shots_per_vector] uint32 basis indices (as written by qcd_run_sv_trajectories) ->
 *   signed_sums [dev] [n_vectors][n_masks] int64 = sum_shots (-1)^popcount(x & mask);  mean = signed / shots.
 * Replaces witnesses.py:40-67 where the reference would need 2^n-entry diagonals.  Asynchronous. */
QCD_API int qcd_parity_outcomes(qcd_ctx* ctx, const uint32_t* outcomes, int n_vectors, uint64_t shots_per_vector,
                                const uint32_t* masks, int n_masks, int per_vector_masks, int64_t* signed_sums,
                                void* stream);

/* Same reduction on exact probabilities (float/double by precision): expvals [dev] [n_vectors][n_masks] double =
 * sum_x p[x] D_mask[x], accumulated in double with a fixed reduction order (deterministic). Asynchronous. */
QCD_API int qcd_parity_probs(qcd_ctx* ctx, int precision, const void* probs, int n_vectors, int n_qubits,
                     const uint32_t* masks, int n_masks, int per_vector_masks, double* expvals, void* stream);

/* Host-only: the 4x4 superoperators (index = ket + 2 * bra, row-major complex as (re, im)) of `count` noisy
 * single-qubit gates: the gate `unitary` (2x2 complex, shared) followed by the calibration-derived error of entry i
 * (depolarizing o thermal relaxation as in qcd_upload_device_noise_programs; (NaN, NaN) = no entry).  Feeds
 * qcd_run_density_matrix_settings with the noisy final basis change of every measurement setting of a batch. */
QCD_API int qcd_device_gate_superops_1q(int count, const double* unitary, const double* gate_error,
                                        const double* gate_length_ns, const double* t1_us, const double* t2_us,
                                        double* superops_out);

/* Replaces: GraphState.sample (graph_state.py:74-122) with pick_subgraphs (:200-215), combine_subgraphs (:176-198)
 * and the vertex-cover screen (networkx min_weighted_vertex_cover, :101-106), for n_graphs graphs at once on native
 * host threads (host-only, no GPU needed).  Undirected graphs.
 *   library: n_lib connected graphs; graph g has lib_orders[g] vertices numbered in node-iteration order and the
 *            edges lib_edges[2e], lib_edges[2e+1] (u < v, sorted by u = networkx edges() order),
 *            e in [lib_edge_offsets[g], lib_edge_offsets[g+1]);
 *   order_graphs[order_offsets[k] .. order_offsets[k+1]) = library graphs with k vertices, k in [0, max_order];
 *   combinations: comb_sizes[comb_offsets[c] .. comb_offsets[c+1]) = sub-graph orders adding up to n_qubits.
 * Graph g (global index first_graph + g) draws from Philox4x32-10 with key = seed, counter = (index lo, index hi,
 * attempt, block); draw j is word j of that stream mapped to [0, k) by (word * k) >> 32, in the order: combination,
 * one library graph per block, then per later block one earlier vertex and order(block) connector targets among
 * the block's first order(block) - 1 vertices.  Up to max(1, max_itr) attempts until cover_size / 2 >=
 * min_vertex_cover.  edges_out [n_graphs][max_edges][2] (max_edges >= n (n - 1) / 2) in networkx edges() order of
 * the joined graph, n_edges_out [n_graphs]. */
QCD_API int qcd_sample_graph_states(const uint8_t* lib_edges, const uint32_t* lib_edge_offsets, const uint8_t* lib_orders,
                                    int n_lib, const uint32_t* order_offsets, const uint32_t* order_graphs, int max_order,
                                    const uint8_t* comb_sizes, const uint32_t* comb_offsets, int n_combs, int n_qubits,
                                    int n_graphs, int min_vertex_cover, int max_itr, uint64_t seed, uint64_t first_graph,
                                    uint8_t* edges_out, uint32_t max_edges, uint32_t* n_edges_out);

/* Replaces: generate_adjacency_tensor (samplers.py:119-192) for a batch of gate lists (host-only).  Walks each
 * circuit's gates in order: a 1-qubit gate on q writes codes[opcode] at [p][q][q], a 2-qubit gate on (a, b) at
 * [p][a][b] (and [p][b][a] unless directed), p = first plane whose [a][b] slot is still 0.  Noise ops are skipped.
 * codes [QCD_OP_COUNT]; out [host] [n_circuits][planes][rows][cols] doubles, overwritten. */
QCD_API int qcd_adjacency_tensors(const qcd_op* ops, const uint32_t* op_offsets, int n_circuits, const int32_t* codes,
                                  int planes, int rows, int cols, int directed, double* out);

/* Host-only: the distinct (gate, qubits) calibration entries of every circuit of a batch -- what
 * DeviceNoiseSampler.build_noise_model collects per circuit before drawing one (gate_error, gate_length) per entry
 * (samplers.py:613-727).  Entries are sorted by (circuit, opcode << 16 | q0 << 8 | q1), i.e. the order of
 * numpy.unique over those keys; op_entry_out[i] = entry of op i.
 *   entry_key_out, entry_circuit_out [host] capacity op_offsets[n_circuits]; *n_entries = entries written. */
QCD_API int qcd_gate_entries(const qcd_op* ops, const uint32_t* op_offsets, int n_circuits, uint32_t* entry_key_out,
                             uint32_t* entry_circuit_out, uint32_t* op_entry_out, size_t* n_entries);

/* Host-only introspection (no GPU needed): lower + schedule ONE circuit and write a JSON description of the sweep
 * program (tile bits, passes, fused matrices, sign terms, site groups) into buf.  mode: 0 = statevector
 * trajectories, 1 = density matrix.  Returns QCD_E_NOMEM if cap is too small (required size in *needed).
 * Used by the CPU tests to check the host logic against the oracle. */
QCD_API int qcd_debug_schedule_json(const qcd_op* ops, uint32_t n_ops, const double* params, size_t n_params, int n_qubits,
                            int mode, int precision, int tile_bits, char* buf, size_t cap, size_t* needed);

/* Host-only introspection (no GPU needed): the sibling runs the engine would form for this batch in statevector mode
 * (option "leaf_cross"): consecutive circuits that share everything but a trailing real 2x2 on one qubit.  JSON:
 * {"runs": [{"lead", "members", "n_high", "reg_bits": [4], "nops", "kbit": [per member: qubit of the trailing op (>= 5),
 * -1 = chunk sums are the shared C0, -2 = keeps its own sweep], "u": [4 per member, row-major]}]}. */
QCD_API int qcd_debug_sibling_runs_json(const qcd_op* ops, const uint32_t* op_offsets, const double* params, size_t n_params,
                                int n_circuits, int n_qubits, int precision, char* buf, size_t cap, size_t* needed);

/* Host-only introspection (no GPU needed): the pass schedule of ONE circuit for the on-chip density-matrix interpreter
 * (option "dm_on_chip").  Pass p acts inside the qubit pair pass_qubits[2 p], pass_qubits[2 p + 1] (0xFF = one qubit
 * only) and executes the next pass_n_ops[p] entries of `order` (indices into ops[], identity gates left out).  Returns
 * QCD_E_NOMEM when a capacity is too small; the sizes needed are always stored in *n_passes / *n_order. */
QCD_API int qcd_debug_dm_schedule(const qcd_op* ops, uint32_t n_ops, const double* params, size_t n_params, int n_qubits,
                                  uint8_t* pass_qubits, uint32_t* pass_n_ops, size_t pass_cap, uint32_t* order,
                                  size_t order_cap, size_t* n_passes, size_t* n_order);

#ifdef __cplusplus
}
#endif
#endif /* QCDSIM_H */
