// host_batch.cpp -- the batched host stages either side of the simulation: random
// graph states for a whole batch and adjacency-tensor encodings of gate lists.  Host-only C++ (threads, no CUDA):
// once the GPU needs ~4 us per 8-qubit circuit, a Python loop over networkx graphs is the bottleneck.
#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/qcdsim.h"

namespace {

// Philox4x32-10, the same generator the kernels use (kernels.cu)
struct Philox {
    uint32_t key[2];
    uint32_t ctr[4];
    uint32_t out[4];
    int have = 0;
    Philox(uint64_t seed, uint64_t graph, uint32_t attempt) {
        key[0] = (uint32_t)seed;
        key[1] = (uint32_t)(seed >> 32);
        ctr[0] = (uint32_t)graph;
        ctr[1] = (uint32_t)(graph >> 32);
        ctr[2] = attempt;
        ctr[3] = 0;
    }
    void block() {
        uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
        for (int r = 0; r < 10; r++) {
            const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
            const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
            const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
        ctr[3]++;
        have = 4;
    }
    // uniform integer in [0, k): word j of the stream -> (word * k) >> 32
    uint32_t below(uint32_t k) {
        if (!have) block();
        const uint32_t w = out[4 - have];
        have--;
        return (uint32_t)(((uint64_t)w * k) >> 32);
    }
};

struct Edge {
    uint8_t u, v;
};

// size of the cover the local-ratio 2-approximation returns on unit weights, visiting the edges in `edges` order
// (networkx min_weighted_vertex_cover, which the reference calls at graph_state.py:101-102)
int local_ratio_cover_size(const std::vector<Edge>& edges, int n) {
    int cost[256];
    bool in[256];
    for (int i = 0; i < n; i++) {
        cost[i] = 1;
        in[i] = false;
    }
    int size = 0;
    for (const Edge& e : edges) {
        if (in[e.u] || in[e.v]) continue;
        if (cost[e.u] <= cost[e.v]) {
            in[e.u] = true;
            cost[e.v] -= cost[e.u];
        } else {
            in[e.v] = true;
            cost[e.u] -= cost[e.v];
        }
        size++;
    }
    return size;
}

}  // namespace

extern "C" {

int qcd_sample_graph_states(const uint8_t* lib_edges, const uint32_t* lib_edge_offsets, const uint8_t* lib_orders,
                            int n_lib, const uint32_t* order_offsets, const uint32_t* order_graphs, int max_order,
                            const uint8_t* comb_sizes, const uint32_t* comb_offsets, int n_combs, int n_qubits,
                            int n_graphs, int min_vertex_cover, int max_itr, uint64_t seed, uint64_t first_graph,
                            uint8_t* edges_out, uint32_t max_edges, uint32_t* n_edges_out) {
    if (!lib_edges || !lib_edge_offsets || !lib_orders || !order_offsets || !order_graphs || !comb_sizes ||
        !comb_offsets || !edges_out || !n_edges_out)
        return QCD_E_INVALID;
    if (n_lib < 1 || n_combs < 1 || n_qubits < 2 || n_qubits > 255 || n_graphs < 0 || max_order < 1) return QCD_E_INVALID;
    if (max_edges < (uint32_t)(n_qubits * (n_qubits - 1) / 2)) return QCD_E_INVALID;
    // every combination must add up to n_qubits out of orders that have library graphs
    for (int c = 0; c < n_combs; c++) {
        int tot = 0;
        if (comb_offsets[c] >= comb_offsets[c + 1]) return QCD_E_INVALID;
        for (uint32_t j = comb_offsets[c]; j < comb_offsets[c + 1]; j++) {
            const int k = comb_sizes[j];
            if (k < 2 || k > max_order || order_offsets[k] == order_offsets[k + 1]) return QCD_E_INVALID;
            tot += k;
        }
        if (tot != n_qubits) return QCD_E_INVALID;
    }
    for (int g = 0; g < n_lib; g++)
        for (uint32_t e = lib_edge_offsets[g]; e < lib_edge_offsets[g + 1]; e++)
            if (lib_edges[2 * e] >= lib_orders[g] || lib_edges[2 * e + 1] >= lib_orders[g] ||
                lib_edges[2 * e] >= lib_edges[2 * e + 1] ||
                (e > lib_edge_offsets[g] && lib_edges[2 * e] < lib_edges[2 * e - 2]))
                return QCD_E_INVALID;   // edges must be (u, v), u < v, sorted by u (node-iteration numbering)
    const int tries = std::max(1, max_itr);
    auto work = [&](int g_lo, int g_hi) {
        std::vector<Edge> edges, merged;
        for (int g = g_lo; g < g_hi; g++) {
            for (int attempt = 0; attempt < tries; attempt++) {
                Philox rng(seed, first_graph + (uint64_t)g, (uint32_t)attempt);
                // 1. a random combination of sub-graph orders, a random library graph per order
                const uint32_t c = rng.below((uint32_t)n_combs);
                const uint32_t nb = comb_offsets[c + 1] - comb_offsets[c];
                uint32_t pick[128];   // blocks have >= 2 vertices and n_qubits <= 255: at most 127 blocks
                for (uint32_t j = 0; j < nb; j++) {
                    const int k = comb_sizes[comb_offsets[c] + j];
                    pick[j] = order_graphs[order_offsets[k] + rng.below(order_offsets[k + 1] - order_offsets[k])];
                }
                // 2. disjoint union, block after block; every later block is tied to ONE random earlier vertex by
                //    order(block) draws from its first order(block) - 1 vertices (duplicates collapse)
                edges.clear();
                int order = 0;
                for (uint32_t j = 0; j < nb; j++) {
                    const uint32_t lg = pick[j];
                    const int k = lib_orders[lg];
                    if (order > 1) {
                        const uint8_t first = (uint8_t)rng.below((uint32_t)order);
                        bool chosen[256] = {false};
                        uint8_t conn[256];
                        int n_conn = 0;
                        for (int d = 0; d < k; d++) {
                            const uint8_t v = (uint8_t)(order + rng.below((uint32_t)(k - 1)));
                            if (!chosen[v]) {
                                chosen[v] = true;
                                conn[n_conn++] = v;
                            }
                        }
                        // edges() lists the edges by their lower endpoint in vertex order: the connectors go behind
                        // the edges `first` already starts
                        merged.clear();
                        size_t i = 0;
                        for (; i < edges.size() && edges[i].u <= first; i++) merged.push_back(edges[i]);
                        for (int d = 0; d < n_conn; d++) merged.push_back(Edge{first, conn[d]});
                        for (; i < edges.size(); i++) merged.push_back(edges[i]);
                        edges.swap(merged);
                    }
                    for (uint32_t e = lib_edge_offsets[lg]; e < lib_edge_offsets[lg + 1]; e++)
                        edges.push_back(Edge{(uint8_t)(lib_edges[2 * e] + order), (uint8_t)(lib_edges[2 * e + 1] + order)});
                    order += k;
                }
                // 3. screen on the size of the approximate minimum vertex cover
                if (local_ratio_cover_size(edges, n_qubits) / 2 >= min_vertex_cover) break;
            }
            const uint32_t ne = (uint32_t)std::min<size_t>(edges.size(), max_edges);
            n_edges_out[g] = ne;
            std::memcpy(edges_out + (size_t)g * max_edges * 2, edges.data(), (size_t)ne * 2);
        }
    };
    int n_thr = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u);
    n_thr = std::max(1, std::min(n_thr, n_graphs / 256));
    if (n_thr == 1) {
        work(0, n_graphs);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < n_thr; t++)
            pool.emplace_back(work, (int)((long long)n_graphs * t / n_thr), (int)((long long)n_graphs * (t + 1) / n_thr));
        for (std::thread& th : pool) th.join();
    }
    return QCD_OK;
}

int qcd_adjacency_tensors(const qcd_op* ops, const uint32_t* op_offsets, int n_circuits, const int32_t* codes,
                          int planes, int rows, int cols, int directed, double* out) {
    if (!op_offsets || !codes || !out || n_circuits < 0 || planes < 1 || rows < 1 || cols < 1) return QCD_E_INVALID;
    if (n_circuits && op_offsets[n_circuits] && !ops) return QCD_E_INVALID;
    const size_t per = (size_t)planes * rows * cols;
    auto work = [&](int c_lo, int c_hi) {
        for (int c = c_lo; c < c_hi; c++) {
            double* adj = out + (size_t)c * per;
            std::memset(adj, 0, per * sizeof(double));
            for (uint32_t i = op_offsets[c]; i < op_offsets[c + 1]; i++) {
                const qcd_op& op = ops[i];
                if (op.opcode >= QCD_OP_KRAUS1Q) continue;   // noise channels are not gates
                const bool two = op.opcode == QCD_OP_CX || op.opcode == QCD_OP_CZ || op.opcode == QCD_OP_SWAP ||
                                 op.opcode == QCD_OP_U2Q;
                const int r = op.q0, cc = two ? op.q1 : op.q0;
                if (r >= rows || cc >= cols || (two && (cc >= rows || r >= cols))) continue;
                for (int p = 0; p < planes; p++) {
                    double& slot = adj[((size_t)p * rows + r) * cols + cc];
                    if (slot == 0) {
                        slot = (double)codes[op.opcode];
                        if (two && !directed) adj[((size_t)p * rows + cc) * cols + r] = (double)codes[op.opcode];
                        break;
                    }
                }
            }
        }
    };
    int n_thr = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u);
    n_thr = std::max(1, std::min(n_thr, n_circuits / 64));
    if (n_thr == 1) {
        work(0, n_circuits);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < n_thr; t++)
            pool.emplace_back(work, (int)((long long)n_circuits * t / n_thr), (int)((long long)n_circuits * (t + 1) / n_thr));
        for (std::thread& th : pool) th.join();
    }
    return QCD_OK;
}

int qcd_gate_entries(const qcd_op* ops, const uint32_t* op_offsets, int n_circuits, uint32_t* entry_key_out,
                     uint32_t* entry_circuit_out, uint32_t* op_entry_out, size_t* n_entries) {
    if (!op_offsets || !entry_key_out || !entry_circuit_out || !op_entry_out || !n_entries || n_circuits < 0) return QCD_E_INVALID;
    const size_t n_ops = op_offsets[n_circuits];
    if (n_ops && !ops) return QCD_E_INVALID;
    // pass 1 (threads): per circuit, the sorted distinct keys go to the circuit's own op range of entry_key_out (a circuit
    // has at most as many entries as ops) and every op learns its entry's rank inside the circuit
    std::vector<uint32_t> count((size_t)n_circuits + 1, 0);
    auto pass1 = [&](int c_lo, int c_hi) {
        std::vector<uint32_t> keys;
        for (int c = c_lo; c < c_hi; c++) {
            const uint32_t b = op_offsets[c], e = op_offsets[c + 1];
            keys.clear();
            for (uint32_t i = b; i < e; i++) keys.push_back(((uint32_t)ops[i].opcode << 16) | ((uint32_t)ops[i].q0 << 8) | ops[i].q1);
            std::sort(keys.begin(), keys.end());
            keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
            for (uint32_t i = b; i < e; i++) {
                const uint32_t k = ((uint32_t)ops[i].opcode << 16) | ((uint32_t)ops[i].q0 << 8) | ops[i].q1;
                op_entry_out[i] = (uint32_t)(std::lower_bound(keys.begin(), keys.end(), k) - keys.begin());
            }
            std::copy(keys.begin(), keys.end(), entry_key_out + b);
            count[c + 1] = (uint32_t)keys.size();
        }
    };
    int n_thr = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u);
    n_thr = std::max(1, std::min(n_thr, n_circuits / 256));
    auto run = [&](auto&& fn) {
        if (n_thr == 1) { fn(0, n_circuits); return; }
        std::vector<std::thread> pool;
        for (int t = 0; t < n_thr; t++)
            pool.emplace_back(fn, (int)((long long)n_circuits * t / n_thr), (int)((long long)n_circuits * (t + 1) / n_thr));
        for (std::thread& th : pool) th.join();
    };
    run(pass1);
    for (int c = 0; c < n_circuits; c++) count[c + 1] += count[c];          // first entry of every circuit
    // pass 2 (in order, in place: the packed position never overtakes the source): compact the keys, globalise the ranks
    for (int c = 0; c < n_circuits; c++) {
        const uint32_t b = op_offsets[c], first = count[c], k = count[c + 1] - count[c];
        for (uint32_t j = 0; j < k; j++) {
            entry_key_out[first + j] = entry_key_out[b + j];
            entry_circuit_out[first + j] = (uint32_t)c;
        }
    }
    auto pass3 = [&](int c_lo, int c_hi) {
        for (int c = c_lo; c < c_hi; c++)
            for (uint32_t i = op_offsets[c]; i < op_offsets[c + 1]; i++) op_entry_out[i] += count[c];
    };
    run(pass3);
    *n_entries = count[n_circuits];
    return QCD_OK;
}

}  // extern "C"
