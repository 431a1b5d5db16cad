// program.h -- the scheduled "sweep program" shared by the host scheduler (lower.cpp) and the CUDA kernels.
// All structs are PODs copied verbatim to the device.
#pragma once
#include <stdint.h>

namespace qcd {

constexpr int REG_BITS = 4;   // a pass keeps 2^4 amplitudes per thread in registers
constexpr int MAX_HIGH = 10;  // scattered (non-contiguous) tile bits
constexpr int MIN_LOW = 5;    // a tile always contains the 2^5 lowest index bits (256 B runs in complex64)
constexpr int MAX_BRANCH_BITS = 16;

enum KOpKind : uint8_t {
    K_M1 = 0,   // 2x2 complex on register bit t0
    K_M2 = 1,   // 4x4 complex on register bits t0 < t1 (index = b(t0) + 2 b(t1))
    K_M4 = 2,   // 16x16 complex on the pass's 4 register bits (index = register index)
    K_SIGNS = 3,// amplitude *= -1 for every sign term with (gidx & mask) == mask   (CZ / Z / CCZ... layers)
    K_D1 = 4    // amplitude *= d[bit(gidx, gbit)]  (2 complex), any global bit, no register residency needed
};

struct KOp {
    uint8_t kind;
    uint8_t t0, t1;     // register-bit indices inside the pass
    uint8_t gbit;       // K_D1: global bit
    uint8_t sel_shift;  // variant = (node.branch >> sel_shift) & sel_mask
    uint8_t sel_mask;
    uint16_t count;     // K_SIGNS: number of masks
    uint32_t off;       // matrices: offset into the complex pool (elements); K_SIGNS: offset into masks[]
    uint32_t stride;    // variant stride in complex elements
};

struct Pass {
    uint8_t nreg;               // number of register bits actually used (<= REG_BITS; always REG_BITS unless k < 4)
    uint8_t rbits[REG_BITS];    // LOCAL bit positions, ascending
    uint8_t pad[3];
    uint32_t op_begin, op_end;  // range in kops[]
};

enum SweepFlags : uint8_t {
    SW_PREP = 1,   // source is the product state prep[] (no read)
    SW_FINAL = 2,  // last sweep of the circuit: also emit 32-amplitude chunk sums of |amp|^2
    SW_CROSS = 4   // DirectSweep only: the ops SHARED by the final sweeps of a run of sibling circuits (everything but a
                   // member's trailing real 2x2 on one qubit k); the launch emits the chunk sums C0 of |amp|^2 and, for the
                   // register bits named in red_q[], the chunk sums X_k of Re(conj(amp[x]) amp[x ^ 2^k]), from which the
                   // chunk sums of every member follow without sweeping the state once per member
};

struct Sweep {
    uint8_t n_bits;             // index bits of the state (n, or 2n in density-matrix mode)
    uint8_t k;                  // tile bits
    uint8_t L;                  // contiguous low bits
    uint8_t n_high;             // k - L
    uint8_t high[MAX_HIGH];     // global bit of local bit L + j, ascending
    uint8_t flags;
    int8_t red_q[2];            // joint populations of these global bits are reduced from the output (-1: none)
    uint8_t pad;
    uint32_t pass_begin, pass_end;
    uint32_t prep_off;          // SW_PREP: pool offset of n_bits (alpha, beta) pairs
    int32_t group;              // site group resolved before this sweep (-1: none)
};

// One stochastic decision inside a group.  data layout (doubles) per branch j: w[4] then T[16] (row-major):
//   p_j = sum_s w[s] P[s],  P' = T P / p_j      with P = joint populations of the group's (q0, q1).
struct SubSite {
    uint8_t nbranch;
    uint8_t shift;     // position of this decision's field in the branch word
    uint8_t nbits;
    uint8_t unit;      // 1: the variant matrices are norm preserving (bare Paulis) -> no 1/sqrt(p_j) rescale
    uint32_t stream;   // Philox stream = ordinal of the stochastic op in the circuit
    uint32_t data_off; // into gdata[]
};

struct Group {
    int8_t q[2];       // global bits whose populations the decisions depend on (-1: unused)
    uint8_t total_bits;
    uint8_t pad;
    uint32_t sub_begin, sub_end;
};

// Fully resolved description of a sweep that consists of ONE pass over <= 4 register bits: executed by the direct
// (register-only, no shared-memory tile) kernel.  ok == 0: the sweep needs the tiled kernel.
constexpr int MAX_DIRECT_OPS = 12;
struct DirectSweep {
    uint8_t ok;
    uint8_t n_bits;
    uint8_t nreg;
    uint8_t nops;
    uint8_t rbit[REG_BITS];     // GLOBAL bit of register bit i, ascending
    int8_t red_q[2];            // SW_CROSS: REGISTER indices of the (up to two) bits k whose cross sums are emitted
    uint8_t flags;              // SW_PREP: the source is the product state at prep_off (no read)
    uint8_t pad;
    uint32_t prep_off;
    KOp ops[MAX_DIRECT_OPS];
};

struct CircuitInfo {
    uint32_t sweep_begin, sweep_end;   // range in sweeps[]
    uint32_t n_sites;
    uint32_t lead;                     // first circuit of the run of consecutive circuits that share this circuit's
                                       // sweeps up to (not including) the final one -- e.g. a graph circuit and its
                                       // n + 1 stabilizer settings (stabilizers.py:194-199).  lead == own index: none.
};
constexpr int MAX_MEMBER_BITS = 6;     // a run holds at most 2^6 circuits (longer runs are cut)

}  // namespace qcd
