// MultiRapProver::prove_rap_constraints: LogUp input layer, GKR, the
// zerocheck / LogUp univariate round 0 (constraint evaluation on cosets of the skip domain) and the
// front-loaded batched MLE sumcheck rounds.
//
// Replaces (reference, relative to /root/reference):
//   crates/cuda-backend/src/logup_zerocheck/mod.rs:119-434           prove_zerocheck_and_logup_gpu
//   crates/cuda-backend/src/logup_zerocheck/rules/mod.rs:27-130      DAG -> three-address rules, register allocation
//   crates/cuda-backend/cuda/src/logup_zerocheck/{gkr_input,zerocheck_round0,logup_round0,mle,batch_mle}.cu
//   crates/cuda-backend/cuda/include/{codec.cuh,dag_entry.cuh}       rule encoding / interpreter
// Semantics = crates/stark-backend/src/prover/logup_zerocheck/{mod.rs,cpu.rs,single.rs,evaluator.rs}.
//
// Design.  Every AIR's SymbolicExpressionDag is compiled once on the host into a linear
// three-address program over a small set of value slots (liveness-based reuse).  Constraint and
// interaction roots are folded into three EF accumulators as soon as they are produced
// (sum_k lambda^k C_k, sum eq3b * count, sum eq3b * beta^j * msg_j), so no per-constraint value is
// ever stored.  The SAME program runs in three kernels: over base-field values at the points of
// the cosets g^c D (round 0; a column's value at a point is the 2^l_skip-term Lagrange combination
// of its chunk), over EF values interpolated at X = 1..D between rows 2y, 2y+1 (MLE rounds), and
// at a single row (the tail terms of the front-loaded batching).  The constant zerofier /
// normalisation factors and all polynomial bookkeeping stay on the host with the transcript.
#include <algorithm>
#include <array>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <type_traits>
#include <vector>

#include "ext.cuh"
#include "hostpoly.hpp"
#include "jit.hpp"
#include "kernels.cuh"
#include "pcs.cuh"
#include "transcript.hpp"

// defined in gkr.cu
extern "C" int swirl_gkr_fractional_sumcheck(swirl_ctx* ctx, swirl_transcript* ts, const uint32_t* d_leaves, int log_n,
                                             int assert_zero, uint32_t h_frac_sum[8], uint32_t* h_claims,
                                             uint32_t* h_polys, uint32_t* h_xi);

namespace swirl {

using bb::ext_add;
using bb::ext_mul;
using bb::ext_mul_base;
using bb::ext_sub;

enum : uint32_t { I_VAR = 0, I_CONST, I_ADD, I_SUB, I_MUL, I_NEG, I_PREF, I_MULACC, I_ACC };
struct Instr {
    uint32_t op_dst;  // op | dst << 8
    uint32_t a, b, c;
};
constexpr int BC_BLOCK = 256;
constexpr int BC_MAX_SLOTS = 256;
constexpr size_t BC_MAX_CHUNKS = 32;   // sub-programs per AIR (result scratch: 256 (AIR, chunk) pairs per round)
constexpr size_t BC_PREFETCH_VARS = 8;  // I_PREF distance, in variables
constexpr size_t BC_R0_SPLIT_BELOW = 8192;  // round 0 walks sub-programs only for traces with fewer hypercube points
constexpr size_t BC_CHUNK_ROOTS = 16;  // constraint / interaction roots per sub-program

// ---- host: DAG -> program ---------------------------------------------------------------------------
struct Program {
    std::vector<Instr> code;
    int n_slots = 0;
};
struct AirLayout {   // how DAG variables map to the row parts [sels, view_mats...]
    uint32_t stride, prep_base, main_base, n_parts;
    std::vector<uint32_t> part_col_off;  // global column offset of each row part
    std::vector<uint32_t> part_width;
};

static int air_layout(const swirl_air_ctx& a, AirLayout* L) {
    L->stride = a.need_rot ? 2 : 1;
    L->prep_base = 1;
    L->main_base = 1 + (a.preprocessed ? L->stride : 0);
    L->part_width.clear();
    L->part_width.push_back(3);
    auto push = [&](const swirl_matrix& m) {
        for (uint32_t k = 0; k < L->stride; k++) L->part_width.push_back((uint32_t)m.width);
    };
    if (a.preprocessed) push(*a.preprocessed);
    for (uint64_t i = 0; i < a.n_cached; i++) push(a.cached_mains[i]);
    push(a.common_main);
    L->n_parts = (uint32_t)L->part_width.size();
    L->part_col_off.assign(L->n_parts, 0);
    for (uint32_t p = 1; p < L->n_parts; p++) L->part_col_off[p] = L->part_col_off[p - 1] + L->part_width[p - 1];
    return 0;
}

struct Root {
    uint32_t node, acc, weight;
};

// Compiles the sub-DAG reachable from `roots`; each root value is added into accumulator
// root.acc with weight index root.weight right after it is computed.
static int compile_program(const swirl_air_ctx& a, const AirLayout& L, const std::vector<Root>& roots, Program* out,
                           size_t prefetch_distance = 0) {
    const size_t n = a.n_nodes;
    std::vector<uint8_t> needed(n, 0);
    std::vector<std::vector<uint32_t>> root_of(n);
    for (size_t i = 0; i < roots.size(); i++) {
        SWIRL_REQUIRE(roots[i].node < n, "DAG root index out of range");
        needed[roots[i].node] = 1;
        root_of[roots[i].node].push_back((uint32_t)i);
    }
    for (size_t i = n; i-- > 0;) {
        if (!needed[i]) continue;
        const swirl_dag_node& nd = a.nodes[i];
        if (nd.op == SWIRL_NODE_ADD || nd.op == SWIRL_NODE_SUB || nd.op == SWIRL_NODE_MUL) {
            SWIRL_REQUIRE(nd.a < i && nd.b < i, "DAG is not in topological order");
            needed[nd.a] = needed[nd.b] = 1;
        } else if (nd.op == SWIRL_NODE_NEG) {
            SWIRL_REQUIRE(nd.a < i, "DAG is not in topological order");
            needed[nd.a] = 1;
        }
    }
    std::vector<int64_t> last_use(n, -1);
    for (size_t i = 0; i < n; i++) {
        if (!needed[i]) continue;
        const swirl_dag_node& nd = a.nodes[i];
        if (nd.op == SWIRL_NODE_ADD || nd.op == SWIRL_NODE_SUB || nd.op == SWIRL_NODE_MUL) {
            last_use[nd.a] = (int64_t)i;
            last_use[nd.b] = (int64_t)i;
        } else if (nd.op == SWIRL_NODE_NEG) {
            last_use[nd.a] = (int64_t)i;
        }
    }
    std::vector<int> slot(n, -1), free_slots;
    int n_slots = 0;
    out->code.clear();
    auto release = [&](uint32_t node, size_t at) {
        if (last_use[node] <= (int64_t)at && slot[node] >= 0) {
            free_slots.push_back(slot[node]);
            slot[node] = -1;
        }
    };
    for (size_t i = 0; i < n; i++) {
        if (!needed[i]) continue;
        const swirl_dag_node& nd = a.nodes[i];
        Instr ins{0, 0, 0, 0};
        uint32_t op = 0;
        switch (nd.op) {
            case SWIRL_NODE_VAR_PREP:
                SWIRL_REQUIRE(a.preprocessed && nd.a < a.preprocessed->width, "PreprocessedIndexOutOfBounds");
                SWIRL_REQUIRE(nd.b < L.stride, "row offset needs need_rot");
                op = I_VAR;
                ins.a = L.prep_base + nd.b;
                ins.b = nd.a;
                break;
            case SWIRL_NODE_VAR_MAIN: {
                SWIRL_REQUIRE(nd.c <= a.n_cached, "MainPartitionIndexOutOfBounds");
                const uint64_t w = nd.c < a.n_cached ? a.cached_mains[nd.c].width : a.common_main.width;
                SWIRL_REQUIRE(nd.a < w, "MainPartitionIndexOutOfBounds");
                SWIRL_REQUIRE(nd.b < L.stride, "row offset needs need_rot");
                op = I_VAR;
                ins.a = L.main_base + nd.c * L.stride + nd.b;
                ins.b = nd.a;
                break;
            }
            case SWIRL_NODE_VAR_PUBLIC:
                SWIRL_REQUIRE(nd.a < a.n_public_values, "PublicValueIndexOutOfBounds");
                op = I_CONST;
                ins.a = a.public_values[nd.a];
                break;
            case SWIRL_NODE_IS_FIRST: op = I_VAR; ins.a = 0; ins.b = 0; break;
            case SWIRL_NODE_IS_TRANSITION: op = I_VAR; ins.a = 0; ins.b = 1; break;
            case SWIRL_NODE_IS_LAST: op = I_VAR; ins.a = 0; ins.b = 2; break;
            case SWIRL_NODE_CONST: op = I_CONST; ins.a = nd.a; break;
            case SWIRL_NODE_ADD: op = I_ADD; break;
            case SWIRL_NODE_SUB: op = I_SUB; break;
            case SWIRL_NODE_MUL: op = I_MUL; break;
            case SWIRL_NODE_NEG: op = I_NEG; break;
            default: SWIRL_REQUIRE(false, "unknown DAG node");
        }
        if (op == I_VAR) ins.c = L.part_col_off[ins.a] + ins.b;  // global column (MLE rounds)
        if (op == I_ADD || op == I_SUB || op == I_MUL) {
            ins.a = (uint32_t)slot[nd.a];
            ins.b = (uint32_t)slot[nd.b];
        } else if (op == I_NEG) {
            ins.a = (uint32_t)slot[nd.a];
        }
        // operands die here: their slots may be reused for the result
        if (nd.op == SWIRL_NODE_ADD || nd.op == SWIRL_NODE_SUB || nd.op == SWIRL_NODE_MUL) {
            release(nd.a, i);
            if (nd.b != nd.a) release(nd.b, i);
        } else if (nd.op == SWIRL_NODE_NEG) {
            release(nd.a, i);
        }
        int s;
        if (!free_slots.empty()) {
            s = free_slots.back();
            free_slots.pop_back();
        } else {
            s = n_slots++;
        }
        slot[i] = s;
        ins.op_dst = op | ((uint32_t)s << 8);
        if (op == I_MUL && last_use[i] < 0 && root_of[i].size() == 1) {
            // a product that only feeds one accumulator (the typical constraint root): multiply and accumulate in one
            // instruction, no slot written.  op_dst = op | acc << 8, a / b = operand slots, c = weight index
            const Root& rt = roots[root_of[i][0]];
            out->code.push_back(Instr{I_MULACC | (rt.acc << 8), ins.a, ins.b, rt.weight});
            release((uint32_t)i, i);
            continue;
        }
        out->code.push_back(ins);
        for (uint32_t ri : root_of[i]) out->code.push_back(Instr{I_ACC, roots[ri].acc, roots[ri].weight, (uint32_t)s});
        if (last_use[i] < 0) release((uint32_t)i, i);  // only consumed by accumulations
    }
    SWIRL_REQUIRE(n_slots <= BC_MAX_SLOTS, "constraint DAG needs more live values than supported");
    out->n_slots = n_slots;
    if (prefetch_distance) {
        // every column load misses L1 (a chunk of a column is read by exactly one warp) and one thread walks its
        // columns serially: I_PREF announces the load `prefetch_distance` variables ahead so the DRAM latency
        // overlaps the evaluation of the variables in between
        std::vector<size_t> vars;
        for (size_t i = 0; i < out->code.size(); i++)
            if ((out->code[i].op_dst & 0xff) == I_VAR) vars.push_back(i);
        auto pref = [&](size_t j) {
            const Instr& v = out->code[vars[j]];
            return Instr{I_PREF, v.a, v.b, v.c};
        };
        std::vector<Instr> code;
        code.reserve(out->code.size() + vars.size());
        for (size_t j = 0; j < std::min(prefetch_distance, vars.size()); j++) code.push_back(pref(j));
        size_t j = 0;
        for (size_t i = 0; i < out->code.size(); i++) {
            if (j < vars.size() && vars[j] == i) {
                if (j + prefetch_distance < vars.size()) code.push_back(pref(j + prefetch_distance));
                j++;
            }
            code.push_back(out->code[i]);
        }
        out->code.swap(code);
    }
    return 0;
}

// ---- device: interpreter ------------------------------------------------------------------------
struct FVal {
    using T = uint32_t;
    static __device__ __forceinline__ T add(T a, T b) { return bb::add(a, b); }
    static __device__ __forceinline__ T sub(T a, T b) { return bb::sub(a, b); }
    static __device__ __forceinline__ T mul(T a, T b) { return bb::mul(a, b); }
    static __device__ __forceinline__ T neg(T a) { return bb::neg(a); }
    static __device__ __forceinline__ T from_base(uint32_t m) { return m; }
    static __device__ __forceinline__ Ext weigh(const Ext& w, T v) { return ext_mul_base(w, v); }
};
struct EVal {
    using T = Ext;
    static __device__ __forceinline__ T add(const T& a, const T& b) { return ext_add(a, b); }
    static __device__ __forceinline__ T sub(const T& a, const T& b) { return ext_sub(a, b); }
    static __device__ __forceinline__ T mul(const T& a, const T& b) { return ext_mul(a, b); }
    static __device__ __forceinline__ T neg(const T& a) { return bb::ext_neg(a); }
    static __device__ __forceinline__ T from_base(uint32_t m) { return bb::ext_from(m); }
    static __device__ __forceinline__ Ext weigh(const Ext& w, const T& v) { return ext_mul(w, v); }
};

// Runs the program on LN independent inputs in lockstep (one decode, LN evaluations per instruction):
// slot values and accumulators carry a lane index.  load_var(part, col, global_col, out[LN]).
// Slot storage in shared memory, one column per thread: slot s lane l of thread t at [(s * LN + l) * blockDim + t]
template <class T, int LN>
struct SharedSlots {
    T* base;  // + threadIdx.x
    int stride;
    struct Row {
        T* p;
        int stride;
        __device__ __forceinline__ T& operator[](int l) const { return p[l * stride]; }
    };
    __device__ __forceinline__ Row operator[](uint32_t s) const { return Row{base + (size_t)s * LN * stride, stride}; }
};

struct PrefetchTag {};  // load_var(part, col, global_col, PrefetchTag{}) announces a later load of that variable
template <class O>
constexpr bool is_prefetch = std::is_same<typename std::decay<O>::type, PrefetchTag>::value;
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

template <class V, int LN, class Slots, class LoadVar>
__device__ __forceinline__ void run_program_on(Slots&& slots, const Instr* __restrict__ code, uint32_t n_instr,
                                               const uint32_t* __restrict__ weights, LoadVar load_var, Ext (&acc)[LN][3]) {
    if (n_instr == 0) return;
    // the next instruction is fetched while the current one executes (the decode is a dependent L1 load)
    uint4 raw = __ldg(reinterpret_cast<const uint4*>(code));
    for (uint32_t pc = 0; pc < n_instr; pc++) {
        const uint4 cur = raw;
        if (pc + 1 < n_instr) raw = __ldg(reinterpret_cast<const uint4*>(code) + pc + 1);
        const uint32_t op = cur.x & 0xff, dst = cur.x >> 8;
        switch (op) {
            case I_VAR: load_var(cur.y, cur.z, cur.w, slots[dst]); break;
            case I_PREF: load_var(cur.y, cur.z, cur.w, PrefetchTag{}); break;
            case I_MULACC: {  // acc[.][dst] += weights[w] * (slots[y] * slots[z])
                const Ext wv = ldg_ext(weights + 4 * cur.w);
#pragma unroll
                for (int l = 0; l < LN; l++) {
                    const Ext t = V::weigh(wv, V::mul(slots[cur.y][l], slots[cur.z][l]));
                    if (dst == 0) acc[l][0] = ext_add(acc[l][0], t);
                    else if (dst == 1) acc[l][1] = ext_add(acc[l][1], t);
                    else acc[l][2] = ext_add(acc[l][2], t);
                }
                break;
            }
            case I_CONST:
#pragma unroll
                for (int l = 0; l < LN; l++) slots[dst][l] = V::from_base(cur.y);
                break;
            case I_ADD:
#pragma unroll
                for (int l = 0; l < LN; l++) slots[dst][l] = V::add(slots[cur.y][l], slots[cur.z][l]);
                break;
            case I_SUB:
#pragma unroll
                for (int l = 0; l < LN; l++) slots[dst][l] = V::sub(slots[cur.y][l], slots[cur.z][l]);
                break;
            case I_MUL:
#pragma unroll
                for (int l = 0; l < LN; l++) slots[dst][l] = V::mul(slots[cur.y][l], slots[cur.z][l]);
                break;
            case I_NEG:
#pragma unroll
                for (int l = 0; l < LN; l++) slots[dst][l] = V::neg(slots[cur.y][l]);
                break;
            default: {  // I_ACC: acc[.][y] += weights[z] * slots[w]; static accumulator indices keep acc in registers
                const Ext wv = ldg_ext(weights + 4 * cur.z);
#pragma unroll
                for (int l = 0; l < LN; l++) {
                    const Ext t = V::weigh(wv, slots[cur.w][l]);
                    if (cur.y == 0) acc[l][0] = ext_add(acc[l][0], t);
                    else if (cur.y == 1) acc[l][1] = ext_add(acc[l][1], t);
                    else acc[l][2] = ext_add(acc[l][2], t);
                }
            }
        }
    }
}

template <class V, int NS, int LN, class LoadVar>
__device__ __forceinline__ void run_program(const Instr* __restrict__ code, uint32_t n_instr,
                                            const uint32_t* __restrict__ weights, LoadVar load_var, Ext (&acc)[LN][3]) {
    typename V::T slots[NS][LN];
    run_program_on<V, LN>(slots, code, n_instr, weights, load_var, acc);
}

struct BasePart {
    const uint32_t* ptr;  // column-major, column stride = height
    uint32_t height;      // power of two
    uint32_t rot;
};

// is_first / is_transition / is_last of the lifted trace as a 3-column matrix (cpu.rs:306-322)
__global__ void sels_kernel(uint32_t* __restrict__ out, size_t lifted, size_t height) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= lifted) return;
    const size_t r = i & (height - 1);
    out[i] = r == 0 ? bb::R1 : 0u;
    out[lifted + i] = r == height - 1 ? 0u : bb::R1;
    out[2 * lifted + i] = r == height - 1 ? bb::R1 : 0u;
}

// LogUp input layer (mod.rs:103-168): one thread per (row, interaction); the interaction's own
// sub-program leaves count in acc[1] (weight 1) and sum beta^j msg_j in acc[2].
struct LeafArgs {
    const Instr* code;
    const uint32_t* prog_off;  // [n_int + 1] instruction ranges
    const BasePart* parts;
    const uint32_t* weights;      // [0] = 1, [1 + j] = beta^j
    const uint32_t* denom_const;  // per interaction: beta^len * (bus + 1) + alpha
    const uint64_t* row_idx;      // per interaction: offset in the stacked leaves
    uint32_t* leaves;
    uint32_t height, reps;  // reps = lifted length / height (cyclic repetition)
    uint32_t norm;          // 1 / reps
};
template <int NS>
__global__ void __launch_bounds__(BC_BLOCK) logup_leaves_kernel(LeafArgs a) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, sigma = blockIdx.y;
    if (i >= a.height) return;
    Ext acc[1][3] = {{bb::ext_zero(), bb::ext_zero(), bb::ext_zero()}};
    const uint32_t p0 = a.prog_off[sigma], p1 = a.prog_off[sigma + 1];
    run_program<FVal, NS, 1>(a.code + p0, p1 - p0, a.weights,
                             [&](uint32_t part, uint32_t col, uint32_t, auto&& out) {
                                 if constexpr (!is_prefetch<decltype(out)>) {
                                     const BasePart bp = a.parts[part];
                                     out[0] = __ldg(bp.ptr + (size_t)col * bp.height + ((i + bp.rot) & (bp.height - 1)));
                                 }
                             },
                             acc);
    const Ext numer = ext_mul_base(acc[0][1], a.norm);
    const Ext denom = ext_add(acc[0][2], ldg_ext(a.denom_const + 4 * sigma));
    for (uint32_t rep = 0; rep < a.reps; rep++) {
        uint32_t* leaf = a.leaves + (a.row_idx[sigma] + (size_t)rep * a.height + i) * 8;
        st_ext(leaf, numer);
        st_ext(leaf + 4, denom);
    }
}
__global__ void leaves_fill_kernel(uint32_t* __restrict__ leaves, size_t n, Ext q) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    st_ext(leaves + 8 * i, bb::ext_zero());
    st_ext(leaves + 8 * i + 4, q);
}

// Round 0: thread = (hypercube point x, coset point p).  partials[block][p * 12 + 4k + c] =
// sum_x eq_xi[x] * acc_k at point p.  All present AIRs share one launch: block -> (AIR, chunk) through
// block_air / first_block.
constexpr int R0_SMEM_PARTS = 24;
struct R0Args {
    const Instr* code;
    uint32_t n_instr;
    const BasePart* parts;
    const uint32_t* weights;
    const uint32_t* lde;    // [P][N]: Lagrange coefficients of D at the P coset points
    const uint32_t* eq_xi;  // 2^n_lift EF
    int l_skip, n_lift, P, x_per_block;
    uint32_t first_block, n_blocks;
    uint32_t n_parts;
    uint32_t* partials;  // this AIR's [n_blocks][P * 12]
    uint32_t* result;    // this AIR's [P * 12]
};
// sum_i lde[i] * c[i] over one 16-element chunk with lazy reduction (4 products per Montgomery reduction)
__device__ __forceinline__ uint32_t chunk_dot16(const uint32_t (&l)[16], const uint32_t (&c)[16]) {
    const uint32_t a0 = bb::dot4(l[0], c[0], l[1], c[1], l[2], c[2], l[3], c[3]);
    const uint32_t a1 = bb::dot4(l[4], c[4], l[5], c[5], l[6], c[6], l[7], c[7]);
    const uint32_t a2 = bb::dot4(l[8], c[8], l[9], c[9], l[10], c[10], l[11], c[11]);
    const uint32_t a3 = bb::dot4(l[12], c[12], l[13], c[13], l[14], c[14], l[15], c[15]);
    return bb::add(bb::add(a0, a1), bb::add(a2, a3));
}

// LOGN = 4: the thread's 16 Lagrange coefficients live in registers and chunks are fetched as four
// 16-byte loads; LOGN = 0: generic l_skip.  Two hypercube points per thread run in lockstep.
template <int NS, int LOGN>
__global__ void __launch_bounds__(BC_BLOCK, 3) batch_round0_kernel(const R0Args* __restrict__ descs,
                                                                const uint16_t* __restrict__ block_air) {
    extern __shared__ uint32_t sm[];  // [blockDim][13] reduction scratch, then [NS][LN][blockDim] value slots
    constexpr int LN = 2;
    const R0Args a = descs[block_air[blockIdx.x]];
    const uint32_t bidx = blockIdx.x - a.first_block;
    // trace-part descriptors in shared memory: one dependent global load less in front of every column load
    __shared__ BasePart sparts[R0_SMEM_PARTS];
    if (threadIdx.x < a.n_parts && threadIdx.x < R0_SMEM_PARTS) sparts[threadIdx.x] = a.parts[threadIdx.x];
    __syncthreads();
    const bool parts_in_smem = a.n_parts <= R0_SMEM_PARTS;
    const int P = a.P, N = 1 << a.l_skip;
    const int G = blockDim.x / P;
    const bool idle = (int)threadIdx.x >= P * G;  // blockDim is fixed; P need not divide it
    const int p = idle ? 0 : threadIdx.x % P, g = idle ? 0 : threadIdx.x / P;
    const size_t nx = size_t(1) << a.n_lift;
    const size_t x0 = (size_t)bidx * a.x_per_block, x1 = idle ? 0 : min(x0 + (size_t)a.x_per_block, nx);
    const uint32_t* lde = a.lde + (size_t)p * N;
    uint32_t lreg[16];
    if (LOGN == 4) {
#pragma unroll
        for (int i = 0; i < 16; i++) lreg[i] = __ldg(lde + i);
    }
    Ext tot[3] = {bb::ext_zero(), bb::ext_zero(), bb::ext_zero()};
    for (size_t xb = x0 + g; xb < x1; xb += (size_t)LN * G) {
        size_t xs[LN];
        bool live[LN];
#pragma unroll
        for (int l = 0; l < LN; l++) {
            live[l] = xb + (size_t)l * G < x1;
            xs[l] = live[l] ? xb + (size_t)l * G : xb;
        }
        Ext acc[LN][3];
#pragma unroll
        for (int l = 0; l < LN; l++)
#pragma unroll
            for (int k = 0; k < 3; k++) acc[l][k] = bb::ext_zero();
        auto load_var = [&](uint32_t part, uint32_t col, uint32_t, auto&& out) {
                                      const BasePart bp = parts_in_smem ? sparts[part] : a.parts[part];
                                      const uint32_t* c = bp.ptr + (size_t)col * bp.height;
                                      if constexpr (is_prefetch<decltype(out)>) {
#pragma unroll
                                          for (int l = 0; l < LN; l++) prefetch_l1(c + (((xs[l] << a.l_skip) + bp.rot) & (bp.height - 1)));
                                      } else {
#pragma unroll
                                      for (int l = 0; l < LN; l++) {
                                          const size_t r0 = xs[l] << a.l_skip;
                                          if (LOGN == 4 && bp.height >= 16) {
                                              const uint4* q = reinterpret_cast<const uint4*>(c + r0);
                                              const uint4 v0 = __ldg(q), v1 = __ldg(q + 1), v2 = __ldg(q + 2), v3 = __ldg(q + 3);
                                              uint32_t ch[16] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w,
                                                                 v2.x, v2.y, v2.z, v2.w, v3.x, v3.y, v3.z, v3.w};
                                              if (bp.rot) {  // rows r0+1 .. r0+16 (cyclic)
                                                  const uint32_t nxt = __ldg(c + ((r0 + 16) & (bp.height - 1)));
#pragma unroll
                                                  for (int i = 0; i < 15; i++) ch[i] = ch[i + 1];
                                                  ch[15] = nxt;
                                              }
                                              out[l] = chunk_dot16(lreg, ch);
                                          } else {
                                              uint32_t v = 0;
                                              for (int i = 0; i < N; i++)
                                                  v = bb::add(v, bb::mul(__ldg(lde + i), __ldg(c + ((r0 + bp.rot + i) & (bp.height - 1)))));
                                              out[l] = v;
                                          }
                                      }
                                      }
                                  };
        if (NS <= 64) {  // value slots in shared memory (short-scoreboard latency instead of local-memory round trips)
            run_program_on<FVal, LN>(SharedSlots<uint32_t, LN>{sm + blockDim.x * 13 + threadIdx.x, (int)blockDim.x}, a.code,
                                     a.n_instr, a.weights, load_var, acc);
        } else {
            run_program<FVal, NS, LN>(a.code, a.n_instr, a.weights, load_var, acc);
        }
#pragma unroll
        for (int l = 0; l < LN; l++) {
            if (!live[l]) continue;
            const Ext e = ldg_ext(a.eq_xi + 4 * xs[l]);
#pragma unroll
            for (int k = 0; k < 3; k++) tot[k] = ext_add(tot[k], ext_mul(e, acc[l][k]));
        }
    }
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
        for (int c = 0; c < 4; c++) sm[threadIdx.x * 13 + 4 * k + c] = tot[k].c[c];
    __syncthreads();
    for (int o = threadIdx.x; o < P * 12; o += blockDim.x) {
        const int pi = o / 12, k = o % 12;
        uint32_t s = 0;
        for (int gg = 0; gg < G; gg++) s = bb::add(s, sm[(gg * P + pi) * 13 + k]);
        a.partials[(size_t)bidx * (P * 12) + o] = s;
    }
}
// per AIR (blockIdx.y): result[o] = sum_b partials[b * nv + o]
__global__ void bc_reduce_multi_kernel(const R0Args* __restrict__ descs) {
    const R0Args a = descs[blockIdx.y];
    const int nv = a.P * 12;
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= nv) return;
    uint32_t s = 0;
    for (uint32_t b = 0; b < a.n_blocks; b++) s = bb::add(s, a.partials[(size_t)b * nv + o]);
    a.result[o] = s;
}
__global__ void bc_reduce_kernel(const uint32_t* __restrict__ partials, size_t nblocks, int nv, uint32_t* __restrict__ result) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= nv) return;
    uint32_t s = 0;
    for (size_t b = 0; b < nblocks; b++) s = bb::add(s, partials[b * nv + o]);
    result[o] = s;
}

// MLE round: thread per hypercube point y; every X in 1..D evaluates the program on the EF values
// t0 + (t1 - t0) X of rows (2y, 2y+1).  result[(X-1)*12 + 4k + c] = sum_y eq_xi[y] acc_k(X, y).
// single: the tables have one row; evaluate it once (D = 1, no eq factor).
struct MleArgs {
    const Instr* code;
    uint32_t n_instr;
    const uint32_t* base;  // EF, all row parts' columns concatenated, column stride h
    size_t h;
    const uint32_t* weights;
    const uint32_t* eq_xi;
    size_t ny;
    int single;
    uint32_t first_block, n_blocks;  // this AIR's blocks inside the shared launch
    uint32_t* partials;
    unsigned int* ticket;
    uint32_t* result;
    uint32_t sub, col_shift, w_shift;  // run-time compiled kernels only: case of the generated switch, operand shifts
};
// every AIR with live tables folds in one launch: out[j] = lerp(in[2j], in[2j+1], r)
struct FoldArgs {
    const uint32_t* in;
    uint32_t* out;
    size_t n_out;
    uint32_t first_block;
};
__global__ void __launch_bounds__(BC_BLOCK)
ef_fold_multi_kernel(const FoldArgs* __restrict__ descs, const uint16_t* __restrict__ block_air, Ext r, RoundLink link) {
    if (!link_wait(link, r)) return;  // linked: the challenge arrives through the mailbox (ext.cuh)
    const FoldArgs a = descs[block_air[blockIdx.x]];
    const size_t j = (size_t)(blockIdx.x - a.first_block) * blockDim.x + threadIdx.x;
    if (j >= a.n_out) return;
    st_ext(a.out + 4 * j, ext_lerp(ldg_ext(a.in + 8 * j), ldg_ext(a.in + 8 * j + 4), r));
}
#ifndef BC_MLE_MIN_BLOCKS
#define BC_MLE_MIN_BLOCKS 4
#endif
template <int NS, int D>
__global__ void __launch_bounds__(128, BC_MLE_MIN_BLOCKS) batch_mle_kernel(const MleArgs* __restrict__ descs,
                                                           const uint16_t* __restrict__ block_air, uint32_t result_tag) {
    const MleArgs a = descs[block_air[blockIdx.x]];
    const uint32_t bidx = blockIdx.x - a.first_block;
    uint32_t v[D * 12];
#pragma unroll
    for (int i = 0; i < D * 12; i++) v[i] = 0;
    for (size_t y = (size_t)bidx * blockDim.x + threadIdx.x; y < a.ny; y += (size_t)a.n_blocks * blockDim.x) {
        const Ext e = a.single ? bb::ext_one() : ldg_ext(a.eq_xi + 4 * y);
        Ext acc[D][3];
#pragma unroll
        for (int X = 0; X < D; X++)
#pragma unroll
            for (int k = 0; k < 3; k++) acc[X][k] = bb::ext_zero();
        // all X = 1..D in lockstep: one pair of loads per variable, values t1, t1 + d, t1 + 2d, ...
        run_program<EVal, NS, D>(a.code, a.n_instr, a.weights,
                                 [&](uint32_t, uint32_t, uint32_t gcol, auto&& out) {
                                     const uint32_t* c = a.base + ((size_t)gcol * a.h) * 4;
                                     if constexpr (is_prefetch<decltype(out)>) {
                                         prefetch_l1(c + (a.single ? 0 : 8 * y));
                                     } else {
                                     if (a.single) {
                                         out[0] = ldg_ext(c);
                                         return;
                                     }
                                     const Ext t0 = ldg_ext(c + 8 * y), t1 = ldg_ext(c + 8 * y + 4);
                                     const Ext d = ext_sub(t1, t0);
                                     out[0] = t1;
#pragma unroll
                                     for (int X = 1; X < D; X++) out[X] = ext_add(out[X - 1], d);
                                     }
                                 },
                                 acc);
#pragma unroll
        for (int X = 0; X < D; X++)
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const Ext t = ext_mul(e, acc[X][k]);
#pragma unroll
                for (int c = 0; c < 4; c++) v[X * 12 + 4 * k + c] = bb::add(v[X * 12 + 4 * k + c], t.c[c]);
            }
    }
    group_sum<D * 12>(v, a.partials, a.ticket, a.result, a.n_blocks, bidx, result_tag);
}

template <int NS>
static void launch_mle(int D, const MleArgs* descs, const uint16_t* block_air, int grid, cudaStream_t st, uint32_t tag) {
    switch (D) {
        case 1: batch_mle_kernel<NS, 1><<<grid, 128, 0, st>>>(descs, block_air, tag); break;
        case 2: batch_mle_kernel<NS, 2><<<grid, 128, 0, st>>>(descs, block_air, tag); break;
        case 3: batch_mle_kernel<NS, 3><<<grid, 128, 0, st>>>(descs, block_air, tag); break;
        case 4: batch_mle_kernel<NS, 4><<<grid, 128, 0, st>>>(descs, block_air, tag); break;
        default: batch_mle_kernel<NS, 5><<<grid, 128, 0, st>>>(descs, block_air, tag); break;
    }
}
// NS buckets keep the per-thread slot array (local memory) as small as the program allows
#define BC_DISPATCH_NS(ns, CALL)          \
    do {                                  \
        if ((ns) <= 16) { CALL(16); }     \
        else if ((ns) <= 32) { CALL(32); }   \
        else if ((ns) <= 64) { CALL(64); }   \
        else if ((ns) <= 128) { CALL(128); } \
        else { CALL(256); }               \
    } while (0)

// ---- host: program -> straight-line CUDA C++ (jit.hpp) -----------------------------------------------------
// One statement per instruction; value slot s lane l is the register variable s<s>_<l>.  The statements are the
// macros of csrc/jit_prelude.cuh (LD = column load at the thread's coset point, PF = prefetch, ACC = weighted
// accumulation into one of the three EF accumulators).
// Programs of real AIRs repeat: the same few instructions applied to consecutive columns / weights (BenchmarkAir:
// assert_bool on every column; range checks, byte decompositions and bus messages in production chips).  A run of r >= 4
// repetitions of a period of p instructions whose column / weight operands advance by constant steps is emitted as ONE loop
// (`it` = repetition), which keeps the generated kernel small enough to compile in seconds and to live in the instruction
// cache; everything else is emitted statement by statement.
struct JitRun {
    size_t start, period, reps;
    std::vector<std::array<int64_t, 3>> delta;  // per instruction of the period: step of (a, b, c)
};
static bool jit_delta_ok(const Instr& x, const Instr& y, std::array<int64_t, 3>* d) {
    if (x.op_dst != y.op_dst) return false;
    (*d)[0] = (int64_t)y.a - (int64_t)x.a;
    (*d)[1] = (int64_t)y.b - (int64_t)x.b;
    (*d)[2] = (int64_t)y.c - (int64_t)x.c;
    switch (x.op_dst & 0xff) {
        case I_VAR:
        case I_PREF: return (*d)[0] == 0;                   // part fixed; column (b) and global column (c) may step
        case I_ACC: return (*d)[0] == 0 && (*d)[2] == 0;     // accumulator and slot fixed; weight (b) may step
        case I_MULACC: return (*d)[0] == 0 && (*d)[1] == 0;  // operand slots fixed; weight (c) may step
        default: return (*d)[0] == 0 && (*d)[1] == 0 && (*d)[2] == 0;
    }
}
static std::vector<JitRun> jit_find_runs(const std::vector<Instr>& code) {
    std::vector<JitRun> runs;
    const size_t n = code.size();
    size_t i = 0;
    while (i < n) {
        JitRun best{i, 0, 0, {}};
        for (size_t p = 1; p <= 48 && i + 2 * p <= n; p++) {
            std::vector<std::array<int64_t, 3>> d(p);
            bool ok = true;
            for (size_t j = 0; j < p && ok; j++) ok = jit_delta_ok(code[i + j], code[i + p + j], &d[j]);
            if (!ok) continue;
            size_t r = 2;
            for (;; r++) {
                if (i + (r + 1) * p > n) break;
                bool same = true;
                for (size_t j = 0; j < p && same; j++) {
                    std::array<int64_t, 3> e;
                    same = jit_delta_ok(code[i + (r - 1) * p + j], code[i + r * p + j], &e) && e == d[j];
                }
                if (!same) break;
            }
            if (r >= 4 && r * p > best.reps * best.period) best = JitRun{i, p, r, d};
        }
        if (best.reps) {
            runs.push_back(best);
            i += best.period * best.reps;
        } else {
            i++;
        }
    }
    return runs;
}

// One statement per instruction; value slot s lane l is the register variable s<s>_<l>.  The statements are the
// macros of csrc/jit_prelude.cuh (LD = column load at the thread's coset point, PF = prefetch, ACC = weighted
// accumulation into one of the three EF accumulators).  Returns "" when the program is too irregular to be worth
// compiling (more than JIT_MAX_STATEMENTS statements after loop detection): NVRTC's time grows much faster than
// linearly with the size of a basic block (measured: 768 statements with inlined loads = 13 minutes).
constexpr size_t JIT_MAX_STATEMENTS = 320;
static std::string generate_round0_source(const std::vector<Instr>& code, int n_slots, const char* name) {
    const std::vector<JitRun> runs = jit_find_runs(code);
    size_t statements = code.size(), loads = 0;
    for (const JitRun& r : runs) statements -= r.period * (r.reps - 1);
    if (statements > JIT_MAX_STATEMENTS) return std::string();
    {
        size_t ri = 0;
        for (size_t i = 0; i < code.size();) {
            const bool in_run = ri < runs.size() && runs[ri].start == i;
            const size_t len = in_run ? runs[ri].period : 1;
            for (size_t j = 0; j < len; j++) loads += (code[i + j].op_dst & 0xff) == I_VAR;
            i += in_run ? runs[ri].period * runs[ri].reps : 1;
            ri += in_run;
        }
    }
    std::string s;
    s.reserve(statements * 128 + 16384);
    // the column load (4 x 16-byte loads + a 16-term dot product) is inlined while the kernel stays small
    s += loads <= 24 ? "#define SW_LOAD_ATTR __device__ __forceinline__\n" : "#define SW_LOAD_ATTR __device__ __noinline__\n";
    {
        int mb = 3;  // 85 registers: 2.98 ms against 3.57 ms at 2 (128 registers) for C2, profiles/r2s_*
        if (const char* env = getenv("SWIRL_JIT_MIN_BLOCKS")) mb = std::max(1, std::min(4, atoi(env)));  // experiment knob
        s += "#define SW_MIN_BLOCKS " + std::to_string(mb) + "\n";
    }
    s += jit_prelude();
    s += "\nSW_R0_SIGNATURE(";
    s += name;
    s += ") {\nSW_R0_PROLOGUE\n";
    for (int i = 0; i < n_slots; i++) s += "uint32_t s" + std::to_string(i) + "_0 = 0, s" + std::to_string(i) + "_1 = 0;\n";
    auto v = [](uint32_t slot, int lane) { return "s" + std::to_string(slot) + "_" + std::to_string(lane); };
    // operand `val` (+ it * step inside a loop)
    auto opnd = [](uint32_t val, int64_t step) {
        if (step == 0) return std::to_string(val);
        return "(" + std::to_string(val) + " + it * (" + std::to_string(step) + "))";
    };
    auto emit = [&](const Instr& in, const std::array<int64_t, 3>& d) {
        const uint32_t op = in.op_dst & 0xff, dst = in.op_dst >> 8;
        auto bin = [&](const char* fn) {
            for (int l = 0; l < 2; l++) s += v(dst, l) + " = " + fn + "(" + v(in.a, l) + ", " + v(in.b, l) + "); ";
            s += "\n";
        };
        switch (op) {
            case I_VAR: s += "LD(" + std::to_string(in.a) + ", " + opnd(in.b, d[1]) + ", " + v(dst, 0) + ", " + v(dst, 1) + ")\n"; break;
            case I_PREF: s += "PF(" + std::to_string(in.a) + ", " + opnd(in.b, d[1]) + ")\n"; break;
            case I_CONST: s += v(dst, 0) + " = " + v(dst, 1) + " = " + std::to_string(in.a) + "u;\n"; break;
            case I_ADD: bin("f_add"); break;
            case I_SUB: bin("f_sub"); break;
            case I_MUL: bin("f_mul"); break;
            case I_NEG:
                for (int l = 0; l < 2; l++) s += v(dst, l) + " = f_neg(" + v(in.a, l) + "); ";
                s += "\n";
                break;
            case I_MULACC:
                s += "ACC(" + std::to_string(dst) + ", " + opnd(in.c, d[2]) + ", f_mul(" + v(in.a, 0) + ", " + v(in.b, 0) + "), f_mul(" +
                     v(in.a, 1) + ", " + v(in.b, 1) + "))\n";
                break;
            default:  // I_ACC: a = accumulator, b = weight index, c = slot
                s += "ACC(" + std::to_string(in.a) + ", " + opnd(in.b, d[1]) + ", " + v(in.c, 0) + ", " + v(in.c, 1) + ")\n";
        }
    };
    const std::array<int64_t, 3> zero{0, 0, 0};
    size_t ri = 0;
    for (size_t i = 0; i < code.size();) {
        if (ri < runs.size() && runs[ri].start == i) {
            const JitRun& r = runs[ri++];
            s += "#pragma unroll 1\nfor (int it = 0; it < " + std::to_string(r.reps) + "; it++) {\n";
            for (size_t j = 0; j < r.period; j++) emit(code[i + j], r.delta[j]);
            s += "}\n";
            i += r.period * r.reps;
        } else {
            emit(code[i], zero);
            i++;
        }
    }
    s += "SW_R0_EPILOGUE\n}\n";
    return s;
}

// ---- host: an AIR's sub-programs -> ONE CUDA C++ kernel for the MLE rounds ------------------------------------------
// The MLE rounds launch a descriptor per (AIR, sub-program).  The generated kernel is a switch over the sub-programs'
// code ("cases"); sub-programs that are equal up to one constant shift of their global columns and one of their weight
// indices -- the same 16 constraints on the next 16 columns -- share a case and carry the shifts in their descriptor.
// Value slots are arrays of D extension-field lanes (csrc/jit_prelude.cuh, SW_D section).
struct MleCaseRef {
    int case_id = -1;
    uint32_t col_shift = 0, w_shift = 0;
};
constexpr size_t JIT_MLE_MAX_STATEMENTS = 256;  // all cases together, after loop detection
constexpr int JIT_MLE_MAX_LANE_SLOTS = 40;      // slots x D of one case: 4 registers each

// y == ref with every global column (I_VAR / I_PREF: c) shifted by *dc and every weight index (I_ACC: b, I_MULACC: c) by *dw
static bool mle_equal_up_to_shift(const std::vector<Instr>& ref, const std::vector<Instr>& y, uint32_t* dc, uint32_t* dw) {
    if (ref.size() != y.size()) return false;
    bool have_c = false, have_w = false;
    *dc = *dw = 0;
    auto shift = [](bool& have, uint32_t* d, uint32_t from, uint32_t to) {
        const uint32_t delta = to - from;  // modulo 2^32, like the device's addition
        if (!have) {
            have = true;
            *d = delta;
        }
        return *d == delta;
    };
    for (size_t i = 0; i < ref.size(); i++) {
        const Instr &r = ref[i], &x = y[i];
        if (r.op_dst != x.op_dst) return false;
        switch (r.op_dst & 0xff) {
            case I_VAR:
            case I_PREF:  // part and column-in-part (a, b) are not read by the MLE rounds
                if (!shift(have_c, dc, r.c, x.c)) return false;
                break;
            case I_ACC:
                if (r.a != x.a || r.c != x.c || !shift(have_w, dw, r.b, x.b)) return false;
                break;
            case I_MULACC:
                if (r.a != x.a || r.b != x.b || !shift(have_w, dw, r.c, x.c)) return false;
                break;
            case I_CONST:
                if (r.a != x.a) return false;
                break;
            case I_NEG:
                if (r.a != x.a) return false;
                break;
            default:
                if (r.a != x.a || r.b != x.b) return false;
        }
    }
    return true;
}

// `codes[i]` / `n_slots[i]`: sub-program i.  refs[i] = its case and shifts.  Returns "" when the kernel would be too
// large to compile in seconds or a case needs more registers than a thread has (the interpreter runs instead).
// With `listing`, every sub-program's instructions are appended as comment lines (tests replay them on the host).
static std::string generate_mle_source(const std::vector<const std::vector<Instr>*>& codes, const std::vector<int>& n_slots, int D,
                                       const char* name, std::vector<MleCaseRef>* refs, bool listing = false) {
    refs->assign(codes.size(), MleCaseRef{});
    std::vector<size_t> case_of;  // case -> representative sub-program
    for (size_t i = 0; i < codes.size(); i++) {
        for (size_t k = 0; k < case_of.size() && (*refs)[i].case_id < 0; k++) {
            uint32_t dc, dw;
            if (mle_equal_up_to_shift(*codes[case_of[k]], *codes[i], &dc, &dw)) (*refs)[i] = MleCaseRef{(int)k, dc, dw};
        }
        if ((*refs)[i].case_id < 0) {
            (*refs)[i] = MleCaseRef{(int)case_of.size(), 0, 0};
            case_of.push_back(i);
        }
    }
    size_t statements = 0;
    std::vector<std::vector<JitRun>> runs(case_of.size());
    for (size_t k = 0; k < case_of.size(); k++) {
        const std::vector<Instr>& code = *codes[case_of[k]];
        if (n_slots[case_of[k]] * D > JIT_MLE_MAX_LANE_SLOTS) return std::string();
        runs[k] = jit_find_runs(code);
        size_t n = code.size();
        for (const JitRun& r : runs[k]) n -= r.period * (r.reps - 1);
        statements += n;
    }
    if (statements > JIT_MLE_MAX_STATEMENTS) return std::string();
    std::string s;
    s.reserve(statements * 96 + 32768);
    s += "#define SW_LOAD_ATTR __device__ __noinline__\n#define SW_MIN_BLOCKS 1\n";  // the round-0 part of the prelude is not used
    s += "#define SW_D " + std::to_string(D) + "\n#define SW_MLE_MIN_BLOCKS 4\n";
    s += jit_prelude();
    s += "\nSW_MLE_SIGNATURE(";
    s += name;
    s += ") {\nSW_MLE_PROLOGUE\n";
    auto v = [](uint32_t slot) { return "s" + std::to_string(slot); };
    auto opnd = [](uint32_t val, int64_t step) {
        if (step == 0) return std::to_string(val) + "u";
        return "(" + std::to_string(val) + "u + (uint32_t)(it * (" + std::to_string(step) + ")))";
    };
    auto emit = [&](const Instr& in, const std::array<int64_t, 3>& d) {
        const uint32_t op = in.op_dst & 0xff, dst = in.op_dst >> 8;
        switch (op) {
            case I_VAR: s += "LDV(" + v(dst) + ", " + opnd(in.c, d[2]) + ")\n"; break;
            case I_PREF: s += "PFV(" + opnd(in.c, d[2]) + ")\n"; break;
            case I_CONST: s += "CST(" + v(dst) + ", " + std::to_string(in.a) + "u)\n"; break;
            case I_ADD: s += "ADD(" + v(dst) + ", " + v(in.a) + ", " + v(in.b) + ")\n"; break;
            case I_SUB: s += "SUB(" + v(dst) + ", " + v(in.a) + ", " + v(in.b) + ")\n"; break;
            case I_MUL: s += "MUL(" + v(dst) + ", " + v(in.a) + ", " + v(in.b) + ")\n"; break;
            case I_NEG: s += "NEG(" + v(dst) + ", " + v(in.a) + ")\n"; break;
            case I_MULACC: s += "MACV(" + std::to_string(dst) + ", " + opnd(in.c, d[2]) + ", " + v(in.a) + ", " + v(in.b) + ")\n"; break;
            default: s += "ACCV(" + std::to_string(in.a) + ", " + opnd(in.b, d[1]) + ", " + v(in.c) + ")\n";  // I_ACC
        }
    };
    const std::array<int64_t, 3> zero{0, 0, 0};
    for (size_t k = 0; k < case_of.size(); k++) {
        const std::vector<Instr>& code = *codes[case_of[k]];
        s += "case " + std::to_string(k) + ": {\n";
        for (int i = 0; i < n_slots[case_of[k]]; i++) s += "SLOT(" + v((uint32_t)i) + ")\n";
        size_t ri = 0;
        for (size_t i = 0; i < code.size();) {
            if (ri < runs[k].size() && runs[k][ri].start == i) {
                const JitRun& r = runs[k][ri++];
                s += "#pragma unroll 1\nfor (int it = 0; it < " + std::to_string(r.reps) + "; it++) {\n";
                for (size_t j = 0; j < r.period; j++) emit(code[i + j], r.delta[j]);
                s += "}\n";
                i += r.period * r.reps;
            } else {
                emit(code[i], zero);
                i++;
            }
        }
        s += "} break;\n";
    }
    s += "SW_MLE_EPILOGUE\n}\n";
    if (listing) {
        for (size_t i = 0; i < codes.size(); i++) {
            s += "// SUB " + std::to_string(i) + " case " + std::to_string((*refs)[i].case_id) + " col_shift " + std::to_string((*refs)[i].col_shift) +
                 " w_shift " + std::to_string((*refs)[i].w_shift) + " n_slots " + std::to_string(n_slots[i]) + "\n";
            for (const Instr& in : *codes[i])
                s += "// I " + std::to_string(i) + " " + std::to_string(in.op_dst & 0xff) + " " + std::to_string(in.op_dst >> 8) + " " +
                     std::to_string(in.a) + " " + std::to_string(in.b) + " " + std::to_string(in.c) + "\n";
        }
    }
    return s;
}

// the sub-programs of an AIR: K chunks shared between constraint roots and interaction roots in proportion to their
// counts, at least one per non-empty class (never mixed: the zerocheck part of round 0 is needed on one coset fewer than
// the LogUp part, cpu.rs:338-361 vs :405-409)
struct ChunkRange {
    size_t r0, r1;
    bool zerocheck;
};
static std::vector<ChunkRange> chunk_ranges(size_t nc, size_t n_roots, size_t n_airs) {
    std::vector<ChunkRange> out;
    const size_t k_max = std::max<size_t>(2, std::min<size_t>(BC_MAX_CHUNKS, 240 / n_airs));
    const size_t K = std::max<size_t>(1, std::min(k_max, (n_roots + BC_CHUNK_ROOTS - 1) / BC_CHUNK_ROOTS));
    size_t k_zc = nc ? std::max<size_t>(1, K * nc / std::max<size_t>(n_roots, 1)) : 0;
    size_t k_lg = n_roots > nc ? std::max<size_t>(1, K - std::min(K, k_zc)) : 0;
    if (k_zc + k_lg > k_max && k_zc > 1) k_zc = k_max - k_lg;
    for (int cls = 0; cls < 2; cls++) {
        const size_t base = cls == 0 ? 0 : nc, cnt = cls == 0 ? nc : n_roots - nc, kk = cls == 0 ? k_zc : k_lg;
        for (size_t k = 0; k < kk; k++) {
            const size_t r0 = base + cnt * k / kk, r1 = base + cnt * (k + 1) / kk;
            if (r0 != r1) out.push_back(ChunkRange{r0, r1, cls == 0});
        }
    }
    if (out.empty()) out.push_back(ChunkRange{0, 0, false});  // no roots at all: one empty program keeps the descriptor logic uniform
    return out;
}

}  // namespace swirl

using namespace swirl;

namespace {

struct TraceState {
    const swirl_air_ctx* air = nullptr;
    int log_height = 0, n = 0, n_lift = 0;
    size_t height = 0, lifted = 0;
    AirLayout L;
    // the constraint program split by roots into independent sub-programs (their accumulators add up): one thread
    // walks ~16 roots instead of the whole DAG, which multiplies the loads in flight per SM and divides the
    // serial latency of the short tail rounds
    struct Jit {  // the program as a run-time compiled kernel (jit.hpp); built at first use by a tall trace
        std::vector<Instr> h_code;
        int n_slots = 0;
        int state = 0;  // 0 = not tried, 1 = ready, -1 = unavailable (interpreter is used)
        JitKernel kernel;
        ~Jit() { jit_release(&kernel); }
    };
    struct Chunk {
        Instr* d_code = nullptr;
        uint32_t n_instr = 0;
        int n_slots = 0;
        bool zerocheck_only = false;  // only constraint roots: round 0 needs it on d - 1 cosets, not d
        std::shared_ptr<Jit> jit;     // whole programs only
        std::shared_ptr<std::vector<Instr>> h_code;  // sub-programs only: host copy for the MLE-round generator
    };
    struct MleJit {  // all sub-programs of the AIR as one run-time compiled kernel for the MLE rounds
        int state = 0;  // 0 = not tried, 1 = ready, -1 = unavailable
        int D = 0;
        JitKernel kernel;
        std::vector<MleCaseRef> refs;  // per sub-program: case and operand shifts
        ~MleJit() { jit_release(&kernel); }
    };
    std::shared_ptr<MleJit> mle_jit;
    std::vector<Chunk> chunks;
    Chunk whole[2];  // the constraint roots / the interaction roots as one program each (round 0 of tall traces)
    BasePart* d_parts = nullptr;
    uint32_t* d_sels = nullptr;
    uint32_t* d_weights = nullptr;
    uint32_t* d_eq_xi = nullptr;
    uint32_t total_cols = 0;
    uint32_t* ef[2] = {nullptr, nullptr};
    int cur = 0;
    size_t h = 0;
    std::vector<Ext> eq_3b;
    Ext denom_const = bb::ext_zero();
    Ext zc_tilde = bb::ext_zero(), lg_tilde[2] = {bb::ext_zero(), bb::ext_zero()};
    uint32_t norm = bb::R1;  // 1 / 2^max(l_skip - log_height, 0)
};

struct ProgramCache {
    struct Entry {
        TraceState::Chunk whole[2];
        std::vector<TraceState::Chunk> chunks;
        std::shared_ptr<TraceState::MleJit> mle_jit;
    };
    std::map<std::string, Entry> entries;
    std::vector<void*> buffers;  // device code of all entries
};
ProgramCache* program_cache(swirl_ctx* ctx) {
    if (!ctx->program_cache) ctx->program_cache = new ProgramCache();
    return (ProgramCache*)ctx->program_cache;
}

int n_logup_of(int l_skip, uint64_t total) {
    if (!total) return 0;
    int bits = 0;
    while (total >> bits) bits++;
    return bits - l_skip;
}

size_t bc_words(int l_skip, int D, const swirl_air_ctx* airs, size_t n_airs, int* L_out, int* n_max_out) {
    uint64_t total = 0;
    int n_max = 0;
    for (size_t t = 0; t < n_airs; t++) {
        const int lh = ilog2(airs[t].common_main.height);
        total += (uint64_t)airs[t].n_interactions << std::max(lh, l_skip);
        n_max = std::max(n_max, lh - l_skip);
    }
    const int L = total ? l_skip + n_logup_of(l_skip, total) : 0;
    size_t n = 1 + 4 + (size_t)L * 16 + (size_t)L * (L > 0 ? L - 1 : 0) / 2 * 12 + n_airs * 8 +
               ((size_t)(D + 1) * ((size_t(1) << l_skip) - 1) + 1) * 4 + (size_t)n_max * (D + 1) * 4;
    for (size_t t = 0; t < n_airs; t++) {
        size_t w = airs[t].common_main.width;
        for (uint64_t i = 0; i < airs[t].n_cached; i++) w += airs[t].cached_mains[i].width;
        if (airs[t].preprocessed) w += airs[t].preprocessed->width;
        n += w * (airs[t].need_rot ? 2 : 1) * 4;
    }
    if (L_out) *L_out = L;
    if (n_max_out) *n_max_out = n_max;
    return n;
}

// a * b (plain convolution; exact product, what the DFT-domain product of mod.rs:208-236 computes)
std::vector<Ext> poly_mul(const std::vector<Ext>& a, const std::vector<Ext>& b, size_t out_len) {
    std::vector<Ext> out(out_len, bb::ext_zero());
    for (size_t i = 0; i < a.size(); i++)
        for (size_t j = 0; j < b.size() && i + j < out_len; j++) out[i + j] = ext_add(out[i + j], ext_mul(a[i], b[j]));
    return out;
}

}  // namespace

namespace swirl {
void program_cache_clear(swirl_ctx* ctx) {
    ProgramCache* c = (ProgramCache*)ctx->program_cache;
    if (!c) return;
    cudaStreamSynchronize(ctx->stream);
    for (void* p : c->buffers) cudaFree(p);
    delete c;
    ctx->program_cache = nullptr;
}
}  // namespace swirl

extern "C" size_t swirl_batch_constraints_proof_words(int l_skip, int max_constraint_degree, const swirl_air_ctx* airs,
                                                      size_t n_airs) {
    if (!airs || !n_airs || l_skip < 0 || max_constraint_degree < 0) return 0;
    for (size_t t = 0; t < n_airs; t++)
        if (!is_pow2(airs[t].common_main.height)) return 0;
    return bc_words(l_skip, max_constraint_degree, airs, n_airs, nullptr, nullptr);
}

extern "C" int swirl_prove_batch_constraints(swirl_ctx* ctx, swirl_transcript* ts, int l_skip, int max_constraint_degree,
                                             int logup_pow_bits, const swirl_air_ctx* airs, size_t n_airs, uint32_t* h_proof,
                                             size_t proof_words, uint32_t* h_r) {
    SWIRL_REQUIRE(ctx && ts && airs && n_airs >= 1 && h_proof && h_r, "null argument");
    SWIRL_REQUIRE(l_skip >= 0 && l_skip <= 6, "l_skip must be in [0, 6]");
    const int D = max_constraint_degree;
    SWIRL_REQUIRE(D >= 1 && D <= 5, "max_constraint_degree must be in [1, 5]");
    SWIRL_CUDA(cudaSetDevice(ctx->device));
    const size_t N = size_t(1) << l_skip;
    Transcript tr(ts);
    RoundScratch* rs;
    SWIRL_TRY(round_scratch_get(ctx, &rs));
    std::vector<void*> to_free;
    struct Cleanup {
        swirl_ctx* ctx;
        std::vector<void*>& v;
        ~Cleanup() {
            for (void* p : v) arena_free_block(ctx, p);
        }
    } cleanup{ctx, to_free};
    auto upload = [&](const void* src, size_t bytes, void** dst) -> int {
        SWIRL_CUDA(arena_alloc(ctx, dst, bytes ? bytes : 4));
        to_free.push_back(*dst);
        if (bytes) SWIRL_CUDA(cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
        return 0;
    };

    // SWIRL_TRACE=1: wall-clock per phase on stderr (each mark synchronises the stream first)
    static const bool trace_on = getenv("SWIRL_TRACE") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    auto mark = [&](const char* name) {
        if (!trace_on) return;
        cudaStreamSynchronize(ctx->stream);
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[swirl trace] bc %-14s %8.3f ms\n", name, std::chrono::duration<double, std::milli>(now - t_prev).count());
        t_prev = now;
    };
    // ---- traces ------------------------------------------------------------------------------------
    std::vector<TraceState> T(n_airs);
    uint64_t total_interactions = 0;
    for (size_t t = 0; t < n_airs; t++) {
        TraceState& s = T[t];
        s.air = &airs[t];
        SWIRL_REQUIRE(is_pow2(airs[t].common_main.height), "trace height must be a power of two");
        s.height = airs[t].common_main.height;
        s.log_height = ilog2(s.height);
        if (t) SWIRL_REQUIRE(s.log_height <= T[t - 1].log_height, "AIRs must be sorted by descending height");
        s.n = s.log_height - l_skip;
        s.n_lift = std::max(s.n, 0);
        s.lifted = std::max(s.height, N);
        SWIRL_REQUIRE((int)airs[t].constraint_degree <= D, "AIR constraint degree exceeds max_constraint_degree");
        SWIRL_REQUIRE(airs[t].constraint_degree * N <= BC_BLOCK, "constraint_degree * 2^l_skip > 256 unsupported");
        for (uint64_t i = 0; i < airs[t].n_cached; i++)
            SWIRL_REQUIRE(airs[t].cached_mains[i].height == s.height, "cached trace height mismatch");
        if (airs[t].preprocessed) SWIRL_REQUIRE(airs[t].preprocessed->height == s.height, "preprocessed trace height mismatch");
        SWIRL_TRY(air_layout(airs[t], &s.L));
        s.total_cols = s.L.part_col_off.back() + s.L.part_width.back();
        s.norm = bb::inv(bb::to_mont((uint32_t)(s.lifted / s.height)));
        total_interactions += (uint64_t)airs[t].n_interactions << std::max(s.log_height, l_skip);
    }
    int L = 0, n_max = 0;
    SWIRL_REQUIRE(proof_words == bc_words(l_skip, D, airs, n_airs, &L, &n_max), "proof buffer size");
    const int n_logup = n_logup_of(l_skip, total_interactions);
    // interactions layout: StackedLayout::new(0, l_skip + n_logup, (num_interactions, log_lifted_height))
    Layout ilayout;
    {
        std::vector<uint64_t> widths(n_airs);
        std::vector<int32_t> lhs(n_airs);
        for (size_t t = 0; t < n_airs; t++) {
            widths[t] = airs[t].n_interactions;
            lhs[t] = std::max(T[t].log_height, l_skip);
        }
        SWIRL_TRY(make_layout(0, l_skip + n_logup, n_airs, widths.data(), lhs.data(), &ilayout));
    }
    // proof sections
    uint32_t* p = h_proof;
    uint32_t* sec_pow = p; p += 1;
    uint32_t* sec_q0 = p; p += 4;
    uint32_t* sec_claims = p; p += (size_t)L * 16;
    uint32_t* sec_gkr_polys = p; p += (size_t)L * (L > 0 ? L - 1 : 0) / 2 * 12;
    uint32_t* sec_numer = p; p += n_airs * 4;
    uint32_t* sec_denom = p; p += n_airs * 4;
    uint32_t* sec_uni = p; p += ((size_t)(D + 1) * (N - 1) + 1) * 4;
    uint32_t* sec_rounds = p; p += (size_t)n_max * (D + 1) * 4;
    uint32_t* sec_open = p;

    SWIRL_TRY(transcript_grind(ctx, ts, logup_pow_bits, sec_pow));
    const Ext alpha = tr.sample_ext(), beta = tr.sample_ext();
    size_t max_len = 0;
    for (size_t t = 0; t < n_airs; t++)
        for (uint64_t i = 0; i < airs[t].n_interactions; i++) max_len = std::max<size_t>(max_len, airs[t].interactions[i].msg_len);
    std::vector<Ext> beta_pows(max_len + 1);
    {
        Ext b = bb::ext_one();
        for (auto& x : beta_pows) {
            x = b;
            b = ext_mul(b, beta);
        }
    }

    mark("grind+setup");
    if (program_cache(ctx)->entries.size() >= 1024) program_cache_clear(ctx);  // bound the cache; nothing of it is in use here
    // ---- per trace: selector matrix, base parts, programs ---------------------------------------------
    for (size_t t = 0; t < n_airs; t++) {
        TraceState& s = T[t];
        const swirl_air_ctx& a = airs[t];
        SWIRL_CUDA(dev_alloc(ctx, &s.d_sels, 3 * s.lifted));
        to_free.push_back(s.d_sels);
        sels_kernel<<<(unsigned)((s.lifted + 255) / 256), 256, 0, ctx->stream>>>(s.d_sels, s.lifted, s.height);
        SWIRL_LAUNCH_CHECK(ctx);
        std::vector<BasePart> parts;
        parts.push_back(BasePart{s.d_sels, (uint32_t)s.lifted, 0});
        auto push = [&](const swirl_matrix& m) {
            parts.push_back(BasePart{m.data, (uint32_t)m.height, 0});
            if (a.need_rot) parts.push_back(BasePart{m.data, (uint32_t)m.height, 1});
        };
        if (a.preprocessed) push(*a.preprocessed);
        for (uint64_t i = 0; i < a.n_cached; i++) push(a.cached_mains[i]);
        push(a.common_main);
        SWIRL_TRY(upload(parts.data(), parts.size() * sizeof(BasePart), (void**)&s.d_parts));
        // full program: constraints (acc 0, weights 0..nc), interactions (acc 1 / acc 2)
        std::vector<Root> roots;
        for (uint64_t k = 0; k < a.n_constraints; k++) roots.push_back(Root{a.constraint_idx[k], 0, (uint32_t)k});
        uint32_t w = (uint32_t)a.n_constraints;
        for (uint64_t i = 0; i < a.n_interactions; i++) {
            const swirl_interaction& it = a.interactions[i];
            roots.push_back(Root{it.count_node, 1, w++});
            for (uint32_t j = 0; j < it.msg_len; j++) roots.push_back(Root{a.msg_nodes[it.msg_offset + j], 2, w++});
        }
        // Compiled programs are cached in the context, keyed by everything they depend on (DAG, roots, public values,
        // part layout, chunk budget): a prover proves the same AIRs again and again, and the reference builds its rules
        // once per proving key (cuda-backend/src/logup_zerocheck/rules/mod.rs:27-130).
        std::string key;
        {
            auto put = [&](const void* p, size_t n) { key.append((const char*)p, n); };
            const uint64_t hdr[6] = {a.n_nodes, a.n_constraints, a.n_interactions, a.n_public_values, (uint64_t)a.need_rot,
                                     std::max<size_t>(2, std::min<size_t>(BC_MAX_CHUNKS, 240 / n_airs))};
            put(hdr, sizeof(hdr));
            put(a.nodes, a.n_nodes * sizeof(swirl_dag_node));
            put(a.constraint_idx, a.n_constraints * 4);
            put(a.interactions, a.n_interactions * sizeof(swirl_interaction));
            for (uint64_t i = 0; i < a.n_interactions; i++)
                put(a.msg_nodes + a.interactions[i].msg_offset, a.interactions[i].msg_len * 4);
            put(a.public_values, a.n_public_values * 4);
            put(s.L.part_width.data(), s.L.part_width.size() * 4);
        }
        ProgramCache* cache = program_cache(ctx);
        auto hit = cache->entries.find(key);
        if (hit != cache->entries.end()) {
            s.whole[0] = hit->second.whole[0];
            s.whole[1] = hit->second.whole[1];
            s.chunks = hit->second.chunks;
            s.mle_jit = hit->second.mle_jit;
        } else {
            // sub-programs never mix constraint and interaction roots: the zerocheck part of round 0 is needed on one
            // coset fewer than the LogUp part (cpu.rs:338-361 vs :405-409)
            const size_t nc = a.n_constraints;
            auto compile_range = [&](size_t r0, size_t r1, bool zc, TraceState::Chunk* c, bool keep_host = false, bool keep_mle = false) -> int {
                Program pr;
                SWIRL_TRY(compile_program(a, s.L, std::vector<Root>(roots.begin() + r0, roots.begin() + r1), &pr, BC_PREFETCH_VARS));
                c->n_instr = (uint32_t)pr.code.size();
                c->n_slots = pr.n_slots;
                c->zerocheck_only = zc;
                if (keep_host) {
                    c->jit = std::make_shared<TraceState::Jit>();
                    c->jit->h_code = pr.code;
                    c->jit->n_slots = pr.n_slots;
                }
                if (keep_mle) c->h_code = std::make_shared<std::vector<Instr>>(pr.code);
                c->d_code = nullptr;  // an empty program is never dereferenced
                if (!pr.code.empty()) {  // owned by the cache, released with the context
                    SWIRL_CUDA(cudaMalloc((void**)&c->d_code, pr.code.size() * sizeof(Instr)));
                    cache->buffers.push_back(c->d_code);
                    SWIRL_CUDA(cudaMemcpyAsync(c->d_code, pr.code.data(), pr.code.size() * sizeof(Instr), cudaMemcpyHostToDevice, ctx->stream));
                    SWIRL_CUDA(swirl::stream_sync(ctx, __FILE__, __LINE__));  // pr.code is a temporary
                }
                return 0;
            };
            SWIRL_TRY(compile_range(0, nc, true, &s.whole[0], true));
            SWIRL_TRY(compile_range(nc, roots.size(), false, &s.whole[1], true));
            for (const ChunkRange& cr : chunk_ranges(nc, roots.size(), n_airs)) {
                TraceState::Chunk c;
                SWIRL_TRY(compile_range(cr.r0, cr.r1, cr.zerocheck, &c, false, true));
                s.chunks.push_back(c);
            }
            s.mle_jit = std::make_shared<TraceState::MleJit>();
            ProgramCache::Entry e;
            e.whole[0] = s.whole[0];
            e.whole[1] = s.whole[1];
            e.chunks = s.chunks;
            e.mle_jit = s.mle_jit;
            cache->entries.emplace(std::move(key), std::move(e));
        }
    }

    mark("programs");
    // ---- LogUp input layer + GKR -----------------------------------------------------------------------
    std::vector<Ext> xi;
    if (total_interactions > 0) {
        // the padding (0, alpha) behind the last interaction block is never materialised: the GKR prover
        // takes the stored prefix and the tail constant (gkr.cu)
        size_t used = 0;
        for (const LayoutCol& lc : ilayout.cols) used = std::max<size_t>(used, lc.row_idx + (size_t(1) << lc.log_height));
        const size_t n_leaves = std::min(size_t(1) << L, (used + 3) & ~size_t(3));
        uint32_t* leaves = nullptr;
        SWIRL_CUDA(dev_alloc(ctx, &leaves, n_leaves * 8));
        to_free.push_back(leaves);
        if (used < n_leaves) {
            leaves_fill_kernel<<<(unsigned)((n_leaves - used + 255) / 256), 256, 0, ctx->stream>>>(leaves + used * 8, n_leaves - used, alpha);
            SWIRL_LAUNCH_CHECK(ctx);
        }
        std::vector<uint32_t> lw((max_len + 2) * 4);
        memcpy(&lw[0], bb::ext_one().c, 16);
        for (size_t j = 0; j <= max_len; j++) memcpy(&lw[4 * (j + 1)], beta_pows[j].c, 16);
        uint32_t* d_lw = nullptr;
        SWIRL_TRY(upload(lw.data(), lw.size() * 4, (void**)&d_lw));
        for (size_t t = 0; t < n_airs; t++) {
            const swirl_air_ctx& a = airs[t];
            if (!a.n_interactions) continue;
            TraceState& s = T[t];
            std::vector<Instr> code;
            std::vector<uint32_t> off{0};
            std::vector<uint32_t> dconst;
            std::vector<uint64_t> row_idx(a.n_interactions, 0);
            int ns = 0;
            for (uint64_t i = 0; i < a.n_interactions; i++) {
                const swirl_interaction& it = a.interactions[i];
                std::vector<Root> roots{Root{it.count_node, 1, 0}};
                for (uint32_t j = 0; j < it.msg_len; j++) roots.push_back(Root{a.msg_nodes[it.msg_offset + j], 2, 1 + j});
                Program pr;
                SWIRL_TRY(compile_program(a, s.L, roots, &pr));
                code.insert(code.end(), pr.code.begin(), pr.code.end());
                off.push_back((uint32_t)code.size());
                ns = std::max(ns, pr.n_slots);
                const Ext dc = ext_add(ext_mul_base(beta_pows[it.msg_len], bb::to_mont(it.bus_index + 1)), alpha);
                dconst.insert(dconst.end(), dc.c, dc.c + 4);
                bool found = false;
                for (const LayoutCol& lc : ilayout.cols)
                    if (lc.mat_idx == t && lc.col_in_mat == i) {
                        row_idx[i] = lc.row_idx;
                        found = true;
                    }
                SWIRL_REQUIRE(found, "InteractionsLayoutMissing");
            }
            LeafArgs la{};
            SWIRL_TRY(upload(code.data(), code.size() * sizeof(Instr), (void**)&la.code));
            SWIRL_TRY(upload(off.data(), off.size() * 4, (void**)&la.prog_off));
            SWIRL_TRY(upload(dconst.data(), dconst.size() * 4, (void**)&la.denom_const));
            SWIRL_TRY(upload(row_idx.data(), row_idx.size() * 8, (void**)&la.row_idx));
            la.parts = s.d_parts;
            la.weights = d_lw;
            la.leaves = leaves;
            la.height = (uint32_t)s.height;
            la.reps = (uint32_t)(s.lifted / s.height);
            la.norm = s.norm;
            const dim3 grid((unsigned)((s.height + BC_BLOCK - 1) / BC_BLOCK), (unsigned)a.n_interactions);
#define BC_LEAVES(NS) logup_leaves_kernel<NS><<<grid, BC_BLOCK, 0, ctx->stream>>>(la)
            BC_DISPATCH_NS(ns, BC_LEAVES);
#undef BC_LEAVES
            SWIRL_LAUNCH_CHECK(ctx);
        }
        mark("leaves");
        uint32_t frac_sum[8];
        std::vector<uint32_t> xi_w((size_t)L * 4);
        SWIRL_TRY(swirl_gkr_fractional_sumcheck_padded(ctx, ts, leaves, n_leaves, alpha.c, L, 1, frac_sum, sec_claims, sec_gkr_polys,
                                                       xi_w.data()));
        memcpy(sec_q0, frac_sum + 4, 16);
        for (int i = 0; i < L; i++) xi.push_back(hp::from_words(&xi_w[4 * i]));
    } else {
        memcpy(sec_q0, bb::ext_one().c, 16);
    }
    const int n_global = std::max(n_max, n_logup);
    while ((int)xi.size() != l_skip + n_global) xi.push_back(tr.sample_ext());

    mark("gkr");
    // ---- batching randomness, weights, eq tables --------------------------------------------------------
    const Ext lambda = tr.sample_ext();
    for (size_t t = 0; t < n_airs; t++) {
        TraceState& s = T[t];
        const swirl_air_ctx& a = airs[t];
        // eq(xi_3, b) per interaction (cpu.rs:247-283)
        for (uint64_t i = 0; i < a.n_interactions; i++) {
            uint64_t stacked_idx = 0;
            for (const LayoutCol& lc : ilayout.cols)
                if (lc.mat_idx == t && lc.col_in_mat == i) stacked_idx = lc.row_idx;
            uint64_t b_int = stacked_idx >> (l_skip + s.n_lift);
            Ext e = bb::ext_one();
            for (int v = l_skip + s.n_lift; v < l_skip + n_logup; v++) {
                e = ext_mul(e, hp::eq1(xi[v], (b_int & 1) != 0));
                b_int >>= 1;
            }
            s.eq_3b.push_back(e);
        }
        std::vector<uint32_t> w;
        Ext lp = bb::ext_one();
        for (uint64_t k = 0; k < a.n_constraints; k++) {
            w.insert(w.end(), lp.c, lp.c + 4);
            lp = ext_mul(lp, lambda);
        }
        s.denom_const = bb::ext_zero();
        for (uint64_t i = 0; i < a.n_interactions; i++) {
            const swirl_interaction& it = a.interactions[i];
            w.insert(w.end(), s.eq_3b[i].c, s.eq_3b[i].c + 4);
            for (uint32_t j = 0; j < it.msg_len; j++) {
                const Ext m = ext_mul(s.eq_3b[i], beta_pows[j]);
                w.insert(w.end(), m.c, m.c + 4);
            }
            s.denom_const = ext_add(s.denom_const,
                                    ext_mul(s.eq_3b[i], ext_mul_base(beta_pows[it.msg_len], bb::to_mont(it.bus_index + 1))));
        }
        SWIRL_TRY(upload(w.data(), w.size() * 4, (void**)&s.d_weights));
    }
    // eq(xi[l_skip..], .) tables are shared by all AIRs of one height class
    std::map<int, uint32_t*> eq_tab;
    auto build_eq = [&](int first_var, int n_lift) -> int {  // table over xi[first_var .. l_skip + n_lift)
        TensorArgs ta;
        const int nv = l_skip + n_lift - first_var;
        for (int b = 0; b < nv; b++) {
            memcpy(ta.w0[b], ext_sub(bb::ext_one(), xi[first_var + b]).c, 16);
            memcpy(ta.w1[b], xi[first_var + b].c, 16);
        }
        return mle_tensor_table(ctx, ta, nv, eq_tab[n_lift]);
    };
    for (size_t t = 0; t < n_airs; t++) {
        TraceState& s = T[t];
        if (!eq_tab.count(s.n_lift)) {
            uint32_t* e = nullptr;
            SWIRL_CUDA(dev_alloc(ctx, &e, (size_t(4) << s.n_lift)));
            to_free.push_back(e);
            eq_tab[s.n_lift] = e;
            SWIRL_TRY(build_eq(l_skip, s.n_lift));
        }
        s.d_eq_xi = eq_tab[s.n_lift];
    }

    mark("weights+eq");
    // ---- round 0 ---------------------------------------------------------------------------------------
    const uint32_t g = bb::to_mont(31), omega_skip = bb::two_adic_generator(l_skip);
    std::vector<uint32_t*> d_lde(D + 1, nullptr);  // per constraint degree d: [d * N][N] Lagrange table
    for (int d = 1; d <= D; d++) {
        bool used = false;
        for (size_t t = 0; t < n_airs; t++) used |= (int)airs[t].constraint_degree == d;
        if (!used) continue;
        std::vector<uint32_t> tab((size_t)d * N * N);
        const uint32_t n_inv = bb::inv(bb::to_mont((uint32_t)N));
        for (int c = 0; c < d; c++) {
            uint32_t z = bb::pow(g, (uint64_t)c + 1);
            for (size_t zi = 0; zi < N; zi++) {
                const uint32_t num = bb::mul(bb::sub(bb::pow(z, N), bb::R1), n_inv);
                uint32_t wi = bb::R1;
                for (size_t i = 0; i < N; i++) {
                    tab[((size_t)c * N + zi) * N + i] = bb::mul(bb::mul(num, wi), bb::inv(bb::sub(z, wi)));
                    wi = bb::mul(wi, omega_skip);
                }
                z = bb::mul(z, omega_skip);
            }
        }
        SWIRL_TRY(upload(tab.data(), tab.size() * 4, (void**)&d_lde[d]));
    }
    // round-0 results: one [cd * N][12] block per (AIR, chunk), summed per AIR on the host
    std::vector<size_t> r0_off(n_airs + 1, 0), r0_chunk_off;
    for (size_t t = 0; t < n_airs; t++) r0_off[t + 1] = r0_off[t] + (size_t)airs[t].constraint_degree * N * 12;
    size_t r0_words = 0;
    int max_slots = 1;
    for (size_t t = 0; t < n_airs; t++)
        for (const auto& c : T[t].chunks) max_slots = std::max(max_slots, c.n_slots);
    int max_slots_r0 = max_slots;
    uint64_t r0_alg_bytes = 0;  // every trace part (and the 3 selector columns) read once
    for (size_t t = 0; t < n_airs; t++)
        if (airs[t].constraint_degree) r0_alg_bytes += (uint64_t)T[t].lifted * 3 * 4 + (uint64_t)T[t].height * (T[t].total_cols - 3) / T[t].L.stride * 4;
    uint32_t* d_r0 = nullptr;
    struct R0Block {
        size_t air, off, len;  // the desc's result block [off, off + len) adds into the first `len` words of its AIR's block
    };
    std::vector<R0Block> r0_desc_air;
    {
        std::vector<R0Args> descs;
        std::vector<void*> desc_kernel;  // run-time compiled kernel of the desc's program, or nullptr = interpreter
        std::vector<uint16_t> block_air;
        size_t part_words = 0;
        const bool use_jit = l_skip == 4 && ctx->jit_mode != 0 && jit_available();
        for (size_t t = 0; t < n_airs; t++) {
            TraceState& s = T[t];
            const int cd = (int)airs[t].constraint_degree;
            if (cd == 0) continue;
            // enough hypercube points to fill the machine: one walk of the whole constraint / interaction program per
            // point (the per-point set-up is paid once); short traces take the sub-programs for their parallelism
            std::vector<TraceState::Chunk> whole{s.whole[0], s.whole[1]};
            const bool split = (size_t(1) << s.n_lift) < BC_R0_SPLIT_BELOW && !(use_jit && ctx->jit_mode == 2);
            for (const auto& ch : split ? s.chunks : whole) {
                // zerocheck: the quotient lives on d - 1 cosets; LogUp numerators / denominators need d
                const int cosets = ch.zerocheck_only ? cd - 1 : cd;
                if (cosets == 0 || ch.n_instr == 0) continue;
                void* jk = nullptr;
                if (use_jit && !split && ch.jit && ch.jit->state >= 0) {
                    TraceState::Jit& j = *ch.jit;
                    if (j.state == 0) {  // first tall trace with this program: compile it (seconds, once per context)
                        const auto tj = std::chrono::steady_clock::now();
                        const std::string src = generate_round0_source(j.h_code, j.n_slots, "swirl_r0_jit");
                        const int jrc = src.empty() ? -1 : jit_compile(ctx, src, "swirl_r0_jit", &j.kernel);
                        j.state = jrc == 0 ? 1 : -1;
                        ctx->jit_stats[0] += jrc == 0;
                        if (src.empty()) {
                            if (trace_on) fprintf(stderr, "[swirl jit] program of %zu instructions is too irregular to compile quickly: interpreter\n", j.h_code.size());
                        } else if (trace_on || jrc != 0)
                            fprintf(stderr, "[swirl jit] round-0 kernel for a program of %zu instructions: %s (%.0f ms)%s%s\n", j.h_code.size(),
                                    jrc == 0 ? "compiled" : "FAILED, using the interpreter",
                                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tj).count(),
                                    jrc == 0 ? "" : ": ", jrc == 0 ? "" : swirl_last_error());
                        j.h_code.clear();
                        j.h_code.shrink_to_fit();
                    }
                    if (j.state == 1) jk = j.kernel.kernel;
                }
                if (!jk) max_slots_r0 = std::max(max_slots_r0, ch.n_slots);
                R0Args ra{};
                ra.code = ch.d_code;
                ra.n_instr = ch.n_instr;
                ra.parts = s.d_parts;
                ra.n_parts = s.L.n_parts;
                ra.weights = s.d_weights;
                ra.lde = d_lde[cd];
                ra.eq_xi = s.d_eq_xi;
                ra.l_skip = l_skip;
                ra.n_lift = s.n_lift;
                ra.P = (int)(cosets * N);
                const int G = std::max(1, BC_BLOCK / ra.P);
                const size_t nx = size_t(1) << s.n_lift;
                ra.x_per_block = G * (split && s.chunks.size() > 1 ? 16 : 4);
                ra.n_blocks = (uint32_t)((nx + ra.x_per_block - 1) / ra.x_per_block);
                ra.first_block = (uint32_t)block_air.size();
                ra.partials = (uint32_t*)(uintptr_t)part_words;  // offsets for now, rebased below
                ra.result = (uint32_t*)(uintptr_t)r0_words;
                r0_desc_air.push_back(R0Block{t, r0_words, (size_t)ra.P * 12});
                r0_words += (size_t)ra.P * 12;
                part_words += (size_t)ra.n_blocks * ra.P * 12;
                SWIRL_REQUIRE(descs.size() < 65535, "too many AIRs");
                block_air.insert(block_air.end(), ra.n_blocks, (uint16_t)descs.size());
                descs.push_back(ra);
                desc_kernel.push_back(jk);
            }
        }
        SWIRL_CUDA(dev_alloc(ctx, &d_r0, r0_words + 4));
        to_free.push_back(d_r0);
        if (!descs.empty()) {
            uint32_t* part = nullptr;
            SWIRL_CUDA(dev_alloc(ctx, &part, part_words + 4));
            to_free.push_back(part);
            for (auto& d : descs) {
                d.partials = part + (size_t)(uintptr_t)d.partials;
                d.result = d_r0 + (size_t)(uintptr_t)d.result;
            }
            // one launch per kernel: the interpreter takes every desc without a compiled program, each compiled program
            // takes its own descs (AIRs that share a DAG share the kernel)
            std::map<void*, std::vector<size_t>> groups;
            for (size_t i = 0; i < descs.size(); i++) groups[desc_kernel[i]].push_back(i);
            bool first_group = true;
            for (const auto& grp : groups) {
                std::vector<R0Args> gd;
                std::vector<uint16_t> gba;
                for (size_t i : grp.second) {
                    R0Args d = descs[i];
                    d.first_block = (uint32_t)gba.size();
                    gba.insert(gba.end(), d.n_blocks, (uint16_t)gd.size());
                    gd.push_back(d);
                }
                R0Args* d_descs = nullptr;
                uint16_t* d_ba = nullptr;
                SWIRL_TRY(upload(gd.data(), gd.size() * sizeof(R0Args), (void**)&d_descs));
                SWIRL_TRY(upload(gba.data(), gba.size() * sizeof(uint16_t), (void**)&d_ba));
                const unsigned blocks = (unsigned)gba.size();
#define BC_R0(NS)                                                                                                     \
    do {                                                                                                              \
        const size_t smem = (size_t)BC_BLOCK * (13 + ((NS) <= 64 ? (NS) * 2 : 0)) * 4;                               \
        if (l_skip == 4) {                                                                                            \
            cudaFuncSetAttribute(batch_round0_kernel<NS, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            batch_round0_kernel<NS, 4><<<blocks, BC_BLOCK, smem, ctx->stream>>>(d_descs, d_ba);                      \
        } else {                                                                                                      \
            cudaFuncSetAttribute(batch_round0_kernel<NS, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            batch_round0_kernel<NS, 0><<<blocks, BC_BLOCK, smem, ctx->stream>>>(d_descs, d_ba);                      \
        }                                                                                                             \
    } while (0)
                {
                    // the family's algorithmic bytes are accounted once (with its first launch)
                    SwirlTimed timed(ctx, SWIRL_T_BC_ROUND0, first_group ? r0_alg_bytes : 0);
                    if (grp.first) {
                        void* kargs[2] = {(void*)&d_descs, (void*)&d_ba};
                        SWIRL_CUDA(cudaLaunchKernel((const void*)grp.first, dim3(blocks), dim3(BC_BLOCK), kargs, 0, ctx->stream));
                        ctx->jit_stats[1]++;
                    } else {
                        BC_DISPATCH_NS(max_slots_r0, BC_R0);
                    }
                }
#undef BC_R0
                SWIRL_LAUNCH_CHECK(ctx);
                first_group = false;
                const int max_nv = D * (int)N * 12;
                bc_reduce_multi_kernel<<<dim3((max_nv + 255) / 256, (unsigned)gd.size()), 256, 0, ctx->stream>>>(d_descs);
                SWIRL_LAUNCH_CHECK(ctx);
            }
        }
    }
    std::vector<uint32_t> h_r0c(r0_words + 4), h_r0(r0_off[n_airs] + 4, 0);
    SWIRL_CUDA(cudaMemcpyAsync(h_r0c.data(), d_r0, r0_words * 4, cudaMemcpyDeviceToHost, ctx->stream));
    SWIRL_CUDA(swirl::stream_sync(ctx, __FILE__, __LINE__));
    for (const R0Block& da : r0_desc_air)
        for (size_t i = 0; i < da.len; i++) h_r0[r0_off[da.air] + i] = bb::add(h_r0[r0_off[da.air] + i], h_r0c[da.off + i]);
    mark("round0 device");
    // host: per-trace s'_0 polynomials (cpu.rs:324-424)
    const size_t sp_0_deg = (size_t)D * (N - 1), s_0_deg = (size_t)(D + 1) * (N - 1);
    std::vector<std::vector<Ext>> sp0(3 * n_airs);  // [2t] numer, [2t+1] denom, [2n + t] zerocheck
    for (size_t t = 0; t < n_airs; t++) {
        const TraceState& s = T[t];
        const int cd = (int)airs[t].constraint_degree;
        if (cd == 0) continue;
        const uint32_t* res = &h_r0[r0_off[t]];
        auto at = [&](int c, size_t zi, int k) { return hp::from_words(res + ((size_t)c * N + zi) * 12 + 4 * k); };
        // zerocheck: quotient on cd - 1 cosets, then s'_0 = (Z^N - 1) q
        {
            std::vector<Ext> qe(N * (cd - 1));
            for (int c = 0; c + 1 < cd; c++) {
                const uint32_t zinv = bb::inv(bb::sub(bb::pow(bb::pow(g, (uint64_t)c + 1), N), bb::R1));
                for (size_t zi = 0; zi < N; zi++) qe[zi * (cd - 1) + c] = ext_mul_base(at(c, zi, 0), zinv);
            }
            std::vector<Ext> q = cd > 1 ? hp::interpolate_geometric_cosets(qe, l_skip, cd - 1) : std::vector<Ext>();
            const size_t deg = (size_t)cd * (N - 1);
            std::vector<Ext> coeffs(deg + 1);
            for (size_t i = 0; i <= deg; i++) {
                Ext c = i < q.size() ? bb::ext_neg(q[i]) : bb::ext_zero();
                if (i >= N) c = ext_add(c, q[i - N]);
                coeffs[i] = c;
            }
            sp0[2 * n_airs + t] = coeffs;
        }
        if (airs[t].n_interactions) {
            std::vector<Ext> ne(N * cd), de(N * cd);
            for (int c = 0; c < cd; c++)
                for (size_t zi = 0; zi < N; zi++) {
                    ne[zi * cd + c] = at(c, zi, 1);
                    de[zi * cd + c] = ext_add(at(c, zi, 2), s.denom_const);
                }
            std::vector<Ext> np = hp::interpolate_geometric_cosets(ne, l_skip, cd);
            for (auto& c : np) c = ext_mul_base(c, s.norm);
            sp0[2 * t] = np;
            sp0[2 * t + 1] = hp::interpolate_geometric_cosets(de, l_skip, cd);
        }
    }
    // s_0 = eq_sharp * s'_0 (logup), eq_D(xi_0, .) * batched s'_0 (zerocheck); sum claims (mod.rs:204-301)
    std::vector<Ext> eq_sharp_coeffs;
    {
        std::vector<Ext> ev(N, bb::ext_zero());
        ev[0] = bb::ext_one();
        for (int i = 0; i < l_skip; i++)
            for (size_t j = 0; j < (size_t(1) << i); j++) {
                ev[(size_t(1) << i) + j] = ext_mul(ev[j], xi[i]);
                ev[j] = ext_mul(ev[j], ext_sub(bb::ext_one(), xi[i]));
            }
        eq_sharp_coeffs = hp::idft_small(ev);
    }
    std::vector<std::vector<Ext>> s0_logup(2 * n_airs);
    for (size_t i = 0; i < 2 * n_airs; i++) {
        std::vector<Ext> c = sp0[i];
        if (c.size() > sp_0_deg + 1) c.resize(sp_0_deg + 1);
        s0_logup[i] = poly_mul(eq_sharp_coeffs, c, s_0_deg + 1);
    }
    for (size_t t = 0; t < n_airs; t++) {
        Ext sums[2];
        for (int d = 0; d < 2; d++) {
            Ext sm = bb::ext_zero();
            for (size_t j = 0; j <= s_0_deg; j += N) sm = ext_add(sm, s0_logup[2 * t + d][j]);
            sums[d] = ext_mul_base(sm, bb::to_mont((uint32_t)N));
        }
        tr.observe_ext(sums[0]);
        tr.observe_ext(sums[1]);
        memcpy(sec_numer + 4 * t, sums[0].c, 16);
        memcpy(sec_denom + 4 * t, sums[1].c, 16);
    }
    const Ext mu = tr.sample_ext();
    std::vector<Ext> mu_pows(3 * n_airs);
    {
        Ext m = bb::ext_one();
        for (auto& x : mu_pows) {
            x = m;
            m = ext_mul(m, mu);
        }
    }
    std::vector<Ext> s0_zc;
    {
        std::vector<Ext> sp(sp_0_deg + 1, bb::ext_zero());
        for (size_t j = 0; j <= sp_0_deg; j++)
            for (size_t t = 0; t < n_airs; t++) {
                const auto& poly = sp0[2 * n_airs + t];
                if (j < poly.size()) sp[j] = ext_add(sp[j], ext_mul(mu_pows[2 * n_airs + t], poly[j]));
            }
        // eq_uni_poly(l_skip, xi[0]) (poly_common.rs:85-102)
        std::vector<Ext> eq(N);
        const uint32_t n_inv = hp::half_pow(l_skip);
        Ext xp = xi[0];
        std::vector<Ext> pw(N);
        for (size_t i = 0; i < N; i++) {
            pw[i] = ext_mul_base(xp, n_inv);
            xp = ext_mul(xp, xi[0]);
        }
        for (size_t i = 0; i < N; i++) eq[i] = pw[N - 1 - i];
        eq[0] = hp::from_base(n_inv);
        s0_zc = poly_mul(eq, sp, s_0_deg + 1);
    }
    std::vector<Ext> s_0(s_0_deg + 1);
    for (size_t j = 0; j <= s_0_deg; j++) {
        Ext c = s0_zc[j];
        for (size_t i = 0; i < 2 * n_airs; i++) c = ext_add(c, ext_mul(mu_pows[i], s0_logup[i][j]));
        tr.observe_ext(c);
        s_0[j] = c;
        memcpy(sec_uni + 4 * j, c.c, 16);
    }
    std::vector<Ext> r{tr.sample_ext()};
    const Ext r_0 = r[0];
    Ext prev_s_eval = hp::horner(s_0, r_0);

    mark("round0 host");
    // ---- fold_ple: all row parts of a trace into one EF buffer -------------------------------------------
    {
        LagrangeArgs la;
        const std::vector<Ext> Lc = hp::lagrange_at(l_skip, r_0);
        for (size_t i = 0; i < N; i++) memcpy(la.L[i], Lc[i].c, 16);
        for (size_t t = 0; t < n_airs; t++) {
            TraceState& s = T[t];
            const swirl_air_ctx& a = airs[t];
            s.h = s.lifted >> l_skip;
            SWIRL_CUDA(dev_alloc(ctx, &s.ef[0], (size_t)s.total_cols * s.h * 4));
            SWIRL_CUDA(dev_alloc(ctx, &s.ef[1], ((size_t)s.total_cols * s.h / 2 + 1) * 4));
            to_free.push_back(s.ef[0]);
            to_free.push_back(s.ef[1]);
            size_t part = 0;
            auto fold = [&](const uint32_t* mat, size_t height, size_t width, bool rot) -> int {
                SWIRL_TRY(fold_ple(ctx, mat, height, width, rot, l_skip, la, s.ef[0] + (size_t)s.L.part_col_off[part] * s.h * 4));
                part++;
                return 0;
            };
            SWIRL_TRY(fold(s.d_sels, s.lifted, 3, false));
            auto both = [&](const swirl_matrix& m) -> int {
                SWIRL_TRY(fold(m.data, m.height, m.width, false));
                if (a.need_rot) SWIRL_TRY(fold(m.data, m.height, m.width, true));
                return 0;
            };
            if (a.preprocessed) SWIRL_TRY(both(*a.preprocessed));
            for (uint64_t i = 0; i < a.n_cached; i++) SWIRL_TRY(both(a.cached_mains[i]));
            SWIRL_TRY(both(a.common_main));
            s.cur = 0;
        }
    }
    std::vector<Ext> eq_ns{hp::eval_eq_uni(l_skip, xi[0], r_0)};
    std::vector<Ext> eq_sharp_ns;
    {
        // eval_eq_sharp_uni (poly_common.rs:134-176)
        std::vector<Ext> ev(N, bb::ext_zero());
        ev[0] = bb::ext_one();
        for (int i = 0; i < l_skip; i++)
            for (size_t j = 0; j < (size_t(1) << i); j++) {
                ev[(size_t(1) << i) + j] = ext_mul(ev[j], xi[i]);
                ev[j] = ext_mul(ev[j], ext_sub(bb::ext_one(), xi[i]));
            }
        Ext res = bb::ext_zero();
        uint32_t wi = bb::R1;
        for (size_t i = 0; i < N; i++) {
            res = ext_add(res, ext_mul(hp::eval_eq_uni(l_skip, r_0, hp::from_base(wi)), ev[i]));
            wi = bb::mul(wi, omega_skip);
        }
        eq_sharp_ns.push_back(res);
    }

    mark("fold_ple");
    // ---- MLE rounds (mod.rs:314-395, cpu.rs:463-597) --------------------------------------------------------
    const int s_deg = D + 1;
    // Plan of all rounds first: table sizes halve deterministically, so every round's kernel descriptors are known now.
    // One upload for all of them (per-round copies from pageable memory were 4 blocking transfers per round), and with
    // the round link (ext.cuh) every round's kernels are enqueued before the first result is read: the fold kernel of
    // round k waits for the challenge in the mailbox, the evaluation kernel of round k + 1 follows it in the stream.
    struct MleJitGroup {  // consecutive AIRs that share a run-time compiled kernel: one launch
        void* kernel;
        uint32_t first_block, n_blocks;
    };
    struct MleRound {
        std::vector<int> mode;  // 0: hypercube sum, 1: single row now, 2: tail multiply, 3: nothing to evaluate
        std::vector<size_t> desc_first, desc_count;
        size_t desc_off = 0, n_descs = 0, ba_off = 0, n_blocks = 0;
        size_t fold_off = 0, n_fold = 0, fba_off = 0, n_fold_blocks = 0;
        std::vector<MleJitGroup> jit_groups;
        RoundLink link{};
    };
    // Run-time compiled evaluation kernels (jit.hpp): one kernel per distinct AIR, a switch over its
    // sub-programs with the value slots as D extension-field lanes in registers.  All or nothing per proof: if one AIR's
    // kernel cannot be built (too irregular, too many live values), every AIR takes the interpreter launch.
    bool mle_jit = ctx->jit_mle && ctx->jit_mode != 0 && jit_available() && n_max >= 1;
    if (mle_jit && ctx->jit_mode != 2) {
        bool tall = false;
        for (size_t t = 0; t < n_airs; t++) tall = tall || T[t].log_height >= 17;
        mle_jit = tall;
    }
    for (size_t t = 0; t < n_airs && mle_jit; t++) {
        TraceState& s = T[t];
        if (airs[t].constraint_degree == 0 && !airs[t].n_interactions && !airs[t].n_constraints) continue;  // never evaluated
        if (!s.mle_jit) {
            mle_jit = false;
            break;
        }
        TraceState::MleJit& j = *s.mle_jit;
        if (j.state == 1 && j.D != D) {  // the same AIR under another max_constraint_degree: rebuild
            jit_release(&j.kernel);
            j.state = 0;
        }
        if (j.state == 0) {
            const auto tj = std::chrono::steady_clock::now();
            std::vector<const std::vector<Instr>*> codes;
            std::vector<int> ns;
            bool have = true;
            for (const auto& ch : s.chunks) {
                have = have && ch.h_code;
                codes.push_back(ch.h_code.get());
                ns.push_back(ch.n_slots);
            }
            const std::string src = have ? generate_mle_source(codes, ns, D, "swirl_mle_jit", &j.refs) : std::string();
            const int jrc = src.empty() ? -1 : jit_compile(ctx, src, "swirl_mle_jit", &j.kernel);
            j.state = jrc == 0 ? 1 : -1;
            j.D = D;
            ctx->jit_stats[2] += jrc == 0;
            if (src.empty()) {
                if (trace_on) fprintf(stderr, "[swirl jit] MLE rounds: %zu sub-programs are too irregular or too wide to compile: interpreter\n", s.chunks.size());
            } else if (trace_on || jrc != 0)
                fprintf(stderr, "[swirl jit] MLE-round kernel for %zu sub-programs: %s (%.0f ms)%s%s\n", s.chunks.size(),
                        jrc == 0 ? "compiled" : "FAILED, using the interpreter",
                        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tj).count(), jrc == 0 ? "" : ": ",
                        jrc == 0 ? "" : swirl_last_error());
        }
        if (j.state != 1 || j.refs.size() != s.chunks.size()) mle_jit = false;
    }
    std::vector<MleRound> plan(n_max + 1);
    std::vector<MleArgs> all_descs;
    std::vector<FoldArgs> all_fd;
    std::vector<uint16_t> all_ba, all_fba;
    for (int round = 1; round <= n_max; round++) {
        MleRound& R = plan[round];
        R.mode.assign(n_airs, 0);
        R.desc_first.assign(n_airs, 0);
        R.desc_count.assign(n_airs, 0);
        R.desc_off = all_descs.size();
        R.ba_off = all_ba.size();
        for (size_t t = 0; t < n_airs; t++) {
            TraceState& s = T[t];
            if (airs[t].constraint_degree == 0 && !airs[t].n_interactions && !airs[t].n_constraints) {
                R.mode[t] = 3;
                continue;
            }
            if (round > s.n_lift + 1) {
                R.mode[t] = 2;
                continue;
            }
            R.desc_first[t] = R.n_descs;
            for (size_t ci = 0; ci < s.chunks.size(); ci++) {
                const auto& ch = s.chunks[ci];
                MleArgs ma{};
                if (mle_jit) {
                    const MleCaseRef& cr = s.mle_jit->refs[ci];
                    ma.sub = (uint32_t)cr.case_id;
                    ma.col_shift = cr.col_shift;
                    ma.w_shift = cr.w_shift;
                }
                ma.code = ch.d_code;
                ma.n_instr = ch.n_instr;
                ma.base = s.ef[s.cur];
                ma.h = s.h;
                ma.weights = s.d_weights;
                ma.eq_xi = s.d_eq_xi;
                SWIRL_REQUIRE(R.n_descs < 256, "too many (AIR, program chunk) pairs for the result scratch");
                ma.ticket = rs->d_ticket + R.n_descs;
                ma.result = rs->d_result + R.n_descs * 64;
                if (round == s.n_lift + 1) {
                    R.mode[t] = 1;
                    ma.single = 1;
                    ma.ny = 1;
                    ma.n_blocks = 1;
                } else {
                    const int log_ny = s.n_lift - round;
                    ma.single = 0;
                    ma.ny = size_t(1) << log_ny;
                    ma.n_blocks = (uint32_t)std::min<size_t>((ma.ny + 127) / 128, std::max<size_t>(1, (size_t)ctx->sm_count * 8 / s.chunks.size()));
                }
                ma.first_block = (uint32_t)R.n_blocks;
                ma.partials = rs->d_partials + (size_t)ma.first_block * 64;
                all_ba.insert(all_ba.end(), ma.n_blocks, (uint16_t)R.n_descs);
                R.n_blocks += ma.n_blocks;
                all_descs.push_back(ma);
                R.n_descs++;
            }
            R.desc_count[t] = R.n_descs - R.desc_first[t];
            if (mle_jit && R.desc_count[t]) {
                const uint32_t fb = all_descs[R.desc_off + R.desc_first[t]].first_block, nb = (uint32_t)R.n_blocks - fb;
                void* k = s.mle_jit->kernel.kernel;
                if (!R.jit_groups.empty() && R.jit_groups.back().kernel == k)
                    R.jit_groups.back().n_blocks += nb;
                else
                    R.jit_groups.push_back(MleJitGroup{k, fb, nb});
            }
        }
        SWIRL_REQUIRE(R.n_blocks <= (size_t)rs->max_blocks, "too many blocks for the reduction scratch");
        R.fold_off = all_fd.size();
        R.fba_off = all_fba.size();
        for (size_t t = 0; t < n_airs; t++) {
            TraceState& s = T[t];
            if (s.h <= 1) continue;
            FoldArgs f{s.ef[s.cur], s.ef[s.cur ^ 1], (size_t)s.total_cols * (s.h / 2), (uint32_t)R.n_fold_blocks};
            const size_t nb = (f.n_out + BC_BLOCK - 1) / BC_BLOCK;
            all_fba.insert(all_fba.end(), nb, (uint16_t)R.n_fold);
            R.n_fold_blocks += nb;
            all_fd.push_back(f);
            R.n_fold++;
            s.cur ^= 1;
            s.h >>= 1;
        }
    }
    MleArgs* d_mle_descs = nullptr;
    FoldArgs* d_fold_descs = nullptr;
    uint16_t *d_mle_ba = nullptr, *d_fold_ba = nullptr;
    SWIRL_CUDA(dev_alloc(ctx, &d_mle_descs, std::max<size_t>(all_descs.size(), 1)));
    SWIRL_CUDA(dev_alloc(ctx, &d_fold_descs, std::max<size_t>(all_fd.size(), 1)));
    SWIRL_CUDA(dev_alloc(ctx, &d_mle_ba, std::max<size_t>(all_ba.size(), 1)));
    SWIRL_CUDA(dev_alloc(ctx, &d_fold_ba, std::max<size_t>(all_fba.size(), 1)));
    to_free.push_back(d_mle_descs);
    to_free.push_back(d_fold_descs);
    to_free.push_back(d_mle_ba);
    to_free.push_back(d_fold_ba);
    if (!all_descs.empty()) SWIRL_CUDA(cudaMemcpyAsync(d_mle_descs, all_descs.data(), all_descs.size() * sizeof(MleArgs), cudaMemcpyHostToDevice, ctx->stream));
    if (!all_ba.empty()) SWIRL_CUDA(cudaMemcpyAsync(d_mle_ba, all_ba.data(), all_ba.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream));
    if (!all_fd.empty()) SWIRL_CUDA(cudaMemcpyAsync(d_fold_descs, all_fd.data(), all_fd.size() * sizeof(FoldArgs), cudaMemcpyHostToDevice, ctx->stream));
    if (!all_fba.empty()) SWIRL_CUDA(cudaMemcpyAsync(d_fold_ba, all_fba.data(), all_fba.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream));
    auto launch_eval = [&](int round) -> int {  // eq tables + evaluation kernel of one round
        const MleRound& R = plan[round];
        for (auto& kv : eq_tab)
            if (round <= kv.first) SWIRL_TRY(build_eq(l_skip + round, kv.first));
        if (R.n_descs && mle_jit) {
            SwirlTimed timed(ctx, SWIRL_T_BC_MLE);
            for (const MleJitGroup& g : R.jit_groups) {
                const MleArgs* dd = d_mle_descs + R.desc_off;
                const uint16_t* ba = d_mle_ba + R.ba_off + g.first_block;
                uint32_t block_base = g.first_block, tag = link_result_tag(R.link.seq);
                void* kargs[4] = {(void*)&dd, (void*)&ba, (void*)&block_base, (void*)&tag};
                SWIRL_CUDA(cudaLaunchKernel((const void*)g.kernel, dim3(g.n_blocks), dim3(128), kargs, 0, ctx->stream));
                ctx->jit_stats[3]++;
                SWIRL_LAUNCH_CHECK(ctx);
            }
        } else if (R.n_descs) {
            const int grid = (int)R.n_blocks;
            // single-row AIRs evaluate lane 0 only; D lanes are still the kernel's width
#define BC_MLE(NS) launch_mle<NS>(D, d_mle_descs + R.desc_off, d_mle_ba + R.ba_off, grid, ctx->stream, link_result_tag(R.link.seq))
            {
                SwirlTimed timed(ctx, SWIRL_T_BC_MLE);
                BC_DISPATCH_NS(max_slots, BC_MLE);
            }
#undef BC_MLE
            SWIRL_LAUNCH_CHECK(ctx);
        }
        return 0;
    };
    auto launch_fold = [&](int round, const Ext& r_val) -> int {
        const MleRound& R = plan[round];
        if (R.n_fold) {
            ef_fold_multi_kernel<<<(unsigned)R.n_fold_blocks, BC_BLOCK, 0, ctx->stream>>>(d_fold_descs + R.fold_off, d_fold_ba + R.fba_off, r_val, R.link);
            SWIRL_LAUNCH_CHECK(ctx);
        }
        return 0;
    };
    // linked rounds need a kernel on both ends of the exchange: one that publishes results and one that takes the challenge
    bool linked = ctx->round_link;
    for (int round = 1; round <= n_max; round++) linked = linked && plan[round].n_descs && plan[round].n_descs <= 64;
    LinkAbortGuard link_guard{ctx, rs};  // an early return must release the kernels that still wait for a challenge
    auto launch_linked = [&](int round) -> int {  // one round ahead of the exchange, see the rule in ext.cuh
        plan[round].link = link_make(rs, true);
        SWIRL_TRY(launch_eval(round));
        return launch_fold(round, bb::ext_zero());
    };
    if (linked && n_max >= 1) {
        link_begin(ctx, rs, 0, 64 * 64);
        link_guard.armed = true;
        SWIRL_TRY(launch_linked(1));
    }
    std::vector<uint32_t> round_words(256 * 64);
    for (int round = 1; round <= n_max; round++) {
        const MleRound& R = plan[round];
        const std::vector<int>& mode = R.mode;
        const std::vector<size_t>&desc_first = R.desc_first, &desc_count = R.desc_count;
        const Ext r_prev = r[round - 1];
        const Ext eq_r_acc = eq_ns.back(), eq_sharp_r_acc = eq_sharp_ns.back();
        if (linked) {
            SWIRL_TRY(link_recv(ctx, rs, R.link.seq, 0, D * 12, R.n_descs, 64, round_words.data()));
        } else {
            SWIRL_TRY(launch_eval(round));
            SWIRL_CUDA(swirl::stream_sync(ctx, __FILE__, __LINE__));
            for (size_t i = 0; i < R.n_descs * 64; i++) round_words[i] = rs->h_result[i];
        }
        // sp_evals[2t] numer, [2t+1] denom, [2n+t] zerocheck: D values (head) or 1 value (tail)
        std::vector<std::vector<Ext>> sp(3 * n_airs);
        for (size_t t = 0; t < n_airs; t++) {
            TraceState& s = T[t];
            uint32_t res[64] = {0};  // sum over this AIR's program chunks
            for (size_t k = 0; k < desc_count[t]; k++)
                for (int i = 0; i < D * 12; i++) res[i] = bb::add(res[i], round_words[(desc_first[t] + k) * 64 + i] % bb::P);
            const bool has_int = airs[t].n_interactions != 0;
            if (mode[t] == 3) {
                sp[2 * n_airs + t].assign(D, bb::ext_zero());
                sp[2 * t].assign(D, bb::ext_zero());
                sp[2 * t + 1].assign(D, bb::ext_zero());
            } else if (mode[t] == 0) {
                for (int X = 0; X < D; X++) {
                    sp[2 * n_airs + t].push_back(hp::from_words(res + X * 12));
                    if (has_int) {
                        sp[2 * t].push_back(ext_mul_base(hp::from_words(res + X * 12 + 4), s.norm));
                        sp[2 * t + 1].push_back(ext_add(hp::from_words(res + X * 12 + 8), s.denom_const));
                    } else {
                        sp[2 * t].push_back(bb::ext_zero());
                        sp[2 * t + 1].push_back(bb::ext_zero());
                    }
                }
            } else {
                if (mode[t] == 1) {
                    s.zc_tilde = ext_mul(eq_r_acc, hp::from_words(res));
                    if (has_int) {
                        s.lg_tilde[0] = ext_mul_base(ext_mul(eq_sharp_r_acc, hp::from_words(res + 4)), s.norm);
                        s.lg_tilde[1] = ext_mul(eq_sharp_r_acc, ext_add(hp::from_words(res + 8), s.denom_const));
                    }
                } else {
                    s.zc_tilde = ext_mul(s.zc_tilde, r_prev);
                    s.lg_tilde[0] = ext_mul(s.lg_tilde[0], r_prev);
                    s.lg_tilde[1] = ext_mul(s.lg_tilde[1], r_prev);
                }
                sp[2 * n_airs + t] = {s.zc_tilde};
                if (has_int) {
                    sp[2 * t] = {s.lg_tilde[0]};
                    sp[2 * t + 1] = {s.lg_tilde[1]};
                } else {
                    sp[2 * t].assign(D, bb::ext_zero());
                    sp[2 * t + 1].assign(D, bb::ext_zero());
                }
            }
        }
        size_t tail_start = n_airs;
        for (size_t t = 0; t < n_airs; t++)
            if (round > T[t].n) {
                tail_start = t;
                break;
            }
        std::vector<Ext> head_zc(D, bb::ext_zero()), head_lg(D, bb::ext_zero());
        Ext sp_tail = bb::ext_zero();
        for (size_t t = 0; t < n_airs; t++) {
            const size_t zc = 2 * n_airs + t, nu = 2 * t, de = nu + 1;
            if (t < tail_start) {
                for (int i = 0; i < D; i++) {
                    head_zc[i] = ext_add(head_zc[i], ext_mul(mu_pows[zc], sp[zc][i]));
                    head_lg[i] = ext_add(head_lg[i], ext_add(ext_mul(mu_pows[nu], sp[nu][i]), ext_mul(mu_pows[de], sp[de][i])));
                }
            } else {
                sp_tail = ext_add(sp_tail, ext_add(ext_mul(mu_pows[zc], sp[zc][0]),
                                                   ext_add(ext_mul(mu_pows[nu], sp[nu][0]), ext_mul(mu_pows[de], sp[de][0]))));
            }
        }
        std::vector<Ext> head(s_deg, bb::ext_zero());
        for (int i = 0; i < D; i++) head[i + 1] = ext_add(ext_mul(eq_ns[round - 1], head_zc[i]), ext_mul(eq_sharp_ns[round - 1], head_lg[i]));
        const Ext xi_cur = xi[l_skip + round - 1];
        head[0] = ext_mul(ext_sub(ext_sub(prev_s_eval, ext_mul(xi_cur, head[1])), sp_tail), bb::ext_inv(ext_sub(bb::ext_one(), xi_cur)));
        std::vector<Ext> coeffs = hp::lagrange_interpolate_0n(head);
        coeffs.push_back(bb::ext_zero());
        {
            const Ext b = ext_sub(bb::ext_one(), xi_cur), a = ext_sub(xi_cur, b);
            for (int i = s_deg - 1; i >= 0; i--) coeffs[i + 1] = ext_add(ext_mul(a, coeffs[i]), ext_mul(b, coeffs[i + 1]));
            coeffs[0] = ext_mul(coeffs[0], b);
            coeffs[1] = ext_add(coeffs[1], sp_tail);
        }
        for (int i = 1; i <= s_deg; i++) {
            const Ext e = hp::horner(coeffs, hp::from_base(bb::to_mont((uint32_t)i)));
            tr.observe_ext(e);
            memcpy(sec_rounds + ((size_t)(round - 1) * s_deg + (i - 1)) * 4, e.c, 16);
        }
        const Ext r_round = tr.sample_ext();
        r.push_back(r_round);
        prev_s_eval = hp::horner(coeffs, r_round);
        if (linked) {
            if (R.n_fold) link_send(rs, R.link.seq, r_round);
            if (round < n_max) SWIRL_TRY(launch_linked(round + 1));
        } else {
            SWIRL_TRY(launch_fold(round, r_round));
        }
        const Ext eq_r = hp::eq1(xi_cur, r_round);
        eq_ns.push_back(ext_mul(eq_ns[round - 1], eq_r));
        eq_sharp_ns.push_back(ext_mul(eq_sharp_ns[round - 1], eq_r));
    }

    link_guard.armed = false;  // every linked launch has received its challenge
    if (linked) SWIRL_CUDA(link_flag_fetch(ctx, rs));
    mark("mle rounds");
    // ---- column openings (cpu.rs:644-694), observed common-main first (mod.rs:404-421) ------------------------
    std::vector<std::vector<uint32_t>> rows(n_airs);
    for (size_t t = 0; t < n_airs; t++) {
        TraceState& s = T[t];
        SWIRL_REQUIRE(s.h == 1, "internal: tables not fully folded");
        rows[t].resize((size_t)s.total_cols * 4);
        SWIRL_CUDA(cudaMemcpyAsync(rows[t].data(), s.ef[s.cur], rows[t].size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    SWIRL_CUDA(swirl::stream_sync(ctx, __FILE__, __LINE__));
    if (linked && link_aborted(rs)) {
        set_error("round link: a fold kernel gave up waiting for its challenge");
        return SWIRL_ERR_INVALID;
    }
    std::vector<std::vector<std::vector<Ext>>> openings(n_airs);
    uint32_t* po = sec_open;
    for (size_t t = 0; t < n_airs; t++) {
        const TraceState& s = T[t];
        const uint32_t stride = s.L.stride;
        auto part_open = [&](uint32_t first_part) {
            std::vector<Ext> v;
            const uint32_t w = s.L.part_width[first_part];
            for (uint32_t c = 0; c < w; c++)
                for (uint32_t k = 0; k < stride; k++)
                    v.push_back(hp::from_words(&rows[t][(size_t)(s.L.part_col_off[first_part + k] + c) * 4]));
            return v;
        };
        openings[t].push_back(part_open(s.L.n_parts - stride));  // common main
        for (uint32_t pt = 1; pt + stride < s.L.n_parts; pt += stride) openings[t].push_back(part_open(pt));
        for (auto& part : openings[t])
            for (const Ext& e : part) {
                memcpy(po, e.c, 16);
                po += 4;
            }
    }
    auto observe_part = [&](const std::vector<Ext>& part, bool need_rot) {
        for (const Ext& e : part) {
            tr.observe_ext(e);
            if (!need_rot) tr.observe_ext(bb::ext_zero());
        }
    };
    for (size_t t = 0; t < n_airs; t++) observe_part(openings[t][0], airs[t].need_rot != 0);
    for (size_t t = 0; t < n_airs; t++)
        for (size_t pt = 1; pt < openings[t].size(); pt++) observe_part(openings[t][pt], airs[t].need_rot != 0);
    for (size_t i = 0; i < r.size(); i++) memcpy(h_r + 4 * i, r[i].c, 16);
    mark("openings");
    return 0;
}

// Debugging aid / documentation: the CUDA C++ the library would compile at run time for one AIR's round-0 program
// (which = 0: the constraint roots, 1: the interaction roots).  Host only, no device needed; matrices of `air` need
// only their shapes.  Returns the source length (excluding the terminator); copies at most cap - 1 characters.
extern "C" size_t swirl_jit_round0_source(const swirl_air_ctx* air, int which, char* out, size_t cap) {
    if (!air) return 0;
    AirLayout L;
    if (air_layout(*air, &L) != 0) return 0;
    std::vector<Root> roots;
    if (which == 0) {
        for (uint64_t k = 0; k < air->n_constraints; k++) roots.push_back(Root{air->constraint_idx[k], 0, (uint32_t)k});
    } else {
        uint32_t w = (uint32_t)air->n_constraints;
        for (uint64_t i = 0; i < air->n_interactions; i++) {
            const swirl_interaction& it = air->interactions[i];
            roots.push_back(Root{it.count_node, 1, w++});
            for (uint32_t j = 0; j < it.msg_len; j++) roots.push_back(Root{air->msg_nodes[it.msg_offset + j], 2, w++});
        }
    }
    Program pr;
    if (compile_program(*air, L, roots, &pr, BC_PREFETCH_VARS) != 0) return 0;
    const std::string src = generate_round0_source(pr.code, pr.n_slots, "swirl_r0_jit");
    if (out && cap) {
        const size_t n = std::min(src.size(), cap - 1);
        memcpy(out, src.data(), n);
        out[n] = 0;
    }
    return src.size();
}

// The MLE-round kernel of one AIR (see swirl_b200.h).  Host only.
extern "C" size_t swirl_jit_mle_source(const swirl_air_ctx* air, int max_constraint_degree, size_t n_airs, char* out, size_t cap) {
    if (!air || max_constraint_degree < 1 || max_constraint_degree > 5 || n_airs == 0) return 0;
    AirLayout L;
    if (air_layout(*air, &L) != 0) return 0;
    std::vector<Root> roots;
    for (uint64_t k = 0; k < air->n_constraints; k++) roots.push_back(Root{air->constraint_idx[k], 0, (uint32_t)k});
    uint32_t w = (uint32_t)air->n_constraints;
    for (uint64_t i = 0; i < air->n_interactions; i++) {
        const swirl_interaction& it = air->interactions[i];
        roots.push_back(Root{it.count_node, 1, w++});
        for (uint32_t j = 0; j < it.msg_len; j++) roots.push_back(Root{air->msg_nodes[it.msg_offset + j], 2, w++});
    }
    std::vector<Program> progs;
    for (const ChunkRange& cr : chunk_ranges(air->n_constraints, roots.size(), n_airs)) {
        Program pr;
        if (compile_program(*air, L, std::vector<Root>(roots.begin() + cr.r0, roots.begin() + cr.r1), &pr, BC_PREFETCH_VARS) != 0) return 0;
        progs.push_back(std::move(pr));
    }
    std::vector<const std::vector<Instr>*> codes;
    std::vector<int> ns;
    for (const Program& pr : progs) {
        codes.push_back(&pr.code);
        ns.push_back(pr.n_slots);
    }
    std::vector<MleCaseRef> refs;
    const std::string src = generate_mle_source(codes, ns, max_constraint_degree, "swirl_mle_jit", &refs, true);
    if (out && cap) {
        const size_t n = std::min(src.size(), cap - 1);
        memcpy(out, src.data(), n);
        out[n] = 0;
    }
    return src.size();
}
