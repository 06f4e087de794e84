// Stacked opening reduction: batch sumcheck that reduces every trace-column opening
// (and rotated opening) at r to openings of the stacked columns at u.
//
// Replaces (reference, relative to /root/reference):
//   crates/cuda-backend/src/stacked_reduction.rs:112-188        StackedReductionGpu (UnstackedSlice descriptors)
//   crates/cuda-backend/cuda/src/stacked_reduction.cu           round-0 / fold / MLE-round kernels
// Semantics = crates/stark-backend/src/prover/stacked_reduction.rs:67-506.
//
// Design.  Round 0 (univariate skip): the summand is LINEAR in the column values q, and extending a
// 2^l_skip chunk from D to the cosets g^j D is linear too, so the hypercube sum is taken on the
// raw evaluations first: per height class, three 2^l_skip-vectors
//     V_A1[i] = sum_t lambda_eq,t  sum_x eq(x) q_t[x,i],   V_B1 likewise with lambda_rot,t,
//     V_B2[i] = sum_t lambda_rot,t sum_x (eq(rot^-1 x) - eq(x)) q_t[x,i]
// leave the device after ONE coalesced sweep over the stacked matrix (4 multiplies per cell); the
// host extends these few values to the cosets, applies the Z-only factors and interpolates s_0.
// The reference instead performs an iDFT and 2 coset DFTs per chunk per column.  Later rounds run
// on the EF matrices obtained by evaluating every chunk's interpolant at u_0 (fold_ple), one
// kernel per round over a (view, chunk) work list + a flat pairwise fold.
#include <algorithm>
#include <cstring>
#include <map>
#include <vector>

#include "ext.cuh"
#include "hostpoly.hpp"
#include "kernels.cuh"
#include "pcs.cuh"
#include "transcript.hpp"

namespace swirl {

using bb::ext_add;
using bb::ext_mul;
using bb::ext_mul_base;
using bb::ext_sub;

constexpr int SR_BLOCK = 256;
constexpr int SR_X_PER_BLOCK = 4096;  // round 0: hypercube points per block
constexpr int SR_Y_PER_BLOCK = 2048;  // MLE rounds: hypercube points per block

struct R0View {
    const uint32_t* q;  // the column's slice of the stacked matrix (2^max(log_height, l_skip) words)
    uint32_t lam_eq[4];
    uint32_t lam_rot[4];  // zero when the trace needs no rotation
};

// One block: view blockIdx.y, hypercube points [blockIdx.x * xc, +xc).  Output: partials[block][i*12 + k],
// k = 0..3 V_A1, 4..7 V_B1, 8..11 V_B2 for point i of D.
__global__ void __launch_bounds__(SR_BLOCK)
sr_round0_kernel(const R0View* __restrict__ views, const uint32_t* __restrict__ eq, int n_lift, int l_skip, int xc,
                 uint32_t* __restrict__ partials) {
    __shared__ uint32_t sm[SR_BLOCK][12 + 1];
    const R0View v = views[blockIdx.y];
    const int N = 1 << l_skip;
    const int i = threadIdx.x & (N - 1), g = threadIdx.x >> l_skip, G = SR_BLOCK >> l_skip;
    const size_t nx = size_t(1) << n_lift;
    const size_t x0 = (size_t)blockIdx.x * xc;
    const size_t x1 = min(x0 + (size_t)xc, nx);
    // a1 = sum_x eq[x] q[x], b = sum_x eq[x-1] q[x] (a2 = b - a1): EF x base products summed four at a time in 64 bits
    // (4 p^2 < 2^64) with one Montgomery reduction per coefficient per group instead of one per product
    Ext a1 = bb::ext_zero(), b1 = bb::ext_zero();
    for (size_t xb = x0 + g; xb < x1; xb += 4 * (size_t)G) {
        uint64_t s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const size_t x = xb + (size_t)j * G;
            if (x < x1) {
                const uint32_t q = __ldg(v.q + (x << l_skip) + i);
                const Ext e = ldg_ext(eq + 4 * x);
                const Ext ep = ldg_ext(eq + 4 * (x == 0 ? nx - 1 : x - 1));
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    s1[k] += (uint64_t)e.c[k] * q;
                    s2[k] += (uint64_t)ep.c[k] * q;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            a1.c[k] = bb::add(a1.c[k], bb::reduce_lazy(s1[k]));
            b1.c[k] = bb::add(b1.c[k], bb::reduce_lazy(s2[k]));
        }
    }
    const Ext a2 = ext_sub(b1, a1);
    const Ext le = Ext{{v.lam_eq[0], v.lam_eq[1], v.lam_eq[2], v.lam_eq[3]}};
    const Ext lr = Ext{{v.lam_rot[0], v.lam_rot[1], v.lam_rot[2], v.lam_rot[3]}};
    const Ext o0 = ext_mul(a1, le), o1 = ext_mul(a1, lr), o2 = ext_mul(a2, lr);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        sm[threadIdx.x][k] = o0.c[k];
        sm[threadIdx.x][4 + k] = o1.c[k];
        sm[threadIdx.x][8 + k] = o2.c[k];
    }
    __syncthreads();
    const size_t blk = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
    for (int o = threadIdx.x; o < N * 12; o += SR_BLOCK) {
        const int pi = o / 12, k = o % 12;
        uint32_t s = 0;
        for (int gg = 0; gg < G; gg++) s = bb::add(s, sm[(gg << l_skip) + pi][k]);
        partials[blk * (size_t)(N * 12) + o] = s;
    }
}
// The same sums for 2^l_skip >= 4 with four consecutive points of D per thread: one 16-byte load of the column per
// hypercube point, and every group of threads walks a contiguous run of points so that eq[x - 1] is the value it already
// holds (one eq load per point instead of two per cell).  Products are summed four at a time in 64 bits (bb::dot4 bound).
__global__ void __launch_bounds__(SR_BLOCK)
sr_round0_vec_kernel(const R0View* __restrict__ views, const uint32_t* __restrict__ eq, int n_lift, int l_skip, int xc,
                     uint32_t* __restrict__ partials) {
    __shared__ uint32_t sm[SR_BLOCK][32 + 1];
    const R0View v = views[blockIdx.y];
    const int N = 1 << l_skip, N4 = N >> 2;
    const int i4 = threadIdx.x % N4, g = threadIdx.x / N4, G = SR_BLOCK / N4;
    const size_t nx = size_t(1) << n_lift;
    const size_t x0 = (size_t)blockIdx.x * xc, x1 = min(x0 + (size_t)xc, nx);
    const size_t run = (size_t)xc / G;  // xc is a multiple of 4 G
    const size_t xs = x0 + (size_t)g * run, xe = min(xs + run, x1);
    uint32_t a1[4][4], b1[4][4];  // [cell][coefficient]
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
        for (int k = 0; k < 4; k++) a1[c][k] = b1[c][k] = 0;
    if (xs < xe) {
        Ext ep = ldg_ext(eq + 4 * (xs == 0 ? nx - 1 : xs - 1));
        for (size_t xb = xs; xb < xe; xb += 4) {
            uint64_t s1[4][4], s2[4][4];
#pragma unroll
            for (int c = 0; c < 4; c++)
#pragma unroll
                for (int k = 0; k < 4; k++) s1[c][k] = s2[c][k] = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const size_t x = xb + j;
                if (x < xe) {
                    const uint4 q4 = __ldg(reinterpret_cast<const uint4*>(v.q + (x << l_skip)) + i4);
                    const Ext e = ldg_ext(eq + 4 * x);
                    const uint32_t q[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
                    for (int c = 0; c < 4; c++)
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            s1[c][k] += (uint64_t)e.c[k] * q[c];
                            s2[c][k] += (uint64_t)ep.c[k] * q[c];
                        }
                    ep = e;
                }
            }
#pragma unroll
            for (int c = 0; c < 4; c++)
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    a1[c][k] = bb::add(a1[c][k], bb::reduce_lazy(s1[c][k]));
                    b1[c][k] = bb::add(b1[c][k], bb::reduce_lazy(s2[c][k]));
                }
        }
    }
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
        for (int k = 0; k < 4; k++) {
            sm[threadIdx.x][c * 8 + k] = a1[c][k];
            sm[threadIdx.x][c * 8 + 4 + k] = b1[c][k];
        }
    __syncthreads();
    // point i = 4 i4 + c of D: sum over the G groups, then the three products with the batching coefficients
    if ((int)threadIdx.x < N) {
        const int i = threadIdx.x, ii4 = i >> 2, c = i & 3;
        Ext sa = bb::ext_zero(), sb = bb::ext_zero();
        for (int gg = 0; gg < G; gg++) {
            const uint32_t* row = sm[gg * N4 + ii4] + c * 8;
            sa = ext_add(sa, Ext{{row[0], row[1], row[2], row[3]}});
            sb = ext_add(sb, Ext{{row[4], row[5], row[6], row[7]}});
        }
        const Ext le = Ext{{v.lam_eq[0], v.lam_eq[1], v.lam_eq[2], v.lam_eq[3]}};
        const Ext lr = Ext{{v.lam_rot[0], v.lam_rot[1], v.lam_rot[2], v.lam_rot[3]}};
        const Ext o0 = ext_mul(sa, le), o1 = ext_mul(sa, lr), o2 = ext_mul(ext_sub(sb, sa), lr);
        const size_t blk = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
        uint32_t* out = partials + blk * (size_t)(N * 12) + (size_t)i * 12;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            out[k] = o0.c[k];
            out[4 + k] = o1.c[k];
            out[8 + k] = o2.c[k];
        }
    }
}
// result[o] = sum_b partials[b * nv + o]
__global__ void sr_reduce_kernel(const uint32_t* __restrict__ partials, size_t nblocks, int nv, uint32_t* __restrict__ result) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= nv) return;
    uint32_t s = 0;
    for (size_t b = 0; b < nblocks; b++) s = bb::add(s, partials[b * nv + o]);
    result[o] = s;
}

struct LagArgs {
    uint32_t L[64][4];
};
// out[c][x] = sum_i L_i * q[c][x * 2^l + i]   (fold_ple_evals, sumcheck.rs:204-251)
__global__ void __launch_bounds__(SR_BLOCK)
sr_fold_ple_kernel(const uint32_t* __restrict__ q, size_t H, size_t total /* W * (H >> l) */, int l_skip, LagArgs la,
                   uint32_t* __restrict__ out) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    // out is column-major with height H >> l: flat index g = c * (H >> l) + x  <->  q + g * 2^l
    const uint32_t* p = q + (g << l_skip);
    Ext acc = bb::ext_zero();
    const int N = 1 << l_skip;
    int i = 0;
    for (; i + 4 <= N; i += 4) {  // four terms per Montgomery reduction (bb::dot4)
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(p + i));
#pragma unroll
        for (int k = 0; k < 4; k++)
            acc.c[k] = bb::add(acc.c[k], bb::dot4(la.L[i][k], q.x, la.L[i + 1][k], q.y, la.L[i + 2][k], q.z, la.L[i + 3][k], q.w));
    }
    for (; i < N; i++) {
        const Ext L = Ext{{la.L[i][0], la.L[i][1], la.L[i][2], la.L[i][3]}};
        acc = ext_add(acc, ext_mul_base(L, __ldg(p + i)));
    }
    st_ext(out + 4 * g, acc);
}
// After u_0: k_rot[x] = ind * (eq_uni_rot * eq[x] + c2 * (eq[rot^-1 x] - eq[x])), eq'[x] = (ind * eq_uni) * eq[x]
struct TabArgs {
    uint32_t ind_eq_uni[4], ind_eq_uni_rot[4], ind_c2[4];
};
__global__ void __launch_bounds__(SR_BLOCK)
sr_tables_kernel(const uint32_t* __restrict__ eq, size_t n, TabArgs t, uint32_t* __restrict__ eq_out, uint32_t* __restrict__ k_out) {
    const size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n) return;
    const Ext e = ldg_ext(eq + 4 * x), ep = ldg_ext(eq + 4 * (x == 0 ? n - 1 : x - 1));
    const Ext a = hp::from_words(t.ind_eq_uni), b = hp::from_words(t.ind_eq_uni_rot), c = hp::from_words(t.ind_c2);
    st_ext(k_out + 4 * x, ext_add(ext_mul(b, e), ext_mul(c, ext_sub(ep, e))));
    st_ext(eq_out + 4 * x, ext_mul(a, e));
}
// out[j] = in[2j] + (in[2j+1] - in[2j]) * r over a flat EF array (columns of even height stay aligned)
__global__ void __launch_bounds__(SR_BLOCK)
ef_fold_flat_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, size_t n_out, Ext r, RoundLink link) {
    if (!link_wait(link, r)) return;  // linked rounds: the challenge arrives through the mailbox (ext.cuh)
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_out) return;
    st_ext(out + 4 * j, ext_lerp(ldg_ext(in + 8 * j), ldg_ext(in + 8 * j + 4), r));
}

struct MleView {
    const uint32_t* q;     // EF column segment: 2 << log_ny entries
    const uint32_t* eq;    // table of the view's height class (2 << log_ny entries, or 1 when exhausted)
    const uint32_t* krot;
    uint32_t wA[4], wB[4];  // lambda_eq * eq_ub, lambda_rot * eq_ub (zero if none)
    uint32_t log_ny;
    uint32_t b;  // exhausted views: the bit of row_idx consumed this round
    uint32_t vidx;  // linked rounds: index of the view (its eq_ub factor lives on the device)
};
struct WorkItem {
    uint32_t view, chunk;
};

__global__ void __launch_bounds__(SR_BLOCK)
sr_mle_round_kernel(const MleView* __restrict__ views, const WorkItem* __restrict__ items, size_t y_per_block,
                    uint32_t* partials, unsigned int* ticket, uint32_t* result, uint32_t tag) {
    const WorkItem it = items[blockIdx.x];
    const MleView v = views[it.view];
    const size_t ny = size_t(1) << v.log_ny;
    const size_t y0 = (size_t)it.chunk * y_per_block, y1 = min(y0 + y_per_block, ny);
    Ext e1 = bb::ext_zero(), e2 = bb::ext_zero(), k1 = bb::ext_zero(), k2 = bb::ext_zero();
    for (size_t y = y0 + threadIdx.x; y < y1; y += SR_BLOCK) {
        const Ext q0 = ldg_ext(v.q + 8 * y), q1 = ldg_ext(v.q + 8 * y + 4);
        const Ext a0 = ldg_ext(v.eq + 8 * y), a1 = ldg_ext(v.eq + 8 * y + 4);
        const Ext b0 = ldg_ext(v.krot + 8 * y), b1 = ldg_ext(v.krot + 8 * y + 4);
        // X = 1: (q1, a1, b1);  X = 2: 2*t1 - t0
        const Ext q2 = ext_sub(ext_add(q1, q1), q0), a2 = ext_sub(ext_add(a1, a1), a0), b2 = ext_sub(ext_add(b1, b1), b0);
        e1 = ext_add(e1, ext_mul(q1, a1));
        e2 = ext_add(e2, ext_mul(q2, a2));
        k1 = ext_add(k1, ext_mul(q1, b1));
        k2 = ext_add(k2, ext_mul(q2, b2));
    }
    const Ext wA = hp::from_words(v.wA), wB = hp::from_words(v.wB);
    const Ext s1 = ext_add(ext_mul(wA, e1), ext_mul(wB, k1)), s2 = ext_add(ext_mul(wA, e2), ext_mul(wB, k2));
    uint32_t o[8] = {s1.c[0], s1.c[1], s1.c[2], s1.c[3], s2.c[0], s2.c[1], s2.c[2], s2.c[3]};
    grid_sum<8>(o, partials, ticket, result, tag);
}
// Views whose own variables are used up (round > n_lift): one thread per view, a 2-entry column.
// eq_ub != null (linked rounds): wA / wB hold the lambda powers only and the view's factor prod eq1(u_j, b_j) over the
// rounds since it was used up comes from the device array that sr_eq_ub_kernel maintains.
__global__ void __launch_bounds__(SR_BLOCK)
sr_mle_exhausted_kernel(const MleView* __restrict__ views, size_t n, uint32_t* partials, unsigned int* ticket,
                        uint32_t* result, const uint32_t* __restrict__ eq_ub, uint32_t tag) {
    Ext s1 = bb::ext_zero(), s2 = bb::ext_zero();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const MleView v = views[i];
        const Ext q0 = ldg_ext(v.q), q1 = ldg_ext(v.q + 4);
        const Ext q2 = ext_sub(ext_add(q1, q1), q0);
        Ext w = ext_add(ext_mul(hp::from_words(v.wA), ldg_ext(v.eq)), ext_mul(hp::from_words(v.wB), ldg_ext(v.krot)));
        if (eq_ub) w = ext_mul(w, ld_ext(eq_ub + 4 * v.vidx));
        // eq(X, b): X = 1 -> b;  X = 2 -> b ? 2 : -1
        if (v.b) {
            s1 = ext_add(s1, ext_mul(w, q1));
            s2 = ext_add(s2, ext_mul(w, ext_add(q2, q2)));
        } else {
            s2 = ext_sub(s2, ext_mul(w, q2));
        }
    }
    uint32_t o[8] = {s1.c[0], s1.c[1], s1.c[2], s1.c[3], s2.c[0], s2.c[1], s2.c[2], s2.c[3]};
    grid_sum<8>(o, partials, ticket, result, tag);
}
// Everything that consumes a round's challenge u, in ONE launch (linked rounds: a single kernel may wait for the mailbox
// per challenge, and it must be the last launch of its round, see the rule in ext.cuh): the pairwise folds of the q
// matrices and of the eq / kappa_rot tables (job.in != null: out[j] = lerp(in[2j], in[2j+1], u)), and the update
// eq_ub[i] *= eq1(u, bit (l_skip + round - 1) of row_idx[i]) of the views used up before this round (job.in == null).
struct SrFoldJob {
    const uint32_t* in;
    uint32_t* out;
    size_t n_out;
    uint32_t first_block;
};
__global__ void __launch_bounds__(SR_BLOCK)
sr_fold_multi_kernel(const SrFoldJob* __restrict__ jobs, const uint16_t* __restrict__ block_job, uint32_t* __restrict__ eq_ub,
                     const int* __restrict__ n_lift, const uint32_t* __restrict__ row_idx, int round, int l_skip, RoundLink link) {
    Ext u = bb::ext_zero();
    if (!link_wait(link, u)) return;
    const SrFoldJob job = jobs[block_job[blockIdx.x]];
    const size_t j = (size_t)(blockIdx.x - job.first_block) * blockDim.x + threadIdx.x;
    if (j >= job.n_out) return;
    if (job.in) {
        st_ext(job.out + 4 * j, ext_lerp(ldg_ext(job.in + 8 * j), ldg_ext(job.in + 8 * j + 4), u));
    } else if (round > n_lift[j]) {
        const bool b = (row_idx[j] >> (l_skip + round - 1)) & 1;
        st_ext(eq_ub + 4 * j, ext_mul(ld_ext(eq_ub + 4 * j), b ? u : ext_one_minus(u)));
    }
}

struct HtTab {  // per distinct log_height: eq(., r) and kappa_rot(., r) tables
    int log_height = 0, n = 0, n_lift = 0;
    size_t len = 0;  // current length
    uint32_t* eq[2] = {nullptr, nullptr};
    uint32_t* krot[2] = {nullptr, nullptr};
    int cur = 0;
};

}  // namespace swirl

using namespace swirl;

extern "C" size_t swirl_stacked_reduction_proof_words(const swirl_pcs* const* pcs, size_t n_commits) {
    if (!pcs || !n_commits || !pcs[0]) return 0;
    const int l_skip = pcs[0]->params.l_skip, n_stack = pcs[0]->params.n_stack;
    size_t n = (2 * ((size_t(1) << l_skip) - 1) + 1) * 4 + (size_t)n_stack * 8;
    for (size_t c = 0; c < n_commits; c++) n += pcs[c]->layout.width * 4;
    return n;
}

extern "C" int swirl_stacked_reduction(swirl_ctx* ctx, swirl_transcript* ts, const swirl_pcs* const* pcs, size_t n_commits,
                                       const uint8_t* const* need_rot, const uint32_t* h_r, size_t r_len, uint32_t* h_proof,
                                       size_t proof_words, uint32_t* h_u) {
    SWIRL_REQUIRE(ctx && ts && pcs && n_commits >= 1 && need_rot && h_r && h_proof && h_u, "null argument");
    SWIRL_REQUIRE(proof_words == swirl_stacked_reduction_proof_words(pcs, n_commits), "proof buffer size");
    SWIRL_CUDA(cudaSetDevice(ctx->device));
    const int l_skip = pcs[0]->params.l_skip, n_stack = pcs[0]->params.n_stack;
    SWIRL_REQUIRE(l_skip <= 6, "l_skip > 6 unsupported");
    const size_t N = size_t(1) << l_skip, H = size_t(1) << (l_skip + n_stack), Hq = H >> l_skip;
    Transcript tr(ts);
    RoundScratch* rs;
    SWIRL_TRY(round_scratch_get(ctx, &rs));

    auto t_prev = std::chrono::steady_clock::now();
    swirl::trace_mark(ctx, "sr", nullptr, &t_prev);
    // ---- views ---------------------------------------------------------------------------------
    struct View {
        size_t com, col_idx, row_idx;
        int log_height;
        size_t lam_eq;
        long lam_rot;
    };
    std::vector<View> views;
    size_t lambda_idx = 0;
    int n_max = 0;
    for (size_t ci = 0; ci < n_commits; ci++) {
        SWIRL_REQUIRE(pcs[ci] && pcs[ci]->layout.height == H && pcs[ci]->params.l_skip == l_skip, "commitments must share parameters");
        for (const LayoutCol& lc : pcs[ci]->layout.cols) {
            View v{ci, (size_t)lc.col_idx, (size_t)lc.row_idx, lc.log_height, lambda_idx, -1};
            lambda_idx++;
            if (need_rot[ci][lc.mat_idx]) v.lam_rot = (long)lambda_idx;
            lambda_idx++;
            views.push_back(v);
            n_max = std::max(n_max, lc.log_height - l_skip);
        }
    }
    SWIRL_REQUIRE(r_len >= (size_t)n_max + 1, "r is shorter than 1 + n_max");
    const Ext lambda = tr.sample_ext();
    std::vector<Ext> lambda_pows(lambda_idx);
    {
        Ext a = bb::ext_one();
        for (auto& x : lambda_pows) {
            x = a;
            a = ext_mul(a, lambda);
        }
    }
    std::vector<Ext> r(r_len);
    for (size_t i = 0; i < r_len; i++) r[i] = hp::from_words(h_r + 4 * i);
    const Ext r_0 = r[0];
    const uint32_t omega_skip = bb::two_adic_generator(l_skip);
    const Ext eq_const = hp::eval_eq_uni_at_one(l_skip, ext_mul_base(r_0, omega_skip));

    // windows of equal height (views are height-sorted inside a commit; a new window starts at every change)
    std::vector<size_t> win;
    for (size_t i = 0; i < views.size(); i++)
        if (i == 0 || views[i].log_height != views[i - 1].log_height) win.push_back(i);
    win.push_back(views.size());

    // ---- eq tables per height class -----------------------------------------------------------------
    std::map<int, HtTab> tabs;
    std::vector<void*> to_free;  // released when the call returns, on every path (declared before the link guard below, so a
                                 // pending round is aborted and the stream drained first)
    struct Cleanup {
        swirl_ctx* ctx;
        std::vector<void*>& v;
        ~Cleanup() {
            for (void* p : v) arena_free_block(ctx, p);
        }
    } cleanup_guard{ctx, to_free};
    auto cleanup = []() {};
    for (const View& v : views) {
        if (tabs.count(v.log_height)) continue;
        HtTab t;
        t.log_height = v.log_height;
        t.n = v.log_height - l_skip;
        t.n_lift = std::max(t.n, 0);
        t.len = size_t(1) << t.n_lift;
        for (int k = 0; k < 2; k++) {
            SWIRL_CUDA(dev_alloc(ctx, &t.eq[k], t.len * 4));
            SWIRL_CUDA(dev_alloc(ctx, &t.krot[k], t.len * 4));
            to_free.push_back(t.eq[k]);
            to_free.push_back(t.krot[k]);
        }
        TensorArgs ta;
        for (int b = 0; b < t.n_lift; b++) {
            const Ext w0 = ext_sub(bb::ext_one(), r[1 + b]);
            memcpy(ta.w0[b], w0.c, 16);
            memcpy(ta.w1[b], r[1 + b].c, 16);
        }
        SWIRL_TRY(mle_tensor_table(ctx, ta, t.n_lift, t.eq[0]));
        tabs[v.log_height] = t;
    }

    swirl::trace_mark(ctx, "sr", "views + eq", &t_prev);
    // ---- round 0 -------------------------------------------------------------------------------
    const size_t s0_len = 2 * (N - 1) + 1;
    std::vector<Ext> total_evals(2 * N, bb::ext_zero());  // [z_idx * 2 + coset]
    {
        std::vector<R0View> hv(views.size());
        for (size_t i = 0; i < views.size(); i++) {
            hv[i].q = pcs[views[i].com]->stacked + views[i].col_idx * H + views[i].row_idx;
            memcpy(hv[i].lam_eq, lambda_pows[views[i].lam_eq].c, 16);
            if (views[i].lam_rot >= 0)
                memcpy(hv[i].lam_rot, lambda_pows[views[i].lam_rot].c, 16);
            else
                memset(hv[i].lam_rot, 0, 16);
        }
        R0View* dv = nullptr;
        SWIRL_CUDA(dev_alloc(ctx, &dv, hv.size()));
        to_free.push_back(dv);
        SWIRL_CUDA(cudaMemcpyAsync(dv, hv.data(), hv.size() * sizeof(R0View), cudaMemcpyHostToDevice, ctx->stream));
        const int nv = (int)(N * 12);
        size_t max_blocks = 0;
        for (size_t wi = 0; wi + 1 < win.size(); wi++) {
            const HtTab& t = tabs[views[win[wi]].log_height];
            const size_t chunks = (t.len + SR_X_PER_BLOCK - 1) / SR_X_PER_BLOCK;
            max_blocks = std::max(max_blocks, chunks * (win[wi + 1] - win[wi]));
        }
        uint32_t *part = nullptr, *d_res = nullptr;
        SWIRL_CUDA(dev_alloc(ctx, &part, max_blocks * nv));
        SWIRL_CUDA(dev_alloc(ctx, &d_res, (win.size() - 1) * (size_t)nv));
        to_free.push_back(part);
        to_free.push_back(d_res);
        for (size_t wi = 0; wi + 1 < win.size(); wi++) {
            const HtTab& t = tabs[views[win[wi]].log_height];
            const size_t nviews = win[wi + 1] - win[wi];
            const size_t chunks = (t.len + SR_X_PER_BLOCK - 1) / SR_X_PER_BLOCK;
            SWIRL_REQUIRE(nviews < 65536, "too many columns of one height");
            if (l_skip >= 2 && l_skip <= 6)
                sr_round0_vec_kernel<<<dim3((unsigned)chunks, (unsigned)nviews), SR_BLOCK, 0, ctx->stream>>>(
                    dv + win[wi], t.eq[0], t.n_lift, l_skip, SR_X_PER_BLOCK, part);
            else
                sr_round0_kernel<<<dim3((unsigned)chunks, (unsigned)nviews), SR_BLOCK, 0, ctx->stream>>>(
                    dv + win[wi], t.eq[0], t.n_lift, l_skip, SR_X_PER_BLOCK, part);
            SWIRL_LAUNCH_CHECK(ctx);
            sr_reduce_kernel<<<(nv + 255) / 256, 256, 0, ctx->stream>>>(part, chunks * nviews, nv, d_res + wi * (size_t)nv);
            SWIRL_LAUNCH_CHECK(ctx);
        }
        std::vector<uint32_t> h_res((win.size() - 1) * (size_t)nv);
        SWIRL_CUDA(cudaMemcpyAsync(h_res.data(), d_res, h_res.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
        SWIRL_CUDA(swirl::stream_sync(ctx, __FILE__, __LINE__));
        // host: extend the three vectors of every window to the cosets g D, g^2 D and apply the Z-only factors
        const uint32_t g = bb::to_mont(31);
        for (size_t wi = 0; wi + 1 < win.size(); wi++) {
            const HtTab& t = tabs[views[win[wi]].log_height];
            std::vector<Ext> V[3];
            for (int k = 0; k < 3; k++) {
                V[k].resize(N);
                for (size_t i = 0; i < N; i++) V[k][i] = hp::from_words(&h_res[wi * (size_t)nv + i * 12 + 4 * k]);
                V[k] = hp::idft_small(V[k]);
            }
            int l = l_skip;
            uint32_t omega = omega_skip;
            Ext r_uni = r_0;
            if (t.n < 0) {
                l = l_skip + t.n;
                for (int i = 0; i < -t.n; i++) omega = bb::sqr(omega);
                r_uni = hp::exp_pow2(r_0, -t.n);
            }
            for (int coset = 0; coset < 2; coset++) {
                const uint32_t shift = bb::pow(g, (uint64_t)coset + 1);
                const std::vector<Ext> A1 = hp::coset_dft_small(V[0], shift), B1 = hp::coset_dft_small(V[1], shift),
                                       B2 = hp::coset_dft_small(V[2], shift);
                uint32_t zb = shift;
                for (size_t zi = 0; zi < N; zi++) {
                    const Ext z = hp::from_base(zb);
                    const Ext ind = hp::eval_in_uni(l_skip, t.n, z);
                    const Ext eq_uni_r0 = hp::eval_eq_uni(l, z, r_uni);
                    const Ext eq_uni_r0_rot = hp::eval_eq_uni(l, z, ext_mul_base(r_uni, omega));
                    const Ext eq_uni_1 = hp::eval_eq_uni_at_one(l_skip, z);
                    const Ext acc0 = ext_mul(ext_mul(ind, eq_uni_r0), A1[zi]);
                    const Ext acc1 = ext_mul(ind, ext_add(ext_mul(eq_uni_r0_rot, B1[zi]), ext_mul(ext_mul(eq_const, eq_uni_1), B2[zi])));
                    total_evals[zi * 2 + coset] = ext_add(total_evals[zi * 2 + coset], ext_add(acc0, acc1));
                    zb = bb::mul(zb, omega_skip);
                }
            }
        }
    }
    std::vector<Ext> s_0 = hp::interpolate_geometric_cosets(total_evals, l_skip, 2);
    s_0.resize(s0_len);
    uint32_t* p_out = h_proof;
    for (const Ext& c : s_0) {
        tr.observe_ext(c);
        memcpy(p_out, c.c, 16);
        p_out += 4;
    }
    std::vector<Ext> u_vec{tr.sample_ext()};
    const Ext u_0 = u_vec[0];

    swirl::trace_mark(ctx, "sr", "round 0", &t_prev);
    // ---- fold_ple: q_evals[ci] = EF matrix (H >> l_skip) x W, ping-pong --------------------------------
    std::vector<uint32_t*> qe[2];
    qe[0].resize(n_commits);
    qe[1].resize(n_commits);
    {
        LagArgs la;
        const std::vector<Ext> L = hp::lagrange_at(l_skip, u_0);
        for (size_t i = 0; i < N; i++) memcpy(la.L[i], L[i].c, 16);
        for (size_t ci = 0; ci < n_commits; ci++) {
            const size_t W = pcs[ci]->layout.width, total = W * Hq;
            SWIRL_CUDA(dev_alloc(ctx, &qe[0][ci], total * 4));
            SWIRL_CUDA(dev_alloc(ctx, &qe[1][ci], (total / 2 + 1) * 4));
            to_free.push_back(qe[0][ci]);
            to_free.push_back(qe[1][ci]);
            sr_fold_ple_kernel<<<(unsigned)((total + SR_BLOCK - 1) / SR_BLOCK), SR_BLOCK, 0, ctx->stream>>>(pcs[ci]->stacked, H, total,
                                                                                                        l_skip, la, qe[0][ci]);
            SWIRL_LAUNCH_CHECK(ctx);
        }
    }
    {
        const Ext eq_uni_u0r0 = hp::eval_eq_uni(l_skip, u_0, r_0);
        const Ext eq_uni_u0r0_rot = hp::eval_eq_uni(l_skip, u_0, ext_mul_base(r_0, omega_skip));
        const Ext eq_uni_u01 = hp::eval_eq_uni_at_one(l_skip, u_0);
        for (auto& kv : tabs) {
            HtTab& t = kv.second;
            const Ext ind = hp::eval_in_uni(l_skip, t.n, u_0);
            Ext eq_uni = eq_uni_u0r0, eq_uni_rot = eq_uni_u0r0_rot;
            if (t.n < 0) {
                uint32_t omega = omega_skip;
                for (int i = 0; i < -t.n; i++) omega = bb::sqr(omega);
                const Ext rr = hp::exp_pow2(r_0, -t.n);
                eq_uni = hp::eval_eq_uni(l_skip + t.n, u_0, rr);
                eq_uni_rot = hp::eval_eq_uni(l_skip + t.n, u_0, ext_mul_base(rr, omega));
            }
            TabArgs ta;
            memcpy(ta.ind_eq_uni, ext_mul(ind, eq_uni).c, 16);
            memcpy(ta.ind_eq_uni_rot, ext_mul(ind, eq_uni_rot).c, 16);
            memcpy(ta.ind_c2, ext_mul(ind, ext_mul(eq_const, eq_uni_u01)).c, 16);
            sr_tables_kernel<<<(unsigned)((t.len + SR_BLOCK - 1) / SR_BLOCK), SR_BLOCK, 0, ctx->stream>>>(t.eq[0], t.len, ta, t.eq[1],
                                                                                                      t.krot[1]);
            SWIRL_LAUNCH_CHECK(ctx);
            t.cur = 1;
        }
    }

    swirl::trace_mark(ctx, "sr", "fold_ple", &t_prev);
    // ---- MLE rounds ------------------------------------------------------------------------------
    std::vector<Ext> eq_ub(views.size(), bb::ext_one());
    MleView *d_views = nullptr, *d_ex = nullptr;
    WorkItem* d_items = nullptr;
    SWIRL_CUDA(dev_alloc(ctx, &d_views, views.size()));
    SWIRL_CUDA(dev_alloc(ctx, &d_ex, views.size()));
    to_free.push_back(d_views);
    to_free.push_back(d_ex);
    // hypercube points per block: SR_Y_PER_BLOCK, doubled until the work items of the first (largest) round fit the
    // reduction scratch (BASELINE configs[3]: 512 views of 2^19 points each)
    size_t y_per_block = SR_Y_PER_BLOCK, max_items = 0;
    for (;; y_per_block *= 2) {
        max_items = 0;
        for (const View& v : views) {
            const int n_lift = std::max(v.log_height - l_skip, 0);
            const size_t ny = n_lift >= 1 ? size_t(1) << (n_lift - 1) : 1;
            max_items += (ny + y_per_block - 1) / y_per_block;
        }
        if (max_items <= (size_t)rs->max_blocks) break;
        SWIRL_REQUIRE(y_per_block < (size_t(1) << 30), "too many views for the reduction scratch");
    }
    SWIRL_CUDA(dev_alloc(ctx, &d_items, max_items + 1));
    to_free.push_back(d_items);
    int qcur = 0;
    size_t qh = Hq;  // current height of the q_evals matrices
    std::vector<MleView> hviews, hex;
    std::vector<WorkItem> items;
    const bool linked = ctx->round_link && views.size() < (size_t(1) << 31);
    if (linked) {
        // ---- linked rounds (ext.cuh): every round's descriptors are planned now (the weights that depend on earlier
        // challenges, eq_ub, live on the device), uploaded once, and the rounds are enqueued one ahead of the exchange
        struct SrRound {
            size_t v_off = 0, n_views = 0, x_off = 0, n_ex = 0, i_off = 0, n_items = 0;
            size_t j_off = 0, n_jobs = 0, b_off = 0, n_blocks = 0;  // the round's fold jobs and their block map
        };
        std::vector<SrRound> plan(n_stack + 1);
        std::vector<MleView> all_views, all_ex;
        std::vector<WorkItem> all_items;
        std::vector<SrFoldJob> all_jobs;
        std::vector<uint16_t> all_block_job;
        {
            int sim_qcur = 0;
            size_t sim_qh = Hq;
            std::map<int, std::pair<int, size_t>> sim;  // log_height -> (cur, len)
            for (auto& kv : tabs) sim[kv.first] = {kv.second.cur, kv.second.len};
            for (int round = 1; round <= n_stack; round++) {
                SrRound& R = plan[round];
                R.v_off = all_views.size();
                R.x_off = all_ex.size();
                R.i_off = all_items.size();
                for (size_t i = 0; i < views.size(); i++) {
                    const View& v = views[i];
                    const HtTab& t = tabs[v.log_height];
                    const int cur = sim[v.log_height].first;
                    const int n_lift = t.n_lift;
                    const int hd = std::max(n_lift - round, 0);
                    MleView mv;
                    const size_t row_start = round <= n_lift ? (v.row_idx >> v.log_height) << (hd + 1) : (v.row_idx >> (l_skip + round)) << 1;
                    mv.q = qe[sim_qcur][v.com] + (v.col_idx * sim_qh + row_start) * 4;
                    mv.eq = t.eq[cur];
                    mv.krot = t.krot[cur];
                    memcpy(mv.wA, lambda_pows[v.lam_eq].c, 16);
                    if (v.lam_rot >= 0)
                        memcpy(mv.wB, lambda_pows[v.lam_rot].c, 16);
                    else
                        memset(mv.wB, 0, 16);
                    mv.log_ny = (uint32_t)hd;
                    mv.b = 0;
                    mv.vidx = (uint32_t)i;
                    if (round > n_lift) {
                        mv.b = (uint32_t)((v.row_idx >> (l_skip + round - 1)) & 1);
                        all_ex.push_back(mv);
                        R.n_ex++;
                    } else {
                        const size_t ny = size_t(1) << hd;
                        for (size_t c = 0; c < (ny + y_per_block - 1) / y_per_block; c++) all_items.push_back(WorkItem{(uint32_t)R.n_views, (uint32_t)c});
                        all_views.push_back(mv);
                        R.n_views++;
                    }
                }
                R.n_items = all_items.size() - R.i_off;
                // the consumers of this round's challenge
                R.j_off = all_jobs.size();
                R.b_off = all_block_job.size();
                auto add_job = [&](const uint32_t* in, uint32_t* out, size_t n_out) {
                    const size_t nb = (n_out + SR_BLOCK - 1) / SR_BLOCK;
                    all_jobs.push_back(SrFoldJob{in, out, n_out, (uint32_t)R.n_blocks});
                    all_block_job.insert(all_block_job.end(), nb, (uint16_t)R.n_jobs);
                    R.n_blocks += nb;
                    R.n_jobs++;
                };
                if (sim_qh > 1) {
                    for (size_t ci = 0; ci < n_commits; ci++) add_job(qe[sim_qcur][ci], qe[sim_qcur ^ 1][ci], pcs[ci]->layout.width * (sim_qh / 2));
                    sim_qcur ^= 1;
                    sim_qh >>= 1;
                }
                for (auto& kv : sim)
                    if (kv.second.second > 1) {
                        const HtTab& t = tabs[kv.first];
                        const int cur = kv.second.first;
                        add_job(t.eq[cur], t.eq[cur ^ 1], kv.second.second / 2);
                        add_job(t.krot[cur], t.krot[cur ^ 1], kv.second.second / 2);
                        kv.second.first ^= 1;
                        kv.second.second /= 2;
                    }
                if (R.n_ex && round < n_stack) add_job(nullptr, nullptr, views.size());
            }
        }
        MleView *d_all_views = nullptr, *d_all_ex = nullptr;
        WorkItem* d_all_items = nullptr;
        SrFoldJob* d_all_jobs = nullptr;
        uint16_t* d_all_block_job = nullptr;
        SWIRL_REQUIRE(n_commits + 2 * tabs.size() + 1 < 65536, "too many fold jobs");
        SWIRL_CUDA(dev_alloc(ctx, &d_all_jobs, std::max<size_t>(all_jobs.size(), 1)));
        SWIRL_CUDA(dev_alloc(ctx, &d_all_block_job, std::max<size_t>(all_block_job.size(), 1)));
        to_free.push_back(d_all_jobs);
        to_free.push_back(d_all_block_job);
        uint32_t *d_eq_ub = nullptr, *d_row_idx = nullptr;
        int* d_n_lift = nullptr;
        SWIRL_CUDA(dev_alloc(ctx, &d_all_views, std::max<size_t>(all_views.size(), 1)));
        SWIRL_CUDA(dev_alloc(ctx, &d_all_ex, std::max<size_t>(all_ex.size(), 1)));
        SWIRL_CUDA(dev_alloc(ctx, &d_all_items, std::max<size_t>(all_items.size(), 1)));
        SWIRL_CUDA(dev_alloc(ctx, &d_eq_ub, views.size() * 4));
        SWIRL_CUDA(dev_alloc(ctx, &d_row_idx, views.size()));
        SWIRL_CUDA(dev_alloc(ctx, &d_n_lift, views.size()));
        to_free.push_back(d_all_views);
        to_free.push_back(d_all_ex);
        to_free.push_back(d_all_items);
        to_free.push_back(d_eq_ub);
        to_free.push_back(d_row_idx);
        to_free.push_back(d_n_lift);
        {
            std::vector<uint32_t> ones(views.size() * 4, 0), rows(views.size());
            std::vector<int> lifts(views.size());
            for (size_t i = 0; i < views.size(); i++) {
                ones[4 * i] = bb::R1;
                SWIRL_REQUIRE(views[i].row_idx < (size_t(1) << 32), "row index");
                rows[i] = (uint32_t)views[i].row_idx;
                lifts[i] = tabs[views[i].log_height].n_lift;
            }
            SWIRL_CUDA(cudaMemcpyAsync(d_eq_ub, ones.data(), ones.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
            SWIRL_CUDA(cudaMemcpyAsync(d_row_idx, rows.data(), rows.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
            SWIRL_CUDA(cudaMemcpyAsync(d_n_lift, lifts.data(), lifts.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
            if (!all_views.empty())
                SWIRL_CUDA(cudaMemcpyAsync(d_all_views, all_views.data(), all_views.size() * sizeof(MleView), cudaMemcpyHostToDevice, ctx->stream));
            if (!all_ex.empty())
                SWIRL_CUDA(cudaMemcpyAsync(d_all_ex, all_ex.data(), all_ex.size() * sizeof(MleView), cudaMemcpyHostToDevice, ctx->stream));
            if (!all_items.empty())
                SWIRL_CUDA(cudaMemcpyAsync(d_all_items, all_items.data(), all_items.size() * sizeof(WorkItem), cudaMemcpyHostToDevice, ctx->stream));
            if (!all_jobs.empty()) {
                SWIRL_CUDA(cudaMemcpyAsync(d_all_jobs, all_jobs.data(), all_jobs.size() * sizeof(SrFoldJob), cudaMemcpyHostToDevice, ctx->stream));
                SWIRL_CUDA(cudaMemcpyAsync(d_all_block_job, all_block_job.data(), all_block_job.size() * sizeof(uint16_t), cudaMemcpyHostToDevice,
                                           ctx->stream));
            }
        }
        std::vector<uint32_t> seqs(n_stack + 2, 0);
        std::vector<char> waits(n_stack + 2, 0);
        // evaluation kernels of a round, then the kernels that consume its challenge: the first of them takes it from the
        // mailbox, the others from its relay
        auto launch_round = [&](int round) -> int {
            const SrRound& R = plan[round];
            const RoundLink pub = link_make(rs, true);
            seqs[round] = pub.seq;
            const uint32_t tag = link_result_tag(pub.seq);
            if (R.n_items) {
                SWIRL_REQUIRE(R.n_items <= (size_t)rs->max_blocks, "too many work items for the reduction scratch");
                sr_mle_round_kernel<<<(unsigned)R.n_items, SR_BLOCK, 0, ctx->stream>>>(d_all_views + R.v_off, d_all_items + R.i_off, y_per_block,
                                                                                    rs->d_partials, rs->d_ticket, rs->d_result, tag);
                SWIRL_LAUNCH_CHECK(ctx);
            }
            if (R.n_ex) {
                const unsigned grid = (unsigned)std::min<size_t>((R.n_ex + SR_BLOCK - 1) / SR_BLOCK, 1024);
                sr_mle_exhausted_kernel<<<grid, SR_BLOCK, 0, ctx->stream>>>(d_all_ex + R.x_off, R.n_ex, rs->d_partials, rs->d_ticket,
                                                                         rs->d_result + 8, d_eq_ub, tag);
                SWIRL_LAUNCH_CHECK(ctx);
            }
            // the one kernel that waits for this round's challenge: last launch of the round
            if (R.n_jobs) {
                sr_fold_multi_kernel<<<(unsigned)R.n_blocks, SR_BLOCK, 0, ctx->stream>>>(d_all_jobs + R.j_off, d_all_block_job + R.b_off, d_eq_ub,
                                                                                      d_n_lift, d_row_idx, round, l_skip, pub);
                SWIRL_LAUNCH_CHECK(ctx);
                waits[round] = 1;
            }
            // host mirror of the state the plan simulated
            if (qh > 1) {
                qcur ^= 1;
                qh >>= 1;
            }
            for (auto& kv : tabs) {
                HtTab& t = kv.second;
                if (t.len <= 1) continue;
                t.cur ^= 1;
                t.len /= 2;
            }
            return 0;
        };
        LinkAbortGuard guard{ctx, rs};
        link_begin(ctx, rs, 0, 16);
        guard.armed = true;
        SWIRL_TRY(launch_round(1));
        for (int round = 1; round <= n_stack; round++) {
            const SrRound& R = plan[round];
            uint32_t w16[16] = {0};
            if (R.n_items && R.n_ex)
                SWIRL_TRY(link_recv(ctx, rs, seqs[round], 0, 8, 2, 8, w16));
            else if (R.n_items)
                SWIRL_TRY(link_recv(ctx, rs, seqs[round], 0, 8, w16));
            else
                SWIRL_TRY(link_recv(ctx, rs, seqs[round], 8, 8, w16 + 8));
            const Ext s1 = ext_add(hp::from_words(w16), hp::from_words(w16 + 8));
            const Ext s2 = ext_add(hp::from_words(w16 + 4), hp::from_words(w16 + 12));
            tr.observe_ext(s1);
            tr.observe_ext(s2);
            memcpy(p_out, s1.c, 16);
            memcpy(p_out + 4, s2.c, 16);
            p_out += 8;
            const Ext u_round = tr.sample_ext();
            u_vec.push_back(u_round);
            // the two result groups are not written by every round (no work items late, no exhausted views early)
            link_expect(rs, 0, 16, seqs[round] + 1);
            if (waits[round]) link_send(rs, seqs[round], u_round);
            if (round < n_stack) SWIRL_TRY(launch_round(round + 1));
        }
        guard.armed = false;
    } else
    for (int round = 1; round <= n_stack; round++) {
        hviews.clear();
        hex.clear();
        items.clear();
        for (size_t i = 0; i < views.size(); i++) {
            const View& v = views[i];
            const HtTab& t = tabs[v.log_height];
            const int n_lift = t.n_lift;
            const int hd = std::max(n_lift - round, 0);
            MleView mv;
            const size_t row_start = round <= n_lift ? (v.row_idx >> v.log_height) << (hd + 1) : (v.row_idx >> (l_skip + round)) << 1;
            mv.q = qe[qcur][v.com] + (v.col_idx * qh + row_start) * 4;
            mv.eq = t.eq[t.cur];
            mv.krot = t.krot[t.cur];
            memcpy(mv.wA, ext_mul(lambda_pows[v.lam_eq], eq_ub[i]).c, 16);
            if (v.lam_rot >= 0)
                memcpy(mv.wB, ext_mul(lambda_pows[v.lam_rot], eq_ub[i]).c, 16);
            else
                memset(mv.wB, 0, 16);
            mv.log_ny = (uint32_t)hd;
            mv.b = 0;
            if (round > n_lift) {
                mv.b = (uint32_t)((v.row_idx >> (l_skip + round - 1)) & 1);
                hex.push_back(mv);
            } else {
                const size_t ny = size_t(1) << hd;
                for (size_t c = 0; c < (ny + y_per_block - 1) / y_per_block; c++)
                    items.push_back(WorkItem{(uint32_t)hviews.size(), (uint32_t)c});
                hviews.push_back(mv);
            }
        }
        Ext s1 = bb::ext_zero(), s2 = bb::ext_zero();
        memset(rs->h_result, 0, 64);
        if (!items.empty()) {
            SWIRL_CUDA(cudaMemcpyAsync(d_views, hviews.data(), hviews.size() * sizeof(MleView), cudaMemcpyHostToDevice, ctx->stream));
            SWIRL_CUDA(cudaMemcpyAsync(d_items, items.data(), items.size() * sizeof(WorkItem), cudaMemcpyHostToDevice, ctx->stream));
            sr_mle_round_kernel<<<(unsigned)items.size(), SR_BLOCK, 0, ctx->stream>>>(d_views, d_items, y_per_block, rs->d_partials,
                                                                                   rs->d_ticket, rs->d_result, 0u);
            SWIRL_LAUNCH_CHECK(ctx);
        }
        if (!hex.empty()) {
            SWIRL_CUDA(cudaMemcpyAsync(d_ex, hex.data(), hex.size() * sizeof(MleView), cudaMemcpyHostToDevice, ctx->stream));
            const unsigned grid = (unsigned)std::min<size_t>((hex.size() + SR_BLOCK - 1) / SR_BLOCK, 1024);
            sr_mle_exhausted_kernel<<<grid, SR_BLOCK, 0, ctx->stream>>>(d_ex, hex.size(), rs->d_partials, rs->d_ticket, rs->d_result + 8,
                                                                         nullptr, 0u);
            SWIRL_LAUNCH_CHECK(ctx);
        }
        SWIRL_CUDA(swirl::stream_sync(ctx, __FILE__, __LINE__));
        if (!items.empty()) {
            s1 = ext_add(s1, hp::from_words(rs->h_result));
            s2 = ext_add(s2, hp::from_words(rs->h_result + 4));
        }
        if (!hex.empty()) {
            s1 = ext_add(s1, hp::from_words(rs->h_result + 8));
            s2 = ext_add(s2, hp::from_words(rs->h_result + 12));
        }
        tr.observe_ext(s1);
        tr.observe_ext(s2);
        memcpy(p_out, s1.c, 16);
        memcpy(p_out + 4, s2.c, 16);
        p_out += 8;
        const Ext u_round = tr.sample_ext();
        u_vec.push_back(u_round);
        // fold q_evals, eq and kappa_rot tables
        if (qh > 1) {
            for (size_t ci = 0; ci < n_commits; ci++) {
                const size_t n_out = pcs[ci]->layout.width * (qh / 2);
                ef_fold_flat_kernel<<<(unsigned)((n_out + SR_BLOCK - 1) / SR_BLOCK), SR_BLOCK, 0, ctx->stream>>>(
                    qe[qcur][ci], qe[qcur ^ 1][ci], n_out, u_round, RoundLink{});
                SWIRL_LAUNCH_CHECK(ctx);
            }
            qcur ^= 1;
            qh >>= 1;
        }
        for (auto& kv : tabs) {
            HtTab& t = kv.second;
            if (t.len <= 1) continue;
            const size_t n_out = t.len / 2;
            ef_fold_flat_kernel<<<(unsigned)((n_out + SR_BLOCK - 1) / SR_BLOCK), SR_BLOCK, 0, ctx->stream>>>(t.eq[t.cur], t.eq[t.cur ^ 1],
                                                                                                         n_out, u_round, RoundLink{});
            ef_fold_flat_kernel<<<(unsigned)((n_out + SR_BLOCK - 1) / SR_BLOCK), SR_BLOCK, 0, ctx->stream>>>(
                t.krot[t.cur], t.krot[t.cur ^ 1], n_out, u_round, RoundLink{});
            ctx->launches++;
            SWIRL_LAUNCH_CHECK(ctx);
            t.cur ^= 1;
            t.len = n_out;
        }
        for (size_t i = 0; i < views.size(); i++) {
            const int n_lift = std::max(views[i].log_height - l_skip, 0);
            if (round > n_lift) {
                const bool b = (views[i].row_idx >> (l_skip + round - 1)) & 1;
                eq_ub[i] = ext_mul(eq_ub[i], hp::eq1(u_round, b));
            }
        }
    }
    swirl::trace_mark(ctx, "sr", "mle rounds", &t_prev);
    // ---- stacking openings ------------------------------------------------------------------------
    if (linked) SWIRL_CUDA(link_flag_fetch(ctx, rs));
    for (size_t ci = 0; ci < n_commits; ci++) {
        const size_t W = pcs[ci]->layout.width;
        SWIRL_REQUIRE(qh == 1, "internal: q_evals not fully folded");
        SWIRL_CUDA(cudaMemcpyAsync(p_out, qe[qcur][ci], W * 16, cudaMemcpyDeviceToHost, ctx->stream));
        SWIRL_CUDA(swirl::stream_sync(ctx, __FILE__, __LINE__));
        if (linked && link_aborted(rs)) {
            set_error("round link: a fold kernel gave up waiting for its challenge");
            cleanup();
            return SWIRL_ERR_INVALID;
        }
        for (size_t j = 0; j < W; j++) tr.observe_ext(hp::from_words(p_out + 4 * j));
        p_out += 4 * W;
    }
    for (size_t i = 0; i < u_vec.size(); i++) memcpy(h_u + 4 * i, u_vec[i].c, 16);
    cleanup();
    return 0;
}
