// TraceCommitter::commit for the stacked PCS: layout -> stacked matrix -> Reed–Solomon codeword ->
// Poseidon2 Merkle tree.  Host orchestration in C++ (the reference's is Rust) + the stacking
// kernel.
//
// Replaces (reference, relative to /root/reference):
//   crates/cuda-backend/src/stacked_pcs.rs:50-88      stacked_commit
//   crates/cuda-backend/src/stacked_pcs.rs:108-220    get_stacked_layout / stack_traces_into_expanded
//   crates/cuda-backend/cuda/src/matrix.cu            batch_expand_pad(_wide) for short traces
//   crates/stark-backend/src/prover/stacked_pcs.rs:144-203  StackedLayout::new (host logic)
// Observation used here: because heights are powers of two sorted in descending order, greedy
// stacking never leaves a gap, so the flat column-major stacked matrix is just the concatenation
// of the flat trace buffers (traces shorter than 2^l_skip are first expanded by striding).  A
// single trace that exactly fills its stacked columns is therefore used in place, with no copy.
#include <algorithm>
#include <cstring>
#include <vector>

#include "kernels.cuh"
#include "pcs.cuh"

namespace swirl {

// sorted = (width, log_height), descending log_height
int make_layout(int l_skip, int log_stacked_height, size_t n, const uint64_t* widths,
                       const int32_t* log_heights, Layout* out) {
    out->l_skip = l_skip;
    out->height = uint64_t(1) << log_stacked_height;
    out->cols.clear();
    uint64_t col = 0, row = 0;
    for (size_t m = 0; m < n; m++) {
        if (widths[m] == 0) continue;
        const int lh = log_heights[m];
        if (lh > log_stacked_height) {
            set_error("LayoutHeightExceeded: trace taller than the stacked height");
            return SWIRL_ERR_LAYOUT;
        }
        const uint64_t slen = uint64_t(1) << (lh > l_skip ? lh : l_skip);
        for (uint64_t j = 0; j < widths[m]; j++) {
            if (row + slen > out->height) {
                if (row != out->height) {
                    set_error("LayoutRowOverflow: traces are not sorted by descending height");
                    return SWIRL_ERR_LAYOUT;
                }
                col++;
                row = 0;
            }
            out->cols.push_back({(uint64_t)m, j, col, row, lh});
            row += slen;
        }
    }
    out->width = col + (row != 0 ? 1 : 0);
    return 0;
}

// dst[i << log_stride] = src[i] for a trace shorter than 2^l_skip (dst pre-zeroed);
// one launch per short trace, all its columns at once: element e of the flat trace buffer lands
// at flat offset e << log_stride because every column is expanded by the same factor.
__global__ void expand_strided_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, size_t n,
                                      int log_stride) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i << log_stride] = src[i];
}

}  // namespace swirl

using namespace swirl;

// columns per group when the codeword is streamed instead of cached (a multiple of the sponge rate)
static constexpr uint64_t STREAM_GROUP = 32;

// layout + stacked matrix (the part of the commitment that involves no encoding or hashing)
static int stack_impl(swirl_ctx* ctx, const swirl_pcs_params* params, const swirl_matrix* traces, size_t n, swirl_pcs* pcs) {
    const int l_skip = params->l_skip, n_stack = params->n_stack;
    SWIRL_REQUIRE(l_skip >= 0 && n_stack >= 0 && l_skip + n_stack <= 27, "l_skip / n_stack");
    SWIRL_REQUIRE(params->log_blowup >= 0 && params->k_whir >= 0, "log_blowup / k_whir");
    std::vector<uint64_t> widths(n);
    std::vector<int32_t> lhs(n);
    uint64_t total_cells = 0;
    for (size_t i = 0; i < n; i++) {
        SWIRL_REQUIRE(is_pow2(traces[i].height), "trace height must be a power of two");
        widths[i] = traces[i].width;
        lhs[i] = ilog2(traces[i].height);
        if (i) SWIRL_REQUIRE(lhs[i] <= lhs[i - 1], "traces must be sorted by descending height");
        const uint64_t lifted = traces[i].height > (uint64_t(1) << l_skip) ? traces[i].height : (uint64_t(1) << l_skip);
        total_cells += lifted * traces[i].width;
    }
    pcs->params = *params;
    SWIRL_TRY(make_layout(l_skip, l_skip + n_stack, n, widths.data(), lhs.data(), &pcs->layout));
    const uint64_t H = pcs->layout.height;
    const uint64_t W = (total_cells + H - 1) / H;
    SWIRL_REQUIRE(W == pcs->layout.width, "layout width mismatch");
    SWIRL_REQUIRE(W > 0, "nothing to commit");

    // ---- stacked matrix ----
    size_t nonempty = 0, only = 0;
    for (size_t i = 0; i < n; i++)
        if (traces[i].width) {
            nonempty++;
            only = i;
        }
    if (nonempty == 1 && lhs[only] >= l_skip && total_cells == W * H) {
        pcs->stacked = traces[only].data;  // the trace *is* the stacked matrix
        pcs->owns_stacked = false;
    } else {
        uint32_t* q = nullptr;
        SWIRL_CUDA(dev_alloc(ctx, &q, W * H));
        pcs->stacked = q;
        pcs->owns_stacked = true;
        uint64_t off = 0;  // flat element offset == col*H + row of the next slice
        for (size_t i = 0; i < n; i++) {
            const uint64_t cells = traces[i].height * traces[i].width;
            if (!cells) continue;
            if (lhs[i] >= l_skip) {
                SWIRL_CUDA(cudaMemcpyAsync(q + off, traces[i].data, cells * 4, cudaMemcpyDeviceToDevice, ctx->stream));
                off += cells;
            } else {
                const int ls = l_skip - lhs[i];
                SWIRL_CUDA(cudaMemsetAsync(q + off, 0, (cells << ls) * 4, ctx->stream));
                expand_strided_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, ctx->stream>>>(traces[i].data, q + off,
                                                                                               cells, ls);
                SWIRL_LAUNCH_CHECK(ctx);
                off += cells << ls;
            }
        }
        if (off < W * H) SWIRL_CUDA(cudaMemsetAsync(q + off, 0, (W * H - off) * 4, ctx->stream));
    }
    const uint64_t N = H << params->log_blowup;
    SWIRL_REQUIRE((uint64_t(1) << params->k_whir) <= N, "MerkleTreeRowsPerQueryExceeded");
    pcs->codeword_height = N;
    pcs->query_stride = N >> params->k_whir;
    return 0;
}

static int commit_impl(swirl_ctx* ctx, const swirl_pcs_params* params, const swirl_matrix* traces, size_t n,
                       uint32_t h_root[8], swirl_pcs* pcs) {
    SWIRL_TRY(stack_impl(ctx, params, traces, n, pcs));
    const int l_skip = params->l_skip;
    const uint64_t H = pcs->layout.height, W = pcs->layout.width, N = pcs->codeword_height;
    // ---- codeword + tree ----
    SWIRL_CUDA(dev_alloc(ctx, &pcs->layers, (2 * pcs->query_stride - 1) * 8 + 8));
    // streaming needs whole sponge blocks per group and the fused leaf kernel's per-row state hand-off
    const bool stream_codeword = !ctx->cache_codeword && W > STREAM_GROUP && (uint64_t(1) << params->k_whir) <= 256;
    if (!stream_codeword) {
        SWIRL_CUDA(dev_alloc(ctx, &pcs->codeword, N * W));
        SWIRL_TRY(rs_encode(ctx, pcs->stacked, H, H, W, l_skip, params->log_blowup, pcs->codeword));
        SWIRL_TRY(merkle_commit(ctx, pcs->codeword, N, W, params->k_whir, pcs->layers));
        if (!ctx->cache_codeword) {  // small matrix: encoded in one piece, still not kept
            dev_free(ctx, pcs->codeword);
            pcs->codeword = nullptr;
        }
    } else {
        // cache_rs_code_matrix = false: the codeword exists only one column group at a time; the per-row sponge states wait in
        // HBM between groups (64 B per codeword row) and the last group finishes the tree
        uint32_t *tmp = nullptr, *state = nullptr;
        SWIRL_CUDA(dev_alloc(ctx, &tmp, N * STREAM_GROUP));
        SWIRL_CUDA(dev_alloc(ctx, &state, N * 16));
        int rc = 0;
        for (uint64_t c0 = 0; c0 < W && rc == 0; c0 += STREAM_GROUP) {
            const uint64_t nc = std::min<uint64_t>(STREAM_GROUP, W - c0);
            rc = rs_encode(ctx, pcs->stacked + c0 * H, H, H, nc, l_skip, params->log_blowup, tmp);
            if (rc == 0) rc = merkle_commit_columns(ctx, tmp, N, nc, params->k_whir, pcs->layers, state, c0 == 0, c0 + nc >= W);
        }
        dev_free(ctx, tmp);
        dev_free(ctx, state);
        SWIRL_TRY(rc);
    }
    SWIRL_CUDA(cudaMemcpyAsync(h_root, pcs->layers + (2 * pcs->query_stride - 2) * 8, 32, cudaMemcpyDeviceToHost,
                               ctx->stream));
    SWIRL_CUDA(swirl::stream_sync(ctx, __FILE__, __LINE__));
    return 0;
}

// Opened rows (out[q][t][c], t < 2^log_rpq strided rows per query) of the commitment's codeword.
int swirl::pcs_open_rows(swirl_ctx* ctx, const swirl_pcs* pcs, int log_rpq, const uint32_t* d_indices, size_t num_queries,
                         uint32_t* d_out) {
    const uint64_t H = pcs->layout.height, W = pcs->layout.width, N = pcs->codeword_height;
    if (pcs->codeword)
        return matrix_open_rows(ctx, pcs->codeword, N, W, pcs->query_stride, log_rpq, d_indices, num_queries, d_out);
    // not cached (reference default, cuda-backend/src/device.rs:113-121): encode again, one column group at a time
    SWIRL_REQUIRE(pcs->stacked, "commitment keeps neither its codeword nor its stacked matrix");
    const uint64_t group = std::min<uint64_t>(W, STREAM_GROUP);
    uint32_t* tmp = nullptr;
    SWIRL_CUDA(dev_alloc(ctx, &tmp, N * group));
    int rc = 0;
    for (uint64_t c0 = 0; c0 < W && rc == 0; c0 += group) {
        const uint64_t nc = std::min(group, W - c0);
        rc = rs_encode(ctx, pcs->stacked + c0 * H, H, H, nc, pcs->params.l_skip, pcs->params.log_blowup, tmp);
        if (rc == 0)
            rc = matrix_open_rows_window(ctx, tmp, N, nc, pcs->query_stride, log_rpq, d_indices, num_queries, d_out, W, c0);
    }
    dev_free(ctx, tmp);
    return rc;
}

static void pcs_release(swirl_ctx* ctx, swirl_pcs* pcs) {
    if (!pcs) return;
    if (pcs->owns_stacked) dev_free(ctx, const_cast<uint32_t*>(pcs->stacked));
    dev_free(ctx, pcs->codeword);
    dev_free(ctx, pcs->layers);
    for (uint32_t* p : pcs->owned_traces) dev_free(ctx, p);
    delete pcs;
}

extern "C" {

int swirl_commit(swirl_ctx* ctx, const swirl_pcs_params* params, const swirl_matrix* d_traces, size_t n_traces,
                 uint32_t h_root[8], swirl_pcs** out) {
    SWIRL_REQUIRE(ctx && params && d_traces && h_root && out, "null argument");
    SWIRL_CUDA(cudaSetDevice(ctx->device));
    swirl_pcs* pcs = new swirl_pcs();
    int rc = commit_impl(ctx, params, d_traces, n_traces, h_root, pcs);
    if (rc != 0) {
        pcs_release(ctx, pcs);
        *out = nullptr;
        return rc;
    }
    *out = pcs;
    return 0;
}

// Pipelined transport + commit of a single host trace that is its own stacked matrix: the H2D copy
// of column group g+1 (copy stream) overlaps the RS encoding and sponge absorption of group g
// (compute stream); per-row sponge states wait in HBM between groups.  Returns 1 when the shape
// does not qualify (caller falls back), 0 on success.
static int commit_host_pipelined(swirl_ctx* ctx, const swirl_pcs_params* params, const swirl_matrix* t, uint32_t h_root[8],
                                 swirl_pcs* pcs) {
    const int l_skip = params->l_skip, n_stack = params->n_stack;
    if (!is_pow2(t->height) || t->height != (uint64_t(1) << (l_skip + n_stack))) return 1;
    const uint64_t H = t->height, W = t->width, N = H << params->log_blowup;
    constexpr uint64_t GROUP = 32;
    if (W <= GROUP || H * W * 4 < (uint64_t(32) << 20) || (uint64_t(1) << params->k_whir) > 256 || N < (uint64_t(1) << params->k_whir))
        return 1;
    SWIRL_REQUIRE(params->log_blowup >= 0 && params->k_whir >= 0 && l_skip + n_stack <= 27, "parameters");
    uint64_t widths[1] = {W};
    int32_t lhs[1] = {ilog2(H)};
    pcs->params = *params;
    SWIRL_TRY(make_layout(l_skip, l_skip + n_stack, 1, widths, lhs, &pcs->layout));
    if (!ctx->copy_stream) SWIRL_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    uint32_t *d_trace = nullptr, *state = nullptr;
    SWIRL_CUDA(dev_alloc(ctx, &d_trace, H * W));
    pcs->owned_traces.push_back(d_trace);
    pcs->stacked = d_trace;
    pcs->owns_stacked = false;
    pcs->codeword_height = N;
    pcs->query_stride = N >> params->k_whir;
    SWIRL_CUDA(dev_alloc(ctx, &pcs->codeword, N * W));
    SWIRL_CUDA(dev_alloc(ctx, &pcs->layers, (2 * pcs->query_stride - 1) * 8 + 8));
    SWIRL_CUDA(dev_alloc(ctx, &state, N * 16));
    // the copy stream must not run ahead of the allocation (stream-ordered pool) on the compute stream
    cudaEvent_t ready;
    SWIRL_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    SWIRL_CUDA(cudaEventRecord(ready, ctx->stream));
    SWIRL_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ready, 0));
    const uint64_t ngroups = (W + GROUP - 1) / GROUP;
    std::vector<cudaEvent_t> ev(ngroups);
    int rc = 0;
    for (uint64_t g = 0; g < ngroups && rc == 0; g++) {
        const uint64_t c0 = g * GROUP, nc = std::min(GROUP, W - c0);
        SWIRL_CUDA(cudaEventCreateWithFlags(&ev[g], cudaEventDisableTiming));
        SWIRL_CUDA(cudaMemcpyAsync(d_trace + c0 * H, t->data + c0 * H, nc * H * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
        SWIRL_CUDA(cudaEventRecord(ev[g], ctx->copy_stream));
        SWIRL_CUDA(cudaStreamWaitEvent(ctx->stream, ev[g], 0));
        rc = rs_encode(ctx, d_trace + c0 * H, H, H, nc, l_skip, params->log_blowup, pcs->codeword + c0 * N);
        if (rc == 0)
            rc = merkle_commit_columns(ctx, pcs->codeword + c0 * N, N, nc, params->k_whir, pcs->layers, state, g == 0,
                                       g == ngroups - 1);
    }
    if (rc == 0) {
        SWIRL_CUDA(cudaMemcpyAsync(h_root, pcs->layers + (2 * pcs->query_stride - 2) * 8, 32, cudaMemcpyDeviceToHost, ctx->stream));
        SWIRL_CUDA(swirl::stream_sync(ctx, __FILE__, __LINE__));
    }
    dev_free(ctx, state);
    cudaEventDestroy(ready);
    for (auto e : ev)
        if (e) cudaEventDestroy(e);
    return rc;
}

int swirl_commit_host(swirl_ctx* ctx, const swirl_pcs_params* params, const swirl_matrix* h_traces, size_t n_traces,
                      uint32_t h_root[8], swirl_pcs** out) {
    SWIRL_REQUIRE(ctx && params && h_traces && h_root && out, "null argument");
    SWIRL_CUDA(cudaSetDevice(ctx->device));
    swirl_pcs* pcs = new swirl_pcs();
    if (n_traces == 1 && h_traces[0].width) {
        const int prc = commit_host_pipelined(ctx, params, &h_traces[0], h_root, pcs);
        if (prc == 0) {
            *out = pcs;
            return 0;
        }
        if (prc != 1) {
            pcs_release(ctx, pcs);
            *out = nullptr;
            return prc;
        }
    }
    std::vector<swirl_matrix> dev(n_traces);
    int rc = 0;
    for (size_t i = 0; i < n_traces && rc == 0; i++) {
        dev[i] = h_traces[i];
        const size_t cells = h_traces[i].height * h_traces[i].width;
        uint32_t* p = nullptr;
        if (cells) {
            cudaError_t e = dev_alloc(ctx, &p, cells);
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(p, h_traces[i].data, cells * 4, cudaMemcpyHostToDevice, ctx->stream);
            if (e != cudaSuccess) rc = cuda_fail(e, "H2D trace transport", __FILE__, __LINE__);
            if (p) pcs->owned_traces.push_back(p);
        }
        dev[i].data = p;
    }
    if (rc == 0) rc = commit_impl(ctx, params, dev.data(), n_traces, h_root, pcs);
    if (rc != 0) {
        pcs_release(ctx, pcs);
        *out = nullptr;
        return rc;
    }
    *out = pcs;
    return 0;
}

int swirl_stack(swirl_ctx* ctx, const swirl_pcs_params* params, const swirl_matrix* d_traces, size_t n_traces, swirl_pcs** out) {
    SWIRL_REQUIRE(ctx && params && d_traces && out, "null argument");
    SWIRL_CUDA(cudaSetDevice(ctx->device));
    swirl_pcs* pcs = new swirl_pcs();
    const int rc = stack_impl(ctx, params, d_traces, n_traces, pcs);
    if (rc != 0) {
        pcs_release(ctx, pcs);
        *out = nullptr;
        return rc;
    }
    *out = pcs;
    return 0;
}

int swirl_pcs_attach_external(swirl_pcs* pcs, const uint32_t root[8], swirl_open_fn fn, void* user) {
    SWIRL_REQUIRE(pcs && root && fn, "null argument");
    SWIRL_REQUIRE(!pcs->codeword && !pcs->layers, "the commitment already has a local tree");
    memcpy(pcs->ext_root, root, 32);
    pcs->open_fn = fn;
    pcs->open_user = user;
    return 0;
}

int swirl_pcs_open_rows(swirl_ctx* ctx, const swirl_pcs* pcs, const uint32_t* d_indices, size_t num_queries, uint32_t* d_out) {
    SWIRL_REQUIRE(ctx && pcs && (num_queries == 0 || (d_indices && d_out)), "null argument");
    SWIRL_CUDA(cudaSetDevice(ctx->device));
    return pcs_open_rows(ctx, pcs, pcs->params.k_whir, d_indices, num_queries, d_out);
}

int swirl_pcs_free(swirl_ctx* ctx, swirl_pcs* pcs) {
    SWIRL_REQUIRE(ctx, "null ctx");
    pcs_release(ctx, pcs);
    return 0;
}

uint64_t swirl_pcs_stacked_height(const swirl_pcs* pcs) { return pcs->layout.height; }
uint64_t swirl_pcs_stacked_width(const swirl_pcs* pcs) { return pcs->layout.width; }
uint64_t swirl_pcs_codeword_height(const swirl_pcs* pcs) { return pcs->codeword_height; }
uint64_t swirl_pcs_query_stride(const swirl_pcs* pcs) { return pcs->query_stride; }
const uint32_t* swirl_pcs_stacked_matrix(const swirl_pcs* pcs) { return pcs->stacked; }
const uint32_t* swirl_pcs_codeword(const swirl_pcs* pcs) { return pcs->codeword; }
const uint32_t* swirl_pcs_layers(const swirl_pcs* pcs) { return pcs->layers; }
uint64_t swirl_pcs_layout(const swirl_pcs* pcs, uint64_t* h_out) {
    const auto& c = pcs->layout.cols;
    if (h_out)
        for (size_t i = 0; i < c.size(); i++) {
            h_out[5 * i + 0] = c[i].mat_idx;
            h_out[5 * i + 1] = c[i].col_in_mat;
            h_out[5 * i + 2] = c[i].col_idx;
            h_out[5 * i + 3] = c[i].row_idx;
            h_out[5 * i + 4] = (uint64_t)c[i].log_height;
        }
    return c.size();
}

int swirl_stacked_layout(int l_skip, int log_stacked_height, size_t n_mats, const uint64_t* widths,
                         const int32_t* log_heights, uint64_t* out_width, uint64_t* out_n, uint64_t* out_cols) {
    SWIRL_REQUIRE(out_width && out_n, "null argument");
    SWIRL_REQUIRE(l_skip >= 0 && l_skip <= log_stacked_height && log_stacked_height < 63, "l_skip / height");
    Layout lay;
    SWIRL_TRY(make_layout(l_skip, log_stacked_height, n_mats, widths, log_heights, &lay));
    *out_width = lay.width;
    *out_n = lay.cols.size();
    if (out_cols)
        for (size_t i = 0; i < lay.cols.size(); i++) {
            out_cols[5 * i + 0] = lay.cols[i].mat_idx;
            out_cols[5 * i + 1] = lay.cols[i].col_in_mat;
            out_cols[5 * i + 2] = lay.cols[i].col_idx;
            out_cols[5 * i + 3] = lay.cols[i].row_idx;
            out_cols[5 * i + 4] = (uint64_t)lay.cols[i].log_height;
        }
    return 0;
}

}  // extern "C"


// ---- sharded commitment: the row exchange as ONE kernel over NVLink peer memory ------------------------------------
// After the per-column RS encode, rank `me` holds columns [col0, col0 + cols) of the codeword.  Rank r must hash the rows
// q + t S of its queries q in [r S/G, (r+1) S/G) over ALL columns.  Instead of packing a send buffer and calling a
// collective, every rank stores its elements straight into the peers' shard buffers (symmetric memory mapped over
// NVLink / NVSwitch): dst[r][(col0 + c) * (N/G) + t * S/G + q'] = src[c * N + t * S + r * S/G + q'], 16 bytes per
// thread, source reads coalesced along q' and so are the remote writes.  Reference counterpart: none (the reference is
// single-GPU); layout = MerkleTree::new's query layout, prover/stacked_pcs.rs:413-485.
namespace swirl {
struct PeerPtrs {
    uint32_t* p[16];
};
__global__ void __launch_bounds__(256)
scatter_rows_peer_kernel(const uint32_t* __restrict__ src, size_t N, size_t cols, size_t col0, int log_rpq, int world, PeerPtrs peers) {
    const size_t S = N >> log_rpq, Sg = S / world, shard_rows = N / world;
    const size_t vec_per_seg = Sg >> 2;                              // uint4 per (column, t, destination) segment
    const size_t total = cols * (size_t(1) << log_rpq) * world * vec_per_seg;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t v = i % vec_per_seg;
        size_t rest = i / vec_per_seg;
        const int r = (int)(rest % world);
        rest /= world;
        const size_t t = rest & ((size_t(1) << log_rpq) - 1), c = rest >> log_rpq;
        const uint4 x = __ldg(reinterpret_cast<const uint4*>(src + c * N + t * S + (size_t)r * Sg) + v);
        reinterpret_cast<uint4*>(peers.p[r] + (col0 + c) * shard_rows + t * Sg)[v] = x;
    }
}
}  // namespace swirl

extern "C" int swirl_scatter_rows_to_peers(swirl_ctx* ctx, const uint32_t* d_src, uint64_t rows, uint64_t cols, uint64_t col_offset,
                                           int log_rows_per_query, int world, void* const* peer_bases) {
    SWIRL_REQUIRE(ctx && d_src && peer_bases, "null argument");
    SWIRL_REQUIRE(world >= 1 && world <= 16, "world size must be in [1, 16]");
    SWIRL_REQUIRE(swirl::is_pow2(rows) && log_rows_per_query >= 0 && (rows >> log_rows_per_query) >= (uint64_t)world,
                  "rows must be a power of two with at least one query per rank");
    const uint64_t Sg = (rows >> log_rows_per_query) / world;
    SWIRL_REQUIRE(((rows >> log_rows_per_query) % world) == 0 && (Sg & 3) == 0, "queries per rank must be a multiple of 4");
    SWIRL_CUDA(cudaSetDevice(ctx->device));
    swirl::PeerPtrs pp{};
    for (int r = 0; r < world; r++) {
        SWIRL_REQUIRE(peer_bases[r] && ((uintptr_t)peer_bases[r] & 15) == 0, "peer buffers must be 16-byte aligned");
        pp.p[r] = (uint32_t*)peer_bases[r];
    }
    if (!cols) return 0;
    const size_t total = cols * (size_t(1) << log_rows_per_query) * world * (Sg >> 2);
    const unsigned grid = (unsigned)std::min<size_t>((total + 255) / 256, (size_t)ctx->sm_count * 16);
    swirl::scatter_rows_peer_kernel<<<grid, 256, 0, ctx->stream>>>(d_src, rows, cols, col_offset, log_rows_per_query, world, pp);
    SWIRL_LAUNCH_CHECK(ctx);
    return 0;
}
