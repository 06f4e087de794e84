// Poseidon2 Merkle commitment of a column-major matrix: fused row (leaf) sponge + the strided
// "rows-per-query" tree levels in one kernel, then multi-level adjacent compression kernels.
//
// Replaces (reference, relative to /root/reference):
//   crates/cuda-backend/cuda/src/merkle_tree.cu:16-79     poseidon2_compressing_row_hashes_kernel
//   crates/cuda-backend/cuda/src/merkle_tree.cu:153-177   poseidon2_*_compress_layer_kernel
//   crates/cuda-backend/cuda/src/merkle_tree.cu:181-304   query_digest_layers
//   crates/cuda-backend/src/merkle_tree.rs:140-197        MerkleTreeGpu::new (one launch per layer)
// Semantics = crates/stark-backend/src/prover/stacked_pcs.rs:413-485 (MerkleTree::new).
//
// Layout: a CTA owns `yb` consecutive query indices y and all 2^k rows {y + t*S} of each
// (S = query stride).  Thread (t, y) walks its row across the columns, 8 columns per sponge
// block; neighbouring threads read neighbouring rows of one column => every load instruction is a
// fully coalesced 128-byte line.  The next 8 columns are prefetched (L1/L2, no registers) while the
// current permutation runs.  The 2^k digests of a query are then folded in shared memory (word-major, conflict-free)
// with the active threads packed into whole warps.  The upper tree is built 9 levels per launch.
#include "kernels.cuh"
#include "poseidon2.cuh"

namespace swirl {

constexpr int MK_BLOCK = 256;

__device__ __forceinline__ void store_digest(uint32_t* dst, const uint32_t* s) {
    uint4* d = reinterpret_cast<uint4*>(dst);
    d[0] = make_uint4(s[0], s[1], s[2], s[3]);
    d[1] = make_uint4(s[4], s[5], s[6], s[7]);
}
__device__ __forceinline__ void load_digest(uint32_t* s, const uint32_t* src) {
    const uint4* p = reinterpret_cast<const uint4*>(src);
    uint4 a = __ldg(p), b = __ldg(p + 1);
    s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w;
    s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w;
}

// Sponge over one row of a column-major matrix; leaves the digest in s[0..8).
__device__ __forceinline__ void hash_row_into(uint32_t s[16], const uint32_t* __restrict__ matrix, size_t height,
                                              uint32_t width, size_t row);
__device__ __forceinline__ void hash_row(uint32_t s[16], const uint32_t* __restrict__ matrix, size_t height,
                                         uint32_t width, size_t row) {
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = 0;
    hash_row_into(s, matrix, height, width, row);
}
// Continues the sponge of one row over `width` more columns (the state so far is in s).
__device__ __forceinline__ void hash_row_into(uint32_t s[16], const uint32_t* __restrict__ matrix, size_t height,
                                              uint32_t width, size_t row) {
    const bool live = row < height;  // rows past the matrix hash as zeros
    const uint32_t* p = matrix + (live ? row : 0);
    const uint32_t nfull = width >> 3, tail = width & 7;
    // No register double buffer: the next block's lines are requested into L1/L2 while this permutation runs and the state
    // words are loaded right before they are used.  That keeps the kernel at 48 registers = 5 CTAs per SM, where the row
    // sponge runs at the permutation's isolated rate (4.20 Gperm/s; 4.02 with eight prefetch registers and 4 CTAs/SM:
    // profiles/r2x_leaf_occupancy.jsonl).
    for (uint32_t c = 0; c < nfull; c++) {
        const uint32_t* q = p + (size_t)c * 8 * height;
        if (c + 1 < nfull) {
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("prefetch.global.L1 [%0];" ::"l"(q + (size_t)(8 + i) * height));
        }
#pragma unroll
        for (int i = 0; i < 8; i++) s[i] = live ? __ldg(q + (size_t)i * height) : 0u;
        p2::permute(s);
    }
    if (tail) {
        const uint32_t* q = p + (size_t)nfull * 8 * height;
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (i < (int)tail) s[i] = live ? __ldg(q + (size_t)i * height) : 0u;
        p2::permute(s);
    }
}

// blockDim.x = yb << log_rpq (<= MK_BLOCK); grid.x = S / yb.
__global__ void __launch_bounds__(MK_BLOCK, 5)
leaf_tree_kernel(const uint32_t* __restrict__ matrix, size_t height, uint32_t width, size_t S, int log_rpq,
                 int log_yb, uint32_t* __restrict__ layer0, const uint32_t* __restrict__ state_in,
                 uint32_t* __restrict__ state_out) {
    __shared__ uint32_t sm[8][MK_BLOCK];
    const int tid = threadIdx.x;
    const int yb = 1 << log_yb;
    const int yl = tid & (yb - 1);
    const int t = tid >> log_yb;
    const size_t y = (size_t)blockIdx.x * yb + yl;
    const size_t row = (size_t)t * S + y, n_rows = S << log_rpq;
    uint32_t s[16];
    if (state_in) {  // sponge states of a previous column group, word-major [16][n_rows]
#pragma unroll
        for (int i = 0; i < 16; i++) s[i] = state_in[(size_t)i * n_rows + row];
        hash_row_into(s, matrix, height, width, row);
    } else {
        hash_row(s, matrix, height, width, row);
    }
    if (state_out) {  // more columns follow: park the sponge state
#pragma unroll
        for (int i = 0; i < 16; i++) state_out[(size_t)i * n_rows + row] = s[i];
        return;
    }

    int active = blockDim.x;
    for (int lvl = 0; lvl < log_rpq; lvl++) {
#pragma unroll
        for (int w = 0; w < 8; w++) sm[w][tid] = s[w];  // only tid < active hold live digests
        __syncthreads();
        active >>= 1;
        if (tid < active) {
            // new[x*S + y] = compress(prev[2x*S + y], prev[(2x+1)*S + y])
            const int x = tid >> log_yb;
            const int l = ((2 * x) << log_yb) + yl;
            const int r = l + yb;
#pragma unroll
            for (int w = 0; w < 8; w++) {
                s[w] = sm[w][l];
                s[8 + w] = sm[w][r];
            }
            p2::permute(s);
        }
        __syncthreads();
    }
    if (tid < yb) store_digest(layer0 + y * 8, s);
}

// Row hashes only (fallback when 2^log_rpq does not fit a CTA): out[row] for row < num_leaves.
__global__ void __launch_bounds__(MK_BLOCK)
leaf_only_kernel(const uint32_t* __restrict__ matrix, size_t height, uint32_t width, size_t num_leaves,
                 uint32_t* __restrict__ out) {
    const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= num_leaves) return;
    uint32_t s[16];
    hash_row(s, matrix, height, width, row);
    store_digest(out + row * 8, s);
}

// out[i] = compress(prev[2x*S + y], prev[(2x+1)*S + y]), i = x*S + y  (one strided level)
__global__ void __launch_bounds__(MK_BLOCK)
strided_compress_kernel(const uint32_t* __restrict__ prev, size_t n_out, size_t S, uint32_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    const size_t x = i / S, y = i % S;
    uint32_t s[16];
    load_digest(s, prev + (2 * x * S + y) * 8);
    load_digest(s + 8, prev + ((2 * x + 1) * S + y) * 8);
    p2::permute(s);
    store_digest(out + i * 8, s);
}

// Adjacent-pair compression, `levels` tree levels per launch.  A CTA of B threads consumes 2B
// digests of `prev` and emits B, B/2, ... digests into the consecutive layers that start at `out`
// (layer l+1 directly follows layer l in memory; `n_prev` digests in prev).
__global__ void __launch_bounds__(MK_BLOCK)
compress_layers_kernel(const uint32_t* __restrict__ prev, size_t n_prev, uint32_t* __restrict__ out, int levels) {
    __shared__ uint32_t sm[8][MK_BLOCK];
    const int tid = threadIdx.x;
    uint32_t s[16];
    const size_t in0 = ((size_t)blockIdx.x * blockDim.x + tid) * 2;
    load_digest(s, prev + in0 * 8);
    load_digest(s + 8, prev + (in0 + 1) * 8);
    p2::permute(s);
    size_t n_out = n_prev >> 1;
    uint32_t* lay = out;
    store_digest(lay + ((size_t)blockIdx.x * blockDim.x + tid) * 8, s);
    int active = blockDim.x;
    for (int lvl = 1; lvl < levels; lvl++) {
#pragma unroll
        for (int w = 0; w < 8; w++) sm[w][tid] = s[w];
        __syncthreads();
        lay += n_out * 8;
        n_out >>= 1;
        active >>= 1;
        if (tid < active) {
#pragma unroll
            for (int w = 0; w < 8; w++) {
                s[w] = sm[w][2 * tid];
                s[8 + w] = sm[w][2 * tid + 1];
            }
            p2::permute(s);
            store_digest(lay + ((size_t)blockIdx.x * active + tid) * 8, s);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(MK_BLOCK) permute_states_kernel(uint32_t* __restrict__ states, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint4* p = reinterpret_cast<uint4*>(states + i * 16);
    uint4 v0 = p[0], v1 = p[1], v2 = p[2], v3 = p[3];
    uint32_t s[16] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w,
                      v2.x, v2.y, v2.z, v2.w, v3.x, v3.y, v3.z, v3.w};
    p2::permute(s);
    p[0] = make_uint4(s[0], s[1], s[2], s[3]);
    p[1] = make_uint4(s[4], s[5], s[6], s[7]);
    p[2] = make_uint4(s[8], s[9], s[10], s[11]);
    p[3] = make_uint4(s[12], s[13], s[14], s[15]);
}

__global__ void __launch_bounds__(MK_BLOCK)
compress_pairs_kernel(const uint32_t* __restrict__ pairs, uint32_t* __restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s[16];
    load_digest(s, pairs + i * 16);
    load_digest(s + 8, pairs + i * 16 + 8);
    p2::permute(s);
    store_digest(out + i * 8, s);
}

// out[q][l] = layers[l][(idx_q >> l) ^ 1], l < depth
__global__ void query_proofs_kernel(const uint32_t* __restrict__ layers, size_t S, int depth,
                                    const uint32_t* __restrict__ indices, size_t num_queries,
                                    uint32_t* __restrict__ out) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= num_queries * (size_t)depth) return;
    const size_t q = g / depth;
    const int l = (int)(g % depth);
    // layer l starts at digest offset 2S - (S >> (l-1)) for l >= 1
    const size_t off = l == 0 ? 0 : 2 * S - (S >> (l - 1));
    // indices are query indices < S (S a power of two): masking is the identity for valid input and keeps a bad one in bounds
    const size_t idx = (((size_t)indices[q] & (S - 1)) >> l) ^ 1;
    uint32_t d[8];
    load_digest(d, layers + (off + idx) * 8);
    store_digest(out + g * 8, d);
}

// out[q][t][col0 + c] = matrix[c*height + t*S + idx_q]   (zero past `height`), c < width; rows of `out_width` words
__global__ void open_rows_kernel(const uint32_t* __restrict__ matrix, size_t height, uint32_t width, size_t S,
                                 int log_rpq, const uint32_t* __restrict__ indices, size_t num_queries,
                                 uint32_t* __restrict__ out, uint32_t out_width, uint32_t col0) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t per_q = (size_t)width << log_rpq;
    if (g >= num_queries * per_q) return;
    const size_t q = g / per_q, rem = g % per_q;
    const size_t t = rem / width, c = rem % width;
    // indices are query indices < S (S a power of two): masking is the identity for valid input and keeps a bad one in bounds
    const size_t row = t * S + ((size_t)indices[q] & (S - 1));
    out[((q << log_rpq) + t) * out_width + col0 + c] = row < height ? __ldg(matrix + c * height + row) : 0u;
}

// ------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------
static int compress_upper_layers(swirl_ctx* ctx, uint32_t* d_layers, size_t S) {
    uint32_t* cur = d_layers;
    size_t n = S;
    while (n > 1) {
        const int bd = (int)(n / 2 < (size_t)MK_BLOCK ? n / 2 : (size_t)MK_BLOCK);
        const size_t grid = n / (2 * (size_t)bd);
        const int levels = ilog2(2 * (size_t)bd);
        compress_layers_kernel<<<(unsigned)grid, bd, 0, ctx->stream>>>(cur, n, cur + n * 8, levels);
        SWIRL_LAUNCH_CHECK(ctx);
        size_t off = 0, m = n;
        for (int i = 0; i < levels; i++) {
            off += m;
            m >>= 1;
        }
        cur += off * 8;
        n >>= levels;
    }
    return 0;
}

int merkle_commit(swirl_ctx* ctx, const uint32_t* d_matrix, size_t height, size_t width, int log_rpq,
                  uint32_t* d_layers) {
    return merkle_commit_columns(ctx, d_matrix, height, width, log_rpq, d_layers, nullptr, true, true);
}

// Streaming form: absorbs `width` more columns (a multiple of 8 unless `last`) into the per-row
// sponge states `d_state` ([16][num_leaves] words; ignored when first && last) and, on the last
// group, finishes the tree.
int merkle_commit_columns(swirl_ctx* ctx, const uint32_t* d_matrix, size_t height, size_t width, int log_rpq,
                          uint32_t* d_layers, uint32_t* d_state, bool first, bool last) {
    SWIRL_REQUIRE(height > 0, "MerkleTreeEmptyMatrix");
    SWIRL_REQUIRE(log_rpq >= 0 && log_rpq < 32, "rows_per_query");
    SWIRL_REQUIRE(width < (size_t(1) << 31), "width");
    SWIRL_REQUIRE(((uintptr_t)d_layers & 15) == 0, "digest buffer must be 16-byte aligned");
    size_t num_leaves = 1;
    while (num_leaves < height) num_leaves <<= 1;
    SWIRL_REQUIRE((size_t(1) << log_rpq) <= num_leaves, "MerkleTreeRowsPerQueryExceeded");
    const size_t S = num_leaves >> log_rpq;
    SWIRL_REQUIRE((first && last) || (d_state && (1 << log_rpq) <= MK_BLOCK), "streaming commit needs a state buffer");
    SWIRL_REQUIRE(last || (width & 7) == 0, "column groups must be multiples of the sponge rate");
    if ((1 << log_rpq) <= MK_BLOCK) {
        int log_yb = ilog2(MK_BLOCK) - log_rpq;
        if ((size_t(1) << log_yb) > S) log_yb = ilog2(S);
        const size_t grid = S >> log_yb;
        SWIRL_REQUIRE(grid < (size_t(1) << 31), "grid too large");
        {
            // algorithmic traffic: the matrix once, digest layer 0 once
            SwirlTimed timed(ctx, SWIRL_T_LEAF, (uint64_t)height * width * 4 + (last ? (uint64_t)S * 32 : 0));
            leaf_tree_kernel<<<(unsigned)grid, 1 << (log_yb + log_rpq), 0, ctx->stream>>>(
                d_matrix, height, (uint32_t)width, S, log_rpq, log_yb, d_layers, first ? nullptr : d_state,
                last ? nullptr : d_state);
        }
        SWIRL_LAUNCH_CHECK(ctx);
        if (!last) return 0;
    } else {
        // rows_per_query larger than a CTA: plain row hashes, then one launch per strided level
        uint32_t *a = nullptr, *b = nullptr;
        SWIRL_CUDA(dev_alloc(ctx, &a, num_leaves * 8));
        SWIRL_CUDA(dev_alloc(ctx, &b, num_leaves * 4));
        leaf_only_kernel<<<(unsigned)((num_leaves + MK_BLOCK - 1) / MK_BLOCK), MK_BLOCK, 0, ctx->stream>>>(
            d_matrix, height, (uint32_t)width, num_leaves, a);
        SWIRL_LAUNCH_CHECK(ctx);
        size_t n = num_leaves;
        uint32_t *src = a, *dst = b;
        for (int lvl = 0; lvl < log_rpq; lvl++) {
            n >>= 1;
            uint32_t* o = (lvl == log_rpq - 1) ? d_layers : dst;
            strided_compress_kernel<<<(unsigned)((n + MK_BLOCK - 1) / MK_BLOCK), MK_BLOCK, 0, ctx->stream>>>(
                src, n, S, o);
            SWIRL_LAUNCH_CHECK(ctx);
            uint32_t* tmp = src;
            src = dst;
            dst = tmp;
        }
        dev_free(ctx, a);
        dev_free(ctx, b);
    }
    SwirlTimed timed(ctx, SWIRL_T_TREE);
    return compress_upper_layers(ctx, d_layers, S);
}

int poseidon2_permute_batch(swirl_ctx* ctx, uint32_t* d_states, size_t n) {
    if (n == 0) return 0;
    SWIRL_REQUIRE(((uintptr_t)d_states & 15) == 0, "states must be 16-byte aligned");
    permute_states_kernel<<<(unsigned)((n + MK_BLOCK - 1) / MK_BLOCK), MK_BLOCK, 0, ctx->stream>>>(d_states, n);
    SWIRL_LAUNCH_CHECK(ctx);
    return 0;
}

int poseidon2_compress_batch(swirl_ctx* ctx, const uint32_t* d_pairs, uint32_t* d_out, size_t n) {
    if (n == 0) return 0;
    SWIRL_REQUIRE((((uintptr_t)d_pairs | (uintptr_t)d_out) & 15) == 0, "buffers must be 16-byte aligned");
    compress_pairs_kernel<<<(unsigned)((n + MK_BLOCK - 1) / MK_BLOCK), MK_BLOCK, 0, ctx->stream>>>(d_pairs, d_out, n);
    SWIRL_LAUNCH_CHECK(ctx);
    return 0;
}

int merkle_query_proofs(swirl_ctx* ctx, const uint32_t* d_layers, size_t query_stride, const uint32_t* d_indices,
                        size_t num_queries, uint32_t* d_out) {
    const int depth = ilog2(query_stride);
    const size_t total = num_queries * (size_t)depth;
    if (total == 0) return 0;
    query_proofs_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(d_layers, query_stride, depth,
                                                                                 d_indices, num_queries, d_out);
    SWIRL_LAUNCH_CHECK(ctx);
    return 0;
}

int matrix_open_rows(swirl_ctx* ctx, const uint32_t* d_matrix, size_t height, size_t width, size_t query_stride,
                     int log_rpq, const uint32_t* d_indices, size_t num_queries, uint32_t* d_out) {
    return matrix_open_rows_window(ctx, d_matrix, height, width, query_stride, log_rpq, d_indices, num_queries, d_out, width, 0);
}

// The same for `width` columns that are columns [col0, col0 + width) of a matrix whose opened rows have `out_width` words
// (the codeword recomputed one column group at a time).
int matrix_open_rows_window(swirl_ctx* ctx, const uint32_t* d_matrix, size_t height, size_t width, size_t query_stride,
                            int log_rpq, const uint32_t* d_indices, size_t num_queries, uint32_t* d_out, size_t out_width,
                            size_t col0) {
    const size_t total = num_queries * (width << log_rpq);
    if (total == 0) return 0;
    open_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(
        d_matrix, height, (uint32_t)width, query_stride, log_rpq, d_indices, num_queries, d_out, (uint32_t)out_width, (uint32_t)col0);
    SWIRL_LAUNCH_CHECK(ctx);
    return 0;
}

}  // namespace swirl
