// Internal (non-ABI) entry points shared between the translation units of libswirl_b200.
#pragma once
#include "bb31.cuh"
#include "common.cuh"

namespace swirl {

// ---- merkle.cu ---------------------------------------------------------------------------
// Digest layers of the tree over `matrix` (column-major, `height` x `width`, rows >= height up to
// the next power of two hash as zero rows).  `d_layers` receives all layers concatenated: layer 0
// (query_stride = num_leaves >> log_rpq digests) first, the root last; 2*query_stride-1 digests.
int merkle_commit(swirl_ctx* ctx, const uint32_t* d_matrix, size_t height, size_t width, int log_rpq,
                  uint32_t* d_layers);
int merkle_commit_columns(swirl_ctx* ctx, const uint32_t* d_matrix, size_t height, size_t width, int log_rpq,
                          uint32_t* d_layers, uint32_t* d_state, bool first, bool last);
int poseidon2_permute_batch(swirl_ctx* ctx, uint32_t* d_states, size_t n);
int poseidon2_compress_batch(swirl_ctx* ctx, const uint32_t* d_pairs, uint32_t* d_out, size_t n);
int merkle_query_proofs(swirl_ctx* ctx, const uint32_t* d_layers, size_t query_stride,
                        const uint32_t* d_indices, size_t num_queries, uint32_t* d_out);
int matrix_open_rows(swirl_ctx* ctx, const uint32_t* d_matrix, size_t height, size_t width,
                     size_t query_stride, int log_rpq, const uint32_t* d_indices, size_t num_queries,
                     uint32_t* d_out);

int matrix_open_rows_window(swirl_ctx* ctx, const uint32_t* d_matrix, size_t height, size_t width, size_t query_stride,
                            int log_rpq, const uint32_t* d_indices, size_t num_queries, uint32_t* d_out, size_t out_width,
                            size_t col0);
// Opened rows of a commitment's codeword: from the cached codeword, or (cache_rs_code_matrix = false) by re-encoding the
// stacked matrix one column group at a time (commit.cu).
int pcs_open_rows(swirl_ctx* ctx, const struct ::swirl_pcs* pcs, int log_rpq, const uint32_t* d_indices, size_t num_queries,
                  uint32_t* d_out);

// ---- sponge.cu ---------------------------------------------------------------------------
void round_scratch_free(swirl_ctx* ctx);

// ---- ntt.cu ------------------------------------------------------------------------------
int ntt_init_twiddles(swirl_ctx* ctx);
// Natural-order (i)DFT of `cols` contiguous columns of length 2^log_n, in place.
int ntt_batch(swirl_ctx* ctx, uint32_t* d_data, int log_n, size_t cols, bool inverse);
// Reed–Solomon encode: every column of the H x W column-major `d_in` (column stride in_stride)
// -> column of length H << log_blowup in `d_out` (column stride H << log_blowup).
int rs_encode(swirl_ctx* ctx, const uint32_t* d_in, size_t in_stride, size_t H, size_t W, int l_skip,
              int log_blowup, uint32_t* d_out);

// 2^l_skip chunk iDFT + subset-zeta of `cols` columns of height H (poly.rs:325-348), src -> dst.
int chunk_coeffs(swirl_ctx* ctx, const uint32_t* src, size_t src_stride, uint32_t* dst, size_t dst_stride, size_t H,
                 size_t cols, int l_skip);

// ---- mle.cu ------------------------------------------------------------------------------
struct TensorArgs {  // per variable b: the factor for bit b clear (w0) / set (w1), EF Montgomery words
    uint32_t w0[28][4];
    uint32_t w1[28][4];
};
int mle_tensor_table(swirl_ctx* ctx, const TensorArgs& t, int n_vars, uint32_t* d_out);
int mle_zeta(swirl_ctx* ctx, uint32_t* d_data, size_t col_stride, int log_n, size_t cols, bool inverse);
struct LagrangeArgs {  // Lagrange coefficients of D = <w_(2^l_skip)> at the folding point
    uint32_t L[64][4];
};
int ef_fold_flat(swirl_ctx* ctx, const uint32_t* in, uint32_t* out, size_t n_out, const bb::Ext& r);
int fold_ple(swirl_ctx* ctx, const uint32_t* mat, size_t height, size_t width, bool is_rot, int l_skip, const LagrangeArgs& la,
             uint32_t* out);
int ext_aos_to_soa(swirl_ctx* ctx, const uint32_t* aos, uint32_t* soa, size_t n, size_t col_stride);
int ext_soa_to_aos(swirl_ctx* ctx, const uint32_t* soa, uint32_t* aos, size_t n, size_t col_stride);

// ---- stacking.cu -------------------------------------------------------------------------
struct StackCopy {  // one unstacked column -> its slot in the stacked matrix
    const uint32_t* src;  // device pointer to the column (height = 1 << log_height)
    uint64_t dst_offset;  // element offset in the stacked matrix (col * H + row)
    uint32_t log_height;
    uint32_t log_stride;  // 0 unless log_height < l_skip
};

}  // namespace swirl
