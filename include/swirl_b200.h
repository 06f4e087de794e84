/*
 * swirl_b200 — C ABI of the B200-native SWIRL prover backend (libswirl_b200.so).
 *
 * This is the drop-in boundary: the symbols a Rust `openvm-b200-backend` crate binds with
 * `extern "C"` to implement the reference's plugin traits
 *     ProverBackend / ProverDevice = TraceCommitter + MultiRapProver + OpeningProver
 *     (reference: crates/stark-backend/src/prover/hal.rs:23-138)
 * in place of crates/cuda-backend/src/cuda/ (one .rs per .cu) (the reference's FFI to its own kernels).
 * INTEGRATION.md shows the Rust-side binding for every entry point.
 *
 * Conventions (same as the reference FFI, crates/cuda-backend/src/cuda/ntt.rs:12-21 and
 * cuda-common/src/error.rs:53-60):
 *   - every function returns int: 0 = ok, a cudaError_t value (1..999) for CUDA failures, or a
 *     SWIRL_ERR_* code (>= 10000); swirl_last_error() has the text for the calling thread;
 *   - plain pointers and sizes only; no C++ / torch types;
 *   - field words are BabyBear in Montgomery form (x * 2^32 mod p), exactly the host
 *     `Vec<BabyBear>` bytes the reference memcpy's (cuda-backend/src/data_transporter.rs:93-106);
 *     EF = 4 words (basis 1,X,X^2,X^3), digest = 8 words; matrices are column-major
 *     `values[col*height + row]` (crates/stark-backend/src/prover/matrix.rs:43-47);
 *   - "d_" parameters are device pointers, "h_" parameters host pointers;
 *   - all work is enqueued on the context's stream; functions that return host data synchronise
 *     that stream before returning, the others are asynchronous (as in the reference).
 */
#ifndef SWIRL_B200_H
#define SWIRL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWIRL_ERR_INVALID 10001     /* bad argument (cf. cudaErrorInvalidValue in merkle_tree.cu:220-222) */
#define SWIRL_ERR_LAYOUT 10002      /* StackedPcsError::Layout* (prover/stacked_pcs.rs:160-174) */
#define SWIRL_ERR_UNSUPPORTED 10003
#define SWIRL_ERR_NO_DEVICE 10004
#define SWIRL_ERR_NONZERO_ROOT_SUM 10005 /* LogupZerocheckError::NonZeroRootSum (fractional_sumcheck_gkr.rs:88-91) */
#define SWIRL_ERR_POW 10006              /* no proof-of-work witness exists in the field */

typedef struct swirl_ctx swirl_ctx; /* one per (device, stream); reference: GpuDeviceCtx, cuda-common/src/stream.rs:132-151 */
typedef struct swirl_pcs swirl_pcs; /* reference: StackedPcsDataGpu, cuda-backend/src/stacked_pcs.rs:30-46 */

/* ---- context ------------------------------------------------------------------------------ */

/* Creates a context on CUDA device `device` with its own non-blocking stream and the NTT twiddle
 * tables (reference: GpuDevice::new, cuda-backend/src/device.rs:52-110; twiddle init ntt.rs:20-50). */
int swirl_ctx_create(int device, swirl_ctx** out);
/* Same, but enqueue on a caller-owned cudaStream_t (passed as void*). */
int swirl_ctx_create_on_stream(int device, void* cuda_stream, swirl_ctx** out);
int swirl_ctx_destroy(swirl_ctx* ctx);
int swirl_ctx_synchronize(swirl_ctx* ctx);
void* swirl_ctx_stream(swirl_ctx* ctx);            /* the cudaStream_t */
uint64_t swirl_ctx_launch_count(swirl_ctx* ctx);   /* kernels launched through this ctx so far */
/* Tuning / test knobs: largest single-pass NTT radix (log2, default 11, range 1..13) and the
 * bytes of inter-pass scratch per column group (default 4 GiB = one launch over all columns of C2: L2-sized groups
 * were measured slower because the passes are multiplier-bound; 0 keeps the current value). */
int swirl_ctx_set_ntt_plan(swirl_ctx* ctx, int max_log_radix, size_t scratch_bytes);
/* GpuProverConfig::cache_rs_code_matrix (reference cuda-backend/src/device.rs:102-121).  on (default here): commitments keep
 * their RS codeword for the WHIR openings.  off (the reference's default): the codeword is streamed through a 32-column
 * scratch at commit time and the opened rows are re-encoded by column groups in the openings -- the large-trace mode. */
int swirl_ctx_set_cache_rs_code_matrix(swirl_ctx* ctx, int on);
/* Device memory the context's scratch arena holds (blocks of 1 MiB and more; the caller's own buffers are not counted):
 * out = { bytes handed out now, high-water mark of that, bytes held (idle + handed out), free bytes on the device }.
 * reset_peak != 0 restarts the high-water mark.  Reference: MemTracker (cuda-common/src/memory_manager). */
int swirl_ctx_mem_stats(swirl_ctx* ctx, int reset_peak, uint64_t out[4]);
const char* swirl_last_error(void);
/* Per-kernel-family device timing with CUDA events on the ctx stream (off by default).
 * enable(on) clears the recorded spans; read() synchronises and returns the summed duration and
 * number of launches of a family: 0 leaf hash + query levels, 1 upper tree layers, 2 chunk
 * iDFT+zeta, 3 strided NTT passes, 4 final NTT pass, 5 stacking, 6 GKR round kernels, 7 batch-constraint
 * round 0 (coset evaluation of the constraint DAG), 8 batch-constraint MLE rounds.
 * Reference: gpu_metrics_span_on, cuda-common/src/stream.rs:278-303. */
int swirl_ctx_timing_enable(swirl_ctx* ctx, int on);
int swirl_ctx_timing_read(swirl_ctx* ctx, int slot, double* total_ms, uint64_t* count);
/* Stream synchronisations issued by the library on this context so far and the wall time spent inside them. */
int swirl_ctx_sync_stats(swirl_ctx* ctx, uint64_t* count, double* wait_ms);
/* Round link: the kernels of all rounds of a sumcheck are enqueued up front and exchange each round's
 * result / next challenge with the host transcript through a mapped pinned mailbox, so neither a launch nor a stream
 * synchronisation sits between two rounds (on = 1, the default; 0 = one launch + cudaStreamSynchronize per round, the
 * reference's pattern: cuda-backend/src/logup_zerocheck/fractional.rs:649-, sponge.rs:267-300).  Same proof either way.
 * swirl_ctx_link_stats: round results received through the mailbox so far. */
int swirl_ctx_set_round_link(swirl_ctx* ctx, int on);
int swirl_ctx_link_stats(swirl_ctx* ctx, uint64_t* count);
/* Algorithmic bytes of the recorded launches of a family; 0 for families without accounting. */
int swirl_ctx_timing_bytes(swirl_ctx* ctx, int slot, uint64_t* bytes);

/* ---- device memory + transport (reference: DeviceDataTransporter, hal.rs:141-207;
 *      DeviceBuffer / cuda_memcpy, cuda-common/src/{d_buffer.rs,copy.rs}) ---------------------- */
int swirl_malloc(swirl_ctx* ctx, size_t bytes, void** d_out); /* stream-ordered; blocks >= 1 MiB come from the ctx arena */
int swirl_free(swirl_ctx* ctx, void* d_ptr);
/* The context keeps large scratch blocks for reuse by the next proof (the role of the reference's VPMM pool,
 * docs/vpmm_spec.md); trim returns the idle ones to the driver. */
int swirl_ctx_trim(swirl_ctx* ctx);
int swirl_memcpy_h2d(swirl_ctx* ctx, void* d_dst, const void* h_src, size_t bytes); /* async on the ctx stream */
int swirl_memcpy_d2h(swirl_ctx* ctx, void* h_dst, const void* d_src, size_t bytes); /* synchronises */

/* ---- kernel-level primitives (reference: the `_name` launchers listed in SURVEY.md §2b) ----- */

/* In-place Poseidon2 permutation of n 16-word states.
 * Reference: poseidon2::poseidon2_mix, cuda-common/include/poseidon2.cuh:184-202. */
int swirl_poseidon2_permute(swirl_ctx* ctx, uint32_t* d_states, size_t n);
/* d_out[i] = compress(d_pairs[2i], d_pairs[2i+1]); reference: _poseidon2_adjacent_compress_layer,
 * cuda-backend/src/cuda/merkle_tree.rs:40-45. */
int swirl_poseidon2_compress(swirl_ctx* ctx, const uint32_t* d_pairs, uint32_t* d_out, size_t n);

/* Natural-order forward / inverse DFT of `cols` contiguous columns of length 2^log_n, in place.
 * Reference: batch_ntt (cuda-backend/src/ntt.rs:111-168) = _bit_rev + _ct_mixed_radix_narrow. */
int swirl_ntt_batch(swirl_ctx* ctx, uint32_t* d_data, int log_n, size_t cols, int inverse);

/* Reed–Solomon encoding of a stacked matrix: d_in is height x width column-major (height a power
 * of two >= 2^l_skip), d_out receives (height << log_blowup) x width.
 * Reference: rs_code_matrix, cuda-backend/src/stacked_pcs.rs:229-337
 * (== prover/stacked_pcs.rs:341-367). */
int swirl_rs_encode(swirl_ctx* ctx, const uint32_t* d_in, size_t height, size_t width, int l_skip,
                    int log_blowup, uint32_t* d_out);

/* Merkle tree over a column-major base-field matrix.  d_layers receives every digest layer
 * concatenated (layer 0 = query_stride digests ... root), 2*query_stride-1 digests of 8 words,
 * query_stride = next_pow2(height) >> log_rows_per_query.
 * Reference: _poseidon2_compressing_row_hashes + _poseidon2_adjacent_compress_layer
 * (cuda-backend/src/cuda/merkle_tree.rs:18-45; MerkleTreeGpu::new merkle_tree.rs:140-197). */
int swirl_merkle_tree(swirl_ctx* ctx, const uint32_t* d_matrix, size_t height, size_t width,
                      int log_rows_per_query, uint32_t* d_layers);
/* d_out[q][l] = sibling digest of query d_indices[q] at layer l, l < log2(query_stride).
 * Reference: _query_digest_layers (cuda/merkle_tree.rs:47-54; stacked_pcs.rs:388-405). */
int swirl_merkle_query_proofs(swirl_ctx* ctx, const uint32_t* d_layers, size_t query_stride,
                              const uint32_t* d_indices, size_t num_queries, uint32_t* d_out);
/* d_out[q][t][c] = matrix[c*height + t*query_stride + d_indices[q]], t < 2^log_rows_per_query.
 * Reference: _matrix_get_rows_fp (cuda/matrix.rs:24-32; stacked_pcs.rs:516-540). */
int swirl_matrix_open_rows(swirl_ctx* ctx, const uint32_t* d_matrix, size_t height, size_t width,
                           size_t query_stride, int log_rows_per_query, const uint32_t* d_indices,
                           size_t num_queries, uint32_t* d_out);

/* Proof-of-work search on an 18-word sponge state (16 state words, absorb_idx, sample_idx).
 * Finds the smallest canonical witness w in [min_w, max_w) with check_witness(bits, w) true and
 * writes it (canonical) to *h_witness, or UINT32_MAX when none exists.
 * Reference: _sponge_grind, cuda-backend/src/cuda/sponge.rs:13-21, sponge.cu:65-117 (which
 * returns *a* witness; the smallest one is among its admissible answers). */
int swirl_sponge_grind(swirl_ctx* ctx, const uint32_t h_state[18], int bits, uint32_t min_w,
                       uint32_t max_w, uint32_t* h_witness);

/* ---- Fiat–Shamir transcript (reference: FiatShamirTranscript, transcript/traits.rs:11-90 over
 *      DuplexSponge, transcript/duplex_sponge.rs:16-115).  The struct is the reference's
 *      DeviceSpongeState (cuda-backend/cuda/src/sponge.cu:13-17): 16 Montgomery state words,
 *      absorb_idx in [0,8), sample_idx in [0,8].  All-zero = a fresh transcript.  Host-resident;
 *      the phase-level prover entry points below read and advance it. -------------------------- */
typedef struct {
    uint32_t state[16];
    uint32_t absorb_idx;
    uint32_t sample_idx;
} swirl_transcript;
int swirl_transcript_observe(swirl_transcript* ts, const uint32_t* words, size_t n); /* Montgomery words */
int swirl_transcript_sample(swirl_transcript* ts, uint32_t* out, size_t n);
int swirl_transcript_sample_bits(swirl_transcript* ts, int bits, uint32_t* out);
/* check_witness (traits.rs:63-69): *ok = 1 iff the canonical witness passes; advances the transcript. */
int swirl_transcript_check_witness(swirl_transcript* ts, int bits, uint32_t witness, int* ok);
/* grind (traits.rs:71-86): finds the smallest canonical witness, observes it, writes it (canonical). */
int swirl_transcript_grind(swirl_ctx* ctx, swirl_transcript* ts, int bits, uint32_t* witness);

/* Linear fold of the low variable of a column-major EF matrix with even height (the MLE / FRI-style fold of every
 * sumcheck round): out[j] = in[2j] + (in[2j+1] - in[2j]) * r over the flat array, n_out = width * height / 2.
 * Reference: fold_mle_evals (prover/sumcheck.rs:395-414), GPU `fold_mle` (cuda-backend/cuda/src/sumcheck.cu). */
int swirl_fold_mle(swirl_ctx* ctx, const uint32_t* d_in, uint32_t* d_out, size_t n_out, const uint32_t r[4]);

/* ---- phase level: LogUp-GKR fractional sumcheck (reference: fractional_sumcheck,
 *      prover/logup_zerocheck/fractional_sumcheck_gkr.rs:60-213; GPU fractional_sumcheck_gpu,
 *      cuda-backend/src/logup_zerocheck/fractional.rs:649-) --------------------------------------
 * d_leaves: 2^log_n fractions Frac<EF> = {p[4], q[4]} (32 bytes, repr(C), :29-35), device.
 * Outputs (host, Montgomery words):
 *   h_frac_sum[8]              (p0, q0) root of the fraction tree
 *   h_claims[log_n][16]        GkrLayerClaims {p_xi_0, q_xi_0, p_xi_1, q_xi_1} per layer
 *   h_polys[sum_{j<log_n} j][12]  s(1), s(2), s(3) per sumcheck round, layer-major (may be NULL iff log_n == 1)
 *   h_xi[log_n][4]             the final evaluation point xi
 * Returns SWIRL_ERR_NONZERO_ROOT_SUM when assert_zero is set and the numerator sum is not zero. */
int swirl_gkr_fractional_sumcheck(swirl_ctx* ctx, swirl_transcript* ts, const uint32_t* d_leaves, int log_n,
                                  int assert_zero, uint32_t h_frac_sum[8], uint32_t* h_claims,
                                  uint32_t* h_polys, uint32_t* h_xi);
/* Same, for a leaf layer whose tail is the constant fraction (0, pad_q): only the first n_stored leaves
 * (2^log_n, or a multiple of 4) are read; the interaction layout of prove_zerocheck_and_logup pads with
 * (0, alpha) (prover/logup_zerocheck/mod.rs:103-168).  The proof is the one the full 2^log_n leaves give. */
int swirl_gkr_fractional_sumcheck_padded(swirl_ctx* ctx, swirl_transcript* ts, const uint32_t* d_leaves, uint64_t n_stored,
                                         const uint32_t pad_q[4], int log_n, int assert_zero, uint32_t h_frac_sum[8],
                                         uint32_t* h_claims, uint32_t* h_polys, uint32_t* h_xi);

/* ---- phase level: TraceCommitter::commit (hal.rs:84-87) ------------------------------------- */

typedef struct {
    int32_t l_skip;
    int32_t n_stack;
    int32_t log_blowup;
    int32_t k_whir; /* rows per query = 2^k_whir */
} swirl_pcs_params;

typedef struct {
    const uint32_t* data; /* column-major, height * width words */
    uint64_t height;      /* power of two */
    uint64_t width;
} swirl_matrix;

/* stacked_commit (reference: cuda-backend/src/stacked_pcs.rs:50-88 == prover/stacked_pcs.rs:116-134):
 * stack the height-sorted device-resident traces, RS-encode, Merkle-commit.  Writes the root to
 * h_root (host) and returns the PCS data that prove_openings consumes.  The traces must stay
 * alive and unmodified while *out is (the stacked matrix may alias them). */
int swirl_commit(swirl_ctx* ctx, const swirl_pcs_params* params, const swirl_matrix* d_traces,
                 size_t n_traces, uint32_t h_root[8], swirl_pcs** out);
/* Same with host-resident traces: H2D transport (reference: transport_matrix_to_device,
 * cuda-backend/src/data_transporter.rs:93-106) + commit in one call. */
int swirl_commit_host(swirl_ctx* ctx, const swirl_pcs_params* params, const swirl_matrix* h_traces,
                      size_t n_traces, uint32_t h_root[8], swirl_pcs** out);
int swirl_pcs_free(swirl_ctx* ctx, swirl_pcs* pcs);
/* Opened rows of the commitment's codeword for `num_queries` query indices (device u32, < query_stride):
 * d_out[q][t][c], t < 2^k_whir strided rows, c < stacked width (reference: MerkleTreeGpu::batch_open_rows,
 * cuda-backend/src/merkle_tree.rs:199-).  Works with and without a cached codeword (swirl_ctx_set_cache_rs_code_matrix). */
int swirl_pcs_open_rows(swirl_ctx* ctx, const swirl_pcs* pcs, const uint32_t* d_indices, size_t num_queries, uint32_t* d_out);

/* ---- commitments whose codeword and Merkle tree live outside this context (a commitment sharded over several GPUs) ----
 * swirl_stack: layout + stacked matrix only (the first half of TraceCommitter::commit; reference
 * stack_traces_into_expanded, cuda-backend/src/stacked_pcs.rs:143-220): every rank of a sharded commitment takes its
 * column slice from swirl_pcs_stacked_matrix.  swirl_pcs_attach_external: gives the handle the commitment's root and a
 * callback that the WHIR opening calls instead of reading a local codeword / digest layers:
 *   fn(user, h_indices[nq] (host), nq, d_rows (device, [nq][2^k_whir][stacked width] words),
 *      d_paths (device, [nq][log2(query_stride)][8] words))  ->  0, with both buffers complete on return.
 * The stream of the context is idle when the callback runs.  No reference counterpart (the reference commits on one GPU);
 * the opened rows and paths must be the ones MerkleTreeGpu would return (merkle_tree.rs:199-, stacked_pcs.rs:388-405). */
typedef int (*swirl_open_fn)(void* user, const uint32_t* h_indices, size_t num_queries, uint32_t* d_rows, uint32_t* d_paths);
int swirl_stack(swirl_ctx* ctx, const swirl_pcs_params* params, const swirl_matrix* d_traces, size_t n_traces, swirl_pcs** out);
int swirl_pcs_attach_external(swirl_pcs* pcs, const uint32_t root[8], swirl_open_fn fn, void* user);

/* PCS data accessors */
uint64_t swirl_pcs_stacked_height(const swirl_pcs* pcs);
uint64_t swirl_pcs_stacked_width(const swirl_pcs* pcs);
uint64_t swirl_pcs_codeword_height(const swirl_pcs* pcs);
uint64_t swirl_pcs_query_stride(const swirl_pcs* pcs);
const uint32_t* swirl_pcs_stacked_matrix(const swirl_pcs* pcs); /* device */
const uint32_t* swirl_pcs_codeword(const swirl_pcs* pcs);       /* device */
const uint32_t* swirl_pcs_layers(const swirl_pcs* pcs);         /* device, concatenated layers */
/* Layout: number of sorted unstacked columns; fills h_out (5 x u64 per column:
 * matrix index, column in matrix, stacked col, stacked row, log_height) when non-NULL.
 * Reference: StackedLayout::sorted_cols, prover/stacked_pcs.rs:34-41. */
uint64_t swirl_pcs_layout(const swirl_pcs* pcs, uint64_t* h_out);

/* Sharded commitment (one commitment over G ranks, column-sharded RS encode): the row exchange in front of the leaf
 * hashing as one kernel over peer memory.  d_src: this rank's `cols` codeword columns (rows x cols, column-major, global
 * columns [col_offset, col_offset + cols)); peer_bases[r]: rank r's shard buffer (W x rows/world words, column-major),
 * mapped into this process (CUDA IPC / symmetric memory over NVLink).  Writes, for every destination r, the 2^k strided
 * row segments of r's queries at [(col_offset + c) * rows/world + t * S/world + q'].  The caller synchronises the ranks
 * before (buffers free) and after (data complete).  No reference counterpart (the reference commits on one GPU). */
int swirl_scatter_rows_to_peers(swirl_ctx* ctx, const uint32_t* d_src, uint64_t rows, uint64_t cols, uint64_t col_offset,
                                int log_rows_per_query, int world, void* const* peer_bases);

/* Host-side layout computation only (StackedLayout::new, prover/stacked_pcs.rs:144-203). */
int swirl_stacked_layout(int l_skip, int log_stacked_height, size_t n_mats, const uint64_t* widths,
                         const int32_t* log_heights, uint64_t* out_width, uint64_t* out_n,
                         uint64_t* out_cols);

/* ---- phase level: WHIR opening (reference: prove_whir_opening, prover/whir.rs:78-341; GPU
 *      prove_whir_opening_gpu, cuda-backend/src/whir.rs:63-560; config: WhirConfig, config.rs:172-197) */
typedef struct {
    int32_t k;               /* folding factor, must equal the commitments' k_whir */
    int32_t num_rounds;      /* WhirConfig::rounds.len(), <= 32 */
    int32_t num_queries[32]; /* WhirRoundConfig::num_queries per round */
    int32_t mu_pow_bits;
    int32_t query_phase_pow_bits;
    int32_t folding_pow_bits;
} swirl_whir_config;
/* Length in 32-bit words of the flat WhirProof for commitments of the given stacked widths.
 * Layout (field elements as Montgomery words, R = num_rounds, m = l_skip + n_stack), in the field
 * order of WhirProof (proof.rs):
 *   mu_pow_witness[1] | whir_sumcheck_polys[R*k][2][4] | codeword_commits[R-1][8] | ood_values[R-1][4]
 *   | folding_pow_witnesses[R*k] | query_phase_pow_witnesses[R]
 *   | initial_round_opened_rows: per commit, per query [2^k][width]
 *   | initial_round_merkle_proofs: per commit, per query [m + log_blowup - k][8]
 *   | codeword_opened_values: per round r = 1..R-1, per query [2^k][4]
 *   | codeword_merkle_proofs: per round r = 1..R-1, per query [m + log_blowup - r - k][8]
 *   | final_poly[2^(m - R*k)][4].   Returns 0 for an invalid configuration. */
size_t swirl_whir_proof_words(const swirl_pcs_params* params, const swirl_whir_config* cfg, size_t n_commits,
                              const uint64_t* widths);
/* Opens all columns of the commitments (common main first, then cached / preprocessed, as in
 * WhirProver::prove_whir) at the point h_u (m EF, = u_cube of cpu_backend.rs:203-210). */
int swirl_whir_open(swirl_ctx* ctx, swirl_transcript* ts, const swirl_whir_config* cfg, const swirl_pcs* const* pcs,
                    size_t n_commits, const uint32_t* h_u, uint32_t* h_proof, size_t proof_words);

/* ---- phase level: stacked opening reduction (reference: prove_stacked_opening_reduction,
 *      prover/stacked_reduction.rs:67-127 with StackedReductionCpu :129-506; GPU
 *      cuda-backend/src/stacked_reduction.rs:188) -------------------------------------------------
 * pcs: common main first, then preprocessed / cached commitments in the order of
 * StackedReductionProver::new (stacked_reduction.rs:36-50).  need_rot[c][mat] != 0 iff matrix
 * `mat` of commitment c is opened with its rotation.  h_r: r_len >= 1 + n_max EF words, the
 * point produced by the batch constraint sumcheck.  Flat proof (Montgomery words), in the field
 * order of StackingProof (proof.rs:155-163):
 *   univariate_round_coeffs[2(2^l_skip - 1) + 1][4] | sumcheck_round_polys[n_stack][2][4]
 *   | stacking_openings: per commitment [stacked width][4].
 * h_u receives u (n_stack + 1 EF). */
size_t swirl_stacked_reduction_proof_words(const swirl_pcs* const* pcs, size_t n_commits);
int swirl_stacked_reduction(swirl_ctx* ctx, swirl_transcript* ts, const swirl_pcs* const* pcs, size_t n_commits,
                            const uint8_t* const* need_rot, const uint32_t* h_r, size_t r_len, uint32_t* h_proof,
                            size_t proof_words, uint32_t* h_u);

/* ---- phase level: MultiRapProver::prove_rap_constraints = LogUp-GKR + batch constraint sumcheck
 *      (reference: prove_zerocheck_and_logup, prover/logup_zerocheck/mod.rs:40-438 with
 *      LogupZerocheckCpu, cpu.rs:72-695; GPU prove_zerocheck_and_logup_gpu,
 *      cuda-backend/src/logup_zerocheck/mod.rs:119-434) ---------------------------------------------
 * AIR constraints cross the boundary as the reference's serialisable DAG (SymbolicConstraintsDag,
 * air_builders/symbolic/dag.rs:17-96), nodes in topological order: */
enum {
    SWIRL_NODE_VAR_PREP = 0,      /* Entry::Preprocessed: a = column index, b = row offset (0 | 1) */
    SWIRL_NODE_VAR_MAIN = 1,      /* Entry::Main: a = column index, b = row offset, c = part_index (cached.., common last) */
    SWIRL_NODE_VAR_PUBLIC = 2,    /* Entry::Public: a = index */
    SWIRL_NODE_IS_FIRST = 3,
    SWIRL_NODE_IS_LAST = 4,
    SWIRL_NODE_IS_TRANSITION = 5,
    SWIRL_NODE_CONST = 6,         /* a = Montgomery word */
    SWIRL_NODE_ADD = 7,           /* a, b = node indices */
    SWIRL_NODE_SUB = 8,
    SWIRL_NODE_NEG = 9,           /* a = node index */
    SWIRL_NODE_MUL = 10
};
typedef struct {
    uint32_t op, a, b, c;
} swirl_dag_node;
typedef struct {       /* Interaction<usize>, interaction/mod.rs: message / count as node indices */
    uint32_t count_node;
    uint32_t bus_index;
    uint32_t msg_offset; /* into swirl_air_ctx::msg_nodes */
    uint32_t msg_len;
} swirl_interaction;
typedef struct {       /* one present AIR: vk constraint data + AirProvingContext (prover/types.rs:18-73) */
    const swirl_dag_node* nodes;
    uint64_t n_nodes;
    const uint32_t* constraint_idx; /* SymbolicExpressionDag::constraint_idx */
    uint64_t n_constraints;
    const swirl_interaction* interactions;
    uint64_t n_interactions;
    const uint32_t* msg_nodes;
    uint32_t constraint_degree;     /* vk.max_constraint_degree of this AIR */
    uint32_t need_rot;              /* vk.params.need_rot */
    const uint32_t* public_values;  /* host, Montgomery words */
    uint64_t n_public_values;
    swirl_matrix common_main;       /* device */
    const swirl_matrix* cached_mains; /* device matrices */
    uint64_t n_cached;
    const swirl_matrix* preprocessed; /* device matrix or NULL */
} swirl_air_ctx;
/* Flat proof (Montgomery words), GkrProof then BatchConstraintProof in field order (proof.rs:70-134):
 *   logup_pow_witness[1] | q0_claim[4] | claims_per_layer[L][16] | sumcheck_polys[L(L-1)/2][12]
 *   | numerator_term_per_air[n][4] | denominator_term_per_air[n][4]
 *   | univariate_round_coeffs[(D+1)(2^l_skip - 1) + 1][4] | sumcheck_round_polys[n_max][D+1][4]
 *   | column_openings: per AIR, per part (common main, preprocessed, cached..) width * (need_rot ? 2 : 1) EF,
 *     rotations interleaved (claim, claim_rot).
 * L = l_skip + n_logup (0 when no AIR has interactions), D = max_constraint_degree, n = n_airs,
 * n_max = max(log2 height) - l_skip clamped at 0.  AIRs must be sorted by descending height (as
 * the Coordinator does, prover/types.rs:144-148).  h_r receives r (n_max + 1 EF).
 * Returns SWIRL_ERR_NONZERO_ROOT_SUM for unbalanced interactions. */
size_t swirl_batch_constraints_proof_words(int l_skip, int max_constraint_degree, const swirl_air_ctx* airs, size_t n_airs);
int swirl_prove_batch_constraints(swirl_ctx* ctx, swirl_transcript* ts, int l_skip, int max_constraint_degree,
                                  int logup_pow_bits, const swirl_air_ctx* airs, size_t n_airs, uint32_t* h_proof,
                                  size_t proof_words, uint32_t* h_r);

/* Per-AIR run-time compiled constraint kernels: for traces of 2^17 rows and more the round-0 program of
 * an AIR is emitted as straight-line CUDA C++, compiled with NVRTC for sm_100a at first use and cached in the context
 * (SWIRL_JIT=0, or a missing libnvrtc, selects the interpreter kernels instead; both are GPU paths and produce the same
 * words).  This returns that source for inspection (which = 0: constraint roots, 1: interaction roots); host only, the
 * matrices of `air` need only their shapes.  Returns the source length; copies at most cap - 1 characters into out.
 * Reference counterpart: the rule compiler + interpreter, cuda-backend/src/logup_zerocheck/rules/mod.rs:27-130. */
/* mode & 3: 0 = interpreter only, 1 = compile the programs of tall traces (default), 2 = compile every program (tests).
 * mode & 4: the MLE rounds run compiled kernels as well (one kernel per AIR: a switch over its sub-programs, value slots
 * as D extension-field lanes in registers), under the same tall-trace rule.  On by default since it was measured
 * (profiles/r3a_*); SWIRL_JIT_MLE=0/1 in the environment sets the context's initial choice, mode 0 switches every
 * compiled kernel off. */
int swirl_ctx_set_jit(swirl_ctx* ctx, int mode);
size_t swirl_jit_round0_source(const swirl_air_ctx* air, int which, char* out, size_t cap);
/* The MLE-round kernel of one AIR for max_constraint_degree D, as it would be compiled when `n_airs` AIRs share the proof
 * (the number of sub-programs per AIR depends on it).  The sub-programs' instructions and their (case, column shift,
 * weight shift) follow the kernel as comment lines.  Host only; returns 0 when the interpreter would be used. */
size_t swirl_jit_mle_source(const swirl_air_ctx* air, int max_constraint_degree, size_t n_airs, char* out, size_t cap);
/* out = {round-0 kernels compiled, round-0 launches of compiled kernels, MLE-round kernels compiled, MLE-round launches of
 * compiled kernels} since the context was created: tells a caller (and the tests) which path really ran. */
int swirl_ctx_jit_stats(swirl_ctx* ctx, uint64_t out[4]);

/* ---- phase level: OpeningProver::prove_openings (hal.rs:118-138; cpu_backend.rs:139-220) =
 *      swirl_stacked_reduction, u_cube = (u_0^(2^i))_{i<l_skip} ++ u[1..], swirl_whir_open.
 *      Buffers as in those two calls. -------------------------------------------------------------- */
int swirl_prove_openings(swirl_ctx* ctx, swirl_transcript* ts, const swirl_whir_config* cfg, const swirl_pcs* const* pcs,
                         size_t n_commits, const uint8_t* const* need_rot, const uint32_t* h_r, size_t r_len,
                         uint32_t* h_stacking_proof, size_t stacking_words, uint32_t* h_whir_proof, size_t whir_words);

#ifdef __cplusplus
}
#endif
#endif /* SWIRL_B200_H */
