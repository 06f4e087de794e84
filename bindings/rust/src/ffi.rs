//! Raw FFI of `libswirl_b200.so` — one declaration per entry point of `include/swirl_b200.h` (generated from the header
//! by the script in `tests/test_abi.py::test_rust_ffi_lists_every_symbol`'s docstring; kept in sync by that test).
//! This crate cannot be built in the image the library was developed in (no Rust toolchain); it is the binding a
//! maintainer of openvm-org/stark-backend adds next to `crates/cuda-backend/src/cuda/*.rs`.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct SwirlCtx { _private: [u8; 0] }
#[repr(C)] pub struct SwirlPcs { _private: [u8; 0] }
/// DeviceSpongeState (crates/cuda-backend/cuda/src/sponge.cu:13-17)
#[repr(C)] #[derive(Clone, Copy)] pub struct SwirlTranscript { pub state: [u32; 16], pub absorb_idx: u32, pub sample_idx: u32 }
#[repr(C)] #[derive(Clone, Copy)] pub struct SwirlPcsParams { pub l_skip: i32, pub n_stack: i32, pub log_blowup: i32, pub k_whir: i32 }
/// DeviceMatrix<F> (crates/cuda-backend/src/base.rs:8-12): column-major Montgomery words on the device
#[repr(C)] #[derive(Clone, Copy)] pub struct SwirlMatrix { pub data: *const u32, pub height: u64, pub width: u64 }
/// WhirConfig (crates/stark-backend/src/config.rs:172-197)
#[repr(C)] #[derive(Clone, Copy)] pub struct SwirlWhirConfig { pub k: i32, pub num_rounds: i32, pub num_queries: [i32; 32],
    pub mu_pow_bits: i32, pub query_phase_pow_bits: i32, pub folding_pow_bits: i32 }
/// SymbolicExpressionNode (air_builders/symbolic/dag.rs:17-45) in the boundary encoding
/// swirl_open_fn of include/swirl_b200.h
pub type SwirlOpenFn = unsafe extern "C" fn(user: *mut c_void, h_indices: *const u32, num_queries: usize, d_rows: *mut u32, d_paths: *mut u32) -> c_int;
/// SWIRL_NODE_* of include/swirl_b200.h
pub const SWIRL_NODE_VAR_PREP: u32 = 0;
pub const SWIRL_NODE_VAR_MAIN: u32 = 1;
pub const SWIRL_NODE_VAR_PUBLIC: u32 = 2;
pub const SWIRL_NODE_IS_FIRST: u32 = 3;
pub const SWIRL_NODE_IS_LAST: u32 = 4;
pub const SWIRL_NODE_IS_TRANSITION: u32 = 5;
pub const SWIRL_NODE_CONST: u32 = 6;
pub const SWIRL_NODE_ADD: u32 = 7;
pub const SWIRL_NODE_SUB: u32 = 8;
pub const SWIRL_NODE_NEG: u32 = 9;
pub const SWIRL_NODE_MUL: u32 = 10;
#[repr(C)] #[derive(Clone, Copy)] pub struct SwirlDagNode { pub op: u32, pub a: u32, pub b: u32, pub c: u32 }
#[repr(C)] #[derive(Clone, Copy)] pub struct SwirlInteraction { pub count_node: u32, pub bus_index: u32, pub msg_offset: u32, pub msg_len: u32 }
#[repr(C)] pub struct SwirlAirCtx { pub nodes: *const SwirlDagNode, pub n_nodes: u64, pub constraint_idx: *const u32, pub n_constraints: u64,
    pub interactions: *const SwirlInteraction, pub n_interactions: u64, pub msg_nodes: *const u32, pub constraint_degree: u32,
    pub need_rot: u32, pub public_values: *const u32, pub n_public_values: u64, pub common_main: SwirlMatrix,
    pub cached_mains: *const SwirlMatrix, pub n_cached: u64, pub preprocessed: *const SwirlMatrix }

#[link(name = "swirl_b200")]
extern "C" {
    pub fn swirl_ctx_create(device: c_int, out: *mut *mut SwirlCtx) -> c_int;
    pub fn swirl_ctx_create_on_stream(device: c_int, cuda_stream: *mut c_void, out: *mut *mut SwirlCtx) -> c_int;
    pub fn swirl_ctx_destroy(ctx: *mut SwirlCtx) -> c_int;
    pub fn swirl_ctx_synchronize(ctx: *mut SwirlCtx) -> c_int;
    pub fn swirl_ctx_stream(ctx: *mut SwirlCtx) -> *mut c_void;
    pub fn swirl_ctx_launch_count(ctx: *mut SwirlCtx) -> u64;
    pub fn swirl_ctx_set_ntt_plan(ctx: *mut SwirlCtx, max_log_radix: c_int, scratch_bytes: usize) -> c_int;
    pub fn swirl_ctx_set_cache_rs_code_matrix(ctx: *mut SwirlCtx, on: c_int) -> c_int;
    pub fn swirl_ctx_mem_stats(ctx: *mut SwirlCtx, reset_peak: c_int, out: *mut u64) -> c_int;
    pub fn swirl_last_error() -> *const c_char;
    pub fn swirl_ctx_timing_enable(ctx: *mut SwirlCtx, on: c_int) -> c_int;
    pub fn swirl_ctx_timing_read(ctx: *mut SwirlCtx, slot: c_int, total_ms: *mut f64, count: *mut u64) -> c_int;
    pub fn swirl_ctx_sync_stats(ctx: *mut SwirlCtx, count: *mut u64, wait_ms: *mut f64) -> c_int;
    pub fn swirl_ctx_set_round_link(ctx: *mut SwirlCtx, on: c_int) -> c_int;
    pub fn swirl_ctx_link_stats(ctx: *mut SwirlCtx, count: *mut u64) -> c_int;
    pub fn swirl_ctx_timing_bytes(ctx: *mut SwirlCtx, slot: c_int, bytes: *mut u64) -> c_int;
    pub fn swirl_malloc(ctx: *mut SwirlCtx, bytes: usize, d_out: *mut *mut c_void) -> c_int;
    pub fn swirl_free(ctx: *mut SwirlCtx, d_ptr: *mut c_void) -> c_int;
    pub fn swirl_ctx_trim(ctx: *mut SwirlCtx) -> c_int;
    pub fn swirl_memcpy_h2d(ctx: *mut SwirlCtx, d_dst: *mut c_void, h_src: *const c_void, bytes: usize) -> c_int;
    pub fn swirl_memcpy_d2h(ctx: *mut SwirlCtx, h_dst: *mut c_void, d_src: *const c_void, bytes: usize) -> c_int;
    pub fn swirl_poseidon2_permute(ctx: *mut SwirlCtx, d_states: *mut u32, n: usize) -> c_int;
    pub fn swirl_poseidon2_compress(ctx: *mut SwirlCtx, d_pairs: *const u32, d_out: *mut u32, n: usize) -> c_int;
    pub fn swirl_ntt_batch(ctx: *mut SwirlCtx, d_data: *mut u32, log_n: c_int, cols: usize, inverse: c_int) -> c_int;
    pub fn swirl_rs_encode(ctx: *mut SwirlCtx, d_in: *const u32, height: usize, width: usize, l_skip: c_int, log_blowup: c_int, d_out: *mut u32) -> c_int;
    pub fn swirl_merkle_tree(ctx: *mut SwirlCtx, d_matrix: *const u32, height: usize, width: usize, log_rows_per_query: c_int, d_layers: *mut u32) -> c_int;
    pub fn swirl_merkle_query_proofs(ctx: *mut SwirlCtx, d_layers: *const u32, query_stride: usize, d_indices: *const u32, num_queries: usize, d_out: *mut u32) -> c_int;
    pub fn swirl_matrix_open_rows(ctx: *mut SwirlCtx, d_matrix: *const u32, height: usize, width: usize, query_stride: usize, log_rows_per_query: c_int, d_indices: *const u32, num_queries: usize, d_out: *mut u32) -> c_int;
    pub fn swirl_sponge_grind(ctx: *mut SwirlCtx, h_state: *const u32, bits: c_int, min_w: u32, max_w: u32, h_witness: *mut u32) -> c_int;
    pub fn swirl_transcript_observe(ts: *mut SwirlTranscript, words: *const u32, n: usize) -> c_int;
    pub fn swirl_transcript_sample(ts: *mut SwirlTranscript, out: *mut u32, n: usize) -> c_int;
    pub fn swirl_transcript_sample_bits(ts: *mut SwirlTranscript, bits: c_int, out: *mut u32) -> c_int;
    pub fn swirl_transcript_check_witness(ts: *mut SwirlTranscript, bits: c_int, witness: u32, ok: *mut c_int) -> c_int;
    pub fn swirl_transcript_grind(ctx: *mut SwirlCtx, ts: *mut SwirlTranscript, bits: c_int, witness: *mut u32) -> c_int;
    pub fn swirl_fold_mle(ctx: *mut SwirlCtx, d_in: *const u32, d_out: *mut u32, n_out: usize, r: *const u32) -> c_int;
    pub fn swirl_gkr_fractional_sumcheck(ctx: *mut SwirlCtx, ts: *mut SwirlTranscript, d_leaves: *const u32, log_n: c_int, assert_zero: c_int, h_frac_sum: *mut u32, h_claims: *mut u32, h_polys: *mut u32, h_xi: *mut u32) -> c_int;
    pub fn swirl_gkr_fractional_sumcheck_padded(ctx: *mut SwirlCtx, ts: *mut SwirlTranscript, d_leaves: *const u32, n_stored: u64, pad_q: *const u32, log_n: c_int, assert_zero: c_int, h_frac_sum: *mut u32, h_claims: *mut u32, h_polys: *mut u32, h_xi: *mut u32) -> c_int;
    pub fn swirl_commit(ctx: *mut SwirlCtx, params: *const SwirlPcsParams, d_traces: *const SwirlMatrix, n_traces: usize, h_root: *mut u32, out: *mut *mut SwirlPcs) -> c_int;
    pub fn swirl_commit_host(ctx: *mut SwirlCtx, params: *const SwirlPcsParams, h_traces: *const SwirlMatrix, n_traces: usize, h_root: *mut u32, out: *mut *mut SwirlPcs) -> c_int;
    pub fn swirl_stack(ctx: *mut SwirlCtx, params: *const SwirlPcsParams, d_traces: *const SwirlMatrix, n_traces: usize, out: *mut *mut SwirlPcs) -> c_int;
    pub fn swirl_pcs_attach_external(pcs: *mut SwirlPcs, root: *const u32, f: SwirlOpenFn, user: *mut c_void) -> c_int;
    pub fn swirl_pcs_free(ctx: *mut SwirlCtx, pcs: *mut SwirlPcs) -> c_int;
    pub fn swirl_pcs_open_rows(ctx: *mut SwirlCtx, pcs: *const SwirlPcs, d_indices: *const u32, num_queries: usize, d_out: *mut u32) -> c_int;
    pub fn swirl_pcs_stacked_height(pcs: *const SwirlPcs) -> u64;
    pub fn swirl_pcs_stacked_width(pcs: *const SwirlPcs) -> u64;
    pub fn swirl_pcs_codeword_height(pcs: *const SwirlPcs) -> u64;
    pub fn swirl_pcs_query_stride(pcs: *const SwirlPcs) -> u64;
    pub fn swirl_pcs_stacked_matrix(pcs: *const SwirlPcs) -> *const u32;
    pub fn swirl_pcs_codeword(pcs: *const SwirlPcs) -> *const u32;
    pub fn swirl_pcs_layers(pcs: *const SwirlPcs) -> *const u32;
    pub fn swirl_pcs_layout(pcs: *const SwirlPcs, h_out: *mut u64) -> u64;
    pub fn swirl_scatter_rows_to_peers(ctx: *mut SwirlCtx, d_src: *const u32, rows: u64, cols: u64, col_offset: u64, log_rows_per_query: c_int, world: c_int, peer_bases: *const *mut c_void) -> c_int;
    pub fn swirl_stacked_layout(l_skip: c_int, log_stacked_height: c_int, n_mats: usize, widths: *const u64, log_heights: *const i32, out_width: *mut u64, out_n: *mut u64, out_cols: *mut u64) -> c_int;
    pub fn swirl_whir_proof_words(params: *const SwirlPcsParams, cfg: *const SwirlWhirConfig, n_commits: usize, widths: *const u64) -> usize;
    pub fn swirl_whir_open(ctx: *mut SwirlCtx, ts: *mut SwirlTranscript, cfg: *const SwirlWhirConfig, pcs: *const *const SwirlPcs, n_commits: usize, h_u: *const u32, h_proof: *mut u32, proof_words: usize) -> c_int;
    pub fn swirl_stacked_reduction_proof_words(pcs: *const *const SwirlPcs, n_commits: usize) -> usize;
    pub fn swirl_stacked_reduction(ctx: *mut SwirlCtx, ts: *mut SwirlTranscript, pcs: *const *const SwirlPcs, n_commits: usize, need_rot: *const *const u8, h_r: *const u32, r_len: usize, h_proof: *mut u32, proof_words: usize, h_u: *mut u32) -> c_int;
    pub fn swirl_ctx_set_jit(ctx: *mut SwirlCtx, mode: c_int) -> c_int;
    pub fn swirl_jit_round0_source(air: *const SwirlAirCtx, which: c_int, out: *mut std::ffi::c_char, cap: usize) -> usize;
    pub fn swirl_ctx_jit_stats(ctx: *mut SwirlCtx, out: *mut u64) -> c_int;
    pub fn swirl_jit_mle_source(air: *const SwirlAirCtx, max_constraint_degree: c_int, n_airs: usize, out: *mut std::ffi::c_char, cap: usize) -> usize;
    pub fn swirl_batch_constraints_proof_words(l_skip: c_int, max_constraint_degree: c_int, airs: *const SwirlAirCtx, n_airs: usize) -> usize;
    pub fn swirl_prove_batch_constraints(ctx: *mut SwirlCtx, ts: *mut SwirlTranscript, l_skip: c_int, max_constraint_degree: c_int, logup_pow_bits: c_int, airs: *const SwirlAirCtx, n_airs: usize, h_proof: *mut u32, proof_words: usize, h_r: *mut u32) -> c_int;
    pub fn swirl_prove_openings(ctx: *mut SwirlCtx, ts: *mut SwirlTranscript, cfg: *const SwirlWhirConfig, pcs: *const *const SwirlPcs, n_commits: usize, need_rot: *const *const u8, h_r: *const u32, r_len: usize, h_stacking_proof: *mut u32, stacking_words: usize, h_whir_proof: *mut u32, whir_words: usize) -> c_int;
}
