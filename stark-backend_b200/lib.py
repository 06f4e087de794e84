"""ctypes binding of ``libswirl_b200.so`` (declarations follow include/swirl_b200.h one to one).

This is the Python stand-in for the Rust ``extern "C"`` block a reference-side crate would hold
(reference: crates/cuda-backend/src/cuda/*.rs); see INTEGRATION.md.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class SwirlError(RuntimeError):
    """A non-zero return code from the C ABI (reference: CudaError::from_result,
    cuda-common/src/error.rs:53-60)."""

    def __init__(self, code, msg):
        super().__init__(f"swirl_b200 error {code}: {msg}")
        self.code = code


def library_path():
    # SWIRL_B200_LIB: an alternative build of the same library (A/B measurements of kernel variants)
    return os.environ.get("SWIRL_B200_LIB") or os.path.join(_HERE, "libswirl_b200.so")


class PcsParamsC(C.Structure):
    _fields_ = [("l_skip", C.c_int32), ("n_stack", C.c_int32), ("log_blowup", C.c_int32), ("k_whir", C.c_int32)]


class TranscriptC(C.Structure):
    """swirl_transcript == the reference's DeviceSpongeState (cuda-backend/cuda/src/sponge.cu:13-17)."""

    _fields_ = [("state", C.c_uint32 * 16), ("absorb_idx", C.c_uint32), ("sample_idx", C.c_uint32)]


class WhirConfigC(C.Structure):
    _fields_ = [("k", C.c_int32), ("num_rounds", C.c_int32), ("num_queries", C.c_int32 * 32), ("mu_pow_bits", C.c_int32),
                ("query_phase_pow_bits", C.c_int32), ("folding_pow_bits", C.c_int32)]


class MatrixC(C.Structure):
    _fields_ = [("data", C.c_void_p), ("height", C.c_uint64), ("width", C.c_uint64)]


class DagNodeC(C.Structure):
    _fields_ = [("op", C.c_uint32), ("a", C.c_uint32), ("b", C.c_uint32), ("c", C.c_uint32)]


class InteractionC(C.Structure):
    _fields_ = [("count_node", C.c_uint32), ("bus_index", C.c_uint32), ("msg_offset", C.c_uint32), ("msg_len", C.c_uint32)]


class AirCtxC(C.Structure):
    _fields_ = [("nodes", C.c_void_p), ("n_nodes", C.c_uint64), ("constraint_idx", C.c_void_p), ("n_constraints", C.c_uint64),
                ("interactions", C.c_void_p), ("n_interactions", C.c_uint64), ("msg_nodes", C.c_void_p),
                ("constraint_degree", C.c_uint32), ("need_rot", C.c_uint32), ("public_values", C.c_void_p),
                ("n_public_values", C.c_uint64), ("common_main", MatrixC), ("cached_mains", C.c_void_p), ("n_cached", C.c_uint64),
                ("preprocessed", C.c_void_p)]


_vp, _sz, _i, _u32, _u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint32, C.c_uint64
# swirl_open_fn(user, h_indices, num_queries, d_rows, d_paths) -> int
OPEN_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint32), C.c_size_t, C.c_void_p, C.c_void_p)

# name -> (restype, argtypes); every prototype of include/swirl_b200.h
PROTOTYPES = {
    "swirl_ctx_create": (_i, [_i, C.POINTER(_vp)]),
    "swirl_ctx_create_on_stream": (_i, [_i, _vp, C.POINTER(_vp)]),
    "swirl_ctx_destroy": (_i, [_vp]),
    "swirl_ctx_synchronize": (_i, [_vp]),
    "swirl_ctx_stream": (_vp, [_vp]),
    "swirl_ctx_launch_count": (_u64, [_vp]),
    "swirl_ctx_sync_stats": (_i, [_vp, C.POINTER(_u64), C.POINTER(C.c_double)]),
    "swirl_ctx_set_round_link": (_i, [_vp, _i]),
    "swirl_ctx_link_stats": (_i, [_vp, C.POINTER(_u64)]),
    "swirl_ctx_timing_bytes": (_i, [_vp, _i, C.POINTER(_u64)]),
    "swirl_ctx_set_ntt_plan": (_i, [_vp, _i, _sz]),
    "swirl_ctx_set_cache_rs_code_matrix": (_i, [_vp, _i]),
    "swirl_ctx_mem_stats": (_i, [_vp, _i, C.POINTER(_u64)]),
    "swirl_last_error": (C.c_char_p, []),
    "swirl_ctx_timing_enable": (_i, [_vp, _i]),
    "swirl_ctx_timing_read": (_i, [_vp, _i, C.POINTER(C.c_double), C.POINTER(_u64)]),
    "swirl_malloc": (_i, [_vp, _sz, C.POINTER(_vp)]),
    "swirl_free": (_i, [_vp, _vp]),
    "swirl_ctx_trim": (_i, [_vp]),
    "swirl_memcpy_h2d": (_i, [_vp, _vp, _vp, _sz]),
    "swirl_memcpy_d2h": (_i, [_vp, _vp, _vp, _sz]),
    "swirl_poseidon2_permute": (_i, [_vp, _vp, _sz]),
    "swirl_poseidon2_compress": (_i, [_vp, _vp, _vp, _sz]),
    "swirl_ntt_batch": (_i, [_vp, _vp, _i, _sz, _i]),
    "swirl_rs_encode": (_i, [_vp, _vp, _sz, _sz, _i, _i, _vp]),
    "swirl_merkle_tree": (_i, [_vp, _vp, _sz, _sz, _i, _vp]),
    "swirl_merkle_query_proofs": (_i, [_vp, _vp, _sz, _vp, _sz, _vp]),
    "swirl_matrix_open_rows": (_i, [_vp, _vp, _sz, _sz, _sz, _i, _vp, _sz, _vp]),
    "swirl_sponge_grind": (_i, [_vp, _vp, _i, _u32, _u32, C.POINTER(_u32)]),
    "swirl_transcript_observe": (_i, [C.POINTER(TranscriptC), _vp, _sz]),
    "swirl_transcript_sample": (_i, [C.POINTER(TranscriptC), _vp, _sz]),
    "swirl_transcript_sample_bits": (_i, [C.POINTER(TranscriptC), _i, C.POINTER(_u32)]),
    "swirl_transcript_check_witness": (_i, [C.POINTER(TranscriptC), _i, _u32, C.POINTER(_i)]),
    "swirl_transcript_grind": (_i, [_vp, C.POINTER(TranscriptC), _i, C.POINTER(_u32)]),
    "swirl_scatter_rows_to_peers": (_i, [_vp, _vp, _u64, _u64, _u64, _i, _i, _vp]),
    "swirl_fold_mle": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "swirl_gkr_fractional_sumcheck": (_i, [_vp, C.POINTER(TranscriptC), _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "swirl_gkr_fractional_sumcheck_padded": (_i, [_vp, C.POINTER(TranscriptC), _vp, _u64, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "swirl_commit": (_i, [_vp, C.POINTER(PcsParamsC), C.POINTER(MatrixC), _sz, _vp, C.POINTER(_vp)]),
    "swirl_commit_host": (_i, [_vp, C.POINTER(PcsParamsC), C.POINTER(MatrixC), _sz, _vp, C.POINTER(_vp)]),
    "swirl_stack": (_i, [_vp, C.POINTER(PcsParamsC), C.POINTER(MatrixC), _sz, C.POINTER(_vp)]),
    "swirl_pcs_attach_external": (_i, [_vp, _vp, _vp, _vp]),
    "swirl_pcs_free": (_i, [_vp, _vp]),
    "swirl_pcs_open_rows": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "swirl_pcs_stacked_height": (_u64, [_vp]),
    "swirl_pcs_stacked_width": (_u64, [_vp]),
    "swirl_pcs_codeword_height": (_u64, [_vp]),
    "swirl_pcs_query_stride": (_u64, [_vp]),
    "swirl_pcs_stacked_matrix": (_vp, [_vp]),
    "swirl_pcs_codeword": (_vp, [_vp]),
    "swirl_pcs_layers": (_vp, [_vp]),
    "swirl_pcs_layout": (_u64, [_vp, _vp]),
    "swirl_whir_proof_words": (_sz, [C.POINTER(PcsParamsC), C.POINTER(WhirConfigC), _sz, _vp]),
    "swirl_whir_open": (_i, [_vp, C.POINTER(TranscriptC), C.POINTER(WhirConfigC), _vp, _sz, _vp, _vp, _sz]),
    "swirl_stacked_reduction_proof_words": (_sz, [_vp, _sz]),
    "swirl_stacked_reduction": (_i, [_vp, C.POINTER(TranscriptC), _vp, _sz, _vp, _vp, _sz, _vp, _sz, _vp]),
    "swirl_batch_constraints_proof_words": (_sz, [_i, _i, _vp, _sz]),
    "swirl_prove_batch_constraints": (_i, [_vp, C.POINTER(TranscriptC), _i, _i, _i, _vp, _sz, _vp, _sz, _vp]),
    "swirl_prove_openings": (_i, [_vp, C.POINTER(TranscriptC), C.POINTER(WhirConfigC), _vp, _sz, _vp, _vp, _sz, _vp, _sz, _vp, _sz]),
    "swirl_ctx_set_jit": (_i, [_vp, _i]),
    "swirl_jit_round0_source": (_sz, [_vp, _i, C.c_char_p, _sz]),
    "swirl_jit_mle_source": (_sz, [_vp, _i, _sz, C.c_char_p, _sz]),
    "swirl_ctx_jit_stats": (_i, [_vp, _vp]),
    "swirl_stacked_layout": (_i, [_i, _i, _sz, _vp, _vp, C.POINTER(_u64), C.POINTER(_u64), _vp]),
}


def load_library():
    """Loads the CUDA library; raises (never falls back) when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C stark-backend_b200`). There is no CPU fallback."
        )
    lib = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load_library().swirl_last_error()
        raise SwirlError(rc, msg.decode() if msg else "")
