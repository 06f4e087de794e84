"""ctypes binding of oracle/_ref/libref_kernels.so — the REFERENCE's own CUDA kernel library compiled from
/root/reference by oracle/Makefile.ref.  TEST INFRASTRUCTURE ONLY (parity anchor + reference-GPU perf bar).

The host orchestration here restates the reference's Rust callers of those launchers, one function per Rust
function, so that what runs on the GPU is the reference's kernel sequence:

* ``rs_code_matrix``   — crates/cuda-backend/src/stacked_pcs.rs:229-337 (batch_expand_pad / batch_ntt_small /
  mle_interpolate_stages / bit_rev / batch_ntt) with ``batch_ntt`` = src/ntt.rs:111-168 and
  ``mle_interpolate_stages`` = src/poly.rs:162-247
* ``merkle_tree``      — src/merkle_tree.rs:140-197 (compress_rows, then adjacent compress layers)
* ``sponge_grind``     — src/sponge.rs:267-300 (one `_sponge_grind` over a witness range)
* GKR tree / WHIR fold — the launchers of cuda/src/logup_zerocheck/gkr.cu and cuda/src/whir.cu

Device memory is torch's (int32 tensors of Montgomery words, same bytes as the reference's DeviceBuffer<F>).
"""
import ctypes as C
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB_PATH = os.path.join(ROOT, "oracle", "_ref", "libref_kernels.so")

_vp, _sz, _i, _u32, _u64, _b, _u16 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint32, C.c_uint64, C.c_bool, C.c_uint16

MAX_LG_DOMAIN_SIZE = 27  # src/cuda/ntt.rs:7
LG_WINDOW_SIZE = -(-MAX_LG_DOMAIN_SIZE // 5)
WINDOW_SIZE = 1 << LG_WINDOW_SIZE
WINDOW_NUM = -(-MAX_LG_DOMAIN_SIZE // LG_WINDOW_SIZE)
RADIX_TWIDDLES_SIZE = 32 + 64 + 128 + 256 + 512  # src/ntt.rs:14-18
DEVICE_NTT_TWIDDLES_SIZE = (1 << 10) - 2  # cuda/include/device_ntt.cuh:21-22
LOG_WARP_SIZE = 5
MLE_SHARED_TILE_LOG_SIZE = 12  # cuda/src/mle_interpolate.cu:387


class FpExtC(C.Structure):
    """FpExt by value (cuda-common/include/fpext.h): 4 Montgomery words, basis 1, X, X^2, X^3."""

    _fields_ = [("c", _u32 * 4)]


def fpext(words):
    return FpExtC((_u32 * 4)(*[int(x) for x in words]))


def available():
    return os.path.exists(REF_LIB_PATH)


class RefKernels:
    def __init__(self, path=REF_LIB_PATH, device=0):
        if not os.path.exists(path):
            raise RuntimeError(f"{path} not built: `make -f Makefile.ref -C oracle` (needs /root/reference)")
        self.L = L = C.CDLL(path)
        self.device = torch.device(f"cuda:{device}")
        protos = {
            "_generate_all_twiddles": [_vp, _b, _vp],
            "_generate_partial_twiddles": [_vp, _b, _vp],
            "_generate_device_ntt_twiddles": [_vp, _vp],
            "_bit_rev": [_vp, _vp, _u32, _u32, _u32, _vp],
            "_ct_mixed_radix_narrow": [_vp, _u32, _u32, _u32, _u32, _u32, _u32, _b, _vp],
            "_batch_ntt_small": [_vp, _sz, _sz, _b, _vp],
            "_batch_expand_pad": [_vp, _vp, _u32, _u32, _u32, _vp],
            "_mle_interpolate_fused_2d": [_vp, _u16, _u32, _u32, _u32, _u32, _b, _b, _vp],
            "_mle_interpolate_shared_2d": [_vp, _u16, _u32, _u32, _u32, _u32, _b, _b, _vp],
            "_mle_interpolate_stage_2d": [_vp, _u16, _u32, _u32, _u32, _b, _vp],
            "_poseidon2_compressing_row_hashes": [_vp, _vp, _sz, _sz, _sz, _vp],
            "_poseidon2_compressing_row_hashes_ext": [_vp, _vp, _sz, _sz, _sz, _vp],
            "_poseidon2_adjacent_compress_layer": [_vp, _vp, _sz, _vp],
            "_poseidon2_strided_compress_layer": [_vp, _vp, _sz, _sz, _vp],
            "_sponge_grind": [_vp, _u32, _u32, _u32, _vp, _vp],
            "_frac_build_tree_layer": [_vp, _sz, _sz, _sz, _b, FpExtC, _b, _vp],
            "_frac_add_alpha": [_vp, _sz, FpExtC, _vp],
            "_whir_fold_coeffs_and_moments": [_vp, _vp, _vp, _vp, FpExtC, _u32, _vp],
        }
        for name, args in protos.items():
            fn = getattr(L, name)
            fn.restype = _i
            fn.argtypes = args
        self._ntt_ready = {False: False, True: False}
        self._small_ready = False

    # -- plumbing -----------------------------------------------------------------------------------------
    def stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    @staticmethod
    def check(rc, what):
        if rc != 0:
            raise RuntimeError(f"reference launcher {what} returned cudaError {rc}")

    def empty(self, n_words):
        return torch.empty(int(n_words), dtype=torch.int32, device=self.device)

    def h2d(self, arr):
        a = np.ascontiguousarray(np.asarray(arr).reshape(-1), dtype=np.uint32)
        return torch.from_numpy(a.view(np.int32)).to(self.device)

    @staticmethod
    def d2h(t):
        return t.cpu().numpy().view(np.uint32)

    # -- NTT (src/ntt.rs) -----------------------------------------------------------------------------------
    def _ensure_initialized(self, inverse):
        """src/ntt.rs:23-50: twiddle tables into __constant__ memory (the launchers synchronise)."""
        if self._ntt_ready[inverse]:
            return
        tw = self.empty(RADIX_TWIDDLES_SIZE)
        pt = self.empty(WINDOW_NUM * WINDOW_SIZE)
        self.check(self.L._generate_all_twiddles(tw.data_ptr(), inverse, self.stream()), "_generate_all_twiddles")
        self.check(self.L._generate_partial_twiddles(pt.data_ptr(), inverse, self.stream()), "_generate_partial_twiddles")
        torch.cuda.synchronize(self.device)
        self._ntt_ready[inverse] = True

    def _ensure_small(self):
        """src/cuda/batch_ntt_small.rs:56-77"""
        if self._small_ready:
            return
        tw = self.empty(DEVICE_NTT_TWIDDLES_SIZE)
        self.check(self.L._generate_device_ntt_twiddles(tw.data_ptr(), self.stream()), "_generate_device_ntt_twiddles")
        torch.cuda.synchronize(self.device)
        self._small_ready = True

    def bit_rev(self, buf, lg_domain_size, padded_poly_size, poly_count):
        self.check(self.L._bit_rev(buf.data_ptr(), buf.data_ptr(), lg_domain_size, padded_poly_size, poly_count, self.stream()),
                   "_bit_rev")

    def batch_ntt(self, buf, log_trace_height, log_blowup, width, bit_reverse, is_intt):
        """src/ntt.rs:111-168.  NOTE (reference semantics): with is_intt the kernels include the 1/n factor
        (`domain_size_inverse`, supra/ntt.cu:240)."""
        if log_trace_height == 0:
            return
        assert log_trace_height <= MAX_LG_DOMAIN_SIZE
        padded = 1 << (log_trace_height + log_blowup)
        if bit_reverse:
            self.bit_rev(buf, log_trace_height, padded, width)
        self._ensure_initialized(is_intt)
        stage = 0

        def step(iterations):
            nonlocal stage
            assert iterations <= 10
            radix = 6 if iterations < 6 else iterations
            self.check(self.L._ct_mixed_radix_narrow(buf.data_ptr(), radix, log_trace_height, stage, iterations, padded, width,
                                                     is_intt, self.stream()), "_ct_mixed_radix_narrow")
            stage += iterations

        n = log_trace_height
        if n <= 10:
            step(n)
        elif n <= 17:
            s = n // 2
            step(s + n % 2)
            step(s)
        else:
            s, rem = n // 3, n % 3
            step(s)
            step(s)
            step(s + rem)

    def batch_ntt_small(self, buf, l_skip, cnt_blocks, is_intt):
        if l_skip == 0 or cnt_blocks == 0:
            return
        self._ensure_small()
        self.check(self.L._batch_ntt_small(buf.data_ptr(), l_skip, cnt_blocks, is_intt, self.stream()), "_batch_ntt_small")

    def mle_interpolate_stages(self, buf, width, padded_height, log_blowup, start_log_step, end_log_step, is_eval_to_coeff,
                               right_pad):
        """src/poly.rs:162-247"""
        if start_log_step > end_log_step:
            return
        cur = start_log_step
        warp_end = min(end_log_step, LOG_WARP_SIZE - 1)
        warp_stages = max(warp_end - cur, 0) + 1  # saturating_sub(..) + 1
        if cur < LOG_WARP_SIZE and warp_stages >= 2:
            self.check(self.L._mle_interpolate_fused_2d(buf.data_ptr(), width, padded_height, log_blowup, 1 << cur, warp_stages,
                                                        is_eval_to_coeff, right_pad, self.stream()), "_mle_interpolate_fused_2d")
            cur = warp_end + 1
        if cur > end_log_step:
            return
        if cur < MLE_SHARED_TILE_LOG_SIZE:
            shared_end = min(end_log_step, MLE_SHARED_TILE_LOG_SIZE - 1)
            self.check(self.L._mle_interpolate_shared_2d(buf.data_ptr(), width, padded_height, log_blowup, cur, shared_end,
                                                         is_eval_to_coeff, right_pad, self.stream()), "_mle_interpolate_shared_2d")
            cur = shared_end + 1
        assert cur > end_log_step or not right_pad
        height = padded_height >> log_blowup
        while cur <= end_log_step:
            self.check(self.L._mle_interpolate_stage_2d(buf.data_ptr(), width, height, padded_height, 1 << cur, is_eval_to_coeff,
                                                        self.stream()), "_mle_interpolate_stage_2d")
            cur += 1

    # -- rs_code_matrix (src/stacked_pcs.rs:229-337, the `stacked_matrix == None` default branch) --------------
    def rs_code_matrix(self, stacked, height, width, l_skip, log_blowup, out=None):
        """stacked: device tensor, column-major height x width evaluations (already stacked).  Returns the
        (height << log_blowup) x width codeword.  The reference stacks the traces straight into the expanded
        buffer with memcpys (`stack_traces_into_expanded`, :143-220); for one full-height trace that is
        memset + one strided copy, which `_batch_expand_pad` performs here (its other branch uses the same launcher)."""
        cw_h = height << log_blowup
        cw = out if out is not None else self.empty(cw_h * width)
        self.check(self.L._batch_expand_pad(cw.data_ptr(), stacked.data_ptr(), width, cw_h, height, self.stream()),
                   "_batch_expand_pad")
        if l_skip > 0:
            self.batch_ntt_small(cw, l_skip, width * (cw_h >> l_skip), True)
            self.mle_interpolate_stages(cw, width, cw_h, log_blowup, 0, l_skip - 1, False, False)
        log_cw = cw_h.bit_length() - 1
        self.bit_rev(cw, log_cw, cw_h, width)
        self.batch_ntt(cw, log_cw, 0, width, False, False)
        return cw

    # -- Merkle tree (src/merkle_tree.rs:140-197) ------------------------------------------------------------
    def merkle_tree(self, matrix, height, width, rows_per_query):
        """Returns the list of digest layers (device tensors of 8-word digests), query layer first, root last."""
        k = rows_per_query.bit_length() - 1
        query_stride = height // rows_per_query
        layer = self.empty(query_stride * 8)
        self.check(self.L._poseidon2_compressing_row_hashes(layer.data_ptr(), matrix.data_ptr(), width, query_stride, k,
                                                            self.stream()), "_poseidon2_compressing_row_hashes")
        layers = [layer]
        while layers[-1].numel() // 8 > 1:
            prev = layers[-1]
            n = prev.numel() // 8 // 2
            nxt = self.empty(n * 8)
            self.check(self.L._poseidon2_adjacent_compress_layer(nxt.data_ptr(), prev.data_ptr(), n, self.stream()),
                       "_poseidon2_adjacent_compress_layer")
            layers.append(nxt)
        return layers

    # -- grind (src/sponge.rs:267-300, cuda/src/sponge.cu:65-117) -------------------------------------------------
    def sponge_grind(self, state18, bits, min_w, max_w):
        """One launch of the reference grind kernel over [min_w, max_w]; returns *a* valid witness or None
        (the kernel keeps the first thread to find one, not necessarily the smallest)."""
        st = self.h2d(np.asarray(state18, dtype=np.uint32))
        res = self.h2d(np.array([0xFFFFFFFF], dtype=np.uint32))
        # the launcher sizes its grid as 2^bits threads starting at min_w
        self.check(self.L._sponge_grind(st.data_ptr(), bits, min_w, max_w, res.data_ptr(), self.stream()), "_sponge_grind")
        w = int(self.d2h(res)[0])
        return None if w == 0xFFFFFFFF else w

    # -- GKR fraction tree (cuda/src/logup_zerocheck/gkr.cu:1228-1258) -----------------------------------------------
    def frac_build_tree_layer(self, layer, layer_size, real_len, logical_len, revert, alpha, apply_alpha):
        self.check(self.L._frac_build_tree_layer(layer.data_ptr(), layer_size, real_len, logical_len, revert, fpext(alpha),
                                                 apply_alpha, self.stream()), "_frac_build_tree_layer")

    # -- WHIR fold (cuda/src/whir.cu:278-290) ----------------------------------------------------------------------
    def whir_fold_coeffs_and_moments(self, f, w, alpha, height):
        f2, w2 = self.empty(height // 2 * 4), self.empty(height // 2 * 4)
        self.check(self.L._whir_fold_coeffs_and_moments(f.data_ptr(), w.data_ptr(), f2.data_ptr(), w2.data_ptr(), fpext(alpha),
                                                        height, self.stream()), "_whir_fold_coeffs_and_moments")
        return f2, w2
