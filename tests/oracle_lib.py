"""ctypes wrapper of the CPU oracle (oracle/capi.cpp).  Test infrastructure only."""
import ctypes as C

import numpy as np

P = 0x78000001
_vp, _sz, _i, _u32, _u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint32, C.c_uint64


def _p(a):
    return a.ctypes.data_as(_vp)


class Oracle:
    def __init__(self, path):
        L = self.L = C.CDLL(path)
        for n in ("orc_from_canonical", "orc_to_canonical", "orc_f_inv", "orc_two_adic_generator", "orc_sponge_grind"):
            getattr(L, n).restype = _u32
        for n in ("orc_f_add", "orc_f_sub", "orc_f_mul"):
            getattr(L, n).restype = _u32
            getattr(L, n).argtypes = [_u32, _u32]
        L.orc_sponge_sample_bits.restype = _u64
        L.orc_poseidon2_permute.argtypes = [_vp, _sz]
        L.orc_hash_slice.argtypes = [_vp, _vp, _sz]
        L.orc_dft.argtypes = [_vp, _sz, _i]
        L.orc_dft_batch.argtypes = [_vp, _sz, _sz, _i]
        L.orc_coset_dft.argtypes = [_vp, _sz, _u32]
        L.orc_stacked_layout.argtypes = [_i, _i, _sz, _vp, _vp, C.POINTER(_u64), C.POINTER(_u64), _vp]
        L.orc_stacked_matrix.argtypes = [_i, _i, _sz, _vp, _vp, _vp, C.POINTER(_u64), _vp]
        L.orc_eval_to_coeff_rs_message.argtypes = [_i, _vp, _sz]
        L.orc_rs_code_matrix.argtypes = [_i, _i, _vp, _sz, _sz, _vp]
        L.orc_merkle_tree.argtypes = [_vp, _sz, _sz, _sz, _vp]
        L.orc_stacked_commit.argtypes = [_i, _i, _i, _i, _sz, _vp, _vp, _vp, _vp, C.POINTER(_u64), _vp, _vp]
        L.orc_sponge_observe.argtypes = [_vp, _vp, _sz]
        L.orc_sponge_sample.argtypes = [_vp, _vp, _sz]
        L.orc_sponge_sample_bits.argtypes = [_vp, _i]
        L.orc_sponge_check_witness.argtypes = [_vp, _i, _u32]
        L.orc_sponge_grind.argtypes = [_vp, _i, _u32]
        L.orc_from_canonical_vec.argtypes = [_vp, _sz]
        L.orc_to_canonical_vec.argtypes = [_vp, _sz]
        L.orc_ef_mul.argtypes = [_vp, _vp, _vp]
        L.orc_ef_inv.argtypes = [_vp, _vp]
        L.orc_gkr_prove.argtypes = [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp]
        L.orc_gkr_verify.argtypes = [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]
        L.orc_eval_mle_evals_at_point.argtypes = [_vp, _i, _vp, _vp]
        L.orc_stacked_reduction_proof_words.restype = _sz
        L.orc_stacked_reduction_proof_words.argtypes = [_i, _i, _sz, _vp]
        L.orc_stacked_reduction_prove.argtypes = [_vp, _i, _i, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _vp, _vp]
        L.orc_column_opening.argtypes = [_i, _vp, _sz, _i, _vp, _vp]
        L.orc_stacked_reduction_verify.argtypes = [_vp, _i, _i, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _vp]
        L.orc_bc_proof_words.restype = _sz
        L.orc_bc_proof_words.argtypes = [_i, _i, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
        L.orc_bc_prove.argtypes = [_vp, _i, _i, _i, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
        L.orc_bc_verify.argtypes = [_vp, _i, _i, _i, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
        L.orc_whir_proof_words.restype = _sz
        L.orc_whir_proof_words.argtypes = [_i, _i, _i, _i, _vp, _sz, _vp]
        L.orc_whir_prove.argtypes = [_vp, _i, _i, _i, _i, _vp, _i, _i, _i, _sz, _vp, _vp, _sz, _vp, _vp, _vp]
        L.orc_whir_stacking_openings.argtypes = [_i, _vp, _sz, _sz, _vp, _vp]
        L.orc_whir_verify.argtypes = [_vp, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _vp, _sz, _sz, _vp, _vp, _vp, _vp]

    # ---- field ----
    def to_mont(self, x):
        a = np.ascontiguousarray(np.asarray(x, dtype=np.uint64) % P, dtype=np.uint32).copy()
        self.L.orc_from_canonical_vec(_p(a), a.size)
        return a

    def from_mont(self, m):
        a = np.ascontiguousarray(m, dtype=np.uint32).copy()
        self.L.orc_to_canonical_vec(_p(a), a.size)
        return a

    def random_field(self, rng, shape):
        """uniform canonical values -> Montgomery words"""
        return self.to_mont(rng.integers(0, P, size=shape, dtype=np.uint64)).reshape(shape)

    def ef_mul(self, a, b):
        a, b = np.ascontiguousarray(a, np.uint32), np.ascontiguousarray(b, np.uint32)
        out = np.zeros(4, np.uint32)
        self.L.orc_ef_mul(_p(out), _p(a), _p(b))
        return out

    def binary_k_fold(self, values, alphas, x):
        """verifier/whir.rs:352-389: values (2^k, 4), alphas (k, 4), x a base-field Montgomery word."""
        values, alphas = np.ascontiguousarray(values, np.uint32), np.ascontiguousarray(alphas, np.uint32)
        out = np.zeros(4, np.uint32)
        self.L.orc_binary_k_fold(_p(out), _p(values), C.c_int(len(alphas)), _p(alphas), C.c_uint32(int(x)))
        return out

    def ef_inv(self, a):
        a = np.ascontiguousarray(a, np.uint32)
        out = np.zeros(4, np.uint32)
        self.L.orc_ef_inv(_p(out), _p(a))
        return out

    # ---- poseidon2 ----
    def permute(self, states):
        s = np.ascontiguousarray(states, dtype=np.uint32).copy()
        self.L.orc_poseidon2_permute(_p(s), s.size // 16)
        return s

    def hash_slice(self, vals):
        v = np.ascontiguousarray(vals, dtype=np.uint32)
        out = np.zeros(8, np.uint32)
        self.L.orc_hash_slice(_p(out), _p(v), v.size)
        return out

    def compress(self, l, r):
        l, r = np.ascontiguousarray(l, np.uint32), np.ascontiguousarray(r, np.uint32)
        out = np.zeros(8, np.uint32)
        self.L.orc_compress(_p(out), _p(l), _p(r))
        return out

    # ---- dft ----
    def dft(self, a, inverse=False):
        v = np.ascontiguousarray(a, dtype=np.uint32).copy()
        assert self.L.orc_dft(_p(v), v.size, 1 if inverse else 0) == 0
        return v

    def dft_batch(self, a, n, cols, inverse=False):
        v = np.ascontiguousarray(a, dtype=np.uint32).copy()
        assert self.L.orc_dft_batch(_p(v), n, cols, 1 if inverse else 0) == 0
        return v

    def coset_dft(self, a, shift):
        v = np.ascontiguousarray(a, dtype=np.uint32).copy()
        assert self.L.orc_coset_dft(_p(v), v.size, int(shift)) == 0
        return v

    # ---- stacking / rs / merkle ----
    def stacked_layout(self, l_skip, log_h, meta):
        n = len(meta)
        w = np.array([m[0] for m in meta] or [0], dtype=np.uint64)
        lh = np.array([m[1] for m in meta] or [0], dtype=np.int32)
        ow, on = _u64(), _u64()
        rc = self.L.orc_stacked_layout(l_skip, log_h, n, _p(w), _p(lh), C.byref(ow), C.byref(on), None)
        if rc:
            return None
        buf = np.zeros(max(5 * on.value, 1), dtype=np.uint64)
        self.L.orc_stacked_layout(l_skip, log_h, n, _p(w), _p(lh), C.byref(ow), C.byref(on), _p(buf))
        return int(ow.value), buf[: 5 * on.value].reshape(-1, 5)

    @staticmethod
    def _trace_args(traces):
        n = len(traces)
        arrs = [np.ascontiguousarray(t[0], dtype=np.uint32) for t in traces]
        ptrs = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in arrs])
        hs = np.array([t[1] for t in traces] or [0], dtype=np.uint64)
        ws = np.array([t[2] for t in traces] or [0], dtype=np.uint64)
        return arrs, ptrs, hs, ws

    def stacked_matrix(self, l_skip, n_stack, traces):
        """traces: list of (flat col-major uint32 array, height, width). Returns (flat, width)."""
        arrs, ptrs, hs, ws = self._trace_args(traces)
        ow = _u64()
        assert self.L.orc_stacked_matrix(l_skip, n_stack, len(traces), ptrs, _p(hs), _p(ws), C.byref(ow), None) == 0
        out = np.zeros((1 << (l_skip + n_stack)) * ow.value, dtype=np.uint32)
        self.L.orc_stacked_matrix(l_skip, n_stack, len(traces), ptrs, _p(hs), _p(ws), C.byref(ow), _p(out))
        return out, int(ow.value)

    def eval_to_coeff_rs_message(self, l_skip, a):
        v = np.ascontiguousarray(a, dtype=np.uint32).copy()
        assert self.L.orc_eval_to_coeff_rs_message(l_skip, _p(v), v.size) == 0
        return v

    def rs_code_matrix(self, l_skip, log_blowup, evals, height, width):
        e = np.ascontiguousarray(evals, dtype=np.uint32)
        out = np.zeros((height << log_blowup) * width, dtype=np.uint32)
        assert self.L.orc_rs_code_matrix(l_skip, log_blowup, _p(e), height, width, _p(out)) == 0
        return out

    def merkle_tree(self, matrix, height, width, rows_per_query):
        """returns list of layers [(n,8) ...] or None on a parameter error"""
        m = np.ascontiguousarray(matrix, dtype=np.uint32)
        leaves = 1
        while leaves < height:
            leaves <<= 1
        if rows_per_query > leaves or height == 0:
            return None
        qs = leaves // rows_per_query
        out = np.zeros((2 * qs - 1) * 8, dtype=np.uint32)
        if self.L.orc_merkle_tree(_p(m), height, width, rows_per_query, _p(out)) != 0:
            return None
        return split_layers(out, qs)

    def stacked_commit(self, l_skip, n_stack, log_blowup, k_whir, traces, want_codeword=True):
        arrs, ptrs, hs, ws = self._trace_args(traces)
        H = 1 << (l_skip + n_stack)
        cells = sum(max(t[1], 1 << l_skip) * t[2] for t in traces)
        W = (cells + H - 1) // H
        N = H << log_blowup
        qs = N >> k_whir
        root = np.zeros(8, np.uint32)
        cw = np.zeros(N * W, np.uint32) if want_codeword else None
        layers = np.zeros((2 * qs - 1) * 8, np.uint32)
        ow = _u64()
        rc = self.L.orc_stacked_commit(
            l_skip, n_stack, log_blowup, k_whir, len(traces), ptrs, _p(hs), _p(ws), _p(root), C.byref(ow),
            _p(cw) if want_codeword else None, _p(layers),
        )
        assert rc == 0
        assert ow.value == W
        return root, cw, split_layers(layers, qs), W

    # ---- transcript ----
    def sponge_new(self):
        return np.zeros(18, dtype=np.uint32)

    def sponge_observe(self, st, vals):
        v = np.ascontiguousarray(vals, dtype=np.uint32)
        self.L.orc_sponge_observe(_p(st), _p(v), v.size)

    def sponge_sample(self, st, n=1):
        out = np.zeros(n, np.uint32)
        self.L.orc_sponge_sample(_p(st), _p(out), n)
        return out

    def sponge_sample_bits(self, st, bits):
        return int(self.L.orc_sponge_sample_bits(_p(st), bits))

    def sponge_check_witness(self, st, bits, w_mont):
        return bool(self.L.orc_sponge_check_witness(_p(st), bits, int(w_mont)))

    def sponge_grind(self, st, bits, start=0):
        return int(self.L.orc_sponge_grind(_p(st), bits, start))


    # ---- LogUp-GKR ----
    def gkr_prove(self, sponge, leaves, log_n, assert_zero):
        """leaves: uint32[2^log_n * 8] (Frac<EF> = p[4], q[4]).  Returns dict or raises
        ValueError('NonZeroRootSum').  `sponge` (uint32[18]) is advanced in place."""
        leaves = np.ascontiguousarray(leaves, np.uint32)
        n_polys = log_n * (log_n - 1) // 2
        out = dict(frac_sum=np.zeros(8, np.uint32), claims=np.zeros((log_n, 16), np.uint32),
                   polys=np.zeros((max(n_polys, 1), 12), np.uint32), xi=np.zeros((log_n, 4), np.uint32))
        rc = self.L.orc_gkr_prove(_p(sponge), _p(leaves), log_n, int(assert_zero), _p(out["frac_sum"]),
                                  _p(out["claims"]), _p(out["polys"]), _p(out["xi"]))
        if rc == 2:
            raise ValueError("NonZeroRootSum")
        assert rc == 0
        out["polys"] = out["polys"][:n_polys]
        return out

    def gkr_verify(self, sponge, log_n, proof):
        numer, denom = np.zeros(4, np.uint32), np.zeros(4, np.uint32)
        xi = np.zeros((log_n, 4), np.uint32)
        polys = np.ascontiguousarray(proof["polys"] if len(proof["polys"]) else np.zeros((1, 12)), np.uint32)
        ok = self.L.orc_gkr_verify(_p(sponge), log_n, _p(np.ascontiguousarray(proof["frac_sum"], np.uint32)),
                                   _p(np.ascontiguousarray(proof["claims"], np.uint32)), _p(polys), _p(numer), _p(denom), _p(xi))
        return bool(ok), numer, denom, xi

    def eval_mle_evals_at_point(self, evals, n, x):
        out = np.zeros(4, np.uint32)
        self.L.orc_eval_mle_evals_at_point(_p(np.ascontiguousarray(evals, np.uint32)), n,
                                           _p(np.ascontiguousarray(x, np.uint32)), _p(out))
        return out

    # ---- WHIR ----
    def whir_proof_words(self, m, log_blowup, cfg, widths):
        nq = np.asarray(cfg["num_queries"], np.int32)
        w = np.asarray(widths, np.uint64)
        return int(self.L.orc_whir_proof_words(m, log_blowup, cfg["k"], len(nq), _p(nq), len(w), _p(w)))

    def whir_prove(self, sponge, l_skip, log_blowup, cfg, mats, height, u):
        """mats: list of (uint32 col-major flat, width).  Returns (roots[n,8], proof words)."""
        nq = np.asarray(cfg["num_queries"], np.int32)
        widths = np.asarray([w for _, w in mats], np.uint64)
        arrs = [np.ascontiguousarray(v, np.uint32) for v, _ in mats]
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        m = height.bit_length() - 1
        proof = np.zeros(self.whir_proof_words(m, log_blowup, cfg, widths), np.uint32)
        roots = np.zeros((len(arrs), 8), np.uint32)
        u = np.ascontiguousarray(u, np.uint32)
        rc = self.L.orc_whir_prove(_p(sponge), l_skip, log_blowup, cfg["k"], len(nq), _p(nq), cfg["mu_pow_bits"],
                                   cfg["query_phase_pow_bits"], cfg["folding_pow_bits"], len(arrs), ptrs, _p(widths),
                                   height, _p(u), _p(roots), _p(proof))
        assert rc == 0
        return roots, proof

    def whir_stacking_openings(self, l_skip, mat, height, width, u):
        out = np.zeros((width, 4), np.uint32)
        self.L.orc_whir_stacking_openings(l_skip, _p(np.ascontiguousarray(mat, np.uint32)), height, width,
                                          _p(np.ascontiguousarray(u, np.uint32)), _p(out))
        return out

    def whir_verify(self, sponge, l_skip, n_stack, log_blowup, cfg, proof, widths, openings, roots, u):
        nq = np.asarray(cfg["num_queries"], np.int32)
        w = np.asarray(widths, np.uint64)
        proof = np.ascontiguousarray(proof, np.uint32)
        return bool(self.L.orc_whir_verify(_p(sponge), l_skip, n_stack, log_blowup, cfg["k"], len(nq), _p(nq),
                                           cfg["mu_pow_bits"], cfg["query_phase_pow_bits"], cfg["folding_pow_bits"],
                                           _p(proof), proof.size, len(w), _p(w),
                                           _p(np.ascontiguousarray(openings, np.uint32)),
                                           _p(np.ascontiguousarray(roots, np.uint32)), _p(np.ascontiguousarray(u, np.uint32))))

    # ---- stacked opening reduction ----
    @staticmethod
    def _commit_meta(commits):
        """commits: list of lists of (vals, height, width, need_rot)."""
        off, hs, ws, rot, arrs = [0], [], [], [], []
        for traces in commits:
            for v, h, w, nr in traces:
                arrs.append(np.ascontiguousarray(v, np.uint32))  # may be empty when only shapes are needed (verifier)
                hs.append(h)
                ws.append(w)
                rot.append(1 if nr else 0)
            off.append(len(hs))
        return (np.asarray(off, np.uint64), arrs, np.asarray(hs, np.uint64), np.asarray(ws, np.uint64),
                np.asarray(rot, np.uint8))

    def stacked_reduction_prove(self, sponge, l_skip, n_stack, commits, r):
        off, arrs, hs, ws, rot = self._commit_meta(commits)
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        r = np.ascontiguousarray(r, np.uint32)
        sw = np.zeros(len(commits), np.uint64)
        # widths of the stacked matrices are needed for the proof size: ceil(sum cells / H)
        H = 1 << (l_skip + n_stack)
        for c, traces in enumerate(commits):
            cells = sum(max(h, 1 << l_skip) * w for _, h, w, _ in traces)
            sw[c] = (cells + H - 1) // H
        n = int(self.L.orc_stacked_reduction_proof_words(l_skip, n_stack, len(commits), _p(sw)))
        proof = np.zeros(n, np.uint32)
        u = np.zeros((n_stack + 1, 4), np.uint32)
        rc = self.L.orc_stacked_reduction_prove(_p(sponge), l_skip, n_stack, len(commits), _p(off), ptrs, _p(hs), _p(ws),
                                                _p(rot), _p(r), r.size // 4, _p(sw), _p(proof), _p(u))
        assert rc == 0
        return proof, u, sw

    def column_opening(self, l_skip, col, is_rot, r):
        out = np.zeros(4, np.uint32)
        col = np.ascontiguousarray(col, np.uint32)
        self.L.orc_column_opening(l_skip, _p(col), col.size, int(is_rot), _p(np.ascontiguousarray(r, np.uint32)), _p(out))
        return out

    def stacked_reduction_verify(self, sponge, l_skip, n_stack, commits, t_claims, r, proof):
        off, _, hs, ws, rot = self._commit_meta(commits)
        u = np.zeros((n_stack + 1, 4), np.uint32)
        r = np.ascontiguousarray(r, np.uint32)
        ok = self.L.orc_stacked_reduction_verify(_p(sponge), l_skip, n_stack, len(commits), _p(off), _p(hs), _p(ws), _p(rot),
                                                 _p(np.ascontiguousarray(t_claims, np.uint32)), _p(r), r.size // 4,
                                                 _p(np.ascontiguousarray(proof, np.uint32)), _p(u))
        return bool(ok), u

    # ---- batch constraints (LogUp-GKR + zerocheck) ----
    @staticmethod
    def _bc_args(flat):
        def pp(a):
            return _p(a) if a.size else None
        hs = np.array([m[1] for m in flat["mats"]], np.uint64)
        ws = np.array([m[2] for m in flat["mats"]], np.uint64)
        arrs = [np.ascontiguousarray(m[0], np.uint32) for m in flat["mats"]]  # unused (may be empty) for the verifier
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        return (pp(flat["meta"]), pp(flat["nodes"]), pp(flat["cidx"]), pp(flat["inter"]), pp(flat["msg"]), pp(flat["pubs"])), ptrs, hs, ws, arrs

    def bc_proof_words(self, l_skip, D, airs_flat, n_airs):
        a, _, hs, ws, _ = self._bc_args(airs_flat)
        return int(self.L.orc_bc_proof_words(l_skip, D, n_airs, *a, _p(hs), _p(ws)))

    def bc_prove(self, sponge, l_skip, D, logup_pow_bits, airs_flat, n_airs, n_max):
        a, ptrs, hs, ws, keep = self._bc_args(airs_flat)
        n = self.bc_proof_words(l_skip, D, airs_flat, n_airs)
        proof = np.zeros(n, np.uint32)
        r = np.zeros((n_max + 1, 4), np.uint32)
        rc = self.L.orc_bc_prove(_p(sponge), l_skip, D, logup_pow_bits, n_airs, *a, ptrs, _p(hs), _p(ws), _p(proof), _p(r))
        if rc == 2:
            raise ValueError("NonZeroRootSum")
        assert rc == 0
        return proof, r

    def bc_verify(self, sponge, l_skip, D, logup_pow_bits, airs_flat, n_airs, n_max, proof):
        a, _, hs, ws, _ = self._bc_args(airs_flat)
        r = np.zeros((n_max + 1, 4), np.uint32)
        ok = self.L.orc_bc_verify(_p(sponge), l_skip, D, logup_pow_bits, n_airs, *a, _p(hs), _p(ws),
                                  _p(np.ascontiguousarray(proof, np.uint32)), _p(r))
        return bool(ok), r

def split_layers(flat, qs):
    out, off, n = [], 0, qs
    while n >= 1:
        out.append(flat[off * 8 : (off + n) * 8].reshape(n, 8).copy())
        off += n
        n >>= 1
    return out
