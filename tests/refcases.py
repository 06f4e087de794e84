"""Seeded cases shared by tools/gen_reference_goldens.py (which runs the REFERENCE's CUDA kernels on a GPU box and
freezes their outputs into tests/golden/reference_gpu_kats.json), tests/test_reference_goldens.py (CPU: oracle vs the
frozen reference outputs) and tests/test_reference_kernels.py (GPU: reference kernels == oracle == product, live).

Inputs are regenerated from the seed (numpy PCG64, canonical values uniform in [0, p), converted to Montgomery form by
the consumer); outputs are stored whole when small and as SHA-256 of the little-endian u32 words otherwise.
"""
import hashlib

import numpy as np

P = 0x78000001

# (width, height, rows_per_query): row hashes (PaddingFreeSponge 16/8/8) + strided levels + adjacent compress layers
MERKLE_CASES = [(1, 8, 1), (7, 8, 2), (8, 16, 1), (9, 16, 4), (16, 32, 8), (17, 64, 16), (64, 64, 4), (256, 32, 16),
                (3, 1024, 16), (40, 4096, 16)]
# (l_skip, log_h, log_blowup, width): rs_code_matrix
RS_CASES = [(1, 1, 1, 2), (2, 2, 1, 3), (2, 5, 2, 9), (3, 6, 1, 1), (4, 4, 3, 2), (4, 10, 1, 5), (4, 12, 1, 19), (5, 13, 1, 4),
            (6, 14, 2, 3), (1, 16, 1, 5), (4, 16, 1, 8), (9, 12, 1, 2), (4, 17, 1, 3), (4, 19, 1, 2)]
# (log_n, cols, inverse): batch_ntt natural -> natural (bit_rev + CT passes, src/ntt.rs:111-168)
NTT_CASES = [(1, 3, False), (2, 3, False), (5, 7, False), (6, 2, True), (9, 5, False), (10, 4, True), (11, 3, False),
             (13, 2, True), (16, 3, False), (17, 2, False), (18, 2, False), (20, 1, True), (21, 1, False)]
# (seed, absorb_idx, sample_idx, bits): grind from an arbitrary sponge state
GRIND_CASES = [(1, 0, 0, 1), (2, 3, 8, 5), (3, 7, 2, 8), (4, 7, 8, 12), (5, 0, 5, 15), (6, 5, 0, 16), (7, 7, 0, 18)]
# (log_n): one fraction-tree layer (non-virtual) / one WHIR coefficient+moment fold
EF_CASES = [1, 4, 9]


def canonical(seed, n):
    return np.random.default_rng(seed).integers(0, P, size=n, dtype=np.uint64).astype(np.uint32)


def sha(words):
    return hashlib.sha256(np.ascontiguousarray(words, dtype="<u4").tobytes()).hexdigest()


def digest_or_words(words, limit=512):
    w = np.ascontiguousarray(words, dtype=np.uint32).reshape(-1)
    return {"sha256": sha(w), "n": int(w.size), **({"words": [int(x) for x in w]} if w.size <= limit else {})}


def merkle_seed(i):
    return 5000 + i


def rs_seed(i):
    return 6000 + i


def ntt_seed(i):
    return 7000 + i


def ef_seed(i):
    return 8000 + i


# ---------------------------------------------------------------------------------------------------------
# running the cases
# ---------------------------------------------------------------------------------------------------------
def merkle_inputs(i, to_mont):
    w, h, _ = MERKLE_CASES[i]
    return to_mont(canonical(merkle_seed(i), w * h))


def rs_inputs(i, to_mont):
    _, log_h, _, w = RS_CASES[i]
    return to_mont(canonical(rs_seed(i), w << log_h))


def ntt_inputs(i, to_mont):
    log_n, cols, _ = NTT_CASES[i]
    return to_mont(canonical(ntt_seed(i), cols << log_n))


def grind_state(i, to_mont):
    seed, a, s, _ = GRIND_CASES[i]
    st = np.zeros(18, np.uint32)
    st[:16] = to_mont(canonical(9000 + seed, 16))
    st[16], st[17] = a, s
    return st


def ef_inputs(i, to_mont):
    n = 1 << EF_CASES[i]
    return to_mont(canonical(ef_seed(i), n * 8)), to_mont(canonical(ef_seed(i) + 50, n * 4)), \
        to_mont(canonical(ef_seed(i) + 99, 4))


def reference_outputs(rk, to_mont):
    """Runs every case through the reference's CUDA kernels (rk: ref_kernels.RefKernels).  Returns
    {family: [np.uint32 array or int, ...]} in Montgomery words, exactly as the kernels wrote them."""
    out = {"merkle": [], "rs": [], "ntt": [], "grind": [], "frac_layer": [], "whir_fold": []}
    for i, (w, h, rpq) in enumerate(MERKLE_CASES):
        layers = rk.merkle_tree(rk.h2d(merkle_inputs(i, to_mont)), h, w, rpq)
        out["merkle"].append(np.concatenate([rk.d2h(l) for l in layers]))
    for i, (l_skip, log_h, lb, w) in enumerate(RS_CASES):
        out["rs"].append(rk.d2h(rk.rs_code_matrix(rk.h2d(rs_inputs(i, to_mont)), 1 << log_h, w, l_skip, lb)))
    for i, (log_n, cols, inv) in enumerate(NTT_CASES):
        buf = rk.h2d(ntt_inputs(i, to_mont))
        rk.batch_ntt(buf, log_n, 0, cols, True, inv)
        out["ntt"].append(rk.d2h(buf))
    for i, (_, _, _, bits) in enumerate(GRIND_CASES):
        out["grind"].append(reference_min_witness(rk, grind_state(i, to_mont), bits))
    for i, log_n in enumerate(EF_CASES):
        fr, f4, alpha = ef_inputs(i, to_mont)
        n = 1 << log_n
        layer = rk.h2d(fr)
        rk.frac_build_tree_layer(layer, n, n, n, False, alpha, False)
        out["frac_layer"].append(rk.d2h(layer)[: n // 2 * 8].copy())
        f2, w2 = rk.whir_fold_coeffs_and_moments(rk.h2d(f4), rk.h2d(fr[: n * 4]), alpha, n)
        out["whir_fold"].append(np.concatenate([rk.d2h(f2), rk.d2h(w2)]))
    return out


def reference_min_witness(rk, state18, bits, limit=1 << 24):
    """Smallest witness the reference's grind kernel accepts: its launcher covers 2^bits candidates from min_witness
    (cuda/src/sponge.cu:101) and keeps whichever thread wins the atomicCAS, so walk the windows in order and, inside the
    first window that has a solution, bisect on max_witness."""
    span = 1 << bits
    start = 0
    while start < limit:
        hi = start + span - 1
        if rk.sponge_grind(state18, bits, start, hi) is not None:
            lo = start
            while lo < hi:  # smallest max_witness for which a solution exists in [start, max_witness]
                mid = (lo + hi) // 2
                if rk.sponge_grind(state18, bits, start, mid) is not None:
                    hi = mid
                else:
                    lo = mid + 1
            assert rk.sponge_grind(state18, bits, lo, lo) == lo
            return lo
        start += span
    return None
