"""CPU tests: pin the oracle against every golden vector / KAT the reference holds for the commit
path, and against independent pure-Python restatements at small sizes."""
import json
import os

import numpy as np
import pytest

P = 0x78000001
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))


# ---------------------------------------------------------------------------------------------
# field
# ---------------------------------------------------------------------------------------------
def test_two_adic_generators_match_reference_table(oracle):
    # fp.h:291-320 holds canonical and Montgomery values of two_adic_generator(k), k = 0..27
    for k, (canon, monty) in enumerate(zip(GOLD["two_adic_generators_canonical"], GOLD["two_adic_generators_monty"])):
        g = oracle.L.orc_two_adic_generator(k)
        assert g == monty
        assert oracle.L.orc_to_canonical(g) == canon
        assert pow(canon, 1 << k, P) == 1 and (k == 0 or pow(canon, 1 << (k - 1), P) != 1)


def test_montgomery_constants(oracle):
    assert oracle.L.orc_from_canonical(1) == GOLD["MONTY_ONE"]
    assert (1 << 64) % P == GOLD["R2"]
    assert GOLD["P"] == P


def test_field_ops_vs_python(oracle):
    rng = np.random.default_rng(1)
    a = rng.integers(0, P, 2000, dtype=np.uint64)
    b = rng.integers(0, P, 2000, dtype=np.uint64)
    am, bm = oracle.to_mont(a), oracle.to_mont(b)
    assert np.array_equal(oracle.from_mont(am), a.astype(np.uint32))
    for x, y, xm, ym in zip(a[:300], b[:300], am[:300], bm[:300]):
        x, y = int(x), int(y)
        assert oracle.L.orc_to_canonical(oracle.L.orc_f_mul(int(xm), int(ym))) == x * y % P
        assert oracle.L.orc_to_canonical(oracle.L.orc_f_add(int(xm), int(ym))) == (x + y) % P
        assert oracle.L.orc_to_canonical(oracle.L.orc_f_sub(int(xm), int(ym))) == (x - y) % P
        if x:
            assert oracle.L.orc_to_canonical(oracle.L.orc_f_inv(int(xm))) == pow(x, P - 2, P)
    assert oracle.L.orc_f_inv(0) == 0


def _ef_mul_py(a, b):
    t = [0] * 7
    for i in range(4):
        for j in range(4):
            t[i + j] += a[i] * b[j]
    return [(t[0] + 11 * t[4]) % P, (t[1] + 11 * t[5]) % P, (t[2] + 11 * t[6]) % P, t[3] % P]


def test_ext_field(oracle):
    rng = np.random.default_rng(2)
    one = [1, 0, 0, 0]
    for _ in range(50):
        a = [int(x) for x in rng.integers(0, P, 4)]
        b = [int(x) for x in rng.integers(0, P, 4)]
        got = oracle.from_mont(oracle.ef_mul(oracle.to_mont(a), oracle.to_mont(b)))
        assert list(got) == _ef_mul_py(a, b)
        inv = oracle.ef_inv(oracle.to_mont(a))
        assert list(oracle.from_mont(oracle.ef_mul(oracle.to_mont(a), inv))) == one


# ---------------------------------------------------------------------------------------------
# Poseidon2: dense-matrix spec form in pure Python vs the oracle's optimised form
# ---------------------------------------------------------------------------------------------
def _poseidon2_spec(state, rc):
    M4 = [[2, 3, 1, 1], [1, 2, 3, 1], [1, 1, 2, 3], [3, 1, 1, 2]]
    ME = [[(2 if i // 4 == j // 4 else 1) * M4[i % 4][j % 4] for j in range(16)] for i in range(16)]
    diag = GOLD["poseidon2_internal_diag_canonical"]
    MI = [[(1 + diag[i]) % P if i == j else 1 for j in range(16)] for i in range(16)]
    mat = lambda M, s: [sum(M[i][j] * s[j] for j in range(16)) % P for i in range(16)]
    s = mat(ME, list(state))
    for r in range(4):
        s = mat(ME, [pow((s[i] + rc["init"][r * 16 + i]) % P, 7, P) for i in range(16)])
    for r in range(13):
        s[0] = pow((s[0] + rc["int"][r]) % P, 7, P)
        s = mat(MI, s)
    for r in range(4):
        s = mat(ME, [pow((s[i] + rc["term"][r * 16 + i]) % P, 7, P) for i in range(16)])
    return s


def _round_constants():
    import re

    src = open(os.path.join(os.path.dirname(__file__), "..", "oracle", "poseidon2_rc.inc")).read()
    g = lambda n: [int(x) for x in re.findall(r"(\d+)u", re.search(n + r"\[\d+\] = \{(.*?)\};", src, re.S).group(1))]
    return {"init": g("P2_RC_EXT_INITIAL"), "int": g("P2_RC_INTERNAL"), "term": g("P2_RC_EXT_TERMINAL")}


def test_poseidon2_spec_form_equals_oracle(oracle):
    rc = _round_constants()
    rng = np.random.default_rng(3)
    cases = [np.zeros(16, np.uint64), np.arange(16, dtype=np.uint64)] + [rng.integers(0, P, 16, dtype=np.uint64) for _ in range(6)]
    for st in cases:
        got = oracle.from_mont(oracle.permute(oracle.to_mont(st)))
        assert list(got) == _poseidon2_spec([int(x) for x in st], rc)


def test_poseidon2_selfcheck_values(oracle):
    z = oracle.from_mont(oracle.permute(np.zeros(16, np.uint32)))
    assert list(z[:4]) == GOLD["poseidon2_selfcheck"]["zeros_first4"]
    i = oracle.from_mont(oracle.permute(oracle.to_mont(np.arange(16))))
    assert list(i[:4]) == GOLD["poseidon2_selfcheck"]["iota_first4"]


def test_hash_slice_and_compress_semantics(oracle):
    rng = np.random.default_rng(4)
    assert not oracle.hash_slice(np.zeros(0, np.uint32)).any()  # empty input: zeros, no permutation
    for n in (1, 7, 8, 9, 16, 19, 64):
        v = oracle.random_field(rng, n)
        st = np.zeros(16, np.uint32)
        for off in range(0, n, 8):
            chunk = v[off : off + 8]
            st[: len(chunk)] = chunk  # overwrite mode, no padding
            st = oracle.permute(st)
        assert np.array_equal(oracle.hash_slice(v), st[:8])
    l, r = oracle.random_field(rng, 8), oracle.random_field(rng, 8)
    assert np.array_equal(oracle.compress(l, r), oracle.permute(np.concatenate([l, r]))[:8])


# ---------------------------------------------------------------------------------------------
# stacking goldens (prover/stacked_pcs.rs:556-619, cuda-backend/src/stacked_pcs.rs:399-517)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", GOLD["stacking"], ids=lambda c: c["name"])
def test_stacked_matrix_golden(oracle, case):
    traces = [(oracle.to_mont(t["values"]), t["height"], t["width"]) for t in case["traces"]]
    flat, width = oracle.stacked_matrix(case["l_skip"], case["n_stack"], traces)
    assert width == case["width"] and flat.size == case["height"] * case["width"]
    got = oracle.from_mont(flat)
    if "expected" in case:
        assert list(got) == case["expected"]
    else:
        pre = case["expected_prefix"]
        assert list(got[: len(pre)]) == pre and not got[len(pre) :].any()


def test_stacked_layout_errors_and_shapes(oracle):
    # LayoutHeightExceeded
    assert oracle.stacked_layout(2, 4, [(1, 5)]) is None
    # unsorted heights overflow a column (LayoutRowOverflow): 3 cols of 2^3 then one of 2^4 in height 2^4
    assert oracle.stacked_layout(0, 4, [(1, 3), (1, 4)]) is None
    w, cols = oracle.stacked_layout(2, 4, [(3, 4), (0, 3), (5, 2), (2, 0)])
    assert w == 3 + 2  # 3 full columns, then 5*4 + 2*4 = 28 rows -> 2 columns
    assert list(cols[0]) == [0, 0, 0, 0, 4] and list(cols[3]) == [2, 0, 3, 0, 2]
    assert list(cols[-1]) == [3, 1, 4, 8, 0]


# ---------------------------------------------------------------------------------------------
# DFT / RS encoding against naive definitions
# ---------------------------------------------------------------------------------------------
def _naive_dft(c, n, inverse=False):
    k = n.bit_length() - 1
    w = GOLD["two_adic_generators_canonical"][k]
    if inverse:
        w = pow(w, P - 2, P)
    out = [sum(int(c[j]) * pow(w, i * j, P) for j in range(len(c))) % P for i in range(n)]
    if inverse:
        ninv = pow(n, P - 2, P)
        out = [x * ninv % P for x in out]
    return out


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 5, 7])
def test_dft_matches_definition(oracle, log_n):
    rng = np.random.default_rng(log_n)
    n = 1 << log_n
    c = rng.integers(0, P, n, dtype=np.uint64)
    assert list(oracle.from_mont(oracle.dft(oracle.to_mont(c)))) == _naive_dft(c, n)
    assert list(oracle.from_mont(oracle.dft(oracle.to_mont(c), inverse=True))) == _naive_dft(c, n, True)
    shift = 31
    got = oracle.from_mont(oracle.coset_dft(oracle.to_mont(c), int(oracle.to_mont([shift])[0])))
    w = GOLD["two_adic_generators_canonical"][log_n]
    want = [sum(int(c[j]) * pow(shift * pow(w, i, P) % P, j, P) for j in range(n)) % P for i in range(n)]
    assert list(got) == want


def _rs_message_py(l_skip, evals):
    # poly.rs:325-348: per chunk iDFT then coeffs_to_evals over the chunk's index bits
    ch = 1 << l_skip
    out = []
    for off in range(0, len(evals), ch):
        a = _naive_dft(evals[off : off + ch], ch, True)
        for b in range(l_skip):
            step = 1 << b
            for i in range(0, ch, 2 * step):
                for j in range(step):
                    a[i + j + step] = (a[i + j + step] + a[i + j]) % P
        out += a
    return out


@pytest.mark.parametrize("l_skip,log_h,log_blowup,width", [(0, 0, 0, 1), (0, 3, 1, 2), (2, 2, 1, 3), (2, 5, 2, 2), (3, 6, 1, 1), (4, 4, 3, 2)])
def test_rs_code_matrix_matches_definition(oracle, l_skip, log_h, log_blowup, width):
    rng = np.random.default_rng(10 * l_skip + log_h)
    H = 1 << log_h
    ev = rng.integers(0, P, H * width, dtype=np.uint64)
    got = oracle.from_mont(oracle.rs_code_matrix(l_skip, log_blowup, oracle.to_mont(ev), H, width))
    N = H << log_blowup
    for c in range(width):
        msg = _rs_message_py(l_skip, [int(x) for x in ev[c * H : (c + 1) * H]])
        assert list(oracle.from_mont(oracle.eval_to_coeff_rs_message(l_skip, oracle.to_mont(ev[c * H : (c + 1) * H])))) == msg
        assert list(got[c * N : (c + 1) * N]) == _naive_dft(msg, N)


def test_rs_code_is_systematic_on_boolean_points(oracle):
    # The RS message is the multilinear-in-index-bits form of the prismalinear polynomial; a
    # size-independent sanity property: with l_skip = 0 the message equals the evaluations.
    rng = np.random.default_rng(5)
    ev = oracle.random_field(rng, 64)
    assert np.array_equal(oracle.eval_to_coeff_rs_message(0, ev), ev)


# ---------------------------------------------------------------------------------------------
# Merkle tree
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("height,width,rpq", [(1, 1, 1), (4, 3, 1), (8, 9, 2), (16, 1, 16), (32, 19, 4), (24, 8, 4), (64, 5, 8)])
def test_merkle_tree_structure(oracle, height, width, rpq):
    rng = np.random.default_rng(height * 100 + width)
    m = oracle.random_field(rng, height * width)
    layers = oracle.merkle_tree(m, height, width, rpq)
    leaves = 1 << (height - 1).bit_length() if height > 1 else 1
    mat = m.reshape(width, height)
    row = lambda r: mat[:, r] if r < height else np.zeros(width, np.uint32)
    cur = [oracle.hash_slice(np.ascontiguousarray(row(r))) for r in range(leaves)]
    S = leaves // rpq
    for _ in range(rpq.bit_length() - 1):
        cur = [oracle.compress(cur[2 * (i // S) * S + i % S], cur[(2 * (i // S) + 1) * S + i % S]) for i in range(len(cur) // 2)]
    assert len(layers[0]) == S and all(np.array_equal(a, b) for a, b in zip(layers[0], cur))
    for l in range(1, len(layers)):
        prev = layers[l - 1]
        assert len(layers[l]) == len(prev) // 2
        for i in range(len(layers[l])):
            assert np.array_equal(layers[l][i], oracle.compress(prev[2 * i], prev[2 * i + 1]))
    assert len(layers[-1]) == 1


def test_merkle_tree_errors(oracle):
    assert oracle.merkle_tree(np.zeros(4, np.uint32), 4, 1, 8) is None  # rows_per_query > leaves
    assert oracle.merkle_tree(np.zeros(0, np.uint32), 0, 1, 1) is None  # empty matrix


def test_stacked_commit_composes(oracle):
    rng = np.random.default_rng(6)
    traces = [(oracle.random_field(rng, 16 * 3), 16, 3), (oracle.random_field(rng, 8 * 2), 8, 2), (oracle.random_field(rng, 2), 2, 1)]
    l_skip, n_stack, lb, k = 2, 2, 1, 2
    root, cw, layers, W = oracle.stacked_commit(l_skip, n_stack, lb, k, traces)
    q, w2 = oracle.stacked_matrix(l_skip, n_stack, traces)
    assert W == w2 == 5
    assert np.array_equal(cw, oracle.rs_code_matrix(l_skip, lb, q, 16, W))
    want = oracle.merkle_tree(cw, 32, W, 1 << k)
    assert all(np.array_equal(a, b) for a, b in zip(layers, want)) and np.array_equal(root, want[-1][0])


# ---------------------------------------------------------------------------------------------
# transcript
# ---------------------------------------------------------------------------------------------
class _PySponge:
    """duplex_sponge.rs:60-83 restated directly on top of the oracle permutation."""

    def __init__(self, oracle):
        self.o, self.state, self.a, self.s = oracle, np.zeros(16, np.uint32), 0, 0

    def observe(self, v):
        self.state[self.a] = v
        self.a += 1
        if self.a == 8:
            self.state = self.o.permute(self.state)
            self.a, self.s = 0, 8

    def sample(self):
        if self.a != 0 or self.s == 0:
            self.state = self.o.permute(self.state)
            self.a, self.s = 0, 8
        self.s -= 1
        return self.state[self.s]


def test_duplex_sponge_matches_direct_restatement(oracle):
    rng = np.random.default_rng(7)
    st, py = oracle.sponge_new(), _PySponge(oracle)
    for step in range(200):
        if rng.integers(0, 2):
            v = oracle.random_field(rng, int(rng.integers(1, 12)))
            oracle.sponge_observe(st, v)
            for x in v:
                py.observe(x)
        else:
            n = int(rng.integers(1, 12))
            got = oracle.sponge_sample(st, n)
            assert list(got) == [py.sample() for _ in range(n)]
        assert st[16] == py.a and st[17] == py.s and np.array_equal(st[:16], py.state)


@pytest.mark.parametrize("bits", [0, 1, 3, 8, 12])
def test_grind_finds_smallest_valid_witness(oracle, bits):
    rng = np.random.default_rng(bits)
    st = oracle.sponge_new()
    oracle.sponge_observe(st, oracle.random_field(rng, 11))
    before = st.copy()
    w = oracle.sponge_grind(st, bits)
    wc = int(oracle.from_mont([w])[0])
    if bits == 0:
        assert w == 0 and np.array_equal(st, before)  # grind(0) leaves the transcript untouched
        return
    # every smaller candidate fails, the returned one passes, and the transcript advanced with it
    for c in range(wc):
        probe = before.copy()
        assert not oracle.sponge_check_witness(probe, bits, int(oracle.to_mont([c])[0]))
    probe = before.copy()
    assert oracle.sponge_check_witness(probe, bits, w)
    assert np.array_equal(probe, st)
    # the sampled word is state[7] after exactly one permutation (what the GPU grind kernel uses)
    s2 = before[:16].copy()
    s2[before[16]] = w
    assert int(oracle.from_mont(oracle.permute(s2))[7]) & ((1 << bits) - 1) == 0
