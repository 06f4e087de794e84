"""Stacked opening reduction: oracle prover vs oracle verifier (CPU); CUDA prover vs
oracle prover through the C ABI (GPU, bit-exact)."""
import numpy as np
import pytest

import stark_backend_b200 as sb

# (l_skip, n_stack, commits); a commit = list of (log_height, width, need_rot), heights descending
CASES = [
    (2, 3, [[(5, 2, True)]]),
    (2, 4, [[(5, 3, True), (4, 2, False), (3, 1, True)]]),
    (3, 3, [[(5, 2, True), (2, 3, True), (1, 2, False), (0, 1, True)]]),   # traces shorter than 2^l_skip (lifted)
    (2, 4, [[(6, 1, False), (4, 5, True)], [(5, 2, True)], [(3, 3, False)]]),  # common main + cached commits
    (4, 2, [[(6, 4, True), (5, 1, True)]]),
]


def make_case(oracle, case, seed):
    l_skip, n_stack, commits = case
    rng = np.random.default_rng(seed)
    data = []
    n_max = 0
    for traces in commits:
        tl = []
        for lh, w, nr in traces:
            h = 1 << lh
            tl.append((oracle.random_field(rng, h * w), h, w, nr))
            n_max = max(n_max, lh - l_skip)
        data.append(tl)
    r = oracle.random_field(rng, (n_max + 1, 4))
    st = np.zeros(18, np.uint32)
    oracle.sponge_observe(st, oracle.to_mont(np.arange(seed, seed + 4)))
    return data, r, st


def t_claims_for(oracle, l_skip, data, r):
    """(claim, rot claim) per commit, per column in layout order (= trace order, column order)."""
    out = []
    for traces in data:
        for v, h, w, nr in traces:
            for c in range(w):
                col = v[c * h : (c + 1) * h]
                cl = oracle.column_opening(l_skip, col, False, r)
                rt = oracle.column_opening(l_skip, col, True, r) if nr else np.zeros(4, np.uint32)
                out.append(np.concatenate([cl, rt]))
    return np.array(out, dtype=np.uint32)


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"l{c[0]}n{c[1]}c{len(c[2])}t{len(c[2][0])}")
def test_oracle_stacked_reduction_accepted_by_oracle_verifier(oracle, case):
    l_skip, n_stack, _ = case
    data, r, st = make_case(oracle, case, 31)
    stv = st.copy()
    proof, u, _ = oracle.stacked_reduction_prove(st, l_skip, n_stack, data, r)
    claims = t_claims_for(oracle, l_skip, data, r)
    ok, uv = oracle.stacked_reduction_verify(stv, l_skip, n_stack, data, claims, r, proof)
    assert ok
    assert np.array_equal(u, uv) and np.array_equal(st, stv)
    bad = claims.copy()
    bad[0, 0] ^= 1
    assert not oracle.stacked_reduction_verify(make_case(oracle, case, 31)[2], l_skip, n_stack, data, bad, r, proof)[0]
    # one negative case per proof section, as the reference has them (tests/stacked_reduction.rs:192-245): univariate round
    # coefficients, sumcheck round polynomials, stacking openings
    n0 = (2 * ((1 << l_skip) - 1) + 1) * 4
    for idx in [0, n0 - 1] + ([n0, n0 + n_stack * 8 - 1] if n_stack else []) + [n0 + n_stack * 8, len(proof) - 1]:
        badp = proof.copy()
        badp[idx] ^= 1
        assert not oracle.stacked_reduction_verify(make_case(oracle, case, 31)[2], l_skip, n_stack, data, claims, r, badp)[0], idx


def gpu_commits(dev, l_skip, n_stack, data):
    params = sb.PcsParams(l_skip, n_stack, 1, 1)
    out = []
    for traces in data:
        _, d = dev.commit(params, [sb.DeviceMatrix(dev.h2d(v), h, w) for v, h, w, _ in traces])
        out.append(d)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"l{c[0]}n{c[1]}c{len(c[2])}t{len(c[2][0])}")
def test_gpu_stacked_reduction_matches_oracle(dev, oracle, case):
    l_skip, n_stack, _ = case
    data, r, st = make_case(oracle, case, 47)
    ts = sb.Transcript(st)
    want, u, _ = oracle.stacked_reduction_prove(st, l_skip, n_stack, data, r)
    pcs = gpu_commits(dev, l_skip, n_stack, data)
    got = dev.stacked_reduction(ts, pcs, [[nr for *_, nr in traces] for traces in data], r)
    assert np.array_equal(got["flat"], want)
    assert np.array_equal(got["u"], u)
    assert np.array_equal(ts.words(), st)


@pytest.mark.gpu
def test_gpu_stacked_reduction_large_accepted_by_oracle_verifier(dev, oracle):
    # 2^14 and 2^12-row traces, 40 columns in total: verified through the oracle verifier
    case = (4, 10, [[(14, 12, True), (14, 9, False), (12, 15, True), (3, 4, True)]])
    l_skip, n_stack, _ = case
    data, r, st = make_case(oracle, case, 3)
    ts = sb.Transcript(st)
    pcs = gpu_commits(dev, l_skip, n_stack, data)
    got = dev.stacked_reduction(ts, pcs, [[nr for *_, nr in traces] for traces in data], r)
    claims = t_claims_for(oracle, l_skip, data, r)
    ok, uv = oracle.stacked_reduction_verify(st, l_skip, n_stack, data, claims, r, got["flat"])
    assert ok and np.array_equal(uv, got["u"]) and np.array_equal(st, ts.words())
