"""WHIR opening: oracle prover vs oracle verifier (CPU); CUDA prover vs oracle
prover through the C ABI (GPU, bit-exact)."""
import numpy as np
import pytest

import stark_backend_b200 as sb

CASES = [
    # l_skip, n_stack, log_blowup, cfg, widths
    (2, 4, 1, dict(k=2, num_queries=[5, 4], mu_pow_bits=3, query_phase_pow_bits=4, folding_pow_bits=2), [3]),
    (4, 4, 1, dict(k=4, num_queries=[7], mu_pow_bits=0, query_phase_pow_bits=0, folding_pow_bits=0), [5]),
    (2, 6, 2, dict(k=2, num_queries=[6, 5, 4], mu_pow_bits=2, query_phase_pow_bits=3, folding_pow_bits=1), [2, 9]),
    (4, 8, 1, dict(k=4, num_queries=[9, 6], mu_pow_bits=4, query_phase_pow_bits=5, folding_pow_bits=3), [4, 1, 3]),
    (0, 5, 1, dict(k=1, num_queries=[4, 4, 3], mu_pow_bits=1, query_phase_pow_bits=1, folding_pow_bits=1), [2]),
]
# the shape grid of the reference's whir_single_fib_{n_stack}_{log_blowup}_{k}_{log_final_poly_len} tests (backend-tests/src/lib.rs:
# 1229-1251, 1673-1700: l_skip = 2, one 2-column commitment, PoW bits 1 / 2 / 3 as test_whir_config_small sets them); rounds =
# ceil((l_skip + n_stack - log_final_poly_len) / k)
_SMALL = dict(mu_pow_bits=3, query_phase_pow_bits=1, folding_pow_bits=2)
CASES += [
    (2, 0, 1, dict(k=1, num_queries=[3, 2], **_SMALL), [2]),
    (2, 2, 1, dict(k=1, num_queries=[4, 3], **_SMALL), [2]),
    (2, 2, 1, dict(k=2, num_queries=[4, 3], **_SMALL), [2]),
    (2, 2, 1, dict(k=3, num_queries=[4], **_SMALL), [2]),
    (2, 2, 1, dict(k=4, num_queries=[4], **_SMALL), [2]),
    (2, 2, 2, dict(k=4, num_queries=[3], **_SMALL), [2]),
    # whir_multiple_commitments (lib.rs:1269-1348): five commitments of 3..12 columns, l_skip = n_stack = 3, two rounds of 6 / 5 queries
    (3, 3, 1, dict(k=2, num_queries=[6, 5], mu_pow_bits=1, query_phase_pow_bits=1, folding_pow_bits=1), [5, 12, 3, 8, 10]),
]


def make_case(oracle, case, seed):
    l_skip, n_stack, log_blowup, cfg, widths = case
    rng = np.random.default_rng(seed)
    H = 1 << (l_skip + n_stack)
    mats = [(oracle.random_field(rng, H * w), w) for w in widths]
    u = oracle.random_field(rng, (l_skip + n_stack, 4))
    st = np.zeros(18, np.uint32)
    oracle.sponge_observe(st, oracle.to_mont(np.arange(seed, seed + 3)))
    return H, mats, u, st


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"l{c[0]}n{c[1]}b{c[2]}k{c[3]['k']}w{len(c[4])}")
def test_oracle_whir_prover_accepted_by_oracle_verifier(oracle, case):
    l_skip, n_stack, log_blowup, cfg, widths = case
    H, mats, u, st = make_case(oracle, case, 11)
    stv = st.copy()
    roots, proof = oracle.whir_prove(st, l_skip, log_blowup, cfg, mats, H, u)
    openings = np.concatenate([oracle.whir_stacking_openings(l_skip, v, H, w, u) for v, w in mats])
    assert oracle.whir_verify(stv, l_skip, n_stack, log_blowup, cfg, proof, widths, openings, roots, u)
    assert np.array_equal(st, stv)
    # wrong opening claim / tampered proof are rejected
    bad = openings.copy()
    bad[0, 0] ^= 1
    st2 = make_case(oracle, case, 11)[3]
    assert not oracle.whir_verify(st2, l_skip, n_stack, log_blowup, cfg, proof, widths, bad, roots, u)
    badp = proof.copy()
    badp[-1] ^= 1
    st3 = make_case(oracle, case, 11)[3]
    assert not oracle.whir_verify(st3, l_skip, n_stack, log_blowup, cfg, badp, widths, openings, roots, u)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"l{c[0]}n{c[1]}b{c[2]}k{c[3]['k']}w{len(c[4])}")
def test_gpu_whir_matches_oracle(dev, oracle, case):
    l_skip, n_stack, log_blowup, cfg, widths = case
    H, mats, u, st = make_case(oracle, case, 23)
    ts = sb.Transcript(st)
    roots, want = oracle.whir_prove(st, l_skip, log_blowup, cfg, mats, H, u)
    params = sb.PcsParams(l_skip, n_stack, log_blowup, cfg["k"])
    pcs = []
    for (v, w), root in zip(mats, roots):
        r, d = dev.commit(params, [sb.DeviceMatrix(dev.h2d(v), H, w)])
        assert np.array_equal(r, root)
        pcs.append(d)
    got = dev.whir_open(ts, sb.WhirConfig(**cfg), params, pcs, u)
    assert got.size == want.size
    assert np.array_equal(got, want)
    assert np.array_equal(ts.words(), st)


@pytest.mark.gpu
def test_gpu_whir_large_accepted_by_oracle_verifier(dev, oracle):
    # 2^16 x 24 with production-like parameters: checked through the oracle verifier
    case = (4, 12, 1, dict(k=4, num_queries=[40, 20, 14], mu_pow_bits=8, query_phase_pow_bits=12, folding_pow_bits=5), [24])
    l_skip, n_stack, log_blowup, cfg, widths = case
    H, mats, u, st = make_case(oracle, case, 5)
    ts = sb.Transcript(st)
    params = sb.PcsParams(l_skip, n_stack, log_blowup, cfg["k"])
    root, d = dev.commit(params, [sb.DeviceMatrix(dev.h2d(mats[0][0]), H, widths[0])])
    proof = dev.whir_open(ts, sb.WhirConfig(**cfg), params, [d], u)
    openings = oracle.whir_stacking_openings(l_skip, mats[0][0], H, widths[0], u)
    assert oracle.whir_verify(st, l_skip, n_stack, log_blowup, cfg, proof, widths, openings, root.reshape(1, 8), u)
    assert np.array_equal(st, ts.words())


# ---- the reference's binary_k_fold identities (backend-tests/src/lib.rs:1191-1227: fold_single, fold_double) -----------
# The expected values are formed here in Python integers (F_p[x] / (x^4 - 11), canonical representatives), independent of
# the oracle's field code; the oracle's verifier-side fold must reproduce them.
_P = 0x78000001


def _ef(*c):
    return tuple(int(x) % _P for x in c)


def _add(a, b):
    return tuple((x + y) % _P for x, y in zip(a, b))


def _sub(a, b):
    return tuple((x - y) % _P for x, y in zip(a, b))


def _mul(a, b):
    r = [0] * 7
    for i in range(4):
        for j in range(4):
            r[i + j] += a[i] * b[j]
    return tuple((r[i] + 11 * (r[i + 4] if i < 3 else 0)) % _P for i in range(4))


def _scale(a, s):
    return tuple(x * s % _P for x in a)


def _fold_step(lo, hi, alpha, x):  # lo + (alpha - x) (lo - hi) / (2 x)
    return _add(lo, _scale(_mul(_sub(alpha, _ef(x, 0, 0, 0)), _sub(lo, hi)), pow(2 * x, -1, _P)))


@pytest.mark.parametrize("seed", range(4))
def test_oracle_binary_k_fold_satisfies_the_reference_identities(oracle, seed):
    rng = np.random.default_rng(seed)
    ef = lambda: _ef(*rng.integers(0, _P, size=4))
    mont = lambda e: oracle.to_mont(np.array(e, dtype=np.uint64))
    canon = lambda w: tuple(int(v) for v in oracle.from_mont(w))
    x = int(rng.integers(1, _P))
    xm = int(oracle.to_mont(np.array([x], dtype=np.uint64))[0])
    # k = 1
    a0, a1, alpha = ef(), ef(), ef()
    got = oracle.binary_k_fold(np.stack([mont(a0), mont(a1)]), np.stack([mont(alpha)]), xm)
    assert canon(got) == _fold_step(a0, a1, alpha, x)
    # k = 2: pairs (a0, a2) at x and (a1, a3) at tw * x, then the two results at x^2
    a = [ef() for _ in range(4)]
    al = [ef(), ef()]
    tw = pow(0x1A427A41, 1 << 25, _P)  # two_adic_generator(2): the 2^27-th root of unity of BabyBear, raised to 2^25
    assert pow(tw, 4, _P) == 1 and pow(tw, 2, _P) != 1
    b0 = _fold_step(a[0], a[2], al[0], x)
    b1 = _fold_step(a[1], a[3], al[0], tw * x % _P)
    want = _fold_step(b0, b1, al[1], x * x % _P)
    got = oracle.binary_k_fold(np.stack([mont(v) for v in a]), np.stack([mont(v) for v in al]), xm)
    assert canon(got) == want
