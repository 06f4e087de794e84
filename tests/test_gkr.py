"""LogUp-GKR fractional sumcheck: oracle prover vs oracle verifier (CPU), and the
CUDA prover vs the oracle prover through the C ABI (GPU, bit-exact)."""
import numpy as np
import pytest

import stark_backend_b200 as sb

P = sb.P


def balanced_leaves(oracle, rng, log_n, alpha=None):
    """2^log_n fractions whose sum is zero as a rational function value: pairs (m, q), (-m, q)."""
    n = 1 << log_n
    leaves = np.zeros((n, 8), np.uint32)
    half = n // 2
    q = oracle.random_field(rng, (half, 4))
    m = oracle.random_field(rng, half)
    perm = rng.permutation(n)
    neg = (P - m.astype(np.int64)) % P
    for k in range(half):
        a, b = perm[2 * k], perm[2 * k + 1]
        leaves[a, 0], leaves[a, 4:] = m[k], q[k]
        leaves[b, 0], leaves[b, 4:] = neg[k], q[k]
    return leaves


def seeded_sponge(oracle, seed):
    st = np.zeros(18, np.uint32)
    oracle.sponge_observe(st, oracle.to_mont(np.arange(seed, seed + 5)))
    return st


@pytest.mark.parametrize("log_n", [1, 2, 3, 5, 8])
def test_oracle_gkr_prover_accepted_by_oracle_verifier(oracle, log_n):
    rng = np.random.default_rng(log_n)
    leaves = balanced_leaves(oracle, rng, log_n)
    st = seeded_sponge(oracle, log_n)
    stv = st.copy()
    proof = oracle.gkr_prove(st, leaves, log_n, True)
    assert not proof["frac_sum"][:4].any()
    ok, numer, denom, xi = oracle.gkr_verify(stv, log_n, proof)
    assert ok
    assert np.array_equal(xi, proof["xi"]) and np.array_equal(st, stv)
    # the final claims are the MLEs of the leaf numerators / denominators at xi
    assert np.array_equal(numer, oracle.eval_mle_evals_at_point(leaves[:, :4], log_n, xi))
    assert np.array_equal(denom, oracle.eval_mle_evals_at_point(leaves[:, 4:], log_n, xi))
    # tampering is rejected
    bad = {k: v.copy() for k, v in proof.items()}
    bad["claims"][log_n - 1, 3] ^= 1
    assert not oracle.gkr_verify(seeded_sponge(oracle, log_n), log_n, bad)[0]


def test_oracle_gkr_unbalanced_is_an_error(oracle):
    rng = np.random.default_rng(5)
    leaves = oracle.random_field(rng, (16, 8))
    with pytest.raises(ValueError):
        oracle.gkr_prove(seeded_sponge(oracle, 1), leaves, 4, True)
    # without assert_zero the numerator is observed and the proof goes through
    pr = oracle.gkr_prove(seeded_sponge(oracle, 1), leaves, 4, False)
    assert pr["frac_sum"][:4].any()


def test_transcript_binding_matches_oracle(oracle):
    # host-only logic of the product library (no GPU needed)
    rng = np.random.default_rng(3)
    vals = oracle.random_field(rng, 37)
    st = np.zeros(18, np.uint32)
    ts = sb.Transcript()
    oracle.sponge_observe(st, vals[:11])
    ts.observe(vals[:11])
    assert np.array_equal(ts.words(), st)
    assert np.array_equal(ts.sample(5), oracle.sponge_sample(st, 5))
    oracle.sponge_observe(st, vals[11:])
    ts.observe(vals[11:])
    assert ts.sample_bits(13) == oracle.sponge_sample_bits(st, 13)
    assert np.array_equal(ts.words(), st)
    w = int(oracle.from_mont([oracle.sponge_grind(st.copy(), 6)])[0])
    assert ts.check_witness(6, w)


@pytest.mark.gpu
@pytest.mark.parametrize("log_n", [1, 2, 3, 4, 7, 10, 13, 14, 16])
def test_gpu_gkr_matches_oracle(dev, oracle, log_n):
    rng = np.random.default_rng(100 + log_n)
    leaves = balanced_leaves(oracle, rng, log_n)
    st = seeded_sponge(oracle, 7)
    ts = sb.Transcript(st)
    want = oracle.gkr_prove(st, leaves, log_n, True)
    got = dev.gkr_fractional_sumcheck(ts, dev.h2d(leaves), log_n, True)
    for k in ("frac_sum", "claims", "polys", "xi"):
        assert np.array_equal(got[k], want[k]), k
    assert np.array_equal(ts.words(), st)


@pytest.mark.gpu
@pytest.mark.parametrize("log_n,n_stored", [(2, 4), (3, 4), (5, 4), (5, 20), (8, 128), (8, 132), (10, 516), (13, 4100), (14, 8192),
                                            (15, 16388), (15, 24580), (16, 40000), (16, 65536)])
def test_gpu_gkr_padded_tail_matches_oracle_on_full_leaves(dev, oracle, log_n, n_stored):
    """Only the first n_stored leaves are given; the rest is the constant (0, pad_q) the interaction layout pads
    with.  The proof must be the one the oracle produces from all 2^log_n leaves."""
    rng = np.random.default_rng(1000 + log_n + n_stored)
    leaves = oracle.random_field(rng, (1 << log_n, 8))
    pad_q = oracle.random_field(rng, 4)
    leaves[n_stored:, :4] = 0
    leaves[n_stored:, 4:] = pad_q
    st = seeded_sponge(oracle, 11)
    ts = sb.Transcript(st)
    want = oracle.gkr_prove(st, leaves, log_n, False)
    got = dev.gkr_fractional_sumcheck(ts, dev.h2d(np.ascontiguousarray(leaves[:n_stored])), log_n, False, n_stored=n_stored, pad_q=pad_q)
    for k in ("frac_sum", "claims", "polys", "xi"):
        assert np.array_equal(got[k], want[k]), k
    assert np.array_equal(ts.words(), st)


@pytest.mark.gpu
def test_gpu_gkr_not_assert_zero_and_error(dev, oracle):
    rng = np.random.default_rng(9)
    leaves = oracle.random_field(rng, (1 << 9, 8))
    st = seeded_sponge(oracle, 2)
    ts = sb.Transcript(st)
    want = oracle.gkr_prove(st, leaves, 9, False)
    got = dev.gkr_fractional_sumcheck(ts, dev.h2d(leaves), 9, False)
    for k in ("frac_sum", "claims", "polys", "xi"):
        assert np.array_equal(got[k], want[k]), k
    with pytest.raises(sb.SwirlError) as e:
        dev.gkr_fractional_sumcheck(sb.Transcript(), dev.h2d(leaves), 9, True)
    assert e.value.code == 10005 and "NonZeroRootSum" in str(e.value)


@pytest.mark.gpu
def test_gpu_gkr_large_accepted_by_oracle_verifier(dev, oracle):
    # 2^20 leaves: too slow for the scalar oracle prover, so check through the verifier and the
    # MLE identity of the final claims (size-independent properties)
    log_n = 20
    rng = np.random.default_rng(20)
    leaves = balanced_leaves(oracle, rng, log_n)
    ts = sb.Transcript(seeded_sponge(oracle, 3))
    got = dev.gkr_fractional_sumcheck(ts, dev.h2d(leaves), log_n, True)
    stv = seeded_sponge(oracle, 3)
    ok, numer, denom, xi = oracle.gkr_verify(stv, log_n, got)
    assert ok and np.array_equal(stv, ts.words())
    assert np.array_equal(numer, oracle.eval_mle_evals_at_point(leaves[:, :4], log_n, xi))
    assert np.array_equal(denom, oracle.eval_mle_evals_at_point(leaves[:, 4:], log_n, xi))


# ---- the reference's own GKR unit tests (crates/stark-backend/tests/fractional_sumcheck_gkr.rs) on the oracle ------------
def _ef(oracle, x):
    return np.array([oracle.to_mont(np.array([x % P], dtype=np.uint64))[0], 0, 0, 0], dtype=np.uint32)


def test_oracle_verifier_rejects_nonzero_base_layer_numerator(oracle):
    """test_gkr_base_layer_numerator_zero (:62-87): claims p = (1, 2), q = (3, 4), q0 = 12: p0 = 1*4 + 2*3 = 10 != 0."""
    claims = np.concatenate([_ef(oracle, 1), _ef(oracle, 3), _ef(oracle, 2), _ef(oracle, 4)]).reshape(1, 16)  # p_xi_0, q_xi_0, p_xi_1, q_xi_1
    proof = dict(frac_sum=np.concatenate([_ef(oracle, 0), _ef(oracle, 12)]), claims=claims, polys=np.zeros((0, 12), np.uint32))
    assert not oracle.gkr_verify(np.zeros(18, np.uint32), 1, proof)[0]
    # the same shape with a vanishing numerator (p = (3, -4): 3*4 - 4*3 = 0) and q0 = 12 passes the base-layer check
    good = np.concatenate([_ef(oracle, 3), _ef(oracle, 3), _ef(oracle, P - 4), _ef(oracle, 4)]).reshape(1, 16)
    proof = dict(frac_sum=np.concatenate([_ef(oracle, 0), _ef(oracle, 12)]), claims=good, polys=np.zeros((0, 12), np.uint32))
    assert oracle.gkr_verify(np.zeros(18, np.uint32), 1, proof)[0]
    # and a wrong q0 does not
    proof["frac_sum"] = np.concatenate([_ef(oracle, 0), _ef(oracle, 13)])
    assert not oracle.gkr_verify(np.zeros(18, np.uint32), 1, proof)[0]


@pytest.mark.parametrize("log_n", [1, 2, 3])
def test_oracle_gkr_trivial_fractions_integration(oracle, log_n):
    """test_gkr_{1,2,3}_round_integration (:89-214): 2^n fractions 0/1 from a fresh transcript; the verifier returns a zero
    numerator claim and a non-zero denominator claim."""
    leaves = np.zeros((1 << log_n, 8), np.uint32)
    leaves[:, 4] = _ef(oracle, 1)[0]
    st, stv = np.zeros(18, np.uint32), np.zeros(18, np.uint32)
    proof = oracle.gkr_prove(st, leaves, log_n, True)
    ok, numer, denom, _ = oracle.gkr_verify(stv, log_n, proof)
    assert ok and not numer.any() and denom.any()
    assert np.array_equal(st, stv)


def test_oracle_gkr_mixed_fractions(oracle):
    """test_gkr_mixed_fractions (:220-251): 5/1 + (-5)/1."""
    leaves = np.zeros((2, 8), np.uint32)
    leaves[0, :4], leaves[1, :4] = _ef(oracle, 5), _ef(oracle, P - 5)
    leaves[:, 4] = _ef(oracle, 1)[0]
    st, stv = np.zeros(18, np.uint32), np.zeros(18, np.uint32)
    proof = oracle.gkr_prove(st, leaves, 1, True)
    ok, _, denom, _ = oracle.gkr_verify(stv, 1, proof)
    assert ok and denom.any()


def test_oracle_verifier_rejects_wrong_shapes(oracle):
    """test_multiple_rounds_shape (:20-60): a proof for one round offered as a two-round proof is rejected (layer count /
    sumcheck polynomial count)."""
    leaves = np.zeros((2, 8), np.uint32)
    leaves[:, 4] = _ef(oracle, 1)[0]
    proof = oracle.gkr_prove(np.zeros(18, np.uint32), leaves, 1, True)
    two = dict(frac_sum=proof["frac_sum"], claims=np.concatenate([proof["claims"], proof["claims"]]), polys=np.zeros((1, 12), np.uint32))
    assert not oracle.gkr_verify(np.zeros(18, np.uint32), 2, two)[0]
