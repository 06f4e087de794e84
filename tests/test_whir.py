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
