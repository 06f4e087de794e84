"""Batch constraint sumcheck = LogUp input layer + GKR + zerocheck/LogUp round 0 + MLE rounds
(SURVEY §8 a5-a8): oracle prover vs oracle verifier (CPU); CUDA prover vs oracle prover (GPU)."""
import numpy as np
import pytest

import airs as A
import stark_backend_b200 as sb


def sorted_airs(airs):
    return sorted(airs, key=lambda a: -a.height)


def case_fib(rng):
    return 2, 2, 0, [A.fibonacci(5)]


def case_fib_short(rng):  # trace shorter than 2^l_skip (lifted)
    return 3, 2, 1, [A.fibonacci(2)]


def case_benchmark(rng):
    return 2, 2, 2, [A.benchmark(5, 6, 6, 2, rng)]


def case_sender_receiver(rng):
    s, r = A.sender_receiver(5, 3, rng)
    return 2, 2, 3, sorted_airs([s, r])


def case_mixed(rng):
    s, r = A.sender_receiver(4, 1, rng)
    return 2, 3, 2, sorted_airs([A.fibonacci(6), A.benchmark(4, 3, 5, 3, rng), s, r, A.with_parts(5, rng)])


def case_parts(rng):
    return 2, 3, 0, [A.with_parts(4, rng)]


CASES = [case_fib, case_fib_short, case_benchmark, case_sender_receiver, case_parts, case_mixed]


def setup(oracle, case, seed=1):
    rng = np.random.default_rng(seed)
    l_skip, D, pow_bits, airs = case(rng)
    n_max = max(max(a.height.bit_length() - 1 - l_skip for a in airs), 0)
    st = np.zeros(18, np.uint32)
    oracle.sponge_observe(st, oracle.to_mont(np.arange(seed, seed + 6)))
    return l_skip, D, pow_bits, airs, n_max, st


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.__name__)
def test_oracle_batch_constraints_accepted_by_oracle_verifier(oracle, case):
    l_skip, D, pow_bits, airs, n_max, st = setup(oracle, case)
    flat = A.flatten(airs)
    stv = st.copy()
    proof, r = oracle.bc_prove(st, l_skip, D, pow_bits, flat, len(airs), n_max)
    ok, rv = oracle.bc_verify(stv, l_skip, D, pow_bits, flat, len(airs), n_max, proof)
    assert ok
    assert np.array_equal(r, rv) and np.array_equal(st, stv)
    bad = proof.copy()
    bad[1] ^= 1  # q0_claim
    assert not oracle.bc_verify(setup(oracle, case)[5], l_skip, D, pow_bits, flat, len(airs), n_max, bad)[0]
    bad = proof.copy()
    bad[5 + 8 * len(airs) + (len(proof) - 5 - 8 * len(airs)) // 3] ^= 1
    rejected = not oracle.bc_verify(setup(oracle, case)[5], l_skip, D, pow_bits, flat, len(airs), n_max, bad)[0]
    assert rejected or case is case_parts  # (an opening no constraint reads is only bound by the stacked reduction)


def test_oracle_violated_constraint_is_rejected(oracle):
    l_skip, D, pow_bits, airs, n_max, st = setup(oracle, case_fib)
    airs[0].common_main[0][7] ^= 1  # break the trace
    flat = A.flatten(airs)
    stv = st.copy()
    proof, _ = oracle.bc_prove(st, l_skip, D, pow_bits, flat, 1, n_max)
    assert not oracle.bc_verify(stv, l_skip, D, pow_bits, flat, 1, n_max, proof)[0]


def test_oracle_unbalanced_logup_is_an_error(oracle):
    rng = np.random.default_rng(4)
    s, r = A.sender_receiver(4, 2, rng, balanced=False)
    st = np.zeros(18, np.uint32)
    with pytest.raises(ValueError):
        oracle.bc_prove(st, 2, 2, 0, A.flatten([s, r]), 2, 2)


# ---- the reference's literal interaction tables (backend-tests/src/lib.rs:844-1018): row-major [count, field] ------------
_RECV = [1, 5, 3, 4, 4, 4, 2, 5, 0, 123, 545, 889, 1, 889, 0, 456]
REFERENCE_INTERACTION_TABLES = {
    # name: ([(table, is_send), ...], balanced)
    "interaction_multi_rows_neg": ([([0, 1, 3, 5, 7, 4, 546, 0], True),
                                    ([1, 5, 3, 4, 4, 4, 2, 5, 0, 123, 545, 0, 0, 0, 0, 456], False)], False),
    "interaction_all_zero_sender": ([([0, 1, 0, 5, 0, 4, 0, 889], True)], True),
    "interaction_multi_senders": ([([0, 1, 3, 5, 6, 4, 333, 889], True), ([1, 4, 213, 889], True), (_RECV, False)], True),
    "interaction_multi_senders_neg": ([([0, 1, 3, 5, 5, 4, 333, 889], True), ([1, 4, 213, 889], True), (_RECV, False)], False),
    "interaction_multi_sender_receiver": ([([0, 1, 3, 5, 6, 4, 333, 889], True), ([1, 4, 213, 889], True),
                                           ([1, 5, 3, 4, 4, 4, 2, 5, 0, 123, 545, 889, 0, 289, 0, 456], False), ([1, 889], False)], True),
}


@pytest.mark.parametrize("name", sorted(REFERENCE_INTERACTION_TABLES))
def test_oracle_on_the_reference_interaction_tables(oracle, name):
    """The oracle must agree with the outcome the reference's test asserts: balanced tables prove and verify (l_skip = 2 as
    default_test_params_small; traces of 8, 4, 2 and 1 rows, i.e. lifted ones included), unbalanced ones are a prover error."""
    tables, balanced = REFERENCE_INTERACTION_TABLES[name]
    airs = sorted([A.dummy_interaction(t, send) for t, send in tables], key=lambda a: -a.height)  # ProvingContext::into_sorted
    l_skip, D, pow_bits = 2, 3, 1
    n_max = max(max(a.height.bit_length() - 1 - l_skip for a in airs), 0)
    st = np.zeros(18, np.uint32)
    oracle.sponge_observe(st, oracle.to_mont(np.arange(3)))
    stv = st.copy()
    flat = A.flatten(airs)
    if not balanced:
        with pytest.raises(ValueError):
            oracle.bc_prove(st, l_skip, D, pow_bits, flat, len(airs), n_max)
        return
    proof, r = oracle.bc_prove(st, l_skip, D, pow_bits, flat, len(airs), n_max)
    ok, rv = oracle.bc_verify(stv, l_skip, D, pow_bits, flat, len(airs), n_max, proof)
    assert ok and np.array_equal(r, rv) and np.array_equal(st, stv)


def to_device_airs(dev, airs):
    out = []
    for a in airs:
        dm = lambda m: sb.DeviceMatrix(dev.h2d(m[0]), m[1], m[2])
        out.append(sb.AirProvingContext(a.nodes, a.constraint_idx, a.interactions, a.constraint_degree, a.need_rot,
                                        dm(a.common_main), a.public_values, [dm(m) for m in a.cached],
                                        dm(a.preprocessed) if a.preprocessed is not None else None))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=lambda c: c.__name__)
def test_gpu_batch_constraints_match_oracle(dev, oracle, case):
    l_skip, D, pow_bits, airs, n_max, st = setup(oracle, case, seed=5)
    ts = sb.Transcript(st)
    want, r = oracle.bc_prove(st, l_skip, D, pow_bits, A.flatten(airs), len(airs), n_max)
    got, rg = dev.prove_batch_constraints(ts, l_skip, D, pow_bits, to_device_airs(dev, airs))
    assert got.size == want.size
    bad = np.nonzero(got != want)[0]
    assert bad.size == 0, f"first mismatch at word {bad[0]} of {got.size}"
    assert np.array_equal(rg, r) and np.array_equal(ts.words(), st)


@pytest.mark.gpu
def test_gpu_unbalanced_logup_is_an_error(dev, oracle):
    rng = np.random.default_rng(4)
    s, r = A.sender_receiver(4, 2, rng, balanced=False)
    with pytest.raises(sb.SwirlError) as e:
        dev.prove_batch_constraints(sb.Transcript(), 2, 2, 0, to_device_airs(dev, [s, r]))
    assert e.value.code == 10005


@pytest.mark.gpu
def test_gpu_batch_constraints_large_accepted_by_oracle_verifier(dev, oracle):
    # 2^12-row BenchmarkAir (24 columns, 24 constraints, 6 interaction pairs) + a Fibonacci trace
    rng = np.random.default_rng(9)
    airs = sorted_airs([A.benchmark(12, 24, 24, 6, rng), A.fibonacci(10)])
    l_skip, D, pow_bits = 4, 2, 6
    n_max = 12 - l_skip
    st = np.zeros(18, np.uint32)
    ts = sb.Transcript(st)
    proof, r = dev.prove_batch_constraints(ts, l_skip, D, pow_bits, to_device_airs(dev, airs))
    ok, rv = oracle.bc_verify(st, l_skip, D, pow_bits, A.flatten(airs), len(airs), n_max, proof)
    assert ok and np.array_equal(rv, r) and np.array_equal(st, ts.words())


@pytest.mark.gpu
def test_gpu_violated_constraint_proof_equals_oracle_and_is_rejected(dev, oracle):
    """A trace that violates its AIR: the prover still runs (the reference's debug builder is a separate check,
    backend-tests/src/lib.rs disable_debug_builder()), the CUDA proof equals the oracle's word for word, and the oracle's
    verifier rejects it."""
    l_skip, D, pow_bits, airs, n_max, st = setup(oracle, case_fib, seed=5)
    airs[0].common_main[0][7] ^= 1  # break the trace
    ts, stv = sb.Transcript(st), st.copy()
    want, r = oracle.bc_prove(st, l_skip, D, pow_bits, A.flatten(airs), 1, n_max)
    got, rg = dev.prove_batch_constraints(ts, l_skip, D, pow_bits, to_device_airs(dev, airs))
    assert np.array_equal(got, want) and np.array_equal(rg, r)
    assert not oracle.bc_verify(stv, l_skip, D, pow_bits, A.flatten(airs), 1, n_max, got)[0]


# ---- zero interactions (backend-tests/src/lib.rs:378-480) -----------------------------------------------------------------
@pytest.mark.parametrize("log_trace_degree", [0, 1, 2, 3, 4, 6])
def test_oracle_zero_interactions_q0_is_one_and_not_malleable(oracle, log_trace_degree):
    """batch_sumcheck_zero_interactions / ..._malleable_q0: Fibonacci alone at l_skip = 2 (default_test_params_small); the
    sampled r has max(log_height - l_skip, 0) + 1 entries; the honest q0_claim is ONE and any other value is rejected."""
    l_skip, D, pow_bits = 2, 3, 1
    air = A.fibonacci(log_trace_degree)
    n_max = max(log_trace_degree - l_skip, 0)
    flat = A.flatten([air])
    st = np.zeros(18, np.uint32)
    stv = st.copy()
    proof, r = oracle.bc_prove(st, l_skip, D, pow_bits, flat, 1, n_max)
    assert r.shape[0] == n_max + 1
    one = oracle.to_mont(np.array([1], dtype=np.uint64))[0]
    assert list(proof[1:5]) == [one, 0, 0, 0]  # [0] = logup_pow_witness, [1:5] = q0_claim
    ok, rv = oracle.bc_verify(stv, l_skip, D, pow_bits, flat, 1, n_max, proof)
    assert ok and np.array_equal(r, rv)
    bad = proof.copy()
    bad[1] = oracle.to_mont(np.array([2], dtype=np.uint64))[0]
    assert not oracle.bc_verify(np.zeros(18, np.uint32), l_skip, D, pow_bits, flat, 1, n_max, bad)[0]
