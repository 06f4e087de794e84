"""Whole proof (Coordinator::prove, prover/mod.rs:104-198): commit -> batch constraints -> stacked
reduction -> WHIR through the C ABI, against the same sequence composed from the oracle, and
through the oracle's verifier chain (verifier/mod.rs order: batch constraints, stacked reduction, WHIR)."""
import numpy as np
import pytest

import airs as A
import stark_backend_b200 as sb

P = sb.P
WHIR = dict(k=2, num_queries=[6, 5], mu_pow_bits=3, query_phase_pow_bits=4, folding_pow_bits=2)
L_SKIP, N_STACK, LOG_BLOWUP, D, LOGUP_POW = 2, 5, 1, 3, 2


def fixture_airs(seed):
    rng = np.random.default_rng(seed)
    s, r = A.sender_receiver(5, 2, rng)
    airs = [A.fibonacci(6), A.with_parts(5, rng), s, r, A.benchmark(3, 4, 4, 2, rng)]
    # air ids = positions; one optional AIR (id 4) is present, ids are kept when sorting
    order = sorted(range(len(airs)), key=lambda i: (-airs[i].height, i))
    return airs, order


def mont1(x):
    return np.array([A.to_mont(x)], dtype=np.uint32)


def oracle_prove(oracle, airs, order, is_required, vk_pre_hash):
    st = np.zeros(18, np.uint32)
    oracle.sponge_observe(st, vk_pre_hash)
    sorted_airs = [airs[i] for i in order]
    root, _, _, _ = oracle.stacked_commit(L_SKIP, N_STACK, LOG_BLOWUP, WHIR["k"], [a.common_main for a in sorted_airs], want_codeword=False)
    oracle.sponge_observe(st, root)
    commit1 = lambda m: oracle.stacked_commit(L_SKIP, N_STACK, LOG_BLOWUP, WHIR["k"], [m], want_codeword=False)[0]
    for air_id, a in enumerate(airs):
        if not is_required[air_id]:
            oracle.sponge_observe(st, mont1(1))
        if a.preprocessed is not None:
            oracle.sponge_observe(st, commit1(a.preprocessed))
        else:
            oracle.sponge_observe(st, mont1(a.height.bit_length() - 1))
        for c in a.cached:
            oracle.sponge_observe(st, commit1(c))
        oracle.sponge_observe(st, a.public_values)
    n_max = max(sorted_airs[0].height.bit_length() - 1 - L_SKIP, 0)
    bc, r = oracle.bc_prove(st, L_SKIP, D, LOGUP_POW, A.flatten(sorted_airs), len(airs), n_max)
    commits = [[m + (a.need_rot,) for a in sorted_airs for m in [a.common_main]]]
    for a in sorted_airs:
        for m in ([a.preprocessed] if a.preprocessed is not None else []) + a.cached:
            commits.append([m + (a.need_rot,)])
    stacking, u, sw = oracle.stacked_reduction_prove(st, L_SKIP, N_STACK, commits, r)
    u_cube = [u[0]]
    for _ in range(L_SKIP - 1):
        u_cube.append(oracle.ef_mul(u_cube[-1], u_cube[-1]))
    u_cube = np.array(u_cube + list(u[1:]), dtype=np.uint32)
    mats = []
    for c in commits:
        flat, w = oracle.stacked_matrix(L_SKIP, N_STACK, [(v, h, wd) for v, h, wd, _ in c])
        mats.append((flat, w))
    roots, whir = oracle.whir_prove(st, L_SKIP, LOG_BLOWUP, WHIR, mats, 1 << (L_SKIP + N_STACK), u_cube)
    return dict(root=root, bc=bc, r=r, stacking=stacking, u=u, u_cube=u_cube, whir=whir, st=st, commits=commits, roots=roots,
                widths=[w for _, w in mats], n_max=n_max, sorted_airs=sorted_airs)


def test_oracle_whole_proof_verifies(oracle):
    import verify_chain

    airs, order = fixture_airs(2)
    is_required = [True, True, True, True, False]
    vk = oracle.to_mont(np.arange(100, 108))
    pr = oracle_prove(oracle, airs, order, is_required, vk)
    commit1 = lambda m: oracle.stacked_commit(L_SKIP, N_STACK, LOG_BLOWUP, WHIR["k"], [m], want_codeword=False)[0]
    pre_cached = [(commit1(a.preprocessed) if a.preprocessed is not None else None, [commit1(c) for c in a.cached]) for a in airs]
    ok, st = verify_chain.verify(oracle, L_SKIP, N_STACK, LOG_BLOWUP, D, LOGUP_POW, WHIR, vk, airs, is_required, pr["root"],
                                 pre_cached, pr["bc"], pr["stacking"], pr["whir"])
    assert ok is True, st
    assert np.array_equal(st, pr["st"])
    # a proof with one flipped word in any part is rejected at that stage
    for key, stage in (("bc", "batch_constraints"), ("stacking", "stacked_reduction"), ("whir", "whir")):
        bad = {k: pr[k].copy() for k in ("bc", "stacking", "whir")}
        bad[key][5] ^= 1
        ok, where = verify_chain.verify(oracle, L_SKIP, N_STACK, LOG_BLOWUP, D, LOGUP_POW, WHIR, vk, airs, is_required,
                                        pr["root"], pre_cached, bad["bc"], bad["stacking"], bad["whir"])
        assert ok is False and where == stage, (key, where)


@pytest.mark.gpu
def test_gpu_whole_proof_matches_oracle(dev, oracle):
    airs, order = fixture_airs(2)
    is_required = [True, True, True, True, False]
    vk = oracle.to_mont(np.arange(100, 108))
    want = oracle_prove(oracle, airs, order, is_required, vk)
    params = sb.SystemParams(L_SKIP, N_STACK, LOG_BLOWUP, sb.WhirConfig(**WHIR), LOGUP_POW, D)
    dm = lambda m: sb.DeviceMatrix(dev.h2d(m[0]), m[1], m[2])

    def committed(m):
        mat = dm(m)
        root, data = dev.commit(params.pcs(), [mat])
        return sb.CommittedTraceData(root, mat, data)

    per_air_pk, per_trace = [], []
    for air_id, a in enumerate(airs):
        prep = committed(a.preprocessed) if a.preprocessed is not None else None
        cached = [committed(c) for c in a.cached]
        per_air_pk.append(sb.AirProvingKey(is_required[air_id], prep))
        ctx = sb.AirProvingContext(a.nodes, a.constraint_idx, a.interactions, a.constraint_degree, a.need_rot, dm(a.common_main),
                                   a.public_values, [c.trace for c in cached], prep.trace if prep else None)
        per_trace.append((air_id, ctx, cached))
    coord = sb.Coordinator(dev, params)
    proof = coord.prove(vk, per_air_pk, per_trace[::-1])  # any input order: the coordinator sorts
    assert np.array_equal(proof.common_main_commit, want["root"])
    assert np.array_equal(proof.constraints_proof, want["bc"])
    assert np.array_equal(proof.r, want["r"])
    assert np.array_equal(proof.stacking_proof, want["stacking"])
    assert np.array_equal(proof.whir_proof, want["whir"])
    assert np.array_equal(coord.transcript.words(), want["st"])


@pytest.mark.gpu
def test_gpu_concurrent_coordinators_on_threads(oracle):
    """Coordinators on separate OS threads, each with its own library context and stream, prove at the same time and every
    proof equals the oracle's (the reference runs concurrent provers the same way: cuda-backend/examples/keccakf.rs with
    NUM_THREADS=3 in CI)."""
    import threading

    airs, order = fixture_airs(2)
    is_required = [True, True, True, True, False]
    vk = oracle.to_mont(np.arange(100, 108))
    want = oracle_prove(oracle, airs, order, is_required, vk)
    params = sb.SystemParams(L_SKIP, N_STACK, LOG_BLOWUP, sb.WhirConfig(**WHIR), LOGUP_POW, D)
    results, errors = {}, []

    def worker(tid):
        try:
            dev = sb.B200Device(0)
            dm = lambda m: sb.DeviceMatrix(dev.h2d(m[0]), m[1], m[2])

            def committed(m):
                mat = dm(m)
                root, data = dev.commit(params.pcs(), [mat])
                return sb.CommittedTraceData(root, mat, data)

            for rep in range(3):
                per_air_pk, per_trace = [], []
                for air_id, a in enumerate(airs):
                    prep = committed(a.preprocessed) if a.preprocessed is not None else None
                    cached = [committed(c) for c in a.cached]
                    per_air_pk.append(sb.AirProvingKey(is_required[air_id], prep))
                    ctx = sb.AirProvingContext(a.nodes, a.constraint_idx, a.interactions, a.constraint_degree, a.need_rot,
                                               dm(a.common_main), a.public_values, [c.trace for c in cached], prep.trace if prep else None)
                    per_trace.append((air_id, ctx, cached))
                results[(tid, rep)] = sb.Coordinator(dev, params).prove(vk, per_air_pk, per_trace).words()
            dev.close()
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(3)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    expect = np.concatenate([want["root"], want["bc"], want["stacking"], want["whir"]])
    assert len(results) == 9 and all(np.array_equal(w, expect) for w in results.values())
