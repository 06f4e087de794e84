"""cache_rs_code_matrix = false (reference cuda-backend/src/device.rs:113-121: the codeword is not kept; streamed at commit
time, re-encoded for the WHIR openings) and the device-memory model behind the choice (memory.py, after the reference's
memory_metering.rs)."""
import numpy as np
import pytest

import stark_backend_b200 as sb
from stark_backend_b200 import memory as M


def test_memory_model_shape():
    # C4 (BASELINE configs[3]): 2^24 x 512, blowup 2 -- the cached codeword alone is 64 GiB; streaming removes it
    counts = M.ProvingMemoryCounts(main_cells_without_rot=(1 << 24) * 512, interaction_cells=(1 << 24) * 8)
    cfg = M.ProvingMemoryConfig(l_skip=4, log_stacked_height=24, log_blowup=1, k_whir=4, cache_rs_code_matrix=True,
                                stacked_aliases_trace=True)
    cached = M.estimate(cfg, counts)
    assert cached.rs_code_matrix == 64 << 30 and cached.main == 32 << 30 and cached.stacked_matrix == 0
    cfg.cache_rs_code_matrix = False
    streamed = M.estimate(cfg, counts)
    assert streamed.rs_code_matrix == 0 and streamed.total < cached.total - (50 << 30)
    # the planner keeps the codeword when it fits and drops it when it does not
    assert M.choose_cache_rs_code_matrix(cfg, counts, 170 << 30)[0] is True
    assert M.choose_cache_rs_code_matrix(cfg, counts, 80 << 30)[0] is False
    # monotone in every count
    more = M.ProvingMemoryCounts(main_cells_without_rot=(1 << 24) * 512, interaction_cells=(1 << 24) * 16)
    assert M.estimate(cfg, more).total >= streamed.total


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(4, 10, 1, 4, 100), (2, 7, 2, 2, 37), (4, 12, 1, 4, 33)])
def test_streamed_commit_equals_cached_commit(oracle, shape):
    """Root and every digest layer of a commitment are the same whether the codeword is kept or streamed through the
    32-column window (widths that are not multiples of the window or of the sponge rate included)."""
    l_skip, n_stack, lb, k, w = shape
    H = 1 << (l_skip + n_stack)
    params = sb.PcsParams(l_skip, n_stack, lb, k)
    dev = sb.B200Device(0)
    try:
        rng = np.random.default_rng(w)
        vals = oracle.random_field(rng, H * w)
        root_c, pcs_c = dev.commit(params, [sb.DeviceMatrix(dev.h2d(vals), H, w)])
        layers_c = np.concatenate([l.reshape(-1) for l in pcs_c.tree.digest_layers()])
        idx = [0, 1, (H << lb >> k) - 1, 5 % (H << lb >> k)]
        rows_c = pcs_c.tree.get_opened_rows(idx)
        dev.set_cache_rs_code_matrix(False)
        root_s, pcs_s = dev.commit(params, [sb.DeviceMatrix(dev.h2d(vals), H, w)])
        assert pcs_s.tree.codeword_ptr in (None, 0)
        assert np.array_equal(root_s, root_c)
        assert np.array_equal(np.concatenate([l.reshape(-1) for l in pcs_s.tree.digest_layers()]), layers_c)
        assert np.array_equal(root_c, oracle.stacked_commit(l_skip, n_stack, lb, k, [(vals, H, w)], want_codeword=False)[0])
        assert np.array_equal(pcs_s.open_rows(idx), rows_c)  # re-encoded by column groups
        pcs_c.free()
        pcs_s.free()
    finally:
        dev.close()


@pytest.mark.gpu
def test_whole_proof_without_cached_codeword_matches_oracle(oracle):
    import test_prove as tp

    airs, order = tp.fixture_airs(2)
    is_required = [True, True, True, True, False]
    vk = oracle.to_mont(np.arange(100, 108))
    want = tp.oracle_prove(oracle, airs, order, is_required, vk)
    params = sb.SystemParams(tp.L_SKIP, tp.N_STACK, tp.LOG_BLOWUP, sb.WhirConfig(**tp.WHIR), tp.LOGUP_POW, tp.D)
    dev = sb.B200Device(0)
    try:
        dev.set_cache_rs_code_matrix(False)
        dm = lambda m: sb.DeviceMatrix(dev.h2d(m[0]), m[1], m[2])

        def committed(m):
            mat = dm(m)
            root, data = dev.commit(params.pcs(), [mat])
            return sb.CommittedTraceData(root, mat, data)

        pks, per_trace = [], []
        for air_id, a in enumerate(airs):
            prep = committed(a.preprocessed) if a.preprocessed is not None else None
            cached = [committed(c) for c in a.cached]
            pks.append(sb.AirProvingKey(is_required[air_id], prep))
            per_trace.append((air_id, sb.AirProvingContext(a.nodes, a.constraint_idx, a.interactions, a.constraint_degree, a.need_rot,
                                                           dm(a.common_main), a.public_values, [c.trace for c in cached],
                                                           prep.trace if prep else None), cached))
        proof = sb.Coordinator(dev, params).prove(vk, pks, per_trace)
        assert np.array_equal(proof.words(), np.concatenate([want["root"], want["bc"], want["stacking"], want["whir"]]))
    finally:
        dev.close()


@pytest.mark.gpu
def test_memory_model_bounds_measured_peak(oracle):
    """One BenchmarkAir 2^16 x 64 proof in both modes: the arena's high-water mark stays below the modelled peak (the model
    is an upper bound used for planning) and streaming lowers the measured peak by about the codeword."""
    import airs as A
    import torch

    log_rows, cols = 16, 64
    air = A.benchmark(3, cols, cols, cols // 8, np.random.default_rng(0))
    g = torch.Generator(device="cuda").manual_seed(1)
    trace = torch.randint(0, 2, ((1 << log_rows) * cols,), dtype=torch.int32, device="cuda", generator=g) * 0x0FFFFFFE
    whir = sb.WhirConfig.new(1, log_rows, 4, 10, 8, 3, 4)
    params = sb.SystemParams(4, log_rows - 4, 1, whir, 4, 3)
    peaks, proofs = {}, {}
    for cache in (True, False):
        dev = sb.B200Device(0)
        try:
            dev.set_cache_rs_code_matrix(cache)
            ctx = sb.AirProvingContext(air.nodes, air.constraint_idx, air.interactions, 2, False, sb.DeviceMatrix(trace, 1 << log_rows, cols))
            dev.mem_stats(reset_peak=True)
            proof = sb.Coordinator(dev, params).prove(np.arange(8, dtype=np.uint32), [sb.AirProvingKey(True, None)], [(0, ctx, [])])
            peaks[cache] = dev.mem_stats()["peak"]
            proofs[cache] = proof.words()
            counts = M.ProvingMemoryCounts.from_airs([ctx], 4)
            est = M.estimate(M.ProvingMemoryConfig(4, log_rows, 1, 4, cache, True), counts, include_main=False)
            assert peaks[cache] <= est.total, (cache, peaks[cache], est)
            assert est.total <= 3 * peaks[cache] + (512 << 20), (cache, peaks[cache], est)
            proof.common_main_pcs.free()
        finally:
            dev.close()
    assert np.array_equal(proofs[True], proofs[False])
    codeword = 4 * 2 * (1 << log_rows) * cols
    assert peaks[True] - peaks[False] >= codeword // 2
