"""Round link (csrc/ext.cuh): the round kernels of a sumcheck enqueued up front, results and
challenges exchanged with the host transcript through a mapped mailbox instead of one launch + cudaStreamSynchronize per
round (the reference's pattern: cuda-backend/src/logup_zerocheck/fractional.rs:649-, sponge.rs:267-300).  The proof must not
change by a bit, and the stream synchronisations must really be gone."""
import numpy as np
import pytest

import airs as A
import stark_backend_b200 as sb


def _prove_benchmark(dev, log_rows, cols, seed):
    import torch

    air = A.benchmark(3, cols, cols, max(cols // 8, 1), np.random.default_rng(0))
    g = torch.Generator(device="cuda").manual_seed(seed)
    trace = torch.randint(0, 2, ((1 << log_rows) * cols,), dtype=torch.int32, device="cuda", generator=g) * 0x0FFFFFFE
    whir = sb.WhirConfig.new(1, log_rows, 4, 8 if log_rows > 12 else 4, 6, 3, 4)
    params = sb.SystemParams(4, log_rows - 4, 1, whir, 4, 3)
    ctx = sb.AirProvingContext(air.nodes, air.constraint_idx, air.interactions, 2, False, sb.DeviceMatrix(trace, 1 << log_rows, cols))
    proof = sb.Coordinator(dev, params).prove(np.arange(8, dtype=np.uint32), [sb.AirProvingKey(True, None)], [(0, ctx, [])])
    words = proof.words().copy()
    proof.common_main_pcs.free()
    return words


@pytest.mark.gpu
@pytest.mark.parametrize("log_rows,cols", [(10, 8), (14, 24), (17, 16)])
def test_linked_rounds_give_the_same_proof_without_stream_syncs(log_rows, cols):
    out, syncs, links = {}, {}, {}
    for on in (False, True):
        dev = sb.B200Device(0)
        try:
            dev.set_round_link(on)
            _prove_benchmark(dev, log_rows, cols, 1)  # warm-up: twiddles, program cache, scratch
            s0, l0 = dev.sync_stats()[0], dev.link_stats()
            out[on] = _prove_benchmark(dev, log_rows, cols, 2)
            syncs[on], links[on] = dev.sync_stats()[0] - s0, dev.link_stats() - l0
        finally:
            dev.close()
    assert np.array_equal(out[False], out[True])
    assert links[False] == 0 and links[True] > 0
    # every linked round is one stream synchronisation less
    assert syncs[True] + links[True] <= syncs[False] + 2, (syncs, links)
    assert syncs[True] < syncs[False]


@pytest.mark.gpu
def test_whole_fixture_proof_linked_equals_oracle(oracle):
    """The 5-AIR fixture (preprocessed + cached commitments, interactions, rotations, an optional AIR; several height
    classes, so AIRs leave the MLE rounds at different times) through the linked rounds: the oracle's proof."""
    import test_prove as tp

    airs, order = tp.fixture_airs(2)
    is_required = [True, True, True, True, False]
    vk = oracle.to_mont(np.arange(100, 108))
    want = tp.oracle_prove(oracle, airs, order, is_required, vk)
    sp = sb.SystemParams(tp.L_SKIP, tp.N_STACK, tp.LOG_BLOWUP, sb.WhirConfig(**tp.WHIR), tp.LOGUP_POW, tp.D)
    for on in (True, False):
        dev = sb.B200Device(0)
        try:
            dev.set_round_link(on)
            dm = lambda m: sb.DeviceMatrix(dev.h2d(m[0]), m[1], m[2])

            def committed(m):
                mat = dm(m)
                r, data = dev.commit(sp.pcs(), [mat])
                return sb.CommittedTraceData(r, mat, data)

            pks, per_trace = [], []
            for air_id, a in enumerate(airs):
                prep = committed(a.preprocessed) if a.preprocessed is not None else None
                cached = [committed(c) for c in a.cached]
                pks.append(sb.AirProvingKey(is_required[air_id], prep))
                per_trace.append((air_id, sb.AirProvingContext(a.nodes, a.constraint_idx, a.interactions, a.constraint_degree, a.need_rot,
                                                               dm(a.common_main), a.public_values, [c.trace for c in cached],
                                                               prep.trace if prep else None), cached))
            proof = sb.Coordinator(dev, sp).prove(vk, pks, per_trace)
            got = [proof.common_main_commit.copy(), proof.constraints_proof.copy(), proof.stacking_proof.copy(), proof.whir_proof.copy()]
        finally:
            dev.close()
        for g, key in zip(got, ("root", "bc", "stacking", "whir")):
            assert np.array_equal(g, want[key]), f"round_link={on}: {key}"
