"""N > 1 host logic on CPU: world size 2, gloo backend (the GPU run uses nccl with the same code)."""
import os

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    import sys

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib
    from stark_backend_b200 import multi

    oracle = oracle_lib.Oracle(os.path.join(ROOT, "oracle", "libswirl_oracle.so"))
    rng = np.random.default_rng(100 + rank)  # every rank commits its own trace
    trace = (oracle.random_field(rng, 64 * 3), 64, 3)
    root = oracle.stacked_commit(2, 4, 1, 2, [trace], want_codeword=False)[0]
    roots = multi.all_gather_commitments(root)
    t = multi.max_over_ranks(10.0 + rank)
    mine = multi.assign_proofs(5, world, rank)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.concatenate(roots + [np.array([t], np.float64).view(np.uint32),
                                                                          np.array(mine, np.uint32)]))
    dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path, oracle):
    world, port = 2, 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    # both ranks see both commitments in rank order, and they differ (independent traces)
    assert np.array_equal(r0[:16], r1[:16]) and not np.array_equal(r0[:8], r0[8:16])
    for rank in range(world):
        rng = np.random.default_rng(100 + rank)
        trace = (oracle.random_field(rng, 64 * 3), 64, 3)
        assert np.array_equal(r0[8 * rank : 8 * rank + 8], oracle.stacked_commit(2, 4, 1, 2, [trace], want_codeword=False)[0])
    assert r0[16:18].view(np.float64)[0] == 11.0 and r1[16:18].view(np.float64)[0] == 11.0  # max over ranks
    assert list(r0[18:]) == [0, 2, 4] and list(r1[18:]) == [1, 3]


class OracleCommitBackend:
    """The compute steps of multi.sharded_commit done by the CPU oracle (CPU tensors, gloo): the test exercises the
    exchange / index logic that the GPU path shares (DeviceCommitBackend does the same steps through the C ABI)."""

    def __init__(self, oracle):
        self.o = oracle

    def rs_encode(self, trace_slice, height, wl, l_skip, log_blowup):
        import torch

        cw = self.o.rs_code_matrix(l_skip, log_blowup, trace_slice.numpy().view(np.uint32), height, wl)
        return torch.from_numpy(cw.view(np.int32)).view(wl, height << log_blowup)

    def merkle_layers(self, shard, log_rpq):
        import torch

        w, rows = shard.shape
        layers = self.o.merkle_tree(shard.numpy().view(np.uint32).reshape(-1), rows, w, 1 << log_rpq)
        return torch.from_numpy(np.concatenate([l.reshape(-1) for l in layers]).view(np.int32))

    def compress_level(self, level):
        import torch

        lv = level.numpy().view(np.uint32).reshape(-1, 8)
        out = np.stack([self.o.compress(lv[2 * i], lv[2 * i + 1]) for i in range(len(lv) // 2)])
        return torch.from_numpy(out.view(np.int32))


def _sharded_worker(rank, world, port, out_dir, width):
    import sys

    import torch

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib
    from stark_backend_b200 import multi

    oracle = oracle_lib.Oracle(os.path.join(ROOT, "oracle", "libswirl_oracle.so"))
    l_skip, n_stack, log_blowup, k = 2, 4, 1, 2
    H = 1 << (l_skip + n_stack)
    full = oracle.random_field(np.random.default_rng(7), H * width)  # the same matrix on every rank; each keeps its columns
    c0, c1 = multi.column_slice(width, world, rank)
    mine = torch.from_numpy(full[c0 * H:c1 * H].view(np.int32).copy())
    res = multi.sharded_commit(OracleCommitBackend(oracle), mine, H, width, l_skip, log_blowup, k, world, rank)
    np.save(os.path.join(out_dir, f"s{rank}.npy"), np.concatenate([res["root"], res["shard"].numpy().view(np.uint32).reshape(-1)]))
    dist.destroy_process_group()


def _run_sharded(tmp_path, oracle, width):
    world, port = 2, 31500 + (os.getpid() + width) % 2000
    mp.spawn(_sharded_worker, args=(world, port, str(tmp_path), width), nprocs=world, join=True)
    l_skip, n_stack, log_blowup, k = 2, 4, 1, 2
    H = 1 << (l_skip + n_stack)
    full = oracle.random_field(np.random.default_rng(7), H * width)
    root, cw, layers, w = oracle.stacked_commit(l_skip, n_stack, log_blowup, k, [(full, H, width)])
    assert w == width
    N, S = H << log_blowup, (H << log_blowup) >> k
    cw = cw.reshape(width, N)
    for rank in range(world):
        got = np.load(tmp_path / f"s{rank}.npy")
        assert np.array_equal(got[:8], root), "sharded commitment differs from the single-device commitment"
        shard = got[8:].reshape(width, N // world)
        # the shard holds, for every column, the rows q + t S of the rank's queries q, as [t][q']
        sg = S // world
        for t in range(1 << k):
            assert np.array_equal(shard[:, t * sg:(t + 1) * sg], cw[:, t * S + rank * sg:t * S + (rank + 1) * sg])


def test_sharded_commit_world_size_2_gloo(tmp_path, oracle):
    _run_sharded(tmp_path, oracle, 6)


def test_sharded_commit_ragged_columns_world_size_2_gloo(tmp_path, oracle):
    _run_sharded(tmp_path, oracle, 5)


@pytest.mark.gpu
def test_sharded_prover_single_rank_matches_plain_proof(oracle):
    """multi.ShardedProver with one rank: the commitment goes through the sharded path (stack -> column slice -> shard
    tree), the WHIR opening through the external-tree callback; the proof must equal Coordinator.prove's word for word
    (the N > 1 runs of tools/sharded_proof.py check the same on 2-8 GPUs)."""
    import test_prove as tp
    import stark_backend_b200 as sb
    from stark_backend_b200 import multi

    airs = [a for a in tp.fixture_airs(2)[0] if a.preprocessed is None and not a.cached]
    dev = sb.B200Device(0)
    try:
        sp = sb.SystemParams(tp.L_SKIP, tp.N_STACK, tp.LOG_BLOWUP, sb.WhirConfig(**tp.WHIR), tp.LOGUP_POW, tp.D)
        dm = lambda m: sb.DeviceMatrix(dev.h2d(m[0]), m[1], m[2])
        pks = [sb.AirProvingKey(True, None) for _ in airs]
        mk = lambda: [(i, sb.AirProvingContext(a.nodes, a.constraint_idx, a.interactions, a.constraint_degree, a.need_rot,
                                               dm(a.common_main), a.public_values), []) for i, a in enumerate(airs)]
        vk = oracle.to_mont(np.arange(100, 108))
        try:
            plain = sb.Coordinator(dev, sp).prove(vk, pks, mk())
        except sb.SwirlError as e:  # the subset of AIRs may not balance its buses: then the sharded path must fail the same way
            with pytest.raises(sb.SwirlError):
                multi.ShardedProver(dev, sp, 1, 0).prove(vk, pks, mk())
            assert e.code == 10005
            return
        sharded = multi.ShardedProver(dev, sp, 1, 0).prove(vk, pks, mk())
        assert np.array_equal(plain.words(), sharded.words())
    finally:
        dev.close()
