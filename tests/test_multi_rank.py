"""N > 1 host logic on CPU: world size 2, gloo backend (the GPU run uses nccl with the same code)."""
import os

import numpy as np
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
