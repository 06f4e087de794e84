"""One commitment sharded by columns over the GPUs of a node (SURVEY §8e commit row, stark-backend_b200/multi.py):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/sharded_commit.py [log_rows cols]
Every rank RS-encodes its column slice, one NCCL all-to-all moves row segments, every rank hashes the rows of its queries,
the G sub-roots are all-gathered.  Rank 0 also commits the whole matrix alone and checks that the roots are equal."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import stark_backend_b200 as sb
from stark_backend_b200 import multi

log_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 20
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 256
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
real_stdout = os.dup(1); os.dup2(2, 1)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
dev = sb.B200Device(local)
L_SKIP, LOG_BLOWUP, K = 4, 1, 4
H = 1 << log_rows

def column(c):  # column c of the common matrix, the same on every rank
    g = torch.Generator(device="cuda").manual_seed(1000 + c)
    return torch.randint(0, sb.P, (H,), dtype=torch.int32, device="cuda", generator=g)

c0, c1 = multi.column_slice(cols, world, rank)
mine = torch.cat([column(c) for c in range(c0, c1)])
backend = multi.DeviceCommitBackend(dev)
use_peer = os.environ.get("SHARDED_EXCHANGE", "peer") == "peer"
px = multi.PeerExchange(dev, cols, H << LOG_BLOWUP, world, rank) if use_peer else None
def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
times = []
stream = dev.torch_stream()
for it in range(6):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    a.record(stream)
    res = multi.sharded_commit(backend, mine, H, cols, L_SKIP, LOG_BLOWUP, K, world, rank, peer_exchange=px)
    torch.cuda.synchronize()
    b.record(stream)
    barrier()
    times.append(a.elapsed_time(b))  # device time on the library's stream; every step is host-synchronised
    if it < 5:
        del res
ms = multi.max_over_ranks(min(times[1:]), dev.torch_device)

# size-independent check on every rank: two of its own queries hash up to the global root (rows from the shard, path =
# local subtree path + the top levels over the gathered sub-roots), recomputed with the CPU oracle
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib
oracle = oracle_lib.Oracle(os.path.join(ROOT, "oracle", "libswirl_oracle.so"))
N = H << LOG_BLOWUP
Sg = (N >> K) // world
shard = res["shard"]
paths_ok = True
for ql in (0, Sg - 1):
    rows = dev.matrix_open_rows(shard.data_ptr(), N // world, cols, Sg, K, [ql])[0]
    path = dev.merkle_query_proofs(res["layers"].data_ptr(), Sg, [ql])[0]
    nodes = [oracle.hash_slice(rows[t]) for t in range(1 << K)]
    while len(nodes) > 1:
        nodes = [oracle.compress(nodes[2 * j], nodes[2 * j + 1]) for j in range(len(nodes) // 2)]
    cur, i = nodes[0], ql
    for sib in path:
        cur = oracle.compress(cur, sib) if i % 2 == 0 else oracle.compress(sib, cur)
        i >>= 1
    level, i = [res["sub_roots"][r] for r in range(world)], rank
    while len(level) > 1:
        cur = oracle.compress(cur, level[i ^ 1]) if i % 2 == 0 else oracle.compress(level[i ^ 1], cur)
        level = [oracle.compress(level[2 * j], level[2 * j + 1]) for j in range(len(level) // 2)]
        i >>= 1
    paths_ok &= bool(np.array_equal(cur, res["root"]))
paths_ok = multi.max_over_ranks(0.0 if paths_ok else 1.0, dev.torch_device) == 0.0
out = {"config": f"sharded commit 2^{log_rows} x {cols}, blowup 2, k_whir 4", "n_gpus": world, "exchange": "peer-memory scatter kernel (NVLink stores)" if use_peer else "pack + NCCL all_to_all_single",
       "sharded_commit_ms": ms,
       "cells_per_s": H * cols / (ms / 1e3)}
out["merkle_paths_verify_on_every_rank"] = paths_ok
if rank == 0 and H * cols * 4 * 4 > 150e9:  # the whole matrix + codeword do not fit one GPU: no single-GPU comparison
    os.write(real_stdout, (json.dumps(out) + "\n").encode())
elif rank == 0:
    full = torch.cat([column(c) for c in range(cols)])
    params = sb.PcsParams(L_SKIP, log_rows - L_SKIP, LOG_BLOWUP, K)
    single = []
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        root, pcs = dev.commit(params, [sb.DeviceMatrix(full, H, cols)])
        dev.synchronize(); single.append(1e3 * (time.perf_counter() - t0))
        pcs.free()
    out.update(single_gpu_commit_ms=min(single), roots_equal=bool(np.array_equal(root, res["root"])),
               speedup=min(single) / ms)
    os.write(real_stdout, (json.dumps(out) + "\n").encode())
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
