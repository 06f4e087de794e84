"""Full prove of the BASELINE configs[1] workload (1 BenchmarkAir, 2^log_rows x cols) with per-phase timing.
   python tools/prove_c2.py [log_rows] [cols]"""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import stark_backend_b200 as sb
import airs as A

log_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 20
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = sb.B200Device(0)
air = A.benchmark(3, cols, cols, cols // 8, np.random.default_rng(0))  # DAG only; the trace is replaced below
g = torch.Generator(device="cuda").manual_seed(42)
trace = torch.randint(0, 2, ((1 << log_rows) * cols,), dtype=torch.int32, device="cuda", generator=g) * 0x0FFFFFFE
ctx = sb.AirProvingContext(air.nodes, air.constraint_idx, air.interactions, 2, False, sb.DeviceMatrix(trace, 1 << log_rows, cols))
whir = sb.WhirConfig.new(1, log_rows, 4, 10, 20, 5, 15)
params = sb.SystemParams(4, log_rows - 4, 1, whir, 18, 3)
print("whir queries", whir.num_queries, flush=True)
vk = np.arange(8, dtype=np.uint32)
for rep in range(3):
    ts = sb.Transcript()
    ts.observe(vk)
    torch.cuda.synchronize()
    s0 = dev.sync_stats()
    t0 = time.perf_counter()
    root, pcs = dev.commit(params.pcs(), [ctx.common_main])
    dev.synchronize()
    t1 = time.perf_counter()
    s1 = dev.sync_stats()
    ts.observe(root)
    ts.observe(np.array([A.to_mont(log_rows)], dtype=np.uint32))
    l0 = dev.launch_count()
    bc, r = dev.prove_batch_constraints(ts, 4, 3, 18, [ctx])
    dev.synchronize()
    t2 = time.perf_counter()
    s2 = dev.sync_stats()
    l1 = dev.launch_count()
    st, wh = dev.prove_openings(ts, whir, [pcs], [[False]], r)
    dev.synchronize()
    t3 = time.perf_counter()
    s3 = dev.sync_stats()
    pcs.free()
    print(json.dumps({"log_rows": log_rows, "cols": cols, "commit_ms": (t1 - t0) * 1e3, "batch_constraints_ms": (t2 - t1) * 1e3,
                      "openings_ms": (t3 - t2) * 1e3, "total_ms": (t3 - t0) * 1e3, "bc_launches": l1 - l0,
                      "syncs": {"commit": s1[0] - s0[0], "bc": s2[0] - s1[0], "open": s3[0] - s2[0]},
                      "sync_wait_ms": {"commit": s1[1] - s0[1], "bc": s2[1] - s1[1], "open": s3[1] - s2[1]},
                      "open_launches": dev.launch_count() - l1, "proof_words": int(bc.size + st.size + wh.size),
                      "cells_per_s": (1 << log_rows) * cols / (t3 - t0)}), flush=True)
