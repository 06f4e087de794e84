import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import stark_backend_b200 as sb
import bench
dev = sb.B200Device(0)
air = bench.benchmark_air_dag(256)
whir = sb.WhirConfig(4, bench.whir_queries(20), 15, 20, 5)
params = sb.SystemParams(4, 16, 1, whir, 18, 3)
host = torch.from_numpy((np.random.default_rng(42).integers(0, 2, size=bench.CELLS, dtype=np.uint64) * bench.R1).astype(np.uint32).view(np.int32)).pin_memory()
air_h = sb.AirProvingContext(air.nodes, air.constraint_idx, air.interactions, 2, False, None)
pk = [sb.AirProvingKey(True, None)]
vk = np.arange(8, dtype=np.uint32)
for rep in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    c = sb.Coordinator(dev, params)
    root, common = dev.commit_host(params.pcs(), [(host, 1 << 20, 256)])
    dev.synchronize()
    t1 = time.perf_counter()
    from stark_backend_b200.backend import PcsTraceView, AirProvingContext
    view = PcsTraceView(common, dev.lib.swirl_pcs_stacked_matrix(common._h), 1 << 20, 256)
    ctx = AirProvingContext(air.nodes, air.constraint_idx, air.interactions, 2, False, view, air.public_values)
    proof = c.prove(vk, pk, [(0, ctx, [])], precommitted=(root, common))
    dev.synchronize()
    t2 = time.perf_counter()
    proof.common_main_pcs.free()
    print(json.dumps({"commit_host_ms": (t1 - t0) * 1e3, "rest_ms": (t2 - t1) * 1e3}), flush=True)
