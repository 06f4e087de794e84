"""Several Coordinators proving concurrently on one GPU, each on its own OS thread, library context and stream — the
reference's concurrency contract ("multiple Coordinators may run concurrently on different OS threads /
streams", crates/cuda-backend/examples/keccakf.rs).  The latency-bound sumcheck phases of one proof overlap the
hash-bound commit of another.   python tools/concurrent_proofs.py [threads] [proofs_per_thread] [log_rows] [cols]"""
import json, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import stark_backend_b200 as sb
import airs as A

n_threads = int(sys.argv[1]) if len(sys.argv) > 1 else 2
per_thread = int(sys.argv[2]) if len(sys.argv) > 2 else 6
log_rows = int(sys.argv[3]) if len(sys.argv) > 3 else 20
cols = int(sys.argv[4]) if len(sys.argv) > 4 else 256
air = A.benchmark(3, cols, cols, cols // 8, np.random.default_rng(0))
whir = sb.WhirConfig.new(1, log_rows, 4, 10, 20, 5, 15)
params = sb.SystemParams(4, log_rows - 4, 1, whir, 18, 3)
vk = np.arange(8, dtype=np.uint32)
g = torch.Generator(device="cuda").manual_seed(42)
trace = torch.randint(0, 2, ((1 << log_rows) * cols,), dtype=torch.int32, device="cuda", generator=g) * 0x0FFFFFFE
torch.cuda.synchronize()

def make_worker():
    dev = sb.B200Device(0)
    ctx = sb.AirProvingContext(air.nodes, air.constraint_idx, air.interactions, 2, False, sb.DeviceMatrix(trace, 1 << log_rows, cols))
    def prove():
        proof = sb.Coordinator(dev, params).prove(vk, [sb.AirProvingKey(True, None)], [(0, ctx, [])])
        proof.common_main_pcs.free()
        return proof
    return dev, prove

def run(k):
    workers = [make_worker() for _ in range(k)]
    for _, prove in workers:      # warm-up (allocations, twiddles)
        for _ in range(2):
            prove()
    torch.cuda.synchronize()
    roots, errs = [None] * k, []
    def body(i):
        try:
            with torch.cuda.stream(workers[i][0].torch_stream()):
                for _ in range(per_thread):
                    roots[i] = workers[i][1]().words()
        except Exception as e:  # noqa
            errs.append(repr(e))
    th = [threading.Thread(target=body, args=(i,)) for i in range(k)]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    for dev, _ in workers:
        dev.close()
    same = all(np.array_equal(roots[0], r) for r in roots)
    return dict(threads=k, proofs=k * per_thread, seconds=dt, ms_per_proof_throughput=1e3 * dt / (k * per_thread),
                cells_per_s=k * per_thread * (1 << log_rows) * cols / dt, identical_proofs=same, errors=errs)

for k in sorted({1, n_threads}):
    print(json.dumps(run(k)), flush=True)
