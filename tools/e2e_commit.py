"""Times the host->commit path: plain H2D, H2D + commit (serial), swirl_commit_host (pipelined)."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import stark_backend_b200 as sb
from stark_backend_b200.lib import check
dev = sb.B200Device(0)
log_rows, cols = 20, 256
n = (1 << log_rows) * cols
host = torch.randint(0, 2, (n,), dtype=torch.int32).pin_memory()
d = torch.empty(n, dtype=torch.int32, device="cuda")
params = sb.PcsParams(4, log_rows - 4, 1, 4)
def t(fn, reps=5):
    out = []
    for _ in range(reps):
        torch.cuda.synchronize(); dev.synchronize()
        t0 = time.perf_counter(); fn(); dev.synchronize(); torch.cuda.synchronize()
        out.append((time.perf_counter() - t0) * 1e3)
    return [round(x, 2) for x in out]
def h2d():
    check(dev.lib.swirl_memcpy_h2d(dev.ctx, d.data_ptr(), host.data_ptr(), 4 * n))
def serial():
    h2d()
    r, p = dev.commit(params, [sb.DeviceMatrix(d, 1 << log_rows, cols)]); p.free()
def piped():
    r, p = dev.commit_host(params, [(host, 1 << log_rows, cols)]); p.free()
def devonly():
    r, p = dev.commit(params, [sb.DeviceMatrix(d, 1 << log_rows, cols)]); p.free()
print(json.dumps({"h2d_ms": t(h2d), "commit_dev_ms": t(devonly), "h2d_then_commit_ms": t(serial), "commit_host_pipelined_ms": t(piped)}))
