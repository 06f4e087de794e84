"""Times swirl_gkr_fractional_sumcheck on 2^log_n random balanced leaves.  python tools/gkr_bench.py [log_n ...]"""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import stark_backend_b200 as sb

dev = sb.B200Device(0)
for log_n in [int(a) for a in sys.argv[1:]] or [16, 20, 24]:
    n = 1 << log_n
    g = torch.Generator(device="cuda").manual_seed(log_n)
    half = torch.randint(0, sb.P, (n // 2, 8), dtype=torch.int32, device="cuda", generator=g)
    half[:, 1:4] = 0
    neg = half.clone()
    neg[:, 0] = (sb.P - half[:, 0]) % sb.P
    leaves = torch.stack([half, neg], dim=1).reshape(n, 8).contiguous()
    for rep in range(3):
        ts = sb.Transcript()
        torch.cuda.synchronize()
        l0 = dev.launch_count()
        t = time.perf_counter()
        out = dev.gkr_fractional_sumcheck(ts, leaves, log_n, True)
        dt = time.perf_counter() - t
    print(json.dumps({"log_n": log_n, "ms": dt * 1e3, "leaves_per_s": n / dt, "launches": dev.launch_count() - l0,
                      "alg_gb_s": (n * 32 * 4) / dt / 1e9}), flush=True)
