"""Times swirl_rs_encode (C2 shape by default) for several inter-pass scratch sizes / radices.
   python tools/ntt_sweep.py [log_h] [width]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import stark_backend_b200 as sb

log_h = int(sys.argv[1]) if len(sys.argv) > 1 else 20
width = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = sb.B200Device(0)
H = 1 << log_h
x = torch.randint(0, sb.P, (H * width,), dtype=torch.int32, device="cuda")
m = sb.DeviceMatrix(x, H, width)
out = dev.alloc(2 * H * width)
for max_r in (10, 11, 12):
    for mb in (16, 32, 48, 64, 96, 128, 256, 4096):
        dev.set_ntt_plan(max_r, mb << 20)
        for _ in range(2):
            dev.rs_encode(m, 4, 1, out=out)
        dev.synchronize()
        dev.timing_enable(True)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st = dev.torch_stream()
        reps = 5
        with torch.cuda.stream(st):
            a.record(st)
            for _ in range(reps):
                dev.rs_encode(m, 4, 1, out=out)
            b.record(st)
        dev.synchronize()
        t = dev.timing_read()
        dev.timing_enable(False)
        print(json.dumps({"max_log_radix": max_r, "scratch_mb": mb, "ms": a.elapsed_time(b) / reps,
                          "pass_ms": t["ntt_pass"][0] / reps, "final_ms": t["ntt_final"][0] / reps,
                          "chunk_ms": t["chunk"][0] / reps, "launches": (t["ntt_pass"][1] + t["ntt_final"][1]) // reps}), flush=True)
