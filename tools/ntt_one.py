"""One rs_encode of the C2 shape (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import stark_backend_b200 as sb
log_h = int(sys.argv[1]) if len(sys.argv) > 1 else 20
width = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = sb.B200Device(0)
H = 1 << log_h
x = torch.randint(0, sb.P, (H * width,), dtype=torch.int32, device="cuda")
m = sb.DeviceMatrix(x, H, width)
out = dev.alloc(2 * H * width)
dev.set_ntt_plan(11, 4096 << 20)
for _ in range(3):
    dev.rs_encode(m, 4, 1, out=out)
dev.synchronize()
