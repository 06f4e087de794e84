import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo/scratch')
import numpy as np, time
import test_prove as tp, test_jit_mle as tm, airs as A
from nvrtc_compile import compile_src
airs, order = tp.fixture_airs(2)
for a in airs:
    src = tm.mle_source(a, a.common_main[1], a.common_main[2], tp.D, len(airs))
    rc, log, dt, cub = compile_src(src)
    regs = [l for l in log.split('\n') if 'registers' in l]
    print(a.common_main[1:], rc, "%.2fs" % dt, regs[-1].strip() if regs else '')
for cols in (24, 16, 20):
    air = A.benchmark(3, cols, cols, max(cols // 8, 1), np.random.default_rng(0))
    src = tm.mle_source(air, 1 << 13, cols, 3, 1)
    rc, log, dt, cub = compile_src(src)
    regs = [l for l in log.split('\n') if 'registers' in l]
    print(cols, rc, "%.2fs" % dt, regs[-1].strip() if regs else '')
