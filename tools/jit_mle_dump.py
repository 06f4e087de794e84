"""The run-time compiled MLE-round kernel of an AIR, without a GPU: prints the generated CUDA C++ (or its sub-program
listing), compiles it with NVRTC for sm_100a and reports ptxas' resource usage and the compile time.
   python tools/jit_mle_dump.py [benchmark COLS | fixture] [D] [--source]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import airs as A
import test_jit_mle as tm


def nvrtc_compile(src):
    nv = C.CDLL("libnvrtc.so.12")
    prog = C.c_void_p()
    assert nv.nvrtcCreateProgram(C.byref(prog), src.encode(), b"jit.cu", 0, None, None) == 0
    o = [b"--gpu-architecture=sm_100a", b"--std=c++17", b"-lineinfo", b"--extra-device-vectorization", b"--ptxas-options=-v"]
    t = time.time()
    rc = nv.nvrtcCompileProgram(prog, len(o), (C.c_char_p * len(o))(*o))
    dt = time.time() - t
    n = C.c_size_t()
    nv.nvrtcGetProgramLogSize(prog, C.byref(n))
    log = C.create_string_buffer(n.value + 1)
    nv.nvrtcGetProgramLog(prog, log)
    size = 0
    if rc == 0:
        nv.nvrtcGetCUBINSize(prog, C.byref(n))
        size = n.value
    return rc, log.value.decode(), dt, size


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    what = args[0] if args else "benchmark"
    if what == "fixture":
        import test_prove as tp

        airs = [(a, a.common_main[1], a.common_main[2]) for a in tp.fixture_airs(2)[0]]
        D = int(args[1]) if len(args) > 1 else tp.D
    else:
        cols = int(args[1]) if len(args) > 1 else 256
        airs = [(A.benchmark(3, cols, cols, max(cols // 8, 1), np.random.default_rng(0)), 1 << 20, cols)]
        D = int(args[2]) if len(args) > 2 else 3
    for air, h, w in airs:
        src = tm.mle_source(air, h, w, D, len(airs))
        if not src:
            print(f"{h} x {w}: the generator declined (interpreter)")
            continue
        subs = tm.listing(src)
        if "--source" in sys.argv:
            print(src[src.index("SW_MLE_SIGNATURE(swirl_mle_jit)"):src.index("// SUB 0")])
        rc, log, dt, size = nvrtc_compile(src)
        usage = [l.strip() for l in log.split("\n") if "registers" in l or "spill" in l]
        print(f"{h} x {w}, D = {D}: {len(subs)} sub-programs, {len({s[0] for s in subs})} cases, NVRTC rc {rc} in {dt:.2f} s, cubin {size} B")
        for u in usage:
            print("   ", u)


if __name__ == "__main__":
    main()
