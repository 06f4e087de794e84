"""Per-AIR run-time compiled constraint kernels (csrc/jit.{hpp,cu}, jit_prelude.cuh): the round-0
program of an AIR emitted as CUDA C++ (repeating instruction runs re-rolled into loops), compiled with NVRTC and launched
in place of the interpreter kernel.  Proofs must not change by a bit."""
import ctypes as C

import numpy as np
import pytest

import airs as A
import stark_backend_b200 as sb
from stark_backend_b200 import lib as L


class _Shape:
    def __init__(self, h, w):
        self.h, self.w = h, w

    def height(self):
        return self.h

    def width(self):
        return self.w

    def ptr(self):
        return 0


def _source(air, h, w, which):
    lib = L.load_library()
    ctx = sb.AirProvingContext(air.nodes, air.constraint_idx, air.interactions, air.constraint_degree, air.need_rot, _Shape(h, w),
                               air.public_values)
    keep = []
    c = ctx.c(keep)
    n = lib.swirl_jit_round0_source(C.byref(c), which, None, 0)
    buf = C.create_string_buffer(n + 1)
    lib.swirl_jit_round0_source(C.byref(c), which, buf, n + 1)
    return buf.value.decode()


def test_generated_source_rerolls_repeating_constraints():
    """BenchmarkAir (256 x assert_bool, 32 send/receive pairs): 768 + 600 instructions become two kernels of a few
    dozen statements -- no GPU needed to generate (or to compile) them."""
    air = A.benchmark(3, 256, 256, 32, np.random.default_rng(0))
    for which in (0, 1):
        src = _source(air, 1 << 20, 256, which)
        body = src[src.index("SW_R0_PROLOGUE\n", src.index("SW_R0_SIGNATURE(swirl_r0_jit)")):]
        assert "for (int it = 0; it <" in body
        assert body.count("\n") < 260, body.count("\n")
    # an AIR without repetition is emitted statement by statement
    fib = A.fibonacci(6)
    src = _source(fib, 64, 2, 0)
    assert "SW_R0_EPILOGUE" in src and "LD(" in src


def test_generated_source_compiles_with_nvrtc():
    nv = None
    for name in ("libnvrtc.so.12", "libnvrtc.so"):
        try:
            nv = C.CDLL(name)
            break
        except OSError:
            pass
    if nv is None:
        pytest.skip("libnvrtc not installed")
    air = A.benchmark(3, 64, 64, 8, np.random.default_rng(0))
    for which in (0, 1):
        src = _source(air, 1 << 18, 64, which).encode()
        prog = C.c_void_p()
        assert nv.nvrtcCreateProgram(C.byref(prog), src, b"jit.cu", 0, None, None) == 0
        opts = (C.c_char_p * 2)(b"--gpu-architecture=sm_100a", b"--std=c++17")
        rc = nv.nvrtcCompileProgram(prog, 2, opts)
        n = C.c_size_t()
        nv.nvrtcGetProgramLogSize(prog, C.byref(n))
        log = C.create_string_buffer(n.value + 1)
        nv.nvrtcGetProgramLog(prog, log)
        assert rc == 0, log.value.decode()[:2000]


@pytest.mark.gpu
def test_whole_proof_with_compiled_kernels_matches_oracle(oracle):
    """The 5-AIR fixture (preprocessed + cached commitments, interactions, rotations, an optional AIR) with every round-0
    program compiled (mode 2): byte-identical to the oracle's proof, and to the interpreter's."""
    import test_prove as tp

    airs, order = tp.fixture_airs(2)
    is_required = [True, True, True, True, False]
    vk = oracle.to_mont(np.arange(100, 108))
    want = tp.oracle_prove(oracle, airs, order, is_required, vk)
    sp = sb.SystemParams(4, 3, tp.LOG_BLOWUP, sb.WhirConfig(**tp.WHIR), tp.LOGUP_POW, tp.D)  # l_skip = 4: the compiled path
    results = {}
    for mode in (0, 2):
        dev = sb.B200Device(0)
        try:
            dev.set_jit(mode)
            dm = lambda m: sb.DeviceMatrix(dev.h2d(m[0]), m[1], m[2])

            def committed(m):
                mat = dm(m)
                r, data = dev.commit(sp.pcs(), [mat])
                return sb.CommittedTraceData(r, mat, data)

            pks, per_trace = [], []
            for air_id, a in enumerate(airs):
                prep = committed(a.preprocessed) if a.preprocessed is not None else None
                cached = [committed(c) for c in a.cached]
                pks.append(sb.AirProvingKey(is_required[air_id], prep))
                per_trace.append((air_id, sb.AirProvingContext(a.nodes, a.constraint_idx, a.interactions, a.constraint_degree, a.need_rot,
                                                               dm(a.common_main), a.public_values, [c.trace for c in cached],
                                                               prep.trace if prep else None), cached))
            results[mode] = sb.Coordinator(dev, sp).prove(vk, pks, per_trace).words()
        finally:
            dev.close()
    assert np.array_equal(results[0], results[2])


@pytest.mark.gpu
@pytest.mark.parametrize("log_rows,cols", [(12, 24), (17, 16)])
def test_benchmark_air_compiled_equals_interpreted(oracle, log_rows, cols):
    import torch

    air = A.benchmark(3, cols, cols, cols // 8, np.random.default_rng(0))
    g = torch.Generator(device="cuda").manual_seed(log_rows)
    trace = torch.randint(0, 2, ((1 << log_rows) * cols,), dtype=torch.int32, device="cuda", generator=g) * 0x0FFFFFFE
    whir = sb.WhirConfig.new(1, log_rows, 4, 8 if log_rows > 12 else 4, 6, 3, 4)
    params = sb.SystemParams(4, log_rows - 4, 1, whir, 4, 3)
    out = {}
    for mode in (0, 2):
        dev = sb.B200Device(0)
        try:
            dev.set_jit(mode)
            ctx = sb.AirProvingContext(air.nodes, air.constraint_idx, air.interactions, 2, False, sb.DeviceMatrix(trace, 1 << log_rows, cols))
            proof = sb.Coordinator(dev, params).prove(np.arange(8, dtype=np.uint32), [sb.AirProvingKey(True, None)], [(0, ctx, [])])
            out[mode] = proof.words()
            proof.common_main_pcs.free()
        finally:
            dev.close()
    assert np.array_equal(out[0], out[2])
