"""Run-time compiled MLE-round kernels (csrc/batch.cu: generate_mle_source, csrc/jit_prelude.cuh SW_D section; SURVEY
section 8f-3): one kernel per AIR, a switch over its sub-programs, sub-programs equal up to a column / weight shift share
a case.

CPU part: the generated CUDA C++ is (a) compiled with NVRTC for sm_100a and (b) compiled with g++ against a small shim
(`SW_HOST_EMU`: threads run one after the other, the grid reduction is a plain sum) and executed on random tables; the
result must equal a replay, in Python integers, of the instruction listing the generator appends -- so the statement
emission, loop re-rolling, case sharing and descriptor handling are checked without a GPU.
GPU part: proofs with the compiled MLE kernels equal the interpreter's and the oracle's bit for bit."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import airs as A
import stark_backend_b200 as sb
from stark_backend_b200 import lib as L
from test_jit import _Shape

P = 0x78000001
R = (1 << 32) % P
RINV = pow(R, -1, P)
I_VAR, I_CONST, I_ADD, I_SUB, I_MUL, I_NEG, I_PREF, I_MULACC, I_ACC = range(9)


def mle_source(air, h, w, D, n_airs=1):
    lib = L.load_library()
    cached = [_Shape(h, m[2]) for m in getattr(air, "cached", [])]
    prep = _Shape(h, air.preprocessed[2]) if getattr(air, "preprocessed", None) is not None else None
    ctx = sb.AirProvingContext(air.nodes, air.constraint_idx, air.interactions, air.constraint_degree, air.need_rot, _Shape(h, w),
                               air.public_values, cached, prep)
    keep = []
    c = ctx.c(keep)
    n = lib.swirl_jit_mle_source(C.byref(c), D, n_airs, None, 0)
    buf = C.create_string_buffer(n + 1)
    lib.swirl_jit_mle_source(C.byref(c), D, n_airs, buf, n + 1)
    return buf.value.decode()


def listing(src):
    """[(case, col_shift, w_shift, [(op, dst, a, b, c), ...])] per sub-program."""
    subs = []
    for line in src.split("\n"):
        f = line.split()
        if line.startswith("// SUB "):
            subs.append((int(f[4]), int(f[6]), int(f[8]), []))
        elif line.startswith("// I "):
            subs[int(f[2])][3].append(tuple(int(x) for x in f[3:8]))
    return subs


# ---- EF arithmetic on Montgomery words, in Python integers ----------------------------------------------------------
def e_add(a, b):
    return tuple((x + y) % P for x, y in zip(a, b))


def e_sub(a, b):
    return tuple((x - y) % P for x, y in zip(a, b))


def e_neg(a):
    return tuple((-x) % P for x in a)


def e_mul(a, b):  # F[x] / (x^4 - 11); Montgomery words: one factor R^-1 per product
    r = [0] * 7
    for i in range(4):
        for j in range(4):
            r[i + j] += a[i] * b[j]
    return tuple((r[i] + 11 * (r[i + 4] if i < 3 else 0)) * RINV % P for i in range(4))


def replay(code, D, col, weights, eq, ny, single):
    """sum_y eq[y] * acc_k(X, y) for X = 1..D as D * 12 words (batch.cu: batch_mle_kernel)."""
    out = [[(0, 0, 0, 0)] * 3 for _ in range(D)]
    for y in range(ny):
        slots = {}
        acc = [[(0, 0, 0, 0)] * 3 for _ in range(D)]
        for op, dst, a, b, c in code:
            if op == I_VAR:
                if single:
                    slots[dst] = [col(c, 0)] * D
                else:
                    t0, t1 = col(c, 2 * y), col(c, 2 * y + 1)
                    d = e_sub(t1, t0)
                    v = [t1]
                    for _ in range(1, D):
                        v.append(e_add(v[-1], d))
                    slots[dst] = v
            elif op == I_PREF:
                pass
            elif op == I_CONST:
                slots[dst] = [(a, 0, 0, 0)] * D
            elif op == I_ADD:
                slots[dst] = [e_add(x, y_) for x, y_ in zip(slots[a], slots[b])]
            elif op == I_SUB:
                slots[dst] = [e_sub(x, y_) for x, y_ in zip(slots[a], slots[b])]
            elif op == I_MUL:
                slots[dst] = [e_mul(x, y_) for x, y_ in zip(slots[a], slots[b])]
            elif op == I_NEG:
                slots[dst] = [e_neg(x) for x in slots[a]]
            elif op == I_MULACC:  # dst = accumulator, a / b = slots, c = weight
                for l in range(D):
                    acc[l][dst] = e_add(acc[l][dst], e_mul(weights[c], e_mul(slots[a][l], slots[b][l])))
            else:  # I_ACC: a = accumulator, b = weight, c = slot
                for l in range(D):
                    acc[l][a] = e_add(acc[l][a], e_mul(weights[b], slots[c][l]))
        e = (R, 0, 0, 0) if single else eq[y]
        for l in range(D):
            for k in range(3):
                out[l][k] = e_add(out[l][k], e_mul(e, acc[l][k]))
    return [w for l in range(D) for k in range(3) for w in out[l][k]]


SHIM = r"""
// host stand-ins for the CUDA constructs the prelude uses (tests/test_jit_mle.py)
#include <cstddef>
#define __device__
#define __global__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __shared__ static
struct uint4 { unsigned x, y, z, w; };
struct Dim3 { unsigned x, y, z; };
static Dim3 threadIdx, blockIdx, blockDim, gridDim;
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline void __syncthreads() {}
template <class T> static inline T min(T a, T b) { return a < b ? a : b; }
"""
DRIVER = r"""
extern "C" void emu_run(const MleArgs* descs, const uint16_t* block_air, unsigned n_blocks, unsigned block_base, unsigned tag) {
    blockDim.x = 128; blockDim.y = blockDim.z = 1;
    gridDim.x = n_blocks; gridDim.y = gridDim.z = 1;
    for (unsigned b = 0; b < n_blocks; b++)
        for (unsigned t = 0; t < 128; t++) {
            blockIdx.x = b; blockIdx.y = blockIdx.z = 0;
            threadIdx.x = t; threadIdx.y = threadIdx.z = 0;
            swirl_mle_jit(descs, block_air, block_base, tag);
        }
}
"""


class MleArgs(C.Structure):  # csrc/batch.cu: MleArgs
    _fields_ = [("code", C.c_void_p), ("n_instr", C.c_uint32), ("base", C.c_void_p), ("h", C.c_size_t), ("weights", C.c_void_p),
                ("eq_xi", C.c_void_p), ("ny", C.c_size_t), ("single", C.c_int), ("first_block", C.c_uint32), ("n_blocks", C.c_uint32),
                ("partials", C.c_void_p), ("ticket", C.c_void_p), ("result", C.c_void_p), ("sub", C.c_uint32), ("col_shift", C.c_uint32),
                ("w_shift", C.c_uint32)]


def build_emulation(src, tmp_path):
    cu = tmp_path / "mle_emu.cpp"
    cu.write_text(SHIM + src + DRIVER)
    so = tmp_path / "libmle_emu.so"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-DSW_HOST_EMU", "-w", "-o", str(so), str(cu)])
    return C.CDLL(str(so))


AIRS = {
    "benchmark48": lambda rng: (A.benchmark(3, 48, 48, 6, rng), 48),
    "benchmark20_c3": lambda rng: (A.benchmark(3, 20, 20, 3, rng), 20),
    "fibonacci": lambda rng: (A.fibonacci(4), 2),
    "with_parts": lambda rng: (A.with_parts(4, rng), None),
    "sender": lambda rng: (A.sender_receiver(4, 2, rng)[0], None),
}


@pytest.mark.parametrize("name,D,n_airs,single", [("benchmark48", 3, 1, False), ("benchmark48", 3, 1, True), ("benchmark20_c3", 2, 32, False),
                                                  ("fibonacci", 3, 1, False), ("with_parts", 4, 5, False), ("sender", 3, 2, False)])
def test_generated_mle_kernel_on_host_equals_replay_of_its_program(tmp_path, name, D, n_airs, single):
    rng = np.random.default_rng(7)
    air, w = AIRS[name](rng)
    if w is None:
        w = air.common_main[2]
    src = mle_source(air, 1 << 18, w, D, n_airs)
    assert src, "the generator declined"
    subs = listing(src)
    assert subs and "SW_MLE_EPILOGUE" in src
    emu = build_emulation(src, tmp_path)
    n_cols = 1 + max([c for s in subs for (op, _, _, _, c) in s[3] if op in (I_VAR, I_PREF)], default=0)
    n_w = 1 + max([(b if op == I_ACC else c) for s in subs for (op, _, _, b, c) in s[3] if op in (I_ACC, I_MULACC)], default=0)
    h = 1 if single else 320  # ny = 160: two blocks per descriptor, the second one partly filled
    ny = 1 if single else h // 2
    table = rng.integers(0, P, size=(n_cols, h, 4), dtype=np.uint32)
    weights = rng.integers(0, P, size=(n_w, 4), dtype=np.uint32)
    eq = rng.integers(0, P, size=(max(ny, 1), 4), dtype=np.uint32)
    n_blocks_each = 1 if single else 2
    descs = (MleArgs * len(subs))()
    results = np.zeros((len(subs), 64), dtype=np.uint32)
    block_air = []
    for i, (case, cs, ws, _) in enumerate(subs):
        d = descs[i]
        d.base, d.h, d.weights, d.eq_xi = table.ctypes.data, h, weights.ctypes.data, eq.ctypes.data
        d.ny, d.single, d.first_block, d.n_blocks = ny, int(single), 100 + len(block_air), n_blocks_each
        d.result = results[i].ctypes.data
        d.sub, d.col_shift, d.w_shift = case, cs, ws
        block_air += [i] * n_blocks_each
    ba = np.array(block_air, dtype=np.uint16)
    tag = 0x80000000
    emu.emu_run(descs, ba.ctypes.data_as(C.c_void_p), C.c_uint(len(block_air)), C.c_uint(100), C.c_uint(tag))
    col = lambda c, r: tuple(int(x) for x in table[c, r])
    wl = [tuple(int(x) for x in weights[i]) for i in range(n_w)]
    el = [tuple(int(x) for x in eq[i]) for i in range(len(eq))]
    lanes = 1 if single else D  # a single-row table is evaluated on lane 0 only (the host reads nothing else)
    for i, (_, _, _, code) in enumerate(subs):
        want = replay(code, D, col, wl, el, ny, single)
        got = [int(x) for x in results[i, :D * 12]]
        assert all(g & tag for g in got), "result words must carry the round's tag"
        assert [g & 0x7FFFFFFF for g in got][:lanes * 12] == want[:lanes * 12], f"sub-program {i}"


def test_sub_programs_that_differ_by_a_shift_share_a_case():
    air = A.benchmark(3, 256, 256, 32, np.random.default_rng(0))
    src = mle_source(air, 1 << 20, 256, 3)
    subs = listing(src)
    assert len(subs) == 24 and len({s[0] for s in subs}) == 2  # 16 x assert_bool on 16 columns, 8 x four bus interactions
    assert [s[1] for s in subs[:16]] == [16 * i for i in range(16)]
    assert sum(line.startswith("case ") for line in src.split("\n")) == 2  # two cases in the switch


def test_generated_mle_source_compiles_with_nvrtc():
    nv = None
    for name in ("libnvrtc.so.12", "libnvrtc.so"):
        try:
            nv = C.CDLL(name)
            break
        except OSError:
            pass
    if nv is None:
        pytest.skip("libnvrtc not installed")
    rng = np.random.default_rng(0)
    for air, w, D in ((A.benchmark(3, 64, 64, 8, rng), 64, 3), (A.with_parts(4, rng), None, 2)):
        src = mle_source(air, 1 << 18, w if w else air.common_main[2], D).encode()
        prog = C.c_void_p()
        assert nv.nvrtcCreateProgram(C.byref(prog), src, b"jit.cu", 0, None, None) == 0
        opts = (C.c_char_p * 2)(b"--gpu-architecture=sm_100a", b"--std=c++17")
        rc = nv.nvrtcCompileProgram(prog, 2, opts)
        n = C.c_size_t()
        nv.nvrtcGetProgramLogSize(prog, C.byref(n))
        log = C.create_string_buffer(n.value + 1)
        nv.nvrtcGetProgramLog(prog, log)
        assert rc == 0, log.value.decode()[:2000]


# ---- GPU: compiled MLE kernels inside whole proofs --------------------------------------------------------------------
@pytest.mark.gpu
def test_whole_proof_with_compiled_mle_kernels_matches_interpreter(oracle):
    """The 5-AIR fixture (preprocessed + cached commitments, interactions, rotations, an optional AIR): interpreter only
    (mode 0) against everything compiled (mode 2 + 4).  Five distinct AIRs: one compiled launch per AIR and round."""
    import test_prove as tp

    airs, order = tp.fixture_airs(2)
    is_required = [True, True, True, True, False]
    vk = oracle.to_mont(np.arange(100, 108))
    sp = sb.SystemParams(4, 3, tp.LOG_BLOWUP, sb.WhirConfig(**tp.WHIR), tp.LOGUP_POW, tp.D)
    results, stats = {}, {}
    for mode in (0, 6):
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
            stats[mode] = dev.jit_stats()
        finally:
            dev.close()
    assert stats[0]["mle_launches"] == 0 and stats[6]["mle_compiled"] >= 1 and stats[6]["mle_launches"] >= 1, stats
    assert np.array_equal(results[0], results[6])


@pytest.mark.gpu
@pytest.mark.parametrize("log_rows,cols,n_copies", [(12, 24, 1), (17, 16, 1), (13, 20, 3)])
def test_benchmark_air_compiled_mle_equals_interpreted(oracle, log_rows, cols, n_copies):
    """BenchmarkAir (sub-programs sharing cases through column / weight shifts), one AIR and several copies of the same AIR
    in one proof (they share one compiled kernel and one launch per round, as BASELINE configs[2] does)."""
    import torch

    air = A.benchmark(3, cols, cols, max(cols // 8, 1), np.random.default_rng(0))
    g = torch.Generator(device="cuda").manual_seed(log_rows)
    traces = [torch.randint(0, 2, ((1 << log_rows) * cols,), dtype=torch.int32, device="cuda", generator=g) * 0x0FFFFFFE
              for _ in range(n_copies)]
    log_stacked = log_rows + (n_copies - 1).bit_length()
    whir = sb.WhirConfig.new(1, log_stacked, 4, 8 if log_rows > 12 else 4, 6, 3, 4)
    params = sb.SystemParams(4, log_stacked - 4, 1, whir, 4, 3)
    out, stats = {}, {}
    for mode in (0, 6):
        dev = sb.B200Device(0)
        try:
            dev.set_jit(mode)
            per_trace = [(i, sb.AirProvingContext(air.nodes, air.constraint_idx, air.interactions, 2, False,
                                                  sb.DeviceMatrix(t, 1 << log_rows, cols)), []) for i, t in enumerate(traces)]
            proof = sb.Coordinator(dev, params).prove(np.arange(8, dtype=np.uint32), [sb.AirProvingKey(True, None)] * n_copies, per_trace)
            out[mode] = proof.words()
            stats[mode] = dev.jit_stats()
            proof.common_main_pcs.free()
        finally:
            dev.close()
    assert stats[6]["mle_compiled"] == 1 and stats[6]["mle_launches"] >= log_rows - 4 - 1, stats
    assert np.array_equal(out[0], out[6])
