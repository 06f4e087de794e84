"""CPU tests of the drop-in boundary: the C-ABI library loads, exports exactly what
include/swirl_b200.h declares, its host-only logic matches the goldens, and it refuses (loudly,
no fallback) to compute without a CUDA device."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest
import torch

import stark_backend_b200 as sb
from stark_backend_b200 import lib as sblib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_kats.json")))


def _declared():
    src = open(os.path.join(ROOT, "include", "swirl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(swirl_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = sb.load_library()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/swirl_b200.h but not exported"
    assert sorted(sblib.PROTOTYPES) == names, "python binding and header disagree"


def test_no_torch_or_cxx_types_in_header():
    src = open(os.path.join(ROOT, "include", "swirl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)  # declarations only
    assert "torch" not in src and "std::" not in src and "at::" not in src and "cudaStream_t" not in src


@pytest.mark.parametrize("case", GOLD["stacking"], ids=lambda c: c["name"])
def test_host_layout_matches_reference_goldens(case, oracle):
    # StackedLayout::new is host logic inside the product library: check it against the oracle and
    # against where the golden matrices put each column.
    meta = [(t["width"], t["height"].bit_length() - 1) for t in case["traces"]]
    lay = sb.StackedLayout.new(case["l_skip"], case["l_skip"] + case["n_stack"], meta)
    ow, ocols = oracle.stacked_layout(case["l_skip"], case["l_skip"] + case["n_stack"], meta)
    assert lay.width == ow == case["width"] and lay.height == case["height"]
    got = [[m, j, s.col_idx, s.row_idx, s.log_height] for m, j, s in lay.sorted_cols]
    assert got == [[int(x) for x in r] for r in ocols]


def test_host_layout_errors():
    with pytest.raises(sb.SwirlError) as e:
        sb.StackedLayout.new(2, 4, [(1, 5)])
    assert e.value.code == 10002 and "LayoutHeightExceeded" in str(e.value)
    with pytest.raises(sb.SwirlError) as e:
        sb.StackedLayout.new(0, 4, [(1, 3), (1, 4)])
    assert "LayoutRowOverflow" in str(e.value)
    lay = sb.StackedLayout.new(2, 4, [(3, 4), (0, 3), (5, 2), (2, 0)])
    assert lay.width == 5 and len(lay.sorted_cols) == 10


def test_no_cpu_fallback_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = sb.load_library()
    h = C.c_void_p()
    rc = lib.swirl_ctx_create(0, C.byref(h))
    assert rc == 10004 and not h.value
    assert b"no CPU fallback" in lib.swirl_last_error()
    with pytest.raises(RuntimeError):
        sb.B200Device(0)


def test_mont_helpers_roundtrip(oracle):
    x = np.array([0, 1, 2, 31, sb.P - 1, 123456789], dtype=np.uint64)
    assert np.array_equal(sb.to_mont(x), oracle.to_mont(x))
    assert np.array_equal(sb.from_mont(sb.to_mont(x)), x.astype(np.uint32))


def test_host_transcript_matches_oracle(oracle):
    """The product's host transcript (AVX2 Poseidon2 in csrc/poseidon2_host.hpp, duplex sponge of transcript.hpp) against
    the oracle's scalar restatement of duplex_sponge.rs:60-83 over random observe / sample sequences."""
    rng = np.random.default_rng(1)
    ts, st = sb.Transcript(), np.zeros(18, np.uint32)
    for _ in range(300):
        w = oracle.random_field(rng, int(rng.integers(1, 40)))
        ts.observe(w)
        oracle.sponge_observe(st, w)
        n = int(rng.integers(1, 12))
        assert np.array_equal(ts.sample(n), oracle.sponge_sample(st, n))
    edge = np.array([0, sb.P - 1, 1, sb.P - 2] * 4, dtype=np.uint32)  # extreme canonical words
    ts.observe(sb.to_mont(edge))
    oracle.sponge_observe(st, sb.to_mont(edge))
    assert np.array_equal(ts.words(), st)


def test_rust_ffi_lists_every_symbol():
    """bindings/rust/src/ffi.rs (the extern block a maintainer links against) declares every entry point of the header."""
    ffi = open(os.path.join(ROOT, "bindings", "rust", "src", "ffi.rs")).read()
    declared = set(re.findall(r"pub fn (swirl_[a-z0-9_]+)\(", ffi))
    assert declared == set(_declared())
