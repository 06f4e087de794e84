"""Freezes outputs of the REFERENCE's own CUDA kernels (oracle/_ref/libref_kernels.so, built from /root/reference by
oracle/Makefile.ref) into tests/golden/reference_gpu_kats.json.  Run on a GPU box:

    gpurun -- 'python tools/gen_reference_goldens.py gpurun_out/reference_gpu_kats.json'

then copy the file to tests/golden/.  The cases and their seeds live in tests/refcases.py; tests/test_reference_goldens.py
checks the CPU oracle against this file without a GPU, tests/test_reference_kernels.py re-runs the three-way comparison live.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import oracle_lib  # noqa: E402
import ref_kernels  # noqa: E402
import refcases as rc  # noqa: E402


def main(out_path):
    oracle = oracle_lib.Oracle(os.path.join(ROOT, "oracle", "libswirl_oracle.so"))  # only for to_mont / from_mont
    rk = ref_kernels.RefKernels()
    outs = rc.reference_outputs(rk, oracle.to_mont)
    gold = {
        "_about": "outputs of openvm-org/stark-backend's own CUDA kernels (crates/cuda-backend/cuda, compiled for sm_100 by "
                  "oracle/Makefile.ref) on the seeded inputs of tests/refcases.py; Montgomery words as written by the kernels",
        "gpu": torch.cuda.get_device_name(0),
        "merkle": [dict(case=list(c), layers=rc.digest_or_words(o)) for c, o in zip(rc.MERKLE_CASES, outs["merkle"])],
        "rs": [dict(case=list(c), codeword=rc.digest_or_words(o)) for c, o in zip(rc.RS_CASES, outs["rs"])],
        "ntt": [dict(case=[c[0], c[1], bool(c[2])], out=rc.digest_or_words(o)) for c, o in zip(rc.NTT_CASES, outs["ntt"])],
        "grind": [dict(case=list(c), witness=int(o)) for c, o in zip(rc.GRIND_CASES, outs["grind"])],
        "frac_layer": [dict(log_n=c, out=rc.digest_or_words(o)) for c, o in zip(rc.EF_CASES, outs["frac_layer"])],
        "whir_fold": [dict(log_n=c, out=rc.digest_or_words(o)) for c, o in zip(rc.EF_CASES, outs["whir_fold"])],
    }
    # a few raw Poseidon2 values for the README of the goldens: digest of the all-zero row of width 8 and of 0..7
    z = rk.merkle_tree(rk.h2d(np.zeros(8, np.uint32)), 1, 8, 1)
    i = rk.merkle_tree(rk.h2d(oracle.to_mont(np.arange(8))), 1, 8, 1)
    gold["poseidon2_kat"] = {"hash_zero8_canonical": [int(x) for x in oracle.from_mont(rk.d2h(z[0]))],
                             "hash_iota8_canonical": [int(x) for x in oracle.from_mont(rk.d2h(i[0]))]}
    with open(out_path, "w") as f:
        json.dump(gold, f, indent=1)
    print("wrote", out_path, os.path.getsize(out_path), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "reference_gpu_kats.json"))
