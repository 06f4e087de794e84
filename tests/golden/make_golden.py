"""Extracts the reference's own known-answer data for the commit path into reference_kats.json.
Run here (needs /root/reference); the GPU box only reads the committed JSON.

Sources (relative to /root/reference):
  crates/cuda-common/include/fp.h:291-320                      TWO_ADIC_GENERATORS (canonical)
  crates/cuda-common/include/fp.h:55-65                        P, R2 and friends
  crates/cuda-common/include/poseidon2.cuh:50-67               internal diagonal (canonical)
  crates/stark-backend/src/prover/stacked_pcs.rs:556-619       stacking goldens (CPU)
  crates/cuda-backend/src/stacked_pcs.rs:399-517               stacking goldens (GPU, incl. Interactions11)
  crates/stark-backend/src/test_utils/mod.rs:227-243           InteractionsFixture11 traces
"""
import json
import re

REF = "/root/reference/crates/"
fp = open(REF + "cuda-common/include/fp.h").read()
gens = [int(x, 16) for x in re.findall(r"Fp\((0x[0-9a-f]+)\),?\s*// Fp\(0x", fp)]
assert len(gens) == 28, len(gens)
gens_mont = [int(x, 16) for x in re.findall(r"// Fp\((0x[0-9a-f]+)u\)", fp)]
assert len(gens_mont) == 28

p2 = open(REF + "cuda-common/include/poseidon2.cuh").read()
blk = p2[p2.index("internal_diag16") :]
blk = blk[blk.index("{") + 1 : blk.index("}")]
diag = [int(x) for x in re.findall(r"^\s*(\d+)\s*,?", re.sub(r"//.*", "", blk), re.M)]
assert len(diag) == 16

# sanity: the literal test vectors below are still what the reference holds
cpu_tests = open(REF + "stark-backend/src/prover/stacked_pcs.rs").read()
assert "[1, 2, 3, 4, 5, 6, 7, 0]" in cpu_tests and "[1, 2, 3, 4, 5, 0, 6, 0, 7, 0, 0, 0]" in cpu_tests
gpu_tests = open(REF + "cuda-backend/src/stacked_pcs.rs").read()
assert "1, 3, 4, 2, 0, 545, 1, 0, 5, 4, 4, 5, 123, 889, 889, 456, 0, 3, 7, 546, 1, 5, 4," in gpu_tests
fixtures = open(REF + "stark-backend/src/test_utils/mod.rs").read()
assert "[0, 1, 3, 5, 7, 4, 546, 889]" in fixtures
assert "[1, 5, 3, 4, 4, 4, 2, 5, 0, 123, 545, 889, 1, 889, 0, 456]" in fixtures

cols = [[1, 2, 3, 4], [5, 6], [7]]
manual = [{"values": c, "height": len(c), "width": 1} for c in cols]
# InteractionsFixture11: row-major width-2 traces -> column-major
snd_rm = [0, 1, 3, 5, 7, 4, 546, 889]
rcv_rm = [1, 5, 3, 4, 4, 4, 2, 5, 0, 123, 545, 889, 1, 889, 0, 456]
cm = lambda rm: rm[0::2] + rm[1::2]
out = {
    "P": 15 * (1 << 27) + 1,
    "R2": 1172168163,
    "MONTY_ONE": 0x0FFFFFFE,
    "two_adic_generators_canonical": gens,
    "two_adic_generators_monty": gens_mont,
    "poseidon2_internal_diag_canonical": diag,
    "stacking": [
        {"name": "manual_0", "l_skip": 0, "n_stack": 2, "traces": manual, "height": 4, "width": 2,
         "expected": [1, 2, 3, 4, 5, 6, 7, 0], "src": "prover/stacked_pcs.rs:556-573"},
        {"name": "manual_strided_0", "l_skip": 2, "n_stack": 0, "traces": manual, "height": 4, "width": 3,
         "expected": [1, 2, 3, 4, 5, 0, 6, 0, 7, 0, 0, 0], "src": "prover/stacked_pcs.rs:575-593"},
        {"name": "manual_strided_1", "l_skip": 3, "n_stack": 0, "traces": manual, "height": 8, "width": 3,
         "expected": [1, 0, 2, 0, 3, 0, 4, 0, 5, 0, 0, 0, 6, 0, 0, 0, 7, 0, 0, 0, 0, 0, 0, 0],
         "src": "prover/stacked_pcs.rs:595-619"},
        {"name": "manual_1_interactions11", "l_skip": 2, "n_stack": 8,
         "traces": [{"values": cm(rcv_rm), "height": 8, "width": 2}, {"values": cm(snd_rm), "height": 4, "width": 2}],
         "height": 1024, "width": 1,
         "expected_prefix": [1, 3, 4, 2, 0, 545, 1, 0, 5, 4, 4, 5, 123, 889, 889, 456, 0, 3, 7, 546, 1, 5, 4, 889],
         "src": "cuda-backend/src/stacked_pcs.rs:425-452"},
    ],
    # SURVEY.md §8c scratch values (independent Python transcription of poseidon2.cuh; NOT validated
    # against Plonky3 — self-consistency only)
    "poseidon2_selfcheck": {
        "zeros_first4": [1168947398, 128782440, 747404447, 883925857],
        "iota_first4": [1906786279, 1737026427, 1959749225, 700325316],
    },
}
json.dump(out, open(__file__.replace("make_golden.py", "reference_kats.json"), "w"), indent=1)
print("wrote reference_kats.json")
