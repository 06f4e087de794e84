"""Runs the GPU comparisons of tests/test_zz_unverified_gpu.py in a process of their own (so that a fault there cannot
take the test session with it) and prints one JSON object {case: "pass" | "fail: ..."} as the last line of stdout."""
import json
import os
import sys
import traceback

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import numpy as np


def interaction_table_case(dev, oracle, name):
    import airs as A
    import stark_backend_b200 as sb
    import test_batch_constraints as tbc

    tables, balanced = tbc.REFERENCE_INTERACTION_TABLES[name]
    airs = sorted([A.dummy_interaction(t, send) for t, send in tables], key=lambda a: -a.height)
    l_skip, D, pow_bits = 2, 3, 1
    n_max = max(max(a.height.bit_length() - 1 - l_skip for a in airs), 0)
    st = np.zeros(18, np.uint32)
    oracle.sponge_observe(st, oracle.to_mont(np.arange(3)))
    ts = sb.Transcript(st)
    if not balanced:
        try:
            dev.prove_batch_constraints(ts, l_skip, D, pow_bits, tbc.to_device_airs(dev, airs))
        except sb.SwirlError as e:
            assert e.code == 10005, e
            return
        raise AssertionError("an unbalanced table was proved")
    want, r = oracle.bc_prove(st, l_skip, D, pow_bits, A.flatten(airs), len(airs), n_max)
    got, rg = dev.prove_batch_constraints(ts, l_skip, D, pow_bits, tbc.to_device_airs(dev, airs))
    assert np.array_equal(got, want) and np.array_equal(rg, r) and np.array_equal(ts.words(), st)


def cases():
    import test_batch_constraints as tbc
    import test_prove_matrix as tpm

    out = [("prove:" + c[0], (lambda dev, oracle, c=c: tpm.gpu_proof_equals_oracle_proof(dev, oracle, c))) for c in tpm.CPU_ONLY_CASES]
    out += [("tables:" + n, (lambda dev, oracle, n=n: interaction_table_case(dev, oracle, n))) for n in sorted(tbc.REFERENCE_INTERACTION_TABLES)]
    return out


def main():
    import oracle_lib
    import stark_backend_b200 as sb

    oracle = oracle_lib.Oracle(os.path.join(os.path.dirname(HERE), "oracle", "libswirl_oracle.so"))
    results = {}
    for name, fn in cases():
        dev = None
        try:
            dev = sb.B200Device(0)
            fn(dev, oracle)
            results[name] = "pass"
        except BaseException as e:  # noqa: BLE001 -- report everything, the caller decides
            results[name] = "fail: " + "".join(traceback.format_exception_only(type(e), e)).strip()[:500]
        finally:
            try:
                if dev is not None:
                    dev.close()
            except Exception:
                pass
        print(json.dumps(results), flush=True)


if __name__ == "__main__":
    main()
