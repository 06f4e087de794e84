"""One full-size CPU proof of the bench workload (BASELINE configs[1]: 2^20 x 256) with the oracle (C++ port of the
reference's col-major prover, all host threads), phase by phase — validates the extrapolation bench.py's
`--impl reference` arm makes from its bounded sample.   python tools/cpu_full_proof.py [log_rows=20]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import airs as A  # noqa: E402
import bench as B  # noqa: E402

log_rows = int(sys.argv[1]) if len(sys.argv) > 1 else B.LOG_ROWS
oracle = B.load_oracle()
rng = np.random.default_rng(42)
air = B.benchmark_air_dag(B.COLS)
h = 1 << log_rows
air.common_main = ((rng.integers(0, 2, size=h * B.COLS, dtype=np.uint64) * B.R1).astype(np.uint32), h, B.COLS)
n_stack = log_rows - B.L_SKIP
cfg = dict(k=B.K_WHIR, num_queries=B.whir_queries(log_rows), mu_pow_bits=B.MU_POW, query_phase_pow_bits=B.QUERY_POW,
           folding_pow_bits=B.FOLD_POW)
t0 = time.perf_counter()
st = np.zeros(18, np.uint32)
root, _, _, _ = oracle.stacked_commit(B.L_SKIP, n_stack, B.LOG_BLOWUP, B.K_WHIR, [air.common_main], want_codeword=False)
t1 = time.perf_counter()
oracle.sponge_observe(st, root)
bc, r = oracle.bc_prove(st, B.L_SKIP, B.MAX_CONSTRAINT_DEGREE, B.LOGUP_POW, A.flatten([air]), 1, n_stack)
t2 = time.perf_counter()
_, u, _ = oracle.stacked_reduction_prove(st, B.L_SKIP, n_stack, [[air.common_main + (False,)]], r)
t3 = time.perf_counter()
u_cube = [u[0]]
for _ in range(B.L_SKIP - 1):
    u_cube.append(oracle.ef_mul(u_cube[-1], u_cube[-1]))
oracle.whir_prove(st, B.L_SKIP, B.LOG_BLOWUP, cfg, [(air.common_main[0], B.COLS)], h, np.array(u_cube + list(u[1:]), dtype=np.uint32))
t4 = time.perf_counter()
print(json.dumps({"workload": f"full proof 2^{log_rows} x {B.COLS}, 100-bit app params", "kind": "port (oracle/)",
                  "cores": os.cpu_count(), "seconds": t4 - t0, "cells_per_s": h * B.COLS / (t4 - t0),
                  "phase_s": {"commit": t1 - t0, "batch_constraints": t2 - t1, "stacked_reduction": t3 - t2,
                              "whir (re-commits the matrix inside)": t4 - t3},
                  "commitment_first_word": int(root[0])}))
