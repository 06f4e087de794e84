#!/bin/bash
# One GPU call: the whole GPU suite with the compiled MLE rounds on, the C2 proof with the MLE rounds interpreted and
# compiled (per-phase wall clock on stderr with SWIRL_TRACE=1), and the mixture fixtures with every program compiled.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=$1
(SWIRL_JIT_MLE=1 timeout 80 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -60) > gpurun_out/${T}_pytest_mle_compiled.log
(SWIRL_TRACE=1 SWIRL_JIT_MLE=0 timeout 30 python tools/prove_c2.py) > gpurun_out/${T}_c2_mle_interpreted.log 2>&1
(SWIRL_TRACE=1 SWIRL_JIT_MLE=1 timeout 40 python tools/prove_c2.py) > gpurun_out/${T}_c2_mle_compiled.log 2>&1
(SWIRL_JIT_MODE=2 SWIRL_JIT_MLE=1 timeout 40 python -m pytest tests/test_prove_matrix.py -m gpu -q -k mixture -p no:cacheprovider 2>&1 | tail -30) > gpurun_out/${T}_pytest_mixture_all_compiled.log
tail -3 gpurun_out/${T}_pytest_mle_compiled.log
grep -h "mle rounds\|MLE-round\|total_ms" gpurun_out/${T}_c2_mle_interpreted.log | cut -c1-200 | tail -7
grep -h "mle rounds\|MLE-round\|total_ms" gpurun_out/${T}_c2_mle_compiled.log | cut -c1-200 | tail -7
tail -3 gpurun_out/${T}_pytest_mixture_all_compiled.log
