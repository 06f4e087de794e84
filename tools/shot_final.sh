#!/bin/bash
# Last GPU call of the round: the GPU suite at HEAD, one bench line (cpu_baseline leg skipped: it is CPU time), and an
# ncu --set full capture of the run-time compiled MLE-round kernel (first two launches = the two largest rounds of C2).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=$1
(timeout 60 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -8) > gpurun_out/${T}_pytest_gpu.log
timeout 50 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/${T}_bench_no_cpu.json 2> gpurun_out/${T}_bench.err
tail -2 gpurun_out/${T}_pytest_gpu.log
cut -c1-400 gpurun_out/${T}_bench_no_cpu.json
timeout 45 ncu --set full --clock-control none --import-source on -k regex:swirl_mle_jit -c 2 -f -o gpurun_out/${T}_mle_jit python tools/prove_c2.py 20 256 > gpurun_out/${T}_ncu.log 2>&1
tail -3 gpurun_out/${T}_ncu.log
