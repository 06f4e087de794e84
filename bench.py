#!/usr/bin/env python
"""bench.py — headline measurement of the SWIRL prover hot path on B200.

One "step" = one full proof (Coordinator::prove, prover/mod.rs:104-198: stacked commit, LogUp-GKR +
batch constraint sumcheck, stacked opening reduction, WHIR opening) of BASELINE.json configs[1]:
a single BenchmarkAir (benchmarks/synthetic/src/bin/uniform_runner.rs:78-113) of 2^20 rows x 256
columns — 256 boolean constraints, 32 self-cancelling send/receive pairs on bus 0 — under
app_params_with_100_bits_security(20) (stark-sdk/src/config/mod.rs:121-138: l_skip 4, n_stack 16,
log_blowup 1, k_whir 4, WHIR queries 193/88/81, PoW bits 18/15/5/20, max constraint degree 3).
Metric: trace cells per second (cells = rows x columns of the common main trace).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl swirl|reference]

`value`   : device-timed (CUDA events on the library's stream), trace resident in HBM.
`e2e`     : the same with the trace in pinned HOST memory: H2D copy inside the timed region; the
            proof (about 3.4 MB) is produced in host memory by every step anyway.
`roofline`: the dominant kernel family, timed live with CUDA events inside the timed region.
`--impl reference`: the CPU restatement of the reference algorithm (oracle/, kind "port": no Rust
            toolchain exists in this image, so the reference crates cannot be built) on a bounded
            sample of the same workload.
N > 1: one process per GPU (torchrun); every rank proves its own trace of the same shape
(independent proofs, "replicas only" — SURVEY §8e fallback; the 32-byte commitments are
all-gathered over NCCL), weak scaling.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# rank 0 prints exactly one JSON line on stdout: NCCL (version banner at WARN and above), torchrun and any other library
# write to file descriptor 1 as well, so fd 1 is pointed at stderr for the whole run and the JSON line goes to the
# saved original descriptor (emit_json)
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit_json(obj):
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

LOG_ROWS, COLS = 20, 256
L_SKIP, LOG_BLOWUP, K_WHIR = 4, 1, 4
MAX_CONSTRAINT_DEGREE, LOGUP_POW, MU_POW, FOLD_POW, QUERY_POW, LOG_FINAL_POLY = 3, 18, 15, 5, 20, 10
CELLS = (1 << LOG_ROWS) * COLS
WORKLOAD = (
    "BASELINE configs[1]: full prove of 1 BenchmarkAir 2^20 rows x 256 cols (256 assert_bool constraints, 32 "
    "send/receive pairs on bus 0; random boolean trace, seed 42), app_params_with_100_bits_security(20): l_skip=4 "
    "n_stack=16 log_blowup=1 k_whir=4, WHIR queries 193/88/81, pow bits logup 18 / mu 15 / fold 5 / query 20, "
    "max_constraint_degree 3 -> commit (LDE 2^21 x 256 + Poseidon2 Merkle) + LogUp-GKR (2^27 leaves) + batch "
    "constraint sumcheck + stacked reduction + WHIR"
)
P = 0x78000001
R1 = 0x0FFFFFFE  # Montgomery 1


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled during the timed region.  In-process NVML (nvidia_ml_py) on a
    thread: a looping `nvidia-smi --query-gpu` child holds the driver lock for its whole multi-field query and
    was seen to stall this sync-heavy step (~350 stream synchronisations per proof) by 0.1-0.6 s at random;
    BENCH_SAMPLER=smi selects that recipe form anyway, BENCH_SAMPLER=none disables sampling."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.rows, self.proc, self.kind = index, [], None, os.environ.get("BENCH_SAMPLER", "nvml")
        self._stop = threading.Event()
        self.period = float(os.environ.get("BENCH_SMI_MS", "100")) / 1e3

    def start(self):
        if self.kind == "none":
            return
        if self.kind == "nvml":
            try:
                import pynvml

                pynvml.nvmlInit()
                self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
                self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
                threading.Thread(target=self._poll_nvml, args=(pynvml,), daemon=True).start()
                return
            except Exception:
                self.kind = "smi"
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self._physical_index()), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                 str(int(self.period * 1e3))], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].strip().isdigit():
                return int(ids[self.index])
        return self.index

    def _poll_nvml(self, nv):
        bits = [nv.nvmlClocksEventReasonHwSlowdown, nv.nvmlClocksEventReasonHwThermalSlowdown,
                nv.nvmlClocksEventReasonSwThermalSlowdown, nv.nvmlClocksEventReasonSwPowerCap]
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.rows.append((time.time(), [str(mhz), str(self.max_mhz), ""] + ["Active" if mask & b else "Not Active" for b in bits]))
            except Exception:
                pass
            self._stop.wait(self.period)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        self._stop.set()
        if self.proc:
            self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2] or [r for _, r in self.rows]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"], "sampler": self.kind}
        sm = sorted(float(r[0]) for r in rows if r[0].replace(".", "").isdigit())
        reasons = [n for i, n in enumerate(self.NAMES) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(rows[0][1]), "reasons": reasons,
                "samples": len(rows), "sampler": self.kind}


def benchmark_air_dag(cols):
    """BenchmarkAir as a DAG: cols assert_bool constraints, cols/8 send+receive pairs (the trace comes separately)."""
    import airs as A

    return A.benchmark(3, cols, cols, cols // 8, np.random.default_rng(0))


def load_oracle():
    import oracle_lib

    path = os.path.join(ROOT, "oracle", "libswirl_oracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    return oracle_lib.Oracle(path)


def whir_queries(log_stacked_height):
    import math

    level, rate, out = 100 - QUERY_POW, LOG_BLOWUP, []
    for _ in range(-(-(log_stacked_height - LOG_FINAL_POLY) // K_WHIR)):
        out.append(math.ceil(level / -math.log2((1.0 + 2.0 ** (-rate)) / 2.0)))
        rate += K_WHIR - 1
    return out


def cpu_prove_sample(oracle, log_rows):
    """The oracle's full proof of the same AIR at 2^log_rows rows (same parameters except that the stacked
    height shrinks with the trace).  Returns (cells/s, seconds, description)."""
    import airs as A

    rng = np.random.default_rng(42)
    air = benchmark_air_dag(COLS)
    h = 1 << log_rows
    air.common_main = ((rng.integers(0, 2, size=h * COLS, dtype=np.uint64) * R1).astype(np.uint32), h, COLS)
    n_stack = log_rows - L_SKIP
    cfg = dict(k=K_WHIR, num_queries=whir_queries(log_rows), mu_pow_bits=MU_POW, query_phase_pow_bits=QUERY_POW,
               folding_pow_bits=FOLD_POW)
    t = time.perf_counter()
    st = np.zeros(18, np.uint32)
    root, _, _, _ = oracle.stacked_commit(L_SKIP, n_stack, LOG_BLOWUP, K_WHIR, [air.common_main], want_codeword=False)
    oracle.sponge_observe(st, root)
    bc, r = oracle.bc_prove(st, L_SKIP, MAX_CONSTRAINT_DEGREE, LOGUP_POW, A.flatten([air]), 1, n_stack)
    commits = [[air.common_main + (False,)]]
    _, u, _ = oracle.stacked_reduction_prove(st, L_SKIP, n_stack, commits, r)
    u_cube = [u[0]]
    for _ in range(L_SKIP - 1):
        u_cube.append(oracle.ef_mul(u_cube[-1], u_cube[-1]))
    u_cube = np.array(u_cube + list(u[1:]), dtype=np.uint32)
    oracle.whir_prove(st, L_SKIP, LOG_BLOWUP, cfg, [(air.common_main[0], COLS)], h, u_cube)
    dt = time.perf_counter() - t
    return h * COLS / dt, dt, (f"full proof of the same AIR at 2^{log_rows} rows x {COLS} cols (1/{1 << (LOG_ROWS - log_rows)} of the "
                               f"workload; n_stack {n_stack}, WHIR queries {cfg['num_queries']}, same PoW bits)")


def reference_sample_log_rows(t14_seconds, steps, budget_s=150.0):
    """Largest sample (rows = 2^k, 14 <= k <= LOG_ROWS) whose `steps` proofs fit the budget; a proof's cost is linear in
    the rows within ~15 % (profiles/r2_cpu_full_proof_*.jsonl: 2^16 -> 2^20 is 14.2x the time for 16x the rows)."""
    k = 14
    while k < LOG_ROWS and steps * t14_seconds * (1 << (k + 1 - 14)) <= budget_s:
        k += 1
    return k


CPU_NOTE = ("C++ restatement of the reference col-major prover (oracle/; the reference is Rust and cannot be built in this image); "
            "every phase (commit, LogUp-GKR, batch constraints, stacked reduction, WHIR, PoW searches) runs on all host threads")
CPU_FULL_SIZE = ("a full-size 2^20 x 256 proof by the same code was timed once per host: profiles/r2_cpu_full_proof_*.jsonl "
                 "(8 cores: 217.8 s = 1.23 M cells/s against 1.09 M cells/s extrapolated from its 2^16-row sample)")


def run_reference(args):
    """The reference arm: the CPU implementation of the same path (oracle port, all host threads) on a bounded sample of the
    same workload: the same AIR and parameters at 2^k rows, k chosen so that the K steps end within a few minutes (2^16 or
    more on a 16-core host for the driver's K)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    oracle = load_oracle()
    cores = os.cpu_count() or 1
    t14 = None
    for _ in range(max(1, args.warmup)):
        _, dt, _ = cpu_prove_sample(oracle, 14)
        t14 = dt if t14 is None else min(t14, dt)
    log_rows = reference_sample_log_rows(t14, max(1, args.steps))
    times = []
    for _ in range(max(1, args.steps)):
        v, dt, sample = cpu_prove_sample(oracle, log_rows)
        times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = (1 << log_rows) * COLS / (ms / 1e3)
    emit_json({
        "impl": "reference", "metric": "trace_cells_per_s", "value": value, "unit": "cells/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 (BabyBear Montgomery)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample, "warmup_sample": "the same at 2^14 rows", "full_size": CPU_FULL_SIZE},
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": cores, "kind": "port", "sample": sample, "note": CPU_NOTE},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernels, from the `ncu --set full` capture of
    this workload that tools/ncu_traffic.py summarises into profiles/ncu_traffic.json (refreshed whenever a kernel changes;
    the file names the capture and the commit it was taken at)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        return {}


def int32_peak():
    """Measured 32-bit integer issue rates of this pool's B200 (tools/int_roofline.cu -> profiles/int32_peaks.json); the
    driver-written MEASURED_PEAKS.json has no integer figure."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "int32_peaks.json")))
    except Exception:
        return {"imad_tops": 17.96, "mixed_issue_tops": 24.1, "source": "fallback: profiles/r1_int_roofline.jsonl"}


def run_swirl(args):
    import torch
    import torch.distributed as dist

    import stark_backend_b200 as sb
    from stark_backend_b200 import multi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import datetime

        # a collective that one rank never enters (an error in the sharded extras below) must end as an exception on the
        # others after two minutes, not hang the launcher for NCCL's default ten
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"), timeout=datetime.timedelta(seconds=120))
    dev = sb.B200Device(local)
    whir = sb.WhirConfig(K_WHIR, whir_queries(LOG_ROWS), MU_POW, QUERY_POW, FOLD_POW)
    params = sb.SystemParams(L_SKIP, LOG_ROWS - L_SKIP, LOG_BLOWUP, whir, LOGUP_POW, MAX_CONSTRAINT_DEGREE)
    air = benchmark_air_dag(COLS)
    rng = np.random.default_rng(42 + rank)
    # one process per GPU: run on, and pin host memory from, the CPUs local to this rank's GPU (N > 1: 8 x 1 GiB per step
    # would otherwise cross the sockets; round 1 measured e2e efficiency 0.84 at N = 8 without it)
    numa_cpus = multi.bind_to_gpu_numa_node(local) if world > 1 else None
    host = torch.from_numpy((rng.integers(0, 2, size=CELLS, dtype=np.uint64) * R1).astype(np.uint32).view(np.int32)).pin_memory()
    trace_dev = host.to(dev.torch_device)
    stream = dev.torch_stream()
    vk_pre_hash = np.arange(8, dtype=np.uint32)
    pk = [sb.AirProvingKey(True, None)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def prove(trace_tensor):
        ctx = sb.AirProvingContext(air.nodes, air.constraint_idx, air.interactions, 2, False,
                                   sb.DeviceMatrix(trace_tensor, 1 << LOG_ROWS, COLS))
        proof = sb.Coordinator(dev, params).prove(vk_pre_hash, pk, [(0, ctx, [])])
        proof.common_main_pcs.free()
        return proof

    def step_device():
        return prove(trace_dev)

    air_h = sb.AirProvingContext(air.nodes, air.constraint_idx, air.interactions, 2, False, None)

    def step_host_latency():
        # one proof in isolation with the trace in pinned host memory: transport (H2D, pipelined by column groups with the
        # RS encoding and leaf hashing inside swirl_commit_host) + the rest of the proof
        proof = sb.Coordinator(dev, params).prove_host(vk_pre_hash, pk, air_h, host, 1 << LOG_ROWS, COLS)
        proof.common_main_pcs.free()
        return proof

    transporter = sb.TraceTransporter(dev, 1 << LOG_ROWS, COLS)
    pending = []

    def step_host():
        # a stream of proofs with every trace starting in pinned host memory: each step's trace is copied host -> device
        # inside the timed region (TraceTransporter, pinned double buffering); the copy of the NEXT step's trace is submitted
        # before this step's proof, so the PCIe transfer overlaps the proof.  The first step of a measurement finds nothing
        # prefetched (`pending` is cleared before the region) and waits for its own copy.
        ticket = pending.pop() if pending else transporter.submit(host)
        pending.append(transporter.submit(host))
        proof = prove(transporter.matrix(ticket).buffer)  # frees the PCS data that aliases the buffer
        transporter.retire(ticket)
        return proof

    def gather(root):
        return multi.all_gather_commitments(root, dev.torch_device)  # only the 32-byte commitments cross NVLink

    stalls = {}

    def timed_once(fn, steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.time()
        a.record(stream)
        walls = []
        proof = None
        for _ in range(steps):
            if proof is not None:
                proof.release()  # explicit: the host proof buffers go back to the pool (what a pooling allocator does)
            ts_ = time.perf_counter()
            proof = fn()
            walls.append(1e3 * (time.perf_counter() - ts_))
            if os.environ.get("BENCH_DEBUG"):
                print(f"step {fn.__name__} {walls[-1]:.1f} ms", file=sys.stderr)
        b.record(stream)
        roots = gather(proof.common_main_commit)
        barrier()
        t1 = time.time()
        ms = multi.max_over_ranks(a.elapsed_time(b), dev.torch_device)
        return ms, roots, t0, t1, proof, walls

    def timed(fn, steps, on_retry=None, label=None):
        """K steps; every step performs identical work (same trace, deterministic transcript), so a step
        that takes > 2.5x the median is a host/box stall (seen sporadically on shared boxes: 0.2-1.2 s):
        like a throttled run it is re-measured once, and both measurements are reported."""
        r = timed_once(fn, steps)
        med = sorted(r[5])[len(r[5]) // 2]
        worst = multi.max_over_ranks(max(r[5]) / med, dev.torch_device)
        if worst > 2.5:
            stalls[label or fn.__name__] = {"first_ms_per_step": r[0] / steps, "first_step_ms": [round(x, 1) for x in r[5]]}
            if on_retry:
                on_retry()
            r = timed_once(fn, steps)
        stalls.setdefault("step_ms", {})[label or fn.__name__] = [round(x, 1) for x in r[5]]
        return r[:5]

    # the sampler starts before the warm-up: the first nvidia-smi start-up on a fresh box contends for the
    # driver lock for about a second, which would otherwise land in the timed region of this sync-heavy step
    sampler = ClockSampler(local)
    if rank == 0:  # one poller per job: every nvidia-smi query briefly takes the driver lock of the whole node
        sampler.start()
    if os.environ.get("BENCH_DEBUG"):
        def wrap(name):
            f = getattr(dev, name)
            def g(*a, **k):
                t_ = time.perf_counter()
                r_ = f(*a, **k)
                dev.synchronize()
                d_ = 1e3 * (time.perf_counter() - t_)
                if d_ > 60:
                    print(f"   slow {name}: {d_:.1f} ms", file=sys.stderr)
                return r_
            setattr(dev, name, g)
        for nm in ("commit", "commit_host", "prove_batch_constraints", "prove_openings"):
            wrap(nm)
    for _ in range(max(args.warmup, 3)):
        step_device().release()
    l0 = dev.launch_count()
    ms, roots, t0, t1, proof = timed(step_device, args.steps)
    launches = dev.launch_count() - l0
    # per-kernel-family durations: the same K steps again with the library's CUDA-event spans on
    # (an event pair per launch costs ~40% on this launch-heavy step, so it is kept out of `value`)
    dev.timing_enable(True)
    timed(step_device, args.steps, on_retry=lambda: dev.timing_enable(True), label="step_device_with_spans")  # enabling clears the spans
    spans = dev.timing_read()
    dev.timing_enable(False)
    clocks = sampler.stop(t0, time.time())
    ms_step = ms / args.steps
    value = world * CELLS / (ms_step / 1e3)

    # end to end through host buffers
    for _ in range(2):
        step_host().release()
    for t_ in pending:
        transporter.retire(t_)
    pending.clear()  # nothing transported before the timed region counts for it
    torch.cuda.synchronize()
    ms_e2e, roots_e2e, _, _, proof_e2e = timed(step_host, args.steps)
    for t_ in pending:
        transporter.retire(t_)
    pending.clear()
    e2e_value = world * CELLS / (ms_e2e / args.steps / 1e3)
    for _ in range(2):
        step_host_latency().release()
    ms_lat, roots_lat, _, _, proof_lat = timed(step_host_latency, args.steps)
    assert np.array_equal(proof.words(), proof_lat.words()), "device and host paths produce different proofs"
    assert all(np.array_equal(a, b) for a, b in zip(roots, roots_e2e)), "device and host paths disagree"
    assert np.array_equal(proof.words(), proof_e2e.words()), "device and host paths produce different proofs"

    # N > 1 also measures ONE commitment of the same shape sharded over the ranks (SURVEY section 8e commit row): column-sharded
    # RS encode, row exchange by one kernel over NVLink peer memory, local fused leaf hash, all-gather of 32-byte sub-roots
    sharded = None
    if world > 1:
        try:
            sharded = multi.sharded_commit_benchmark(dev, LOG_ROWS, COLS, L_SKIP, LOG_BLOWUP, K_WHIR, world, rank)
        except Exception as e:  # the replica measurement above stands on its own
            sharded = {"error": f"{type(e).__name__}: {e}"}

    # ... and ONE PROOF over the N GPUs (multi.ShardedProver: commitment sharded over the ranks, sumcheck phases on rank 0,
    # WHIR openings gathered from the ranks that hold the queried rows), for BASELINE configs[1] and configs[2]; the proof
    # bytes must equal the single-GPU proof's
    sharded_proofs = None
    if world > 1:
        sharded_proofs = []
        for which in ("c2", "c3"):
            try:
                sharded_proofs.append(sharded_proof_benchmark(dev, world, rank, which))
            except Exception as e:
                sharded_proofs.append({"workload": which, "error": f"{type(e).__name__}: {e}"})

    # Throughput with two provers on one GPU (N = 1 only): Coordinators on separate OS threads, each with its own library
    # context and stream -- the reference's concurrency contract (cuda-backend/examples/keccakf.rs) -- so that one proof's
    # latency-bound sumcheck rounds overlap the other's hash-bound commit.  Reported beside the headline, which stays one
    # proof at a time.
    concurrent = None
    if world == 1:
        try:
            concurrent = concurrent_provers_benchmark(sb, torch, params, air, trace_dev, vk_pre_hash, pk, 2, max(3, min(args.steps, 6)), proof.words())
        except Exception as e:
            concurrent = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        pk_, pk_kind = peaks()
        # per family: avg ms per launch, launches per step, ms per step, algorithmic bytes per step (accounted by the
        # library for exactly the launches that were timed: leaf hashing and round 0)
        fam = {k: (v[0] / max(v[1], 1), v[1] // args.steps, v[0] / args.steps, v[2] / args.steps) for k, v in spans.items()}
        dom = max((k for k in fam if fam[k][3] > 0), key=lambda k: fam[k][2])
        dom_ms = fam[dom][0]
        ach = fam[dom][3] / (fam[dom][2] / 1e3) / 1e9 if fam[dom][2] else 0.0
        leaf_perms = (1 << (LOG_ROWS + LOG_BLOWUP)) * (COLS // 8) + ((1 << (LOG_ROWS + LOG_BLOWUP)) - ((1 << (LOG_ROWS + LOG_BLOWUP)) >> K_WHIR))
        names = {"leaf": "leaf_tree_kernel (fused Poseidon2 row sponge + 2^k_whir strided tree levels)",
                 "bc_round0": "batch_round0_kernel (constraint DAG evaluation on the cosets of the skip domain)"}
        lde_ms = sum(fam[k][2] for k in ("chunk", "ntt_pass", "ntt_final"))
        lde_bytes = 4 * (1 << LOG_ROWS) * COLS * (1 + (1 << LOG_BLOWUP))
        out = {
            "metric": "trace_cells_per_s", "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 (BabyBear Montgomery)", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "l2": "per-step working set (1 GiB trace, 2 GiB codeword, 8 GiB GKR tree) exceeds the 126 MB L2; no flush needed",
                       "parallelism": f"{world} independent proofs (one per GPU), commitments all-gathered" if world > 1 else "single GPU"},
            "e2e": {"value": e2e_value, "unit": "cells/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": 4 * CELLS, "d2h_bytes_per_step": int(proof.words().size * 4),
                    "host_affinity": (f"rank bound to the {len(numa_cpus)} CPUs local to its GPU before pinning the trace" if numa_cpus
                                      else "default"),
                    "path": "TraceTransporter (pinned double buffering: the next step's trace is copied while this step proves; the "
                            "first step waits for its own copy) + Coordinator.prove; K + 1 copies of 1 GiB in the K timed steps",
                    "single_proof_ms": ms_lat / args.steps,
                    "single_proof_note": "Coordinator.prove_host: one proof in isolation, its H2D pipelined with RS encode and leaf "
                                         "hashing by column groups inside swirl_commit_host"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline_block(names[dom], dom, fam, ach, pk_, pk_kind, leaf_perms, clocks),
            "phases_ms_per_step": {k: v[2] for k, v in fam.items()},
            "phases_note": "kernel families with CUDA-event spans only; GKR tree/leaves, stacked reduction, WHIR and host latency are the rest",
            "lde": {"ms_per_step": lde_ms, "algorithmic_gb_s": lde_bytes / (lde_ms / 1e3) / 1e9 if lde_ms else 0.0,
                    "frac_of_hbm": (lde_bytes / (lde_ms / 1e3) / 1e9) / pk_["hbm_gbs"] if lde_ms else 0.0},
            "sharded_commit": sharded,
            "sharded_proof": sharded_proofs,
            "concurrent_provers": concurrent,
            "proof_bytes": len(proof.encode()),  # Proof::encode_to_vec() wire format (stark-backend_b200/codec.py)
            "host_step_ms": stalls.pop("step_ms"),
            "remeasured_after_host_stall": stalls or None,
        }
        if world == 1 and not args.no_cpu:
            oracle = load_oracle()
            cpu_prove_sample(oracle, 12)
            v, dt, sample = cpu_prove_sample(oracle, 16)
            out["cpu_baseline"] = {"value": v, "unit": "cells/s", "cores": os.cpu_count(), "kind": "port", "sample": sample,
                                   "seconds": dt, "note": CPU_NOTE, "full_size": CPU_FULL_SIZE}
        emit_json(out)
    del trace_dev
    torch.cuda.synchronize()
    dev.close()
    if world > 1:
        # the line is out; leave without the collective tear-down (a rank that failed in an extra must not keep the
        # launcher waiting on the others' communicator destructors)
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def concurrent_provers_benchmark(sb, torch, params, air, trace_dev, vk_pre_hash, pk, n_provers, proofs_each, expect_words):
    import threading

    devs = [sb.B200Device(trace_dev.device.index) for _ in range(n_provers)]
    ctx = sb.AirProvingContext(air.nodes, air.constraint_idx, air.interactions, 2, False, sb.DeviceMatrix(trace_dev, 1 << LOG_ROWS, COLS))
    last, errors = [None] * n_provers, []

    def prove(i):
        proof = sb.Coordinator(devs[i], params).prove(vk_pre_hash, pk, [(0, ctx, [])])
        proof.common_main_pcs.free()
        return proof

    def body(i, n):
        try:
            with torch.cuda.stream(devs[i].torch_stream()):
                for _ in range(n):
                    last[i] = prove(i).words().copy()
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    def run(n):
        th = [threading.Thread(target=body, args=(i, n)) for i in range(n_provers)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    try:
        run(2)  # warm-up: twiddles, compiled programs, scratch of the extra contexts
        dt = run(proofs_each)
    finally:
        for d in devs:
            d.close()
    n = n_provers * proofs_each
    return {"provers": n_provers, "proofs": n, "ms_per_proof": 1e3 * dt / n, "cells_per_s": n * CELLS / dt,
            "timing": "host wall clock around the threads, device synchronised on both sides",
            "proofs_identical_to_single_prover": bool(all(w is not None and np.array_equal(w, expect_words) for w in last)),
            "errors": errors}


def sharded_proof_benchmark(dev, world, rank, which):
    """One proof over all ranks (strong scaling of a single proof): returns, on rank 0, ms per proof (wall clock around the
    collective call, best of 3 after a warm-up; the ranks enter together) beside the same proof on one GPU."""
    import torch
    import torch.distributed as dist

    import airs as A
    import stark_backend_b200 as sb
    from stark_backend_b200 import multi

    if which == "c2":
        name, log_stack = "BASELINE configs[1]: BenchmarkAir 2^20 x 256", 20
        specs = [(benchmark_air_dag(COLS), 1 << LOG_ROWS, COLS)]
    else:
        name, log_stack = "BASELINE configs[2]: 32 BenchmarkAirs 2^17 x 20 with LogUp (2^22 rows), stacked height 2^24", 24
        specs = [(A.benchmark(3, 20, 20, 3, np.random.default_rng(i)), 1 << 17, 20) for i in range(32)]
    params = sb.SystemParams(L_SKIP, log_stack - L_SKIP, LOG_BLOWUP, sb.WhirConfig(K_WHIR, whir_queries(log_stack), MU_POW, QUERY_POW, FOLD_POW),
                             LOGUP_POW, MAX_CONSTRAINT_DEGREE)
    stacked_width = -(-sum(h * w for _, h, w in specs) >> log_stack)
    if stacked_width < world:  # the commitment is sharded by stacked columns: every rank needs at least one (same answer on all ranks)
        return {"workload": name, "n_gpus": world, "skipped": f"stacked matrix has {stacked_width} columns, fewer than the {world} ranks"} if rank == 0 else None
    g = torch.Generator(device=dev.torch_device).manual_seed(4242)  # the same traces on every rank
    per_trace = []
    for i, (air, h, w) in enumerate(specs):
        t = torch.randint(0, 2, (h * w,), dtype=torch.int32, device=dev.torch_device, generator=g) * R1
        per_trace.append((i, sb.AirProvingContext(air.nodes, air.constraint_idx, air.interactions, 2, False, sb.DeviceMatrix(t, h, w)), []))
    pk = [sb.AirProvingKey(True, None) for _ in specs]
    vk = np.arange(8, dtype=np.uint32)
    sp = multi.ShardedProver(dev, params, world, rank)
    best, words = None, None
    for rep in range(4):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        proof = sp.prove(vk, pk, per_trace)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if rep:
            best = dt if best is None else min(best, dt)
        if proof is not None:
            words = proof.words()
            proof.common_main_pcs.free()
            proof.release()
    out = None
    if rank == 0:
        single = []
        for rep in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ref = sb.Coordinator(dev, params).prove(vk, pk, per_trace)
            dev.synchronize()
            single.append(time.perf_counter() - t0)
            same = bool(np.array_equal(ref.words(), words))
            ref.common_main_pcs.free()
            ref.release()
        cells = sum(h * w for _, h, w in specs)
        out = {"workload": name, "n_gpus": world, "ms_per_proof": best * 1e3, "single_gpu_ms_per_proof": min(single[1:]) * 1e3,
               "speedup": min(single[1:]) / best, "cells_per_s": cells / best, "proof_identical_to_single_gpu": same,
               "sharded_commit_ms": sp.timings.get("sharded_commit_ms"), "rank0_rest_ms": sp.timings.get("rank0_rest_ms"),
               "scaling": "strong", "sharded": "commitment (RS encode by columns, leaf hashing by query ranges); sumcheck phases on rank 0"}
    dist.barrier()
    return out


def roofline_block(kernel_name, dom, fam, hbm_ach, pk_, pk_kind, leaf_perms, clocks):
    """The dominant kernel family against BOTH rooflines; `bound` names the binding one.  The Poseidon2 leaf kernel is bound by
    32-bit integer issue, so `achieved`/`peak`/`frac` are its integer figures and the HBM figures sit beside
    them as `hbm`; for an HBM-bound dominant kernel it is the other way round."""
    ip = int32_peak()
    tr = ncu_traffic().get(dom, {})
    hbm = {"achieved": hbm_ach, "peak": pk_["hbm_gbs"], "unit": "GB/s", "frac": hbm_ach / pk_["hbm_gbs"], "peak_source": pk_kind + " (MEASURED_PEAKS.json)"}
    common = {"kernel": kernel_name, "ms_per_launch": fam[dom][0], "launches_per_step": fam[dom][1],
              "algorithmic_bytes_per_step": fam[dom][3], "traffic": tr.get("bytes_per_launch"), "traffic_note": tr.get("note"),
              "achieved_definition": "algorithmic units of all launches of the family in a step / their summed duration",
              "timing": "CUDA events around every launch of the family, over a repeat of the K timed steps"}
    if dom != "leaf":
        return {**common, "bound": "hbm", **hbm}
    # one permutation = 564 S-box Montgomery products (IMAD.WIDE + IMAD + IMAD.HI = 5 multiplier-pipe passes) + 91 IMADs of
    # the internal diagonal = 2911 passes of the fma-heavy pipe: the multiplier roofline is the measured IMAD rate / 2911
    passes = 564 * 5 + 91
    ms = fam["leaf"][2]
    ach = leaf_perms / (ms / 1e3) / 1e9 if ms else 0.0
    peak = ip["imad_tops"] * 1e12 / passes / 1e9
    return {**common, "bound": "int32", "achieved": ach, "peak": peak, "unit": "Gperm/s", "frac": ach / peak,
            "peak_source": f"measured IMAD issue rate {ip['imad_tops']} Tops/s ({ip.get('source', 'profiles/int32_peaks.json')}) / {passes} "
                           "multiplier passes per permutation",
            "perms_per_step": leaf_perms,
            "issue_bound": {"note": "the permutation is ~5800 thread instructions; the measured mixed IMAD+ALU issue rate of the chip "
                                    "bounds it more tightly than the multiplier pipe alone",
                            "peak_gperm_s": ip.get("mixed_issue_tops", 24.1) * 1e12 / 5800 / 1e9,
                            "frac": ach / (ip.get("mixed_issue_tops", 24.1) * 1e12 / 5800 / 1e9)},
            "hbm": hbm,
            "reference_gpu": "the reference's own Poseidon2 kernels compiled for sm_100 reach 3.0 Gperm/s on the same box "
                             "(profiles/r2a_reference_gpu_kernels_vs_swirl.jsonl)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="swirl", choices=["swirl", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_swirl(args)


if __name__ == "__main__":
    main()
