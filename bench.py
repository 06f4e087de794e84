#!/usr/bin/env python
"""bench.py — headline measurement of the SWIRL commit hot path on B200.

One "step" = one stacked_commit (stacking + Reed–Solomon LDE + Poseidon2 Merkle tree) of
BASELINE.json configs[1]: a single AIR of 2^20 rows x 256 columns of uniform random BabyBear
elements, app parameters (l_skip 4, log_blowup 1, k_whir 4), stacked height 2^20 (W = 256,
codeword 2^21 x 256).  Metric: trace cells per second.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl swirl|reference]

`value`   : device-timed (CUDA events on the library's stream), traces resident in HBM.
`e2e`     : the same through swirl_commit_host with pinned HOST buffers (H2D inside the timing).
`roofline`: the dominant kernel (fused leaf hash) timed live with CUDA events inside the region.
`--impl reference`: the CPU restatement of the reference algorithm (oracle/, kind "port": no Rust
toolchain exists in this image so the reference itself cannot be built) on all host cores.
N > 1: one process per GPU (torchrun), every rank commits its own AIR trace (independent
commitments, weak scaling), the 32-byte roots are all-gathered over NCCL.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

LOG_ROWS, COLS = 20, 256
L_SKIP, LOG_BLOWUP, K_WHIR = 4, 1, 4
N_STACK = LOG_ROWS - L_SKIP
CELLS = (1 << LOG_ROWS) * COLS
WORKLOAD = (
    "BASELINE configs[1]: 1 AIR 2^20 rows x 256 cols (uniform random BabyBear, seed 42), app params "
    "l_skip=4 log_blowup=1 k_whir=4 n_stack=16 -> stacked 2^20 x 256, codeword 2^21 x 256: "
    "stack + RS/LDE + Poseidon2 Merkle commit (TraceCommitter::commit)"
)
P = 0x78000001


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc:
            self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2] or [r for _, r in self.rows]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(rows[0][1]), "reasons": reasons,
                "samples": len(rows)}


def gen_trace(seed):
    rng = np.random.default_rng(seed)
    canon = rng.integers(0, P, size=CELLS, dtype=np.uint64)
    return ((canon << np.uint64(32)) % np.uint64(P)).astype(np.uint32)  # Montgomery words


def load_oracle():
    import oracle_lib

    path = os.path.join(ROOT, "oracle", "libswirl_oracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    return oracle_lib.Oracle(path)


def cpu_commit_sample(oracle, log_rows, reps=1):
    """Times the oracle's stacked_commit on a bounded sample: 2^log_rows rows x 256 cols, same
    parameters (n_stack shrunk with the height).  Returns (cells/s, seconds, description)."""
    rng = np.random.default_rng(42)
    h = 1 << log_rows
    canon = rng.integers(0, P, size=h * COLS, dtype=np.uint64)
    vals = ((canon << np.uint64(32)) % np.uint64(P)).astype(np.uint32)
    best = None
    for _ in range(reps):
        t = time.perf_counter()
        oracle.stacked_commit(L_SKIP, log_rows - L_SKIP, LOG_BLOWUP, K_WHIR, [(vals, h, COLS)], want_codeword=False)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    return h * COLS / best, best, f"2^{log_rows} rows x {COLS} cols (1/{1 << (LOG_ROWS - log_rows)} of the workload), same params"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    oracle = load_oracle()
    cores = os.cpu_count() or 1
    log_rows = 15
    for _ in range(args.warmup):
        cpu_commit_sample(oracle, 12)
    times = []
    for _ in range(args.steps):
        v, dt, sample = cpu_commit_sample(oracle, log_rows)
        times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = (1 << log_rows) * COLS / (ms / 1e3)
    print(json.dumps({
        "impl": "reference", "metric": "trace_cells_per_s", "value": value, "unit": "cells/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 (BabyBear Montgomery)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def int_pipe_roofline(perms, ms, clocks):
    """Integer-multiplier roofline of a Poseidon2 kernel.  Measured on B200 (tools/int_roofline.cu,
    profiles/r1_p2_iterate_ncu.txt): 32-bit integer multiplies issue only on the fma-heavy pipe,
    64 lanes/clk/SM, IMAD 1 pass, IMAD.WIDE / IMAD.HI 2 passes.  A Montgomery product needs
    IMAD.WIDE + IMAD + IMAD.HI = 5 passes; one permutation has 564 S-box products and 91 IMADs of
    the internal diagonal => 2911 passes minimum."""
    passes = 564 * 5 + 91
    mhz = (clocks or {}).get("sm_max_mhz") or 1965.0
    peak = 64 * 148 * mhz * 1e6 / passes / 1e9
    ach = perms / (ms / 1e3) / 1e9 if ms else 0.0
    return {"perms_per_launch": perms, "gperm_per_s": ach, "peak_gperm_per_s": peak, "frac": ach / peak,
            "model": "fma-heavy pipe, 64 lanes/clk/SM x 148 SM x sm_max_mhz / 2911 multiplier passes per permutation"}


def run_swirl(args):
    import torch
    import torch.distributed as dist

    import stark_backend_b200 as sb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = sb.B200Device(local)
    params = sb.PcsParams(L_SKIP, N_STACK, LOG_BLOWUP, K_WHIR)
    host = torch.from_numpy(gen_trace(42 + rank).view(np.int32)).pin_memory()
    trace = sb.DeviceMatrix(host.to(dev.torch_device), 1 << LOG_ROWS, COLS)
    stream = dev.torch_stream()
    roots_dev = torch.zeros(8, dtype=torch.int32, device=dev.torch_device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        root, pcs = dev.commit(params, [trace])
        pcs.free()
        return root

    def step_host():
        root, pcs = dev.commit_host(params, [(host, 1 << LOG_ROWS, COLS)])
        pcs.free()
        return root

    def gather(root):
        if world > 1:  # only the 32-byte roots cross NVLink
            roots_dev.copy_(torch.from_numpy(root.view(np.int32)))
            out = [torch.empty_like(roots_dev) for _ in range(world)]
            dist.all_gather(out, roots_dev)
            return [o.cpu().numpy().view(np.uint32) for o in out]
        return [root]

    def timed(fn, steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.time()
        a.record(stream)
        for _ in range(steps):
            root = fn()
        b.record(stream)
        roots = gather(root)
        barrier()
        t1 = time.time()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], device=dev.torch_device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, roots, t0, t1

    for _ in range(max(args.warmup, 3)):
        step_device()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    dev.timing_enable(True)
    l0 = dev.launch_count()
    ms, roots, t0, t1 = timed(step_device, args.steps)
    launches = dev.launch_count() - l0
    spans = dev.timing_read()
    dev.timing_enable(False)
    clocks = sampler.stop(t0, t1)
    ms_step = ms / args.steps
    value = world * CELLS / (ms_step / 1e3)

    # end to end through host buffers
    step_host()
    ms_e2e, roots_e2e, _, _ = timed(step_host, args.steps)
    e2e_value = world * CELLS / (ms_e2e / args.steps / 1e3)
    assert all(np.array_equal(a, b) for a, b in zip(roots, roots_e2e)), "device and host paths disagree"

    if rank == 0:
        pk, pk_kind = peaks()
        N, W = 1 << (LOG_ROWS + LOG_BLOWUP), COLS
        leaf_ms, leaf_n = spans["leaf"]
        leaf_avg = leaf_ms / max(leaf_n, 1)
        # algorithmic bytes of the dominant kernel: read the codeword once + write layer 0
        leaf_bytes = 4 * N * W + 32 * (N >> K_WHIR)
        leaf_perms = N * (W // 8) + (N - (N >> K_WHIR))
        ach = leaf_bytes / (leaf_avg / 1e3) / 1e9 if leaf_avg else 0.0
        lde_ms = sum(spans[k][0] for k in ("chunk", "ntt_pass", "ntt_final")) / args.steps
        lde_bytes = 4 * (1 << LOG_ROWS) * W * (1 + (1 << LOG_BLOWUP))
        out = {
            "metric": "trace_cells_per_s", "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 (BabyBear Montgomery)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "l2": "inputs (1 GiB trace, 2 GiB codeword) exceed the 126 MB L2; no flush needed",
                       "phase": "commit only (LDE + Merkle); full prove not yet in the timed step",
                       "parallelism": f"{world} independent per-AIR commits, roots all-gathered" if world > 1 else "single GPU"},
            "e2e": {"value": e2e_value, "unit": "cells/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": 4 * CELLS, "d2h_bytes_per_step": 32},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {
                "kernel": "leaf_tree_kernel (fused Poseidon2 row sponge + 2^k_whir strided tree levels)",
                "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                "peak_source": pk_kind, "traffic": None, "ms_per_launch": leaf_avg, "launches": leaf_n,
                "note": "kernel is INT32-issue bound (Poseidon2), not HBM bound; see int_pipe",
                "int_pipe": int_pipe_roofline(leaf_perms, leaf_avg, clocks),
            },
            "phases_ms_per_step": {k: v[0] / args.steps for k, v in spans.items()},
            "lde": {"ms_per_step": lde_ms, "algorithmic_gb_s": lde_bytes / (lde_ms / 1e3) / 1e9 if lde_ms else 0.0,
                    "frac_of_hbm": (lde_bytes / (lde_ms / 1e3) / 1e9) / pk["hbm_gbs"] if lde_ms else 0.0},
        }
        if world == 1 and not args.no_cpu:
            oracle = load_oracle()
            cpu_commit_sample(oracle, 12)
            v, dt, sample = cpu_commit_sample(oracle, 16)
            out["cpu_baseline"] = {"value": v, "unit": "cells/s", "cores": os.cpu_count(), "kind": "port",
                                   "sample": sample, "seconds": dt}
        print(json.dumps(out))
    dev.close()
    if world > 1:
        dist.destroy_process_group()


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
