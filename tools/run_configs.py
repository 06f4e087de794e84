"""Runs the BASELINE.json configs other than the bench line on one B200 and writes JSON lines:
  C1  Fibonacci AIR 2^16 rows, SystemParams::new_for_testing(16)            full proof, oracle verifier chain
  C2  BenchmarkAir 2^20 x 256, app_params_with_100_bits_security(20)         full proof, oracle verifier chain
  C3  32 BenchmarkAirs 2^17 x 20 with LogUp, app params, stacked height 2^24 full proof, oracle verifier chain
  C4  2^24 x 512 trace, log_blowup 1: LDE + Poseidon2 Merkle commit          (size-independent checks only)
  C5  micro sweeps: batched NTT, Poseidon2 permute / compress / leaf hash
The verifier never reads traces, so the oracle's verifier chain (tests/verify_chain.py) is run on the
full-size GPU proofs.   python tools/run_configs.py [c1 c2 c3 c4 c5]"""
import json, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import stark_backend_b200 as sb
import airs as A, oracle_lib, verify_chain

which = [a.lower() for a in sys.argv[1:]] or ["c1", "c2", "c3", "c4", "c5"]
dev = sb.B200Device(0)
oracle = oracle_lib.Oracle(os.path.join(ROOT, "oracle", "libswirl_oracle.so"))
R1 = 0x0FFFFFFE
def emit(**kw):
    print(json.dumps(kw), flush=True)

def whir_cfg(log_blowup, log_h, k, lfp, qpow, fpow, mpow, sec, list_from=None, m=3):
    level, rate, nq = max(sec - qpow, 0), log_blowup, []
    for rnd in range(-(-(log_h - lfp) // k)):
        rho = 2.0 ** (-rate)
        agree = (1 + rho) / 2 if (list_from is None or rnd < list_from) else math.sqrt(rho) * (1 + 1 / (2 * m))
        nq.append(math.ceil(level / -math.log2(agree)))
        rate += k - 1
    return dict(k=k, num_queries=nq, mu_pow_bits=mpow, query_phase_pow_bits=qpow, folding_pow_bits=fpow)

def prove_and_verify(name, airs, traces, l_skip, n_stack, log_blowup, D, logup_pow, whir, reps=3, cache=None, plan_memory=False):
    """airs: A.Air with shape-only common_main; traces: CUDA int32 tensors."""
    params = sb.SystemParams(l_skip, n_stack, log_blowup, sb.WhirConfig(**whir), logup_pow, D)
    vk = np.arange(8, dtype=np.uint32)
    pk = [sb.AirProvingKey(True, None) for _ in airs]
    per_trace = [(i, sb.AirProvingContext(a.nodes, a.constraint_idx, a.interactions, a.constraint_degree, a.need_rot,
                                          sb.DeviceMatrix(t, a.common_main[1], a.common_main[2]), a.public_values), [])
                 for i, (a, t) in enumerate(zip(airs, traces))]
    cells = sum(a.common_main[1] * a.common_main[2] for a in airs)
    best = None
    if cache is not None:
        dev.set_cache_rs_code_matrix(cache)
    dev.trim()
    dev.mem_stats(reset_peak=True)
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        coord = sb.Coordinator(dev, params, plan_memory=plan_memory)
        proof = coord.prove(vk, pk, per_trace)
        dev.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        if _ != reps - 1:
            proof.common_main_pcs.free()
    t0 = time.perf_counter()
    ok, where = verify_chain.verify(oracle, l_skip, n_stack, log_blowup, D, logup_pow, whir, vk, airs, [True] * len(airs),
                                    proof.common_main_commit, [(None, [])] * len(airs), proof.constraints_proof,
                                    proof.stacking_proof, proof.whir_proof)
    tv = time.perf_counter() - t0
    proof.common_main_pcs.free()
    mem = dev.mem_stats()
    dev.set_cache_rs_code_matrix(True)
    emit(config=name, arena_peak_gib=mem["peak"] / 2**30, cache_rs_code_matrix=cache, memory_plan=coord.memory_estimate,
         prove_ms=best * 1e3, trace_cells=cells, cells_per_s=cells / best, proof_bytes=int(proof.words().size * 4),
         phase_ms=coord.phase_ms, oracle_verifier_accepts=bool(ok is True), failed_stage=None if ok is True else where, verify_s=tv, whir=whir,
         params=dict(l_skip=l_skip, n_stack=n_stack, log_blowup=log_blowup, max_constraint_degree=D, logup_pow_bits=logup_pow))

def shape_only(a, h, w):
    a.common_main = (np.zeros(0, np.uint32), h, w)
    return a

if "c1" in which:
    fib = A.fibonacci(16)
    t = dev.h2d(fib.common_main[0])
    prove_and_verify("C1 Fibonacci 2^16 (new_for_testing(16))", [shape_only(fib, 1 << 16, 2)], [t], 4, 12, 1, 4, 2,
                     whir_cfg(1, 16, 4, 0, 1, 2, 3, 5, list_from=1))
if "c2" in which:
    air = shape_only(A.benchmark(3, 256, 256, 32, np.random.default_rng(0)), 1 << 20, 256)
    g = torch.Generator(device="cuda").manual_seed(42)
    t = torch.randint(0, 2, ((1 << 20) * 256,), dtype=torch.int32, device="cuda", generator=g) * R1
    prove_and_verify("C2 BenchmarkAir 2^20 x 256 (app params, stacked height 2^20)", [air], [t], 4, 16, 1, 3, 18,
                     whir_cfg(1, 20, 4, 10, 20, 5, 15, 100))
    del t
if "c2h24" in which:
    # uniform_runner's own defaults for configs[1]: --log-stacked-height 24 (H = 2^24, W = 16 stacked columns; the 256 trace
    # columns of 2^20 rows are stacked 16 to a column) and the all-zero trace (benchmarks/synthetic/src/bin/uniform_runner.rs:72,250-267)
    air = shape_only(A.benchmark(3, 256, 256, 32, np.random.default_rng(0)), 1 << 20, 256)
    t = torch.zeros((1 << 20) * 256, dtype=torch.int32, device="cuda")
    prove_and_verify("C2 BenchmarkAir 2^20 x 256, all-zero trace, stacked height 2^24 (uniform_runner defaults)", [air], [t], 4, 20, 1, 3, 18,
                     whir_cfg(1, 24, 4, 10, 20, 5, 15, 100))
    prove_and_verify("C2 BenchmarkAir 2^20 x 256, all-zero trace, stacked height 2^20", [air], [t], 4, 16, 1, 3, 18,
                     whir_cfg(1, 20, 4, 10, 20, 5, 15, 100))
    del t
if "c4p" in which:
    # BASELINE configs[3] as a FULL proof: 2^24 x 512 (32 GiB trace, 64 GiB codeword), 512 boolean constraints, 3 send/receive
    # pairs = 6 x 2^24 LogUp leaves, the most that fits the 2^27-leaf GKR layout here (the uniform_runner default of 0.25 interactions per column would need 64 GiB of LogUp leaves and as much again
    # for the fraction tree at this height -- more than the device next to the trace; reported as out of memory, not shrunk silently)
    h, w = 1 << 24, 512
    air = shape_only(A.benchmark(3, w, w, 3, np.random.default_rng(0)), h, w)
    g = torch.Generator(device="cuda").manual_seed(4)
    t = torch.randint(0, 2, (h * w,), dtype=torch.int32, device="cuda", generator=g) * R1
    for cache in (False, True):
        prove_and_verify(f"C4 BenchmarkAir 2^24 x 512 full proof (app params, stacked height 2^24), cache_rs_code_matrix={cache}", [air], [t],
                         4, 20, 1, 3, 18, whir_cfg(1, 24, 4, 10, 20, 5, 15, 100), reps=2, cache=cache)
    prove_and_verify("C4 full proof, planner decides (Coordinator plan_memory)", [air], [t], 4, 20, 1, 3, 18,
                     whir_cfg(1, 24, 4, 10, 20, 5, 15, 100), reps=1, plan_memory=True)
    del t
if "c3" in which:
    airs, traces = [], []
    g = torch.Generator(device="cuda").manual_seed(7)
    for i in range(32):
        airs.append(shape_only(A.benchmark(3, 20, 20, 3, np.random.default_rng(i)), 1 << 17, 20))
        traces.append(torch.randint(0, 2, ((1 << 17) * 20,), dtype=torch.int32, device="cuda", generator=g) * R1)
    prove_and_verify("C3 32 BenchmarkAirs 2^17 x 20, 6 interactions each (app params, stacked height 2^24)", airs, traces,
                     4, 20, 1, 3, 18, whir_cfg(1, 24, 4, 10, 20, 5, 15, 100))
    del traces
if "c4" in which:
    h, w = 1 << 24, 512
    g = torch.Generator(device="cuda").manual_seed(4)
    t = torch.randint(0, sb.P, (h * w,), dtype=torch.int32, device="cuda", generator=g)
    params = sb.PcsParams(4, 20, 1, 4)
    best = None
    for _ in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        root, pcs = dev.commit(params, [sb.DeviceMatrix(t, h, w)])
        dev.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        # size-independent checks: opened rows of a few queries hash up to the root through their Merkle paths;
        # codeword restricted to the even positions of the RS domain of one column == its half-size encoding
        idx = [0, 1, 12345, pcs.tree.query_stride() - 1]
        rows = pcs.tree.get_opened_rows(idx)
        paths = pcs.tree.query_merkle_proofs(idx)
        ok = True
        for q, i in enumerate(idx):
            leaves = [oracle.hash_slice(rows[q, tt]) for tt in range(16)]
            while len(leaves) > 1:
                leaves = [oracle.compress(leaves[2 * j], leaves[2 * j + 1]) for j in range(len(leaves) // 2)]
            cur, ii = leaves[0], i
            for sib in paths[q]:
                cur = oracle.compress(cur, sib) if ii % 2 == 0 else oracle.compress(sib, cur)
                ii >>= 1
            ok &= bool(np.array_equal(cur, root))
        pcs.free()
    emit(config="C4 2^24 x 512 (blowup 2): LDE + Poseidon2 Merkle commit", commit_ms=best * 1e3, trace_cells=h * w,
         cells_per_s=h * w / best, merkle_paths_verify=ok, codeword_bytes=2 * h * w * 4)
    del t
if "c5" in which:
    def timeit(fn, reps=5):
        fn(); dev.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        dev.synchronize()
        return (time.perf_counter() - t0) / reps
    for log_n in (16, 18, 20, 22, 24, 26):
        for cols in (1, 16, 256):
            if (1 << log_n) * cols > (1 << 31):
                continue
            x = torch.randint(0, sb.P, ((1 << log_n) * cols,), dtype=torch.int32, device="cuda")
            dt = timeit(lambda: dev.ntt_batch(x, log_n, cols, False), 3)
            emit(config="C5 batched NTT", log_n=log_n, poly_count=cols, ms=dt * 1e3, elements_per_s=(1 << log_n) * cols / dt,
                 gb_s_alg=(1 << log_n) * cols * 8 / dt / 1e9)
            del x
    for log_n in (16, 20, 24, 26):
        s = torch.randint(0, sb.P, ((1 << log_n) * 16,), dtype=torch.int32, device="cuda")
        dt = timeit(lambda: dev.poseidon2_permute(s))
        emit(config="C5 Poseidon2 permute", log_n=log_n, ms=dt * 1e3, gperm_per_s=(1 << log_n) / dt / 1e9)
        dt = timeit(lambda: dev.poseidon2_compress(s))
        emit(config="C5 Poseidon2 compress", log_pairs=log_n, ms=dt * 1e3, gperm_per_s=(1 << log_n) / dt / 1e9)
        del s
    for width in (8, 64, 256, 512):
        h = 1 << 20
        m = sb.DeviceMatrix(torch.randint(0, sb.P, (h * width,), dtype=torch.int32, device="cuda"), h, width)
        out = dev.alloc((2 * (h >> 4) - 1) * 8 + 8)
        dt = timeit(lambda: dev.merkle_tree(m, 4, out=out), 3)
        perms = h * (-(-width // 8)) + h - 1
        emit(config="C5 Poseidon2 leaf hash + tree", rows=h, width=width, ms=dt * 1e3, gperm_per_s=perms / dt / 1e9)
    # FRI/WHIR-style fold of an EF table by one challenge (fold_mle): 32 B read + 16 B written per output
    for log_n in (16, 20, 24, 26):
        t = torch.randint(0, sb.P, ((1 << log_n) * 4,), dtype=torch.int32, device="cuda")
        o = dev.alloc((1 << (log_n - 1)) * 4)
        r = np.array([5, 6, 7, 8], dtype=np.uint32)
        dt = timeit(lambda: dev.fold_mle(t, r, out=o), 5)
        emit(config="C5 fold (fold_mle of 2^log_n EF)", log_n=log_n, ms=dt * 1e3, ef_per_s=(1 << log_n) / dt,
             gb_s_alg=(1 << log_n) * 24 / dt / 1e9)
        del t, o
