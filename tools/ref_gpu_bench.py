"""The perf bar SURVEY §2b names: the REFERENCE's own CUDA kernels, recompiled for sm_100 (oracle/_ref/libref_kernels.so),
timed on the same B200 beside the product's kernels, family by family.  JSON lines on stdout.

  commit  rs_code_matrix (stacked_pcs.rs:229-337) + MerkleTreeGpu::new (merkle_tree.rs:140-197) vs swirl_rs_encode +
          swirl_merkle_tree, at C2 (2^20 x 256) and, with `c4`, C4 (2^24 x 512)
  ntt     batch_ntt natural->natural (ntt.rs:111-168: bit_rev + CT passes) vs swirl_ntt_batch, lg n 16..26
  p2      _poseidon2_compressing_row_hashes (leaf sponge, widths 8/64/256/512) and _poseidon2_adjacent_compress_layer
          vs swirl_merkle_tree at rows_per_query 1 / the product's compress

    python tools/ref_gpu_bench.py [commit] [c4] [ntt] [p2]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import ref_kernels  # noqa: E402
import stark_backend_b200 as sb  # noqa: E402

P = sb.P


def emit(**kw):
    print(json.dumps(kw), flush=True)


def time_ms(fn, stream, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def bench_commit(dev, rk, log_h, width, l_skip=4, lb=1, k=4, reps=5):
    H = 1 << log_h
    trace = torch.randint(0, P, (H * width,), dtype=torch.int32, device="cuda")
    cw = torch.empty((H << lb) * width, dtype=torch.int32, device="cuda")
    cur = torch.cuda.current_stream()
    # the reference's mle_interpolate_fused_2d_kernel indexes the buffer with 32 bits (`uint32_t base_idx = col * padded_height
    # + physical_idx`, cuda/src/mle_interpolate.cu:89): a codeword of 2^32 elements or more (C4: 2^34) wraps around and the
    # root comes out wrong, so the reference is run on column groups below 2^32 elements
    cwh = H << lb
    grp = width if cwh * width < (1 << 32) else max(1, ((1 << 32) - 1) // cwh // 2)

    def ref_rs():
        for c0 in range(0, width, grp):
            nc = min(grp, width - c0)
            rk.rs_code_matrix(trace[c0 * H:(c0 + nc) * H], H, nc, l_skip, lb, out=cw[c0 * cwh:(c0 + nc) * cwh])

    ref_lde = time_ms(ref_rs, cur, reps)
    ref_tree = time_ms(lambda: rk.merkle_tree(cw, H << lb, width, 1 << k), cur, reps)
    ref_root = rk.d2h(rk.merkle_tree(cw, H << lb, width, 1 << k)[-1]).tolist()
    m = sb.DeviceMatrix(trace, H, width)
    st = dev.torch_stream()
    qs = (H << lb) >> k
    layers = dev.alloc((2 * qs - 1) * 8)
    our_lde = time_ms(lambda: dev.rs_encode(m, l_skip, lb, out=cw), st, reps)
    cwm = sb.DeviceMatrix(cw, H << lb, width)
    our_tree = time_ms(lambda: dev.merkle_tree(cwm, k, out=layers), st, reps)
    dev.synchronize()
    our_root = layers.cpu().numpy().view("uint32")[-8:].tolist()
    cells = H * width
    emit(bench="commit", log_h=log_h, width=width, l_skip=l_skip, log_blowup=lb, k_whir=k,
         reference_sm100_ms=dict(rs_code_matrix=ref_lde, merkle_tree=ref_tree, total=ref_lde + ref_tree),
         swirl_ms=dict(rs_encode=our_lde, merkle_tree=our_tree, total=our_lde + our_tree),
         speedup=dict(lde=ref_lde / our_lde, merkle=ref_tree / our_tree, total=(ref_lde + ref_tree) / (our_lde + our_tree)),
         reference_gcells_s=cells / (ref_lde + ref_tree) / 1e6, swirl_gcells_s=cells / (our_lde + our_tree) / 1e6,
         reference_column_groups=-(-width // grp), roots_equal=bool(ref_root == our_root))
    del trace, cw


def bench_ntt(dev, rk):
    for log_n in range(16, 27, 2):
        for cols in (1, 16, 256):
            if (cols << log_n) > (1 << 32):
                continue
            x = torch.randint(0, P, (cols << log_n,), dtype=torch.int32, device="cuda")
            ref = time_ms(lambda: rk.batch_ntt(x, log_n, 0, cols, True, False), torch.cuda.current_stream(), 3, 1)
            ours = time_ms(lambda: dev.ntt_batch(x, log_n, cols, False), dev.torch_stream(), 3, 1)
            n = cols << log_n
            emit(bench="ntt", log_n=log_n, cols=cols, reference_sm100_ms=ref, swirl_ms=ours, speedup=ref / ours,
                 reference_gelem_s=n / ref / 1e6, swirl_gelem_s=n / ours / 1e6)
            del x


def bench_p2(dev, rk):
    log_rows = 21
    rows = 1 << log_rows
    for width in (8, 64, 256, 512):
        m = torch.randint(0, P, (rows * width,), dtype=torch.int32, device="cuda")
        out = torch.empty(rows * 8, dtype=torch.int32, device="cuda")
        ref = time_ms(lambda: rk.check(rk.L._poseidon2_compressing_row_hashes(out.data_ptr(), m.data_ptr(), width, rows, 0,
                                                                                rk.stream()), "row_hashes"),
                      torch.cuda.current_stream(), 3, 1)
        dm = sb.DeviceMatrix(m, rows, width)
        layers = dev.alloc((2 * rows - 1) * 8)
        ours = time_ms(lambda: dev.merkle_tree(dm, 0, out=layers), dev.torch_stream(), 3, 1)  # leaves + the whole tree above
        perms = rows * -(-width // 8)
        emit(bench="p2_leaf", rows=rows, width=width, reference_sm100_ms=ref, swirl_ms_incl_tree=ours,
             reference_gperm_s=perms / ref / 1e6, swirl_gperm_s=(perms + rows - 1) / ours / 1e6)
        del m, out, layers
    n = 1 << 24
    prev = torch.randint(0, P, (n * 16,), dtype=torch.int32, device="cuda")
    nxt = torch.empty(n * 8, dtype=torch.int32, device="cuda")
    ref = time_ms(lambda: rk.check(rk.L._poseidon2_adjacent_compress_layer(nxt.data_ptr(), prev.data_ptr(), n, rk.stream()),
                                   "compress"), torch.cuda.current_stream(), 3, 1)
    ours = time_ms(lambda: dev.poseidon2_compress(prev), dev.torch_stream(), 3, 1)
    emit(bench="p2_compress", pairs=n, reference_sm100_ms=ref, swirl_ms=ours, reference_gperm_s=n / ref / 1e6,
         swirl_gperm_s=n / ours / 1e6)


if __name__ == "__main__":
    which = sys.argv[1:] or ["commit", "ntt", "p2"]
    dev = sb.B200Device(0)
    rk = ref_kernels.RefKernels()
    emit(gpu=torch.cuda.get_device_name(0), reference_build="oracle/Makefile.ref: nvcc -O3 -gencode arch=compute_100,code=sm_100")
    if "commit" in which:
        bench_commit(dev, rk, 20, 256)
        bench_commit(dev, rk, 24, 16)  # C2 at uniform_runner's default --log-stacked-height 24
    if "c4" in which:
        bench_commit(dev, rk, 24, 512, reps=2)
    if "ntt" in which:
        bench_ntt(dev, rk)
    if "p2" in which:
        bench_p2(dev, rk)
