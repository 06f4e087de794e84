"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI, against the CPU
oracle on the same seeded inputs — bit-exact (all arithmetic is integer)."""
import json
import os

import numpy as np
import pytest
import torch

import stark_backend_b200 as sb

pytestmark = pytest.mark.gpu
P = sb.P
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))


def _dm(dev, vals, h, w):
    return sb.DeviceMatrix(dev.h2d(vals), h, w)


# ---------------------------------------------------------------------------------------------
# Poseidon2
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 31, 256, 257, 5000])
def test_poseidon2_permute(dev, oracle, n):
    rng = np.random.default_rng(n)
    st = oracle.random_field(rng, n * 16)
    if n >= 2:
        st[:16] = 0
        st[16:32] = oracle.to_mont(np.arange(16))
    d = dev.h2d(st)
    dev.poseidon2_permute(d)
    dev.synchronize()
    got = d.cpu().numpy().view(np.uint32)
    assert np.array_equal(got, oracle.permute(st))
    if n >= 2:
        assert list(sb.from_mont(got[:4])) == GOLD["poseidon2_selfcheck"]["zeros_first4"]
        assert list(sb.from_mont(got[16:20])) == GOLD["poseidon2_selfcheck"]["iota_first4"]


def test_poseidon2_compress(dev, oracle):
    rng = np.random.default_rng(11)
    pairs = oracle.random_field(rng, 300 * 16)
    out = dev.poseidon2_compress(dev.h2d(pairs))
    dev.synchronize()
    got = out.cpu().numpy().view(np.uint32).reshape(300, 8)
    for i in (0, 1, 150, 299):
        assert np.array_equal(got[i], oracle.compress(pairs[i * 16 : i * 16 + 8], pairs[i * 16 + 8 : i * 16 + 16]))


# ---------------------------------------------------------------------------------------------
# NTT
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 4, 5, 8, 10, 11, 12, 13, 16])
@pytest.mark.parametrize("inverse", [False, True])
def test_ntt_batch_default_plan(dev, oracle, log_n, inverse):
    rng = np.random.default_rng(100 + log_n)
    cols = 3 if log_n > 12 else 37
    n = 1 << log_n
    x = oracle.random_field(rng, n * cols)
    d = dev.h2d(x)
    dev.ntt_batch(d, log_n, cols, inverse)
    dev.synchronize()
    assert np.array_equal(d.cpu().numpy().view(np.uint32), oracle.dft_batch(x, n, cols, inverse))


@pytest.mark.parametrize("max_radix,log_n", [(2, 4), (2, 5), (2, 6), (3, 7), (3, 9), (4, 12), (5, 11), (6, 17), (7, 20)])
def test_ntt_multi_pass_plans(oracle, max_radix, log_n):
    # force two- and three-pass decompositions at sizes the oracle finishes quickly
    dev = sb.B200Device(0)
    try:
        dev.set_ntt_plan(max_radix, 1 << 20)  # tiny scratch: also exercises column grouping
        rng = np.random.default_rng(7 * log_n + max_radix)
        cols = 2 if log_n >= 17 else 5
        n = 1 << log_n
        x = oracle.random_field(rng, n * cols)
        for inverse in (False, True):
            d = dev.h2d(x)
            dev.ntt_batch(d, log_n, cols, inverse)
            dev.synchronize()
            assert np.array_equal(d.cpu().numpy().view(np.uint32), oracle.dft_batch(x, n, cols, inverse))
    finally:
        dev.close()


def test_ntt_roundtrip_large(dev, oracle):
    # reference: tests/ntt_roundtrip.rs — size-independent property at a size the oracle skips
    rng = np.random.default_rng(5)
    log_n, cols = 22, 3
    x = oracle.random_field(rng, cols << log_n)
    d = dev.h2d(x)
    dev.ntt_batch(d, log_n, cols, False)
    dev.synchronize()
    fwd = d.cpu().numpy().view(np.uint32).copy()
    assert not np.array_equal(fwd, x)
    # DFT[0] = sum of inputs (cheap linear checksum of the forward transform)
    for c in range(cols):
        s = int(np.sum(sb.from_mont(x[c << log_n : (c + 1) << log_n]).astype(np.uint64)) % P)
        assert int(sb.from_mont(fwd[c << log_n : (c << log_n) + 1])[0]) == s
    dev.ntt_batch(d, log_n, cols, True)
    dev.synchronize()
    assert np.array_equal(d.cpu().numpy().view(np.uint32), x)


# ---------------------------------------------------------------------------------------------
# RS encode
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize(
    "l_skip,log_h,log_blowup,width",
    [(0, 0, 0, 1), (0, 0, 2, 3), (0, 3, 1, 2), (2, 2, 1, 3), (2, 5, 2, 9), (3, 6, 1, 1), (4, 4, 3, 2), (4, 12, 1, 19),
     (5, 13, 1, 4), (6, 14, 2, 3), (1, 16, 1, 5), (4, 16, 1, 8), (9, 12, 1, 2), (11, 11, 1, 2)],
)
def test_rs_encode(dev, oracle, l_skip, log_h, log_blowup, width):
    rng = np.random.default_rng(1000 * l_skip + 10 * log_h + log_blowup)
    H = 1 << log_h
    ev = oracle.random_field(rng, H * width)
    out = dev.rs_encode(_dm(dev, ev, H, width), l_skip, log_blowup)
    assert out.height() == H << log_blowup and out.width() == width
    dev.synchronize()
    assert np.array_equal(out.to_host(), oracle.rs_code_matrix(l_skip, log_blowup, ev, H, width))


def test_rs_encode_multi_pass_and_grouping(oracle):
    dev = sb.B200Device(0)
    try:
        dev.set_ntt_plan(4, 1 << 16)
        rng = np.random.default_rng(77)
        for l_skip, log_h, lb, width in [(2, 8, 1, 7), (4, 9, 2, 5), (3, 10, 1, 3), (0, 7, 3, 4)]:
            H = 1 << log_h
            ev = oracle.random_field(rng, H * width)
            out = dev.rs_encode(_dm(dev, ev, H, width), l_skip, lb)
            dev.synchronize()
            assert np.array_equal(out.to_host(), oracle.rs_code_matrix(l_skip, lb, ev, H, width))
    finally:
        dev.close()


def test_rs_encode_linearity_large(dev, oracle):
    # size-independent property at 2^20: RS(a) + RS(b) == RS(a + b), and column 0 of an all-zero
    # input encodes to zero
    rng = np.random.default_rng(8)
    log_h, l_skip, lb = 20, 4, 1
    H = 1 << log_h
    a = rng.integers(0, P, H, dtype=np.uint64)
    b = rng.integers(0, P, H, dtype=np.uint64)
    c = (a + b) % P
    ev = np.concatenate([sb.to_mont(a), sb.to_mont(b), sb.to_mont(c), np.zeros(H, np.uint32)])
    out = dev.rs_encode(_dm(dev, ev, H, 4), l_skip, lb)
    dev.synchronize()
    o = sb.from_mont(out.to_host()).astype(np.uint64).reshape(4, H << lb)
    assert np.array_equal((o[0] + o[1]) % P, o[2])
    assert not o[3].any()


# ---------------------------------------------------------------------------------------------
# Merkle
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize(
    "height,width,log_rpq",
    [(1, 1, 0), (2, 3, 1), (4, 3, 0), (8, 9, 1), (16, 1, 4), (32, 19, 2), (24, 8, 2), (64, 5, 3), (512, 8, 4),
     (1024, 256, 4), (4096, 7, 0), (4096, 17, 8), (2048, 3, 9), (2048, 2, 10), (1000, 16, 3), (1 << 14, 33, 4)],
)
def test_merkle_tree(dev, oracle, height, width, log_rpq):
    rng = np.random.default_rng(height + 31 * width + log_rpq)
    m = oracle.random_field(rng, height * width)
    layers_dev = dev.merkle_tree(_dm(dev, m, height, width), log_rpq)
    dev.synchronize()
    want = oracle.merkle_tree(m, height, width, 1 << log_rpq)
    got = layers_dev.cpu().numpy().view(np.uint32)
    assert np.array_equal(got, np.concatenate([l.reshape(-1) for l in want]))


def test_merkle_errors(dev):
    m = _dm(dev, np.zeros(4, np.uint32), 4, 1)
    with pytest.raises(sb.SwirlError) as e:
        dev.merkle_tree(m, 3, out=dev.alloc(8))
    assert "MerkleTreeRowsPerQueryExceeded" in str(e.value)


def test_merkle_queries_and_opened_rows(dev, oracle):
    rng = np.random.default_rng(21)
    traces = [(oracle.random_field(rng, 256 * 5), 256, 5), (oracle.random_field(rng, 64 * 3), 64, 3)]
    params = sb.PcsParams(l_skip=2, n_stack=6, log_blowup=1, k_whir=3)
    root, pcs = dev.commit(params, [_dm(dev, *t) for t in traces])
    oroot, cw, layers, W = oracle.stacked_commit(2, 6, 1, 3, traces)
    assert np.array_equal(root, oroot)
    qs = pcs.tree.query_stride()
    idx = [0, 1, qs - 1, 17, 17]
    proofs = pcs.tree.query_merkle_proofs(idx)
    rows = pcs.tree.get_opened_rows(idx)
    N = 512
    cwm = cw.reshape(W, N)
    for qi, q in enumerate(idx):
        i = q
        for l in range(pcs.tree.proof_depth()):
            assert np.array_equal(proofs[qi, l], layers[l][i ^ 1])  # stacked_pcs.rs:388-405
            i >>= 1
        for t in range(8):
            assert np.array_equal(rows[qi, t], cwm[:, t * qs + q])  # stacked_pcs.rs:516-540
        # the opened rows + proof authenticate against the root
        d = [oracle.hash_slice(np.ascontiguousarray(rows[qi, t])) for t in range(8)]
        while len(d) > 1:
            d = [oracle.compress(d[2 * j], d[2 * j + 1]) for j in range(len(d) // 2)]
        node, i = d[0], q
        for l in range(pcs.tree.proof_depth()):
            node = oracle.compress(node, proofs[qi, l]) if i % 2 == 0 else oracle.compress(proofs[qi, l], node)
            i >>= 1
        assert np.array_equal(node, root)
    pcs.free()


# ---------------------------------------------------------------------------------------------
# stacked_commit
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", GOLD["stacking"], ids=lambda c: c["name"])
def test_stacked_matrix_golden_on_device(dev, case):
    # the reference's own stacking vectors (prover/stacked_pcs.rs:556-619, cuda-backend/src/stacked_pcs.rs:399-517)
    traces = [_dm(dev, sb.to_mont(t["values"]), t["height"], t["width"]) for t in case["traces"]]
    params = sb.PcsParams(case["l_skip"], case["n_stack"], 1, 0)
    _, pcs = dev.commit(params, traces)
    got = sb.from_mont(pcs.matrix())
    assert pcs.height == case["height"] and pcs.width == case["width"]
    if "expected" in case:
        assert list(got) == case["expected"]
    else:
        pre = case["expected_prefix"]
        assert list(got[: len(pre)]) == pre and not got[len(pre) :].any()
    pcs.free()


COMMIT_CASES = [
    # (l_skip, n_stack, log_blowup, k_whir, [(height, width), ...])   heights sorted descending
    (2, 8, 1, 3, [(1 << 10, 2)]),                       # Fib-like, default_test_params_small
    (2, 8, 1, 3, [(8, 2), (4, 2)]),                     # Interactions11 shape: W=1, mostly padding
    (2, 3, 1, 2, [(32, 3), (16, 1), (4, 5), (2, 1), (1, 3)]),  # short traces strided below 2^l_skip
    (0, 4, 2, 1, [(16, 2), (8, 3), (1, 1)]),
    (3, 5, 1, 4, [(256, 4), (128, 1), (8, 8), (2, 2)]),
    (4, 12, 1, 4, [(1 << 16, 2)]),                      # BASELINE config 1: Fib 2^16 x 2, new_for_testing(16)
    (5, 7, 2, 2, [(1 << 12, 3), (1 << 9, 5), (16, 1)]),
    (6, 6, 1, 4, [(1 << 12, 1), (32, 2)]),
    (4, 8, 1, 4, [(1 << 12, 16), (1 << 12, 0), (1 << 10, 7)]),  # zero-width trace in the middle
    (2, 8, 3, 5, [(1 << 10, 9)]),
]


@pytest.mark.parametrize("case", COMMIT_CASES, ids=lambda c: f"l{c[0]}_n{c[1]}_b{c[2]}_k{c[3]}_{len(c[4])}tr")
def test_stacked_commit_matches_oracle(dev, oracle, case):
    l_skip, n_stack, lb, k, shapes = case
    rng = np.random.default_rng(hash(case[:4]) % 1000)
    traces = [(oracle.random_field(rng, h * w), h, w) for h, w in shapes]
    root, pcs = dev.commit(sb.PcsParams(l_skip, n_stack, lb, k), [_dm(dev, *t) for t in traces])
    oroot, cw, layers, W = oracle.stacked_commit(l_skip, n_stack, lb, k, traces)
    assert pcs.width == W
    assert np.array_equal(pcs.matrix(), oracle.stacked_matrix(l_skip, n_stack, traces)[0])
    assert np.array_equal(pcs.tree.backing_matrix(), cw)
    for a, b in zip(pcs.tree.digest_layers(), layers):
        assert np.array_equal(a, b)
    assert np.array_equal(root, oroot) and np.array_equal(pcs.commit(), oroot)
    # host-buffer entry point gives the same commitment
    root2, pcs2 = dev.commit_host(sb.PcsParams(l_skip, n_stack, lb, k), traces)
    assert np.array_equal(root2, oroot)
    pcs.free()
    pcs2.free()


def test_commit_errors(dev, oracle):
    rng = np.random.default_rng(3)
    t = lambda h, w: _dm(dev, oracle.random_field(rng, h * w), h, w)
    with pytest.raises(sb.SwirlError) as e:  # unsorted
        dev.commit(sb.PcsParams(2, 4, 1, 1), [t(8, 1), t(16, 1)])
    assert "sorted" in str(e.value)
    with pytest.raises(sb.SwirlError) as e:  # taller than the stacked height
        dev.commit(sb.PcsParams(2, 2, 1, 1), [t(32, 1)])
    assert e.value.code == 10002
    with pytest.raises(sb.SwirlError):  # rows_per_query > codeword height
        dev.commit(sb.PcsParams(0, 1, 0, 4), [t(2, 1)])


def test_commit_all_zero_trace_baseline_shape(dev, oracle):
    # uniform_runner commits all-zero traces (benchmarks/synthetic/src/bin/uniform_runner.rs:260-267)
    h, w = 1 << 12, 20
    z = np.zeros(h * w, np.uint32)
    root, pcs = dev.commit(sb.PcsParams(4, 8, 1, 4), [_dm(dev, z, h, w)])
    oroot, _, _, _ = oracle.stacked_commit(4, 8, 1, 4, [(z, h, w)], want_codeword=False)
    assert np.array_equal(root, oroot)
    pcs.free()


def test_commit_full_size_properties(dev, oracle):
    # BASELINE config 2 shape scaled to what a test may spend: 2^20 x 16 (n_stack 16, blowup 2,
    # k_whir 4).  The oracle is too slow here, so check structure: every sampled query opens
    # rows that hash + compress to the committed root, and the codeword's even-indexed rows
    # are the DFT of the message of size H (RS code is an extension of the rate-1 code).
    rng = np.random.default_rng(42)
    h, w = 1 << 20, 16
    vals = sb.to_mont(rng.integers(0, P, h * w, dtype=np.uint64))
    params = sb.PcsParams(4, 16, 1, 4)
    root, pcs = dev.commit(params, [_dm(dev, vals, h, w)])
    qs = pcs.tree.query_stride()
    idx = [0, 1, qs - 1] + [int(x) for x in rng.integers(0, qs, 5)]
    proofs = pcs.tree.query_merkle_proofs(idx)
    rows = pcs.tree.get_opened_rows(idx)
    for qi, q in enumerate(idx):
        d = [oracle.hash_slice(np.ascontiguousarray(rows[qi, t])) for t in range(16)]
        while len(d) > 1:
            d = [oracle.compress(d[2 * j], d[2 * j + 1]) for j in range(len(d) // 2)]
        node, i = d[0], q
        for l in range(pcs.tree.proof_depth()):
            node = oracle.compress(node, proofs[qi, l]) if i % 2 == 0 else oracle.compress(proofs[qi, l], node)
            i >>= 1
        assert np.array_equal(node, root)
    # column 0 of the codeword against the oracle (one column of 2^21 is affordable)
    cw0 = dev._d2h(pcs.tree.codeword_ptr, 2 * h)
    assert np.array_equal(cw0, oracle.rs_code_matrix(4, 1, vals[:h], h, 1))
    pcs.free()


# ---------------------------------------------------------------------------------------------
# grind
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("bits,absorbed", [(0, 3), (1, 0), (5, 11), (10, 7), (16, 8), (20, 5)])
def test_sponge_grind_matches_oracle(dev, oracle, bits, absorbed):
    rng = np.random.default_rng(bits * 17 + absorbed)
    st = oracle.sponge_new()
    if absorbed:
        oracle.sponge_observe(st, oracle.random_field(rng, absorbed))
    w = dev.sponge_grind(st, bits)
    if bits <= 16:
        ow = oracle.sponge_grind(st.copy(), bits)
        assert w == int(oracle.from_mont([ow])[0])
    else:
        probe = st.copy()
        assert oracle.sponge_check_witness(probe, bits, int(oracle.to_mont([w])[0]))
    # a window that excludes the answer finds the next one, or nothing
    if bits == 5:
        w2 = dev.sponge_grind(st, bits, min_w=w + 1)
        assert w2 > w and oracle.sponge_check_witness(st.copy(), bits, int(oracle.to_mont([w2])[0]))
        assert dev.sponge_grind(st, bits, min_w=w + 1, max_w=w2) is None


def test_commit_host_pipelined_matches_device_commit(dev, oracle):
    # wide enough (>= 32 MiB, > 32 columns) to take the pipelined transport path of swirl_commit_host
    rng = np.random.default_rng(77)
    h, w = 1 << 17, 72
    vals = oracle.random_field(rng, h * w)
    params = sb.PcsParams(4, 13, 1, 4)
    root_d, pcs_d = dev.commit(params, [_dm(dev, vals, h, w)])
    host = torch.from_numpy(vals.view(np.int32)).pin_memory()
    root_h, pcs_h = dev.commit_host(params, [(host, h, w)])
    assert np.array_equal(root_d, root_h)
    assert np.array_equal(pcs_d.tree.backing_matrix(), pcs_h.tree.backing_matrix())
    assert np.array_equal(pcs_h.matrix(), vals)
    pcs_d.free()
    pcs_h.free()


@pytest.mark.gpu
@pytest.mark.parametrize("log_n", [1, 4, 11, 16])
def test_fold_mle_matches_oracle(dev, oracle, log_n):
    """swirl_fold_mle = fold_mle_evals (prover/sumcheck.rs:395-414): one fold checked entry by entry with the oracle's
    field arithmetic, log_n folds checked against the oracle's MLE evaluation at the same point."""
    rng = np.random.default_rng(300 + log_n)
    evals = oracle.random_field(rng, (1 << log_n, 4))
    point = oracle.random_field(rng, (log_n, 4))
    t = dev.h2d(evals)
    first = dev._d2h(dev.fold_mle(t, point[0]).data_ptr(), (1 << (log_n - 1)) * 4).reshape(-1, 4)
    for j in list(range(min(4, len(first)))) + [len(first) - 1]:
        d = (evals[2 * j + 1].astype(np.int64) - evals[2 * j].astype(np.int64)) % sb.P
        want = (evals[2 * j].astype(np.int64) + oracle.ef_mul(d.astype(np.uint32), point[0]).astype(np.int64)) % sb.P
        assert np.array_equal(first[j], want.astype(np.uint32))
    for i in range(log_n):
        t = dev.fold_mle(t, point[i])
    got = dev._d2h(t.data_ptr(), 4)
    assert np.array_equal(got, oracle.eval_mle_evals_at_point(evals, log_n, point))


@pytest.mark.gpu
def test_sharded_commit_device_backend_single_rank(dev, oracle):
    """multi.sharded_commit through the C-ABI primitives (world size 1: no collective; the exchange / layout logic for
    world size 2 is covered with gloo in tests/test_multi_rank.py, the NCCL run by tools/sharded_commit.py)."""
    from stark_backend_b200 import multi

    l_skip, n_stack, log_blowup, k, width = 3, 7, 1, 3, 9
    H = 1 << (l_skip + n_stack)
    full = oracle.random_field(np.random.default_rng(77), H * width)
    res = multi.sharded_commit(multi.DeviceCommitBackend(dev), dev.h2d(full), H, width, l_skip, log_blowup, k, 1, 0)
    root = oracle.stacked_commit(l_skip, n_stack, log_blowup, k, [(full, H, width)], want_codeword=False)[0]
    assert np.array_equal(res["root"], root)
    # the same with the row exchange done by swirl_scatter_rows_to_peers (the only peer is this rank)
    px = multi.PeerExchange(dev, width, H << log_blowup, 1, 0)
    res2 = multi.sharded_commit(multi.DeviceCommitBackend(dev), dev.h2d(full), H, width, l_skip, log_blowup, k, 1, 0, peer_exchange=px)
    assert np.array_equal(res2["root"], root)
    assert np.array_equal(res2["shard"].cpu().numpy(), res["shard"].cpu().numpy())


@pytest.mark.gpu
def test_scatter_rows_to_peers_layout(dev, oracle):
    """swirl_scatter_rows_to_peers with 4 destination buffers on one device: every buffer must hold, for all columns, the rows
    q + t S of its quarter of the queries as [column][t][q'] (the layout multi.pack_codeword_slice + all-to-all produce)."""
    import ctypes as C

    import torch

    from stark_backend_b200 import multi
    from stark_backend_b200.lib import check

    rows, cols, col0, width, k, world = 1 << 9, 3, 2, 7, 2, 4
    src = oracle.random_field(np.random.default_rng(5), rows * cols)
    bufs = [torch.zeros(width * (rows // world), dtype=torch.int32, device=dev.torch_device) for _ in range(world)]
    ptrs = (C.c_void_p * world)(*[b.data_ptr() for b in bufs])
    d_src = dev.h2d(src)
    torch.cuda.synchronize()
    check(dev.lib.swirl_scatter_rows_to_peers(dev.ctx, d_src.data_ptr(), rows, cols, col0, k, world, ptrs))
    dev.synchronize()
    want = multi.pack_codeword_slice(torch.from_numpy(src.view(np.int32)).view(cols, rows), k, world)  # (world, cols, 2^k, Sg)
    for r in range(world):
        got = bufs[r].cpu().view(width, rows // world)
        assert torch.equal(got[col0:col0 + cols], want[r].reshape(cols, -1))
        assert int(got[:col0].abs().sum()) == 0 and int(got[col0 + cols:].abs().sum()) == 0


@pytest.mark.gpu
def test_trace_transporter_double_buffering(dev, oracle):
    """TraceTransporter: traces submitted from pinned host memory (two in flight, buffers reused round-robin) commit to the
    same roots as the oracle."""
    import torch

    l_skip, n_stack, log_blowup, k, w = 2, 6, 1, 2, 5
    H = 1 << (l_skip + n_stack)
    params = sb.PcsParams(l_skip, n_stack, log_blowup, k)
    tp = sb.TraceTransporter(dev, H, w)
    rng = np.random.default_rng(21)
    hosts = [oracle.random_field(rng, H * w) for _ in range(5)]
    pinned = [torch.from_numpy(h.view(np.int32)).pin_memory() for h in hosts]
    tickets = [tp.submit(pinned[0])]
    for i in range(5):
        if i + 1 < 5:
            tickets.append(tp.submit(pinned[i + 1]))  # next trace in flight while this one is committed
        root, pcs = dev.commit(params, [tp.matrix(tickets[i])])
        assert np.array_equal(root, oracle.stacked_commit(l_skip, n_stack, log_blowup, k, [(hosts[i], H, w)], want_codeword=False)[0])
        pcs.free()
        tp.retire(tickets[i])  # its readers are enqueued: the buffer may be overwritten by the submit after next
    # the depth is enforced: a third outstanding ticket is refused instead of overwriting a trace in use
    a, b = tp.submit(pinned[0]), tp.submit(pinned[1])
    with pytest.raises(RuntimeError):
        tp.submit(pinned[2])
    tp.retire(a)
    tp.retire(b)
