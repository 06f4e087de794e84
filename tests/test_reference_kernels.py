"""GPU three-way parity: the REFERENCE's own CUDA kernels (oracle/_ref/libref_kernels.so, compiled from
/root/reference/crates/cuda-backend/cuda by oracle/Makefile.ref) == the CPU oracle == the product library, bit for bit.

This is what pins the oracle (and the product) to reference *outputs* for the third-party arithmetic the Rust code
takes from Plonky3: Poseidon2 digests / Merkle layers (merkle_tree.cu:214-283), the RS codeword and the NTT
(stacked_pcs.rs:229-337, ntt.rs:111-168, supra/ntt.cu:206, ntt_bitrev.cu:206, batch_ntt_small.cu:134), the duplex-sponge
proof of work (sponge.cu:65-117) and extension-field arithmetic (gkr.cu:82-170, whir.cu:141-166).  Full-size cases
(BASELINE configs[1]: 2^20 x 256) compare the product with the reference kernels directly.
"""
import numpy as np
import pytest
import torch

import ref_kernels
import refcases as rc
import stark_backend_b200 as sb

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_kernels.available(), reason="oracle/_ref/libref_kernels.so not built")]


@pytest.fixture(scope="module")
def rk():
    return ref_kernels.RefKernels()


@pytest.fixture(scope="module")
def ref_out(rk, oracle):
    return rc.reference_outputs(rk, oracle.to_mont)


def ef_add(a, b):
    return ((a.astype(np.uint64) + b.astype(np.uint64)) % rc.P).astype(np.uint32)


def ef_sub(a, b):
    return ((a.astype(np.uint64) + rc.P - b.astype(np.uint64)) % rc.P).astype(np.uint32)


@pytest.mark.parametrize("i", range(len(rc.MERKLE_CASES)))
def test_merkle_layers_three_way(dev, oracle, ref_out, i):
    w, h, rpq = rc.MERKLE_CASES[i]
    m = rc.merkle_inputs(i, oracle.to_mont)
    want = ref_out["merkle"][i]
    assert np.array_equal(np.concatenate([l.reshape(-1) for l in oracle.merkle_tree(m, h, w, rpq)]), want), "oracle != reference"
    got = dev.merkle_tree(sb.DeviceMatrix(dev.h2d(m), h, w), rpq.bit_length() - 1)
    dev.synchronize()
    assert np.array_equal(got.cpu().numpy().view(np.uint32), want), "product != reference"


@pytest.mark.parametrize("i", range(len(rc.RS_CASES)))
def test_rs_code_matrix_three_way(dev, oracle, ref_out, i):
    l_skip, log_h, lb, w = rc.RS_CASES[i]
    ev = rc.rs_inputs(i, oracle.to_mont)
    want = ref_out["rs"][i]
    assert np.array_equal(oracle.rs_code_matrix(l_skip, lb, ev, 1 << log_h, w), want), "oracle != reference"
    out = dev.rs_encode(sb.DeviceMatrix(dev.h2d(ev), 1 << log_h, w), l_skip, lb)
    dev.synchronize()
    assert np.array_equal(out.to_host(), want), "product != reference"


@pytest.mark.parametrize("i", range(len(rc.NTT_CASES)))
def test_ntt_three_way(dev, oracle, ref_out, i):
    log_n, cols, inv = rc.NTT_CASES[i]
    x = rc.ntt_inputs(i, oracle.to_mont)
    want = ref_out["ntt"][i]
    assert np.array_equal(oracle.dft_batch(x, 1 << log_n, cols, inv), want), "oracle != reference"
    d = dev.h2d(x)
    dev.ntt_batch(d, log_n, cols, inv)
    dev.synchronize()
    assert np.array_equal(d.cpu().numpy().view(np.uint32), want), "product != reference"


@pytest.mark.parametrize("i", range(len(rc.GRIND_CASES)))
def test_grind_three_way(dev, oracle, ref_out, i):
    bits = rc.GRIND_CASES[i][3]
    st = rc.grind_state(i, oracle.to_mont)
    want = ref_out["grind"][i]
    assert want is not None
    assert int(oracle.from_mont([oracle.sponge_grind(st.copy(), bits)])[0]) == want, "oracle != reference"
    assert dev.sponge_grind(st, bits) == want, "product != reference"


def test_reference_accepts_product_witnesses_on_transcript_states(dev, oracle, rk):
    # sponge states reached by a real transcript (observe/sample sequences), 16-20 bits as the 100-bit parameter sets use
    rng = np.random.default_rng(3)
    for bits in (16, 18, 20):
        st = oracle.sponge_new()
        oracle.sponge_observe(st, oracle.random_field(rng, 11))
        oracle.sponge_sample(st, 3)
        oracle.sponge_observe(st, oracle.random_field(rng, 2))
        w = dev.sponge_grind(st, bits)
        assert rk.sponge_grind(st, bits, w, w) == w  # the reference kernel accepts it ...
        assert rc.reference_min_witness(rk, st, bits) == w  # ... and finds nothing smaller


@pytest.mark.parametrize("i", range(len(rc.EF_CASES)))
def test_extension_field_kernels_vs_oracle(oracle, ref_out, i):
    n = 1 << rc.EF_CASES[i]
    fr, f4, alpha = rc.ef_inputs(i, oracle.to_mont)
    fr = fr.reshape(n, 2, 4)
    half = n // 2
    want = np.zeros((half, 2, 4), np.uint32)
    for j in range(half):  # frac_add: (p1 q2 + p2 q1, q1 q2)
        (p1, q1), (p2, q2) = fr[j], fr[j + half]
        want[j, 0] = ef_add(oracle.ef_mul(p1, q2), oracle.ef_mul(p2, q1))
        want[j, 1] = oracle.ef_mul(q1, q2)
    assert np.array_equal(want.reshape(-1), ref_out["frac_layer"][i])
    f = f4.reshape(n, 4)
    wm = fr.reshape(-1)[: n * 4].reshape(n, 4)
    one = oracle.to_mont([1, 0, 0, 0])
    wantf = np.zeros((half, 4), np.uint32)
    wantw = np.zeros((half, 4), np.uint32)
    for y in range(half):  # whir.cu:141-166
        wantf[y] = ef_add(f[2 * y], oracle.ef_mul(alpha, f[2 * y + 1]))
        wantw[y] = ef_add(oracle.ef_mul(ef_sub(one, alpha), wm[2 * y]),
                          oracle.ef_mul(ef_sub(ef_add(alpha, alpha), one), wm[2 * y + 1]))
    assert np.array_equal(np.concatenate([wantf.reshape(-1), wantw.reshape(-1)]), ref_out["whir_fold"][i])


def test_commit_config2_full_size_product_vs_reference(dev, rk, oracle):
    """BASELINE configs[1] at full size (2^20 x 256, blowup 2, k_whir 4): codeword and every digest layer of the product
    commit equal the reference kernels' (the oracle would need minutes here; it is pinned to both at the small sizes)."""
    H, W, l_skip, lb, k = 1 << 20, 256, 4, 1, 4
    g = torch.Generator(device="cuda").manual_seed(42)
    trace = torch.randint(0, rc.P, (H * W,), dtype=torch.int32, device="cuda", generator=g)
    cw_ref = rk.rs_code_matrix(trace, H, W, l_skip, lb)
    layers_ref = rk.merkle_tree(cw_ref, H << lb, W, 1 << k)
    torch.cuda.synchronize()
    root, pcs = dev.commit(sb.PcsParams(l_skip=l_skip, n_stack=16, log_blowup=lb, k_whir=k), [sb.DeviceMatrix(trace, H, W)])
    dev.synchronize()
    assert np.array_equal(pcs.tree.backing_matrix(), rk.d2h(cw_ref)), "codeword != reference rs_code_matrix"
    got_layers = np.concatenate([l.reshape(-1) for l in pcs.tree.digest_layers()])
    want_layers = np.concatenate([rk.d2h(l) for l in layers_ref])
    assert np.array_equal(got_layers, want_layers), "digest layers != reference merkle tree"
    assert np.array_equal(root, rk.d2h(layers_ref[-1]))
    pcs.free()
