"""Proof wire format (reference: `impl Encode/Decode for Proof`, crates/stark-backend/src/proof.rs:204-705;
primitives crates/stark-backend/src/codec.rs:191-310).

The product encoder (stark-backend_b200/codec.py) walks the *flat* C-ABI sections.  The checker below is
written the other way round: it first rebuilds the reference's nested `Proof` struct (Vec<Vec<..>> fields of
proof.rs:20-200) from the flat words and then serialises that struct field by field the way the Rust impls
iterate it, so a wrong flat-layout <-> struct-field correspondence shows up as a byte difference."""
import struct

import numpy as np
import pytest

import airs as A
import stark_backend_b200 as sb
import test_prove as tp
from stark_backend_b200 import codec

P = sb.P


# ---- checker: nested reference struct + field-by-field serialiser -------------------------------------------
def canon(words):
    return [int(v) for v in sb.from_mont(np.asarray(words, dtype=np.uint32).reshape(-1))]


def nested_proof(shape, root, bc, st, wh):
    """Nested lists shaped like Proof / GkrProof / BatchConstraintProof / StackingProof / WhirProof."""
    take = lambda it, n: [next(it) for _ in range(n)]
    ef = lambda it: take(it, 4)
    L, n, D = shape.gkr_layers, len(shape.airs), shape.max_constraint_degree
    it = iter(canon(bc))
    gkr = dict(logup_pow_witness=next(it), q0_claim=ef(it),
               # flat order = transcript order (fractional_sumcheck_gkr.rs:185-193)
               claims_per_layer=[dict(zip(("p_xi_0", "q_xi_0", "p_xi_1", "q_xi_1"), [ef(it) for _ in range(4)])) for _ in range(L)],
               sumcheck_polys=[[[ef(it) for _ in range(3)] for _ in range(j)] for j in range(1, L)])
    bcp = dict(numerator_term_per_air=[ef(it) for _ in range(n)], denominator_term_per_air=[ef(it) for _ in range(n)],
               univariate_round_coeffs=[ef(it) for _ in range((D + 1) * ((1 << shape.l_skip) - 1) + 1)],
               sumcheck_round_polys=[[ef(it) for _ in range(D + 1)] for _ in range(shape.n_max)],
               column_openings=[[[ef(it) for _ in range(w * (2 if a.need_rot else 1))] for w in a.part_widths] for a in shape.airs])
    assert next(it, None) is None
    it = iter(canon(st))
    stp = dict(univariate_round_coeffs=[ef(it) for _ in range(2 * ((1 << shape.l_skip) - 1) + 1)],
               sumcheck_round_polys=[[ef(it), ef(it)] for _ in range(shape.n_stack)],
               stacking_openings=[[ef(it) for _ in range(w)] for w in shape.commit_widths])
    assert next(it, None) is None
    it = iter(canon(wh))
    k, R, m = shape.k_whir, len(shape.num_queries), shape.l_skip + shape.n_stack
    dig = lambda it: take(it, 8)
    whp = dict(mu_pow_witness=next(it), whir_sumcheck_polys=[[ef(it), ef(it)] for _ in range(R * k)],
               codeword_commits=[dig(it) for _ in range(R - 1)], ood_values=[ef(it) for _ in range(R - 1)],
               folding_pow_witnesses=take(it, R * k), query_phase_pow_witnesses=take(it, R))
    q0 = shape.num_queries[0]
    whp["initial_round_opened_rows"] = [[[take(it, w) for _ in range(1 << k)] for _ in range(q0)] for w in shape.commit_widths]
    whp["initial_round_merkle_proofs"] = [[[dig(it) for _ in range(m + shape.log_blowup - k)] for _ in range(q0)]
                                          for _ in shape.commit_widths]
    whp["codeword_opened_values"] = [[[ef(it) for _ in range(1 << k)] for _ in range(shape.num_queries[r])] for r in range(1, R)]
    whp["codeword_merkle_proofs"] = [[[dig(it) for _ in range(m + shape.log_blowup - r - k)] for _ in range(shape.num_queries[r])]
                                     for r in range(1, R)]
    whp["final_poly"] = [ef(it) for _ in range(1 << (m - R * k))]
    assert next(it, None) is None
    return dict(common_main_commit=canon(root), trace_vdata=shape.trace_vdata, public_values=[canon(p) for p in shape.public_values],
                gkr_proof=gkr, batch_constraint_proof=bcp, stacking_proof=stp, whir_proof=whp)


def serialise_nested(p):
    out = bytearray()
    u32 = lambda x: out.extend(struct.pack("<I", x))
    f = u32                                    # encode_prime_field32: canonical LE u32
    ef = lambda e: [f(c) for c in e]           # 4 basis coefficients
    dig = lambda d: [f(c) for c in d]
    u32(3)                                     # CODEC_VERSION
    dig(p["common_main_commit"])
    tv = p["trace_vdata"]
    u32(len(tv))
    for i in range(0, len(tv), 8):
        out.append(sum((1 << j) for j, v in enumerate(tv[i:i + 8]) if v is not None))
    for v in tv:
        if v is not None:
            u32(v[0]); u32(len(v[1]))
            for d in v[1]:
                dig(canon(d))
    u32(len(p["public_values"]))
    for pv in p["public_values"]:
        u32(len(pv)); [f(x) for x in pv]
    g = p["gkr_proof"]
    f(g["logup_pow_witness"]); ef(g["q0_claim"])
    u32(len(g["claims_per_layer"]))
    for c in g["claims_per_layer"]:  # GkrLayerClaims::encode, proof.rs:211-218
        ef(c["p_xi_0"]); ef(c["p_xi_1"]); ef(c["q_xi_0"]); ef(c["q_xi_1"])
    for rnd in g["sumcheck_polys"]:
        for arr in rnd:
            [ef(e) for e in arr]
    b = p["batch_constraint_proof"]
    u32(len(b["numerator_term_per_air"])); [ef(e) for e in b["numerator_term_per_air"]]
    [ef(e) for e in b["denominator_term_per_air"]]
    u32(len(b["univariate_round_coeffs"])); [ef(e) for e in b["univariate_round_coeffs"]]
    u32(len(b["sumcheck_round_polys"]))
    if b["sumcheck_round_polys"]:
        u32(len(b["sumcheck_round_polys"][0]))
        for rp in b["sumcheck_round_polys"]:
            [ef(e) for e in rp]
    for parts in b["column_openings"]:
        u32(len(parts))
        for col in parts:
            u32(len(col)); [ef(e) for e in col]
    s = p["stacking_proof"]
    u32(len(s["univariate_round_coeffs"])); [ef(e) for e in s["univariate_round_coeffs"]]
    u32(len(s["sumcheck_round_polys"]))
    for arr in s["sumcheck_round_polys"]:
        [ef(e) for e in arr]
    u32(len(s["stacking_openings"]))
    for o in s["stacking_openings"]:
        u32(len(o)); [ef(e) for e in o]
    w = p["whir_proof"]
    f(w["mu_pow_witness"])
    u32(len(w["whir_sumcheck_polys"]))
    for arr in w["whir_sumcheck_polys"]:
        [ef(e) for e in arr]
    u32(len(w["codeword_commits"])); [dig(d) for d in w["codeword_commits"]]
    [ef(e) for e in w["ood_values"]]
    [f(x) for x in w["folding_pow_witnesses"]]
    [f(x) for x in w["query_phase_pow_witnesses"]]
    rows = w["initial_round_opened_rows"]
    u32(len(rows)); u32(len(rows[0]))
    if len(rows[0]) > 0:
        u32(len(w["initial_round_merkle_proofs"][0][0]))
        for commit_rows in rows:
            u32(len(commit_rows[0][0]))
        for commit_rows in rows:
            for q in commit_rows:
                for row in q:
                    [f(x) for x in row]
        for mp in w["initial_round_merkle_proofs"]:
            for proof in mp:
                [dig(d) for d in proof]
    for rnd in w["codeword_opened_values"]:
        u32(len(rnd))
        for q in rnd:
            [ef(e) for e in q]
    first = len(w["codeword_merkle_proofs"][0][0]) if (len(w["codeword_commits"]) > 0 and len(rows[0]) > 0) else 0
    u32(first)
    for rp in w["codeword_merkle_proofs"]:
        for proof in rp:
            [dig(d) for d in proof]
    u32(len(w["final_poly"])); [ef(e) for e in w["final_poly"]]
    return bytes(out)


# ---- fixtures -------------------------------------------------------------------------------------------------
def oracle_proof_and_shape(oracle):
    airs, order = tp.fixture_airs(2)
    is_required = [True, True, True, True, False]
    vk = oracle.to_mont(np.arange(100, 108))
    pr = tp.oracle_prove(oracle, airs, order, is_required, vk)
    commit1 = lambda m: oracle.stacked_commit(tp.L_SKIP, tp.N_STACK, tp.LOG_BLOWUP, tp.WHIR["k"], [m], want_codeword=False)[0]
    sa = pr["sorted_airs"]
    lifted = lambda a: max(a.height, 1 << tp.L_SKIP)
    total = sum(len(a.interactions) * lifted(a) for a in sa)
    shape = codec.ProofShape(
        l_skip=tp.L_SKIP, n_stack=tp.N_STACK, log_blowup=tp.LOG_BLOWUP, max_constraint_degree=tp.D, k_whir=tp.WHIR["k"],
        num_queries=tp.WHIR["num_queries"],
        airs=[codec.AirShape([a.common_main[2]] + ([a.preprocessed[2]] if a.preprocessed is not None else []) + [c[2] for c in a.cached],
                             a.need_rot) for a in sa],
        gkr_layers=total.bit_length() if total else 0, n_max=pr["n_max"], commit_widths=pr["widths"],
        trace_vdata=[(a.height.bit_length() - 1, [commit1(c) for c in a.cached]) for a in airs],
        public_values=[a.public_values for a in airs])
    return pr, shape


def test_encode_matches_field_by_field_serialisation_and_round_trips(oracle):
    pr, shape = oracle_proof_and_shape(oracle)
    data = codec.encode_proof(shape, pr["root"], pr["bc"], pr["stacking"], pr["whir"])
    want = serialise_nested(nested_proof(shape, pr["root"], pr["bc"], pr["stacking"], pr["whir"]))
    assert data == want
    assert data[:4] == b"\x03\x00\x00\x00"
    dec = codec.decode_proof(data)
    flat = codec.flat_to_montgomery(dec)
    assert np.array_equal(flat["constraints_proof"], pr["bc"])
    assert np.array_equal(flat["stacking_proof"], pr["stacking"])
    assert np.array_equal(flat["whir_proof"], pr["whir"])
    assert np.array_equal(sb.to_mont(dec["common_main_commit"]), pr["root"])
    assert dec["gkr_layers"] == shape.gkr_layers and dec["n_max"] == shape.n_max and dec["n_stack"] == shape.n_stack
    assert dec["k_whir"] == shape.k_whir and dec["num_queries"] == list(shape.num_queries)
    assert dec["commit_widths"] == list(shape.commit_widths) == dec["stacking_widths"]
    assert [v[0] if v else None for v in dec["trace_vdata"]] == [v[0] for v in shape.trace_vdata]
    assert all(np.array_equal(sb.to_mont(a), b) for a, b in zip(dec["public_values"], shape.public_values))
    assert [[len(p) for p in parts] for parts in dec["column_openings"]] == \
        [[w * (2 if a.need_rot else 1) for w in a.part_widths] for a in shape.airs]


def test_decode_rejects_malformed(oracle):
    pr, shape = oracle_proof_and_shape(oracle)
    data = bytearray(codec.encode_proof(shape, pr["root"], pr["bc"], pr["stacking"], pr["whir"]))
    with pytest.raises(ValueError):
        codec.decode_proof(bytes(data[:-1]))                       # truncated
    with pytest.raises(ValueError):
        codec.decode_proof(bytes(data) + b"\x00")                  # trailing
    bad = bytearray(data); bad[0] = 2
    with pytest.raises(ValueError):
        codec.decode_proof(bytes(bad))                             # CODEC_VERSION (proof.rs:447-453)
    bad = bytearray(data); bad[4:8] = struct.pack("<I", P)
    with pytest.raises(ValueError):
        codec.decode_proof(bytes(bad))                             # non-canonical field element (codec.rs:218-230)
    # 5 AIRs -> one bitmap byte at offset 4 + 32 + 4; bit 5 is padding (proof.rs:464-470)
    bad = bytearray(data); bad[40] |= 1 << 5
    with pytest.raises(ValueError):
        codec.decode_proof(bytes(bad))
    with pytest.raises(ValueError):                                # flat section inconsistent with its shape
        codec.encode_proof(shape, pr["root"], pr["bc"][:-4], pr["stacking"], pr["whir"])


def test_absent_air_and_hand_checked_prefix():
    """Header bytes written out by hand from proof.rs:226-251: version, digest, num_airs, bitmap, TraceVData of the
    present AIRs only, public values."""
    shape = codec.ProofShape(l_skip=1, n_stack=1, log_blowup=1, max_constraint_degree=1, k_whir=1, num_queries=[0],
                             airs=[codec.AirShape([1], False)], gkr_layers=0, n_max=0, commit_widths=[1],
                             trace_vdata=[None, (1, [sb.to_mont(np.arange(8) + 20)])] + [None] * 7,
                             public_values=[np.zeros(0, np.uint32), sb.to_mont([7, 9])] + [np.zeros(0, np.uint32)] * 7)
    root = sb.to_mont(np.arange(8) + 1)
    # flat sections with canonical values 1..: bc = pow, q0, num, den, uni[(1+1)*1+1 = 3], openings[1]
    bc = sb.to_mont(np.arange(1 + 4 + 4 + 4 + 12 + 4) + 1)
    st = sb.to_mont(np.arange(12 + 8 + 4) + 100)
    # whir: mu | polys[1][2][4] | (no commits/ood) | fold pow[1] | query pow[1] | (no rows) | final_poly[2][4]
    wh = sb.to_mont(np.arange(1 + 8 + 1 + 1 + 8) + 200)
    data = codec.encode_proof(shape, root, bc, st, wh)
    le = lambda *xs: b"".join(struct.pack("<I", x) for x in xs)
    head = le(3) + le(*range(1, 9)) + le(9) + bytes([0b10, 0]) + le(1, 1) + le(*range(20, 28)) \
        + le(9) + le(0) + le(2, 7, 9) + le(0) * 7
    assert data[:len(head)] == head
    gkr_bc = le(1) + le(2, 3, 4, 5) + le(0) + le(1) + le(6, 7, 8, 9) + le(10, 11, 12, 13) + le(3) + le(*range(14, 26)) \
        + le(0) + le(1) + le(1) + le(26, 27, 28, 29)
    assert data[len(head):len(head) + len(gkr_bc)] == gkr_bc
    stacking = le(3) + le(*range(100, 112)) + le(1) + le(*range(112, 120)) + le(1) + le(1) + le(*range(120, 124))
    off = len(head) + len(gkr_bc)
    assert data[off:off + len(stacking)] == stacking
    whir = le(200) + le(1) + le(*range(201, 209)) + le(0) + le(209) + le(210) + le(1) + le(0) + le(0) + le(2) + le(*range(211, 219))
    assert data[off + len(stacking):] == whir
    dec = codec.decode_proof(data)
    assert [v is not None for v in dec["trace_vdata"]] == [False, True] + [False] * 7


@pytest.mark.gpu
def test_gpu_proof_bytes_equal_oracle_proof_bytes(dev, oracle):
    pr, shape = oracle_proof_and_shape(oracle)
    want = codec.encode_proof(shape, pr["root"], pr["bc"], pr["stacking"], pr["whir"])
    airs, order = tp.fixture_airs(2)
    is_required = [True, True, True, True, False]
    vk = oracle.to_mont(np.arange(100, 108))
    params = sb.SystemParams(tp.L_SKIP, tp.N_STACK, tp.LOG_BLOWUP, sb.WhirConfig(**tp.WHIR), tp.LOGUP_POW, tp.D)
    dm = lambda m: sb.DeviceMatrix(dev.h2d(m[0]), m[1], m[2])

    def committed(m):
        mat = dm(m)
        root, data = dev.commit(params.pcs(), [mat])
        return sb.CommittedTraceData(root, mat, data)

    per_air_pk, per_trace = [], []
    for air_id, a in enumerate(airs):
        prep = committed(a.preprocessed) if a.preprocessed is not None else None
        cached = [committed(c) for c in a.cached]
        per_air_pk.append(sb.AirProvingKey(is_required[air_id], prep))
        ctx = sb.AirProvingContext(a.nodes, a.constraint_idx, a.interactions, a.constraint_degree, a.need_rot, dm(a.common_main),
                                   a.public_values, [c.trace for c in cached], prep.trace if prep else None)
        per_trace.append((air_id, ctx, cached))
    proof = sb.Coordinator(dev, params).prove(vk, per_air_pk, per_trace)
    got = proof.encode()
    assert got == want
    assert serialise_nested(nested_proof(proof.shape, proof.common_main_commit, proof.constraints_proof, proof.stacking_proof,
                                         proof.whir_proof)) == got
