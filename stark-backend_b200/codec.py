"""Byte serialisation of a whole proof in the reference's wire format.

reference: `impl Encode for Proof` and its parts (crates/stark-backend/src/proof.rs:204-420), primitives in
crates/stark-backend/src/codec.rs:191-310 — CODEC_VERSION = 3 (proof.rs:224); `usize` / `u32` as 4 LE bytes
(codec.rs:191-202); a base-field element as the 4 LE bytes of its *canonical* value (codec.rs:213-215); an
extension element as its 4 basis coefficients (codec.rs:234-243); a digest as 8 base elements; `*_slice`
writes a u32 length prefix, `*_iter` does not (codec.rs:60-118).

The C ABI returns the proof parts as flat Montgomery words in the field order of the reference structs
(include/swirl_b200.h); this module only converts them to canonical form and inserts the length prefixes the
reference writes, so a proof produced here is byte-comparable with `Proof::encode_to_vec()`.
`decode_proof` follows `impl Decode for Proof` (proof.rs:444-705) and needs no side information.
"""
import struct
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

P = 0x78000001
_RINV = pow((1 << 32) % P, P - 2, P)
CODEC_VERSION = 3


def _canon(words):
    a = np.ascontiguousarray(words, dtype=np.uint32).astype(np.uint64).reshape(-1)
    return ((a * np.uint64(_RINV)) % np.uint64(P)).astype("<u4")


def _mont(canon):
    a = np.asarray(canon, dtype=np.uint64)
    return ((a << np.uint64(32)) % np.uint64(P)).astype(np.uint32)


@dataclass
class AirShape:
    """One present AIR in sorted order: widths of its trace parts in the order of
    BatchConstraintProof::column_openings (common main, preprocessed if any, cached..; proof.rs:122-133)."""

    part_widths: Sequence[int]
    need_rot: bool


@dataclass
class ProofShape:
    l_skip: int
    n_stack: int
    log_blowup: int
    max_constraint_degree: int
    k_whir: int
    num_queries: Sequence[int]           # per WHIR round
    airs: List[AirShape]                 # present AIRs, sorted (height desc, air id asc)
    gkr_layers: int                      # L = l_skip + n_logup, 0 without interactions
    n_max: int
    commit_widths: Sequence[int]         # stacked width per commitment (common main first)
    # Proof::trace_vdata / public_values, in vk order (proof.rs:24-31)
    trace_vdata: List[Optional[Tuple[int, list]]] = field(default_factory=list)  # (log_height, [cached commitments])
    public_values: List[np.ndarray] = field(default_factory=list)


class _W:
    def __init__(self):
        self.parts = []

    def u32(self, x):
        if not 0 <= int(x) < (1 << 32):  # usize::encode fails above u32 (codec.rs:197-201)
            raise ValueError("length does not fit the codec's u32")
        self.parts.append(struct.pack("<I", int(x)))

    def u8(self, x):
        self.parts.append(struct.pack("<B", int(x)))

    def raw(self, canon_words):
        self.parts.append(np.ascontiguousarray(canon_words, dtype="<u4").tobytes())

    def bytes(self):
        return b"".join(self.parts)


class _Cursor:
    """Sequential reader over one flat section of canonical words."""

    def __init__(self, words):
        self.w, self.pos = words, 0

    def take(self, n):
        if self.pos + n > len(self.w):
            raise ValueError("flat proof section shorter than its shape")
        out = self.w[self.pos:self.pos + n]
        self.pos += n
        return out

    def done(self):
        if self.pos != len(self.w):
            raise ValueError(f"flat proof section has {len(self.w) - self.pos} trailing words")


def _claims_flat_to_wire(words):
    """Per layer: EF elements (0, 1, 2, 3) -> (0, 2, 1, 3)."""
    a = np.asarray(words).reshape(-1, 4, 4)
    return a[:, [0, 2, 1, 3], :].reshape(-1)


def encode_gkr_and_batch_constraints(w, shape, bc_words):
    """GkrProof then BatchConstraintProof (proof.rs:259-304) from the flat section of
    swirl_prove_batch_constraints."""
    c = _Cursor(_canon(bc_words))
    L, n, D = shape.gkr_layers, len(shape.airs), shape.max_constraint_degree
    w.raw(c.take(1))                       # logup_pow_witness
    w.raw(c.take(4))                       # q0_claim
    w.u32(L)                               # Vec<GkrLayerClaims>
    # the flat section holds a layer's claims in transcript order p(xi,0), q(xi,0), p(xi,1), q(xi,1)
    # (fractional_sumcheck_gkr.rs:185-193); GkrLayerClaims::encode writes p_xi_0, p_xi_1, q_xi_0, q_xi_1 (proof.rs:211-218)
    w.raw(_claims_flat_to_wire(c.take(16 * L)))
    w.raw(c.take(12 * (L * (L - 1) // 2)))  # sumcheck_polys: no prefixes (proof.rs:264-270)
    w.u32(n)                               # numerator_term_per_air (slice)
    w.raw(c.take(4 * n))
    w.raw(c.take(4 * n))                   # denominator_term_per_air (iter)
    n_uni = (D + 1) * ((1 << shape.l_skip) - 1) + 1
    w.u32(n_uni)
    w.raw(c.take(4 * n_uni))
    w.u32(shape.n_max)
    if shape.n_max > 0:
        w.u32(D + 1)
        w.raw(c.take(4 * (D + 1) * shape.n_max))
    for a in shape.airs:
        w.u32(len(a.part_widths))
        for width in a.part_widths:
            m = width * (2 if a.need_rot else 1)
            w.u32(m)
            w.raw(c.take(4 * m))
    c.done()


def encode_stacking(w, shape, st_words):
    """StackingProof (proof.rs:306-321)."""
    c = _Cursor(_canon(st_words))
    n_uni = 2 * ((1 << shape.l_skip) - 1) + 1
    w.u32(n_uni)
    w.raw(c.take(4 * n_uni))
    w.u32(shape.n_stack)
    w.raw(c.take(8 * shape.n_stack))
    w.u32(len(shape.commit_widths))
    for width in shape.commit_widths:
        w.u32(width)
        w.raw(c.take(4 * width))
    c.done()


def encode_whir(w, shape, wh_words):
    """WhirProof (proof.rs:323-420)."""
    c = _Cursor(_canon(wh_words))
    k, R = shape.k_whir, len(shape.num_queries)
    m = shape.l_skip + shape.n_stack
    w.raw(c.take(1))                       # mu_pow_witness
    w.u32(R * k)
    w.raw(c.take(8 * R * k))               # whir_sumcheck_polys
    w.u32(R - 1)                           # codeword_commits (digest slice)
    w.raw(c.take(8 * (R - 1)))
    w.raw(c.take(4 * (R - 1)))             # ood_values
    w.raw(c.take(R * k))                   # folding_pow_witnesses
    w.raw(c.take(R))                       # query_phase_pow_witnesses
    nc, q0 = len(shape.commit_widths), shape.num_queries[0]
    w.u32(nc)
    w.u32(q0)
    depth0 = m + shape.log_blowup - k
    rows = c.take(sum(q0 * (1 << k) * width for width in shape.commit_widths))
    paths = c.take(nc * q0 * depth0 * 8)
    if q0 > 0:
        w.u32(depth0)
        for width in shape.commit_widths:
            w.u32(width)
        w.raw(rows)
        w.raw(paths)
    for r in range(1, R):
        w.u32(shape.num_queries[r])
        w.raw(c.take(shape.num_queries[r] * (1 << k) * 4))
    w.u32(m + shape.log_blowup - 1 - k if (R > 1 and q0 > 0) else 0)
    for r in range(1, R):
        w.raw(c.take(shape.num_queries[r] * (m + shape.log_blowup - r - k) * 8))
    n_final = 1 << (m - R * k)
    w.u32(n_final)
    w.raw(c.take(4 * n_final))
    c.done()


def encode_proof(shape, common_main_commit, bc_words, stacking_words, whir_words):
    """Bytes of `Proof::encode` (proof.rs:226-257)."""
    w = _W()
    w.u32(CODEC_VERSION)
    w.raw(_canon(common_main_commit))
    tv = shape.trace_vdata
    w.u32(len(tv))
    for i in range(0, len(tv), 8):
        byte = 0
        for j, v in enumerate(tv[i:i + 8]):
            byte |= (1 if v is not None else 0) << j
        w.u8(byte)
    for v in tv:
        if v is not None:
            log_height, cached = v
            w.u32(log_height)
            w.u32(len(cached))
            for d in cached:
                w.raw(_canon(d))
    w.u32(len(shape.public_values))
    for pv in shape.public_values:
        pv = np.asarray(pv, dtype=np.uint32).reshape(-1)
        w.u32(len(pv))
        w.raw(_canon(pv))
    encode_gkr_and_batch_constraints(w, shape, bc_words)
    encode_stacking(w, shape, stacking_words)
    encode_whir(w, shape, whir_words)
    return w.bytes()


class _R:
    def __init__(self, data):
        self.b, self.pos = memoryview(data), 0

    def u32(self):
        if self.pos + 4 > len(self.b):
            raise ValueError("unexpected end of proof bytes")
        (x,) = struct.unpack_from("<I", self.b, self.pos)
        self.pos += 4
        return x

    def u8(self):
        if self.pos + 1 > len(self.b):
            raise ValueError("unexpected end of proof bytes")
        x = self.b[self.pos]
        self.pos += 1
        return x

    def field(self, n):
        """n canonical base-field words; values >= p are rejected (codec.rs:218-230)."""
        if self.pos + 4 * n > len(self.b):
            raise ValueError("unexpected end of proof bytes")
        a = np.frombuffer(self.b, dtype="<u4", count=n, offset=self.pos).astype(np.uint32)
        self.pos += 4 * n
        if n and int(a.max()) >= P:
            raise ValueError("field element out of range")
        return a


def decode_proof(data):
    """`Proof::decode` (proof.rs:444-705).  Returns a dict of canonical numpy arrays shaped like the reference
    struct fields plus `flat` = the three flat sections (canonical words) in the C-ABI layouts."""
    r = _R(data)
    if r.u32() != CODEC_VERSION:
        raise ValueError("CODEC_VERSION mismatch")
    out = {"common_main_commit": r.field(8)}
    num_airs = r.u32()
    bitmap = [r.u8() for _ in range((num_airs + 7) // 8)]
    present = []
    for bi, byte in enumerate(bitmap):
        nbits = min(8, num_airs - 8 * bi)
        if byte >> nbits:
            raise ValueError("trace_vdata bitmap padding bits set")
        present += [(byte >> i) & 1 for i in range(nbits)]
    tv = []
    for bit in present:
        if bit:
            lh = r.u32()
            tv.append((lh, [r.field(8) for _ in range(r.u32())]))
        else:
            tv.append(None)
    out["trace_vdata"] = tv
    out["public_values"] = [r.field(r.u32()) for _ in range(r.u32())]
    # GkrProof
    bc = [r.field(1), r.field(4)]
    L = r.u32()
    bc.append(_claims_flat_to_wire(r.field(16 * L)))  # the permutation (p0, q0, p1, q1) <-> (p0, p1, q0, q1) is an involution
    bc.append(r.field(12 * (L * (L - 1) // 2)))
    # BatchConstraintProof
    n = r.u32()
    bc += [r.field(4 * n), r.field(4 * n)]
    n_uni = r.u32()
    bc.append(r.field(4 * n_uni))
    n_max = r.u32()
    deg1 = r.u32() if n_max > 0 else 0
    bc.append(r.field(4 * deg1 * n_max))
    col_open = []
    for _ in range(n):
        parts = [r.field(4 * r.u32()) for _ in range(r.u32())]
        col_open.append(parts)
        bc += parts
    out.update(gkr_layers=L, n_present=n, n_max=n_max, bc_uni_len=n_uni, bc_round_len=deg1,
               column_openings=[[p.reshape(-1, 4) for p in parts] for parts in col_open])
    # StackingProof
    st = [r.field(4 * r.u32())]
    n_stack = r.u32()
    st.append(r.field(8 * n_stack))
    widths_st = []
    for _ in range(r.u32()):
        widths_st.append(r.u32())
        st.append(r.field(4 * widths_st[-1]))
    out.update(n_stack=n_stack, stacking_widths=widths_st)
    # WhirProof
    wh = [r.field(1)]
    n_sc = r.u32()
    wh.append(r.field(8 * n_sc))
    n_commits_cw = r.u32()
    wh.append(r.field(8 * n_commits_cw))
    R = n_commits_cw + 1
    if n_sc % R:
        raise ValueError("num_whir_sumcheck_rounds must be a multiple of num_whir_rounds")
    k = n_sc // R
    wh += [r.field(4 * (R - 1)), r.field(n_sc), r.field(R)]
    nc = r.u32()
    if nc == 0:
        raise ValueError("num_commits must be nonzero")
    q0 = r.u32()
    depth0, widths = 0, [0] * nc
    if q0 > 0:
        depth0 = r.u32()
        widths = [r.u32() for _ in range(nc)]
    for width in widths:
        wh.append(r.field(q0 * (1 << k) * width))
    wh.append(r.field(nc * q0 * depth0 * 8))
    nq = []
    for _ in range(R - 1):
        nq.append(r.u32())
        wh.append(r.field(nq[-1] * (1 << k) * 4))
    depth = r.u32()
    for q in nq:
        wh.append(r.field(q * depth * 8))
        depth -= 1
    wh.append(r.field(4 * r.u32()))
    if r.pos != len(r.b):
        raise ValueError("trailing bytes after the proof")
    out.update(k_whir=k, whir_rounds=R, num_queries=[q0] + nq, commit_widths=widths, initial_merkle_depth=depth0)
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.uint32)
    out["flat"] = {"constraints_proof": cat(bc), "stacking_proof": cat(st), "whir_proof": cat(wh)}
    return out


def flat_to_montgomery(decoded):
    """The three flat sections of a decoded proof as Montgomery words (what the C ABI produces)."""
    return {k: _mont(v) for k, v in decoded["flat"].items()}
