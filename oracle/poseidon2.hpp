// ORACLE — TEST INFRASTRUCTURE ONLY (see bb31.hpp header).
//
// Poseidon2 over BabyBear, width 16, S-box x^7, 4 + 13 + 4 rounds, and the two hash modes the
// SWIRL commitment uses.  CPU restatement of:
//   crates/cuda-common/include/poseidon2.cuh:14-202   permutation structure + constants
//   crates/cuda-backend/cuda/src/merkle_tree.cu:32-51  leaf sponge (PaddingFreeSponge<16,8,8>)
//   crates/cuda-backend/cuda/src/merkle_tree.cu:166-172 2-to-1 compress (TruncatedPermutation)
//   crates/stark-sdk/src/config/baby_bear_poseidon2.rs:21-42,74-78  which instance is used
// The algorithm itself is p3-poseidon2 / p3-baby-bear / p3-symmetric 0.4.3 (crates.io; not
// vendored in /root/reference).
//
// PARITY UNPINNED for digests: the reference tree holds no known-answer vector for a BabyBear
// Poseidon2 permutation output or a Merkle root.  What pins this file:
//  (1) the round constants / diagonal are the reference's numeric tables,
//  (2) dense-matrix spec form == optimised form (tests/test_oracle_poseidon2.py),
//  (3) SURVEY.md scratch self-consistency values perm(0..0)[0..4], perm(0..15)[0..4].
#pragma once
#include "bb31.hpp"

namespace orc {

#include "poseidon2_rc.inc"

constexpr int P2_WIDTH = 16;
constexpr int P2_RATE = 8;
constexpr int P2_DIGEST = 8;

struct Poseidon2Tables {
    F ext_init[4][16];
    F internal[13];
    F ext_term[4][16];
    F diag[16];  // the "d" in  M_I = J + diag(d)
    Poseidon2Tables() {
        for (int r = 0; r < 4; r++)
            for (int i = 0; i < 16; i++) {
                ext_init[r][i] = from_canonical(P2_RC_EXT_INITIAL[r * 16 + i]);
                ext_term[r][i] = from_canonical(P2_RC_EXT_TERMINAL[r * 16 + i]);
            }
        for (int r = 0; r < 13; r++) internal[r] = from_canonical(P2_RC_INTERNAL[r]);
        // d = (-2, 1, 2, 1/2, 3, 4, -1/2, -3, -4, 2^-8, 1/4, 1/8, 2^-27, -2^-8, -1/16, -2^-27)
        // (poseidon2.cuh:50-67).  Built from field operations rather than literals.
        F one = f_one(), two = f_two(), inv2 = f_inv(two);
        auto inv2pow = [&](int k) { return f_pow(inv2, (uint64_t)k); };
        diag[0] = -two;
        diag[1] = one;
        diag[2] = two;
        diag[3] = inv2;
        diag[4] = two + one;
        diag[5] = two + two;
        diag[6] = -inv2;
        diag[7] = -(two + one);
        diag[8] = -(two + two);
        diag[9] = inv2pow(8);
        diag[10] = inv2pow(2);
        diag[11] = inv2pow(3);
        diag[12] = inv2pow(27);
        diag[13] = -inv2pow(8);
        diag[14] = -inv2pow(4);
        diag[15] = -inv2pow(27);
    }
};

inline const Poseidon2Tables& p2_tables() {
    static const Poseidon2Tables t;
    return t;
}

inline F p2_sbox(F x) {
    F x2 = x * x;
    F x3 = x2 * x;
    F x4 = x2 * x2;
    return x3 * x4;
}

// External linear layer: circ(2*M4, M4, M4, M4), M4 = circ-like [[2,3,1,1],[1,2,3,1],[1,1,2,3],[3,1,1,2]].
inline void p2_external_linear(F s[16]) {
    F colsum[4] = {};
    for (int b = 0; b < 4; b++) {
        F* x = s + 4 * b;
        F all = x[0] + x[1] + x[2] + x[3];
        // row i of M4 . x  =  all + x[i] + 2*x[i+1]   (indices mod 4)
        F y[4];
        for (int i = 0; i < 4; i++) {
            F nxt = x[(i + 1) & 3];
            y[i] = all + x[i] + nxt + nxt;
        }
        for (int i = 0; i < 4; i++) {
            x[i] = y[i];
            colsum[i] += y[i];
        }
    }
    for (int i = 0; i < 16; i++) s[i] += colsum[i & 3];
}

inline void p2_internal_linear(F s[16], const Poseidon2Tables& t) {
    F total = f_zero();
    for (int i = 0; i < 16; i++) total += s[i];
    for (int i = 0; i < 16; i++) s[i] = total + t.diag[i] * s[i];
}

inline void poseidon2_permute(F s[16]) {
    const Poseidon2Tables& t = p2_tables();
    p2_external_linear(s);
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 16; i++) s[i] = p2_sbox(s[i] + t.ext_init[r][i]);
        p2_external_linear(s);
    }
    for (int r = 0; r < 13; r++) {
        s[0] = p2_sbox(s[0] + t.internal[r]);
        p2_internal_linear(s, t);
    }
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 16; i++) s[i] = p2_sbox(s[i] + t.ext_term[r][i]);
        p2_external_linear(s);
    }
}

struct Digest {
    F w[8];
    bool operator==(const Digest& o) const {
        for (int i = 0; i < 8; i++)
            if (w[i] != o.w[i]) return false;
        return true;
    }
};

// PaddingFreeSponge<16, 8, 8>::hash_slice — overwrite the first len<=8 words, permute, repeat;
// no padding, no permutation at all for empty input.
inline Digest hash_slice(const F* vals, size_t n) {
    F st[16] = {};
    for (size_t off = 0; off < n; off += P2_RATE) {
        size_t len = n - off < (size_t)P2_RATE ? n - off : (size_t)P2_RATE;
        for (size_t i = 0; i < len; i++) st[i] = vals[off + i];
        poseidon2_permute(st);
    }
    Digest d;
    for (int i = 0; i < 8; i++) d.w[i] = st[i];
    return d;
}

// TruncatedPermutation<_, 2, 8, 16>::compress
inline Digest compress(const Digest& l, const Digest& r) {
    F st[16];
    for (int i = 0; i < 8; i++) {
        st[i] = l.w[i];
        st[8 + i] = r.w[i];
    }
    poseidon2_permute(st);
    Digest d;
    for (int i = 0; i < 8; i++) d.w[i] = st[i];
    return d;
}

}  // namespace orc
