// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the shipped product
// path; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load it, and only as the checker (or the timed CPU baseline), never as the thing
// shipped.
//
// BabyBear base field and its degree-4 binomial extension, plain scalar C++ restatement.
//
// Follows (paths relative to /root/reference):
//   crates/cuda-common/include/fp.h:52-65        p, R2, TWO_ADICITY, Montgomery-in-memory
//   crates/cuda-common/include/fp.h:291-320      TWO_ADIC_GENERATORS (canonical) — KAT table
//   crates/cuda-common/include/fpext.h:37-121    EF = F[X]/(X^4 - 11), basis (1,X,X^2,X^3)
//   crates/cuda-common/include/ff/baby_bear.hpp:24,78-80   M0, RR, ONE, beta = 11
// The arithmetic itself is p3-baby-bear / p3-monty-31 / p3-field 0.4.3 (crates.io, not
// vendored in the reference): p = 15*2^27+1, Montgomery radix 2^32, words stored as x*2^32 mod p.
#pragma once
#include <cstdint>
#include <cstddef>

namespace orc {

constexpr uint32_t P = 0x78000001u;          // 2013265921
constexpr uint32_t MONTY_NEG_PINV = 0x77ffffffu;  // -p^{-1} mod 2^32
constexpr uint32_t MONTY_R2 = 1172168163u;   // 2^64 mod p
constexpr uint32_t MONTY_ONE = 0x0ffffffeu;  // 2^32 mod p
constexpr int TWO_ADICITY = 27;

// A base-field element, held as its Montgomery word (exactly the in-memory format the
// reference moves across the host/device boundary, data_transporter.rs:93-106).
struct F {
    uint32_t v;  // Montgomery form, always canonical: v < P  (value-initialise: F{} == 0)
    static constexpr F raw(uint32_t m) { return F{m}; }
    bool operator==(const F& o) const { return v == o.v; }
    bool operator!=(const F& o) const { return v != o.v; }
};

inline uint32_t monty_reduce(uint64_t x) {
    // x < p * 2^32  ->  x * 2^-32 mod p, canonical
    uint32_t m = (uint32_t)x * MONTY_NEG_PINV;
    uint64_t t = (x + (uint64_t)m * P) >> 32;
    return t >= P ? (uint32_t)(t - P) : (uint32_t)t;
}

inline F operator+(F a, F b) {
    uint32_t s = a.v + b.v;  // < 2^32 since both < 2^31
    return F::raw(s >= P ? s - P : s);
}
inline F operator-(F a, F b) { return F::raw(a.v >= b.v ? a.v - b.v : a.v + P - b.v); }
inline F operator-(F a) { return F::raw(a.v ? P - a.v : 0); }
inline F operator*(F a, F b) { return F::raw(monty_reduce((uint64_t)a.v * b.v)); }
inline F& operator+=(F& a, F b) { a = a + b; return a; }
inline F& operator-=(F& a, F b) { a = a - b; return a; }
inline F& operator*=(F& a, F b) { a = a * b; return a; }

inline F from_canonical(uint64_t x) { return F::raw(monty_reduce((uint64_t)(x % P) * MONTY_R2)); }
inline uint32_t to_canonical(F a) { return monty_reduce(a.v); }
inline F f_zero() { return F::raw(0); }
inline F f_one() { return F::raw(MONTY_ONE); }
inline F f_two() { return f_one() + f_one(); }
inline F halve(F a) { return F::raw((a.v & 1) ? (uint32_t)(((uint64_t)a.v + P) >> 1) : a.v >> 1); }

inline F f_pow(F b, uint64_t e) {
    F r = f_one();
    while (e) {
        if (e & 1) r *= b;
        b *= b;
        e >>= 1;
    }
    return r;
}
inline F f_inv(F a) { return f_pow(a, P - 2); }  // 0 -> 0

// Multiplicative generator 31 (F::GENERATOR; used at prover/sumcheck.rs:82-86).
inline F f_generator() { return from_canonical(31); }

// two_adic_generator(bits): canonical values pinned by fp.h:291-320.  Derived here from the
// top entry (0x1a427a41, order 2^27) by repeated squaring, as p3-baby-bear does.
inline F two_adic_generator(int bits) {
    F g = from_canonical(0x1a427a41u);
    for (int i = bits; i < TWO_ADICITY; i++) g *= g;
    return g;
}

// ---------------------------------------------------------------------------------------------
// EF = F[X]/(X^4 - 11)
// ---------------------------------------------------------------------------------------------
struct EF {
    F c[4];
    bool operator==(const EF& o) const {
        return c[0] == o.c[0] && c[1] == o.c[1] && c[2] == o.c[2] && c[3] == o.c[3];
    }
    bool operator!=(const EF& o) const { return !(*this == o); }
};

inline F f_beta() { return from_canonical(11); }
inline EF ef_zero() { return EF{}; }
inline EF ef_from(F a) { EF r{}; r.c[0] = a; return r; }
inline EF ef_one() { return ef_from(f_one()); }
inline EF operator+(EF a, EF b) { EF r; for (int i = 0; i < 4; i++) r.c[i] = a.c[i] + b.c[i]; return r; }
inline EF operator-(EF a, EF b) { EF r; for (int i = 0; i < 4; i++) r.c[i] = a.c[i] - b.c[i]; return r; }
inline EF operator-(EF a) { EF r; for (int i = 0; i < 4; i++) r.c[i] = -a.c[i]; return r; }
inline EF operator*(EF a, F s) { EF r; for (int i = 0; i < 4; i++) r.c[i] = a.c[i] * s; return r; }
inline EF operator*(EF a, EF b) {
    // schoolbook, then fold X^4 = 11
    F t[7] = {};
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) t[i + j] += a.c[i] * b.c[j];
    F w = f_beta();
    EF r;
    r.c[0] = t[0] + w * t[4];
    r.c[1] = t[1] + w * t[5];
    r.c[2] = t[2] + w * t[6];
    r.c[3] = t[3];
    return r;
}
inline EF& operator+=(EF& a, EF b) { a = a + b; return a; }
inline EF& operator-=(EF& a, EF b) { a = a - b; return a; }
inline EF& operator*=(EF& a, EF b) { a = a * b; return a; }
inline EF& operator*=(EF& a, F b) { a = a * b; return a; }
inline EF ef_add_base(EF a, F b) { a.c[0] += b; return a; }

inline EF ef_pow(EF b, uint64_t e) {
    EF r = ef_one();
    while (e) {
        if (e & 1) r *= b;
        b *= b;
        e >>= 1;
    }
    return r;
}

// Inverse through the norm to the quadratic subfield F[Y]/(Y^2-11), Y = X^2.
// a = A(Y) + X*B(Y) with A = a0 + a2 Y, B = a1 + a3 Y;  a * (A - X B) = A^2 - Y B^2 =: N(Y).
inline EF ef_inv(EF a) {
    F w = f_beta();
    F a0 = a.c[0], a1 = a.c[1], a2 = a.c[2], a3 = a.c[3];
    // A^2 = (a0^2 + w a2^2) + (2 a0 a2) Y ;  B^2 = (a1^2 + w a3^2) + (2 a1 a3) Y
    F A2_0 = a0 * a0 + w * a2 * a2, A2_1 = (a0 * a2) + (a0 * a2);
    F B2_0 = a1 * a1 + w * a3 * a3, B2_1 = (a1 * a3) + (a1 * a3);
    // Y * B^2 = w*B2_1 + B2_0 Y
    F n0 = A2_0 - w * B2_1, n1 = A2_1 - B2_0;
    // 1/N = (n0 - n1 Y) / (n0^2 - w n1^2)
    F d = f_inv(n0 * n0 - w * n1 * n1);
    F i0 = n0 * d, i1 = -(n1 * d);
    // result = (A - X B) * (i0 + i1 Y)
    //   A*(i0 + i1 Y) = (a0 i0 + w a2 i1) + (a0 i1 + a2 i0) Y
    //   B*(i0 + i1 Y) = (a1 i0 + w a3 i1) + (a1 i1 + a3 i0) Y
    EF r;
    r.c[0] = a0 * i0 + w * a2 * i1;
    r.c[2] = a0 * i1 + a2 * i0;
    r.c[1] = -(a1 * i0 + w * a3 * i1);
    r.c[3] = -(a1 * i1 + a3 * i0);
    return r;
}

}  // namespace orc
