// Host-side polynomial helpers of the phase orchestration: everything here works on a handful of
// field elements per sumcheck round (univariate-skip factors, interpolation of round polynomials)
// and therefore stays on the CPU next to the transcript; the data-sized work is in the kernels.
//
// Replaces (reference, relative to /root/reference):
//   crates/stark-backend/src/poly_common.rs:7-140,232-283    eval_eq_mle / eval_eq_uni / eval_eq_uni_at_one /
//                                                            eval_in_uni / eq_uni_poly / horner / interpolate_*
//   crates/stark-backend/src/prover/poly.rs:619-750          from_geometric_cosets_evals_idft
//   crates/stark-backend/src/prover/poly.rs:349-420          lagrange_interpolate
#pragma once
#include <vector>

#include "bb31.cuh"

namespace swirl {
namespace hp {

using bb::Ext;
using bb::ext_add;
using bb::ext_mul;
using bb::ext_sub;

__host__ __device__ inline Ext from_words(const uint32_t* w) { return Ext{{w[0], w[1], w[2], w[3]}}; }
inline Ext one() { return bb::ext_one(); }
inline Ext zero() { return bb::ext_zero(); }
inline Ext from_base(uint32_t m) { return bb::ext_from(m); }
inline Ext mul_base(const Ext& a, uint32_t m) { return bb::ext_mul_base(a, m); }
inline bool is_zero(const Ext& a) { return !(a.c[0] | a.c[1] | a.c[2] | a.c[3]); }
inline uint32_t half_pow(int l) { return bb::pow(bb::halve(bb::R1), (uint64_t)l); }
inline Ext exp_pow2(Ext x, int k) {
    for (int i = 0; i < k; i++) x = ext_mul(x, x);
    return x;
}
// poly_common.rs:60-66
inline Ext eval_eq_uni(int l_skip, Ext x, Ext y) {
    Ext res = one();
    for (int i = 0; i < l_skip; i++) {
        res = ext_add(ext_mul(ext_add(x, y), res), ext_mul(ext_sub(x, one()), ext_sub(y, one())));
        x = ext_mul(x, x);
        y = ext_mul(y, y);
    }
    return mul_base(res, half_pow(l_skip));
}
// poly_common.rs:71-77
inline Ext eval_eq_uni_at_one(int l_skip, Ext x) {
    Ext res = one();
    for (int i = 0; i < l_skip; i++) {
        res = ext_mul(res, ext_add(x, one()));
        x = ext_mul(x, x);
    }
    return mul_base(res, half_pow(l_skip));
}
// poly_common.rs:104-114
inline Ext eval_in_uni(int l_skip, int n, Ext z) {
    if (n < 0) return eval_eq_uni_at_one(-n, exp_pow2(z, l_skip + n));
    return one();
}
// eq(x, b) for one variable and a boolean b (eval_eq_mle on a single coordinate)
inline Ext eq1(const Ext& x, bool b) { return b ? x : ext_sub(one(), x); }
// eq(x, y) for one EF coordinate each
inline Ext eq1(const Ext& x, const Ext& y) {
    const Ext xy = ext_mul(x, y);
    return ext_add(ext_sub(ext_sub(one(), y), x), ext_add(xy, xy));
}
inline Ext horner(const std::vector<Ext>& c, const Ext& x) {
    Ext acc = zero();
    for (size_t i = c.size(); i-- > 0;) acc = ext_add(ext_mul(acc, x), c[i]);
    return acc;
}

// Coefficients (size N = 2^log_n) of the interpolant of `evals` over the subgroup <w_N>, natural order.
inline std::vector<Ext> idft_small(const std::vector<Ext>& evals) {
    const size_t n = evals.size();
    int log_n = 0;
    while ((size_t(1) << log_n) < n) log_n++;
    const uint32_t w_inv = bb::inv(bb::two_adic_generator(log_n));
    const uint32_t n_inv = bb::inv(bb::to_mont((uint32_t)n));
    std::vector<Ext> out(n);
    uint32_t x = bb::R1;
    for (size_t k = 0; k < n; k++) {
        Ext acc = zero();
        for (size_t i = n; i-- > 0;) acc = ext_add(mul_base(acc, x), evals[i]);
        out[k] = mul_base(acc, n_inv);
        x = bb::mul(x, w_inv);
    }
    return out;
}
// evaluations of `coeffs` at shift * w_N^i
inline std::vector<Ext> coset_dft_small(const std::vector<Ext>& coeffs, uint32_t shift) {
    const size_t n = coeffs.size();
    int log_n = 0;
    while ((size_t(1) << log_n) < n) log_n++;
    const uint32_t w = bb::two_adic_generator(log_n);
    std::vector<Ext> out(n);
    uint32_t x = shift;
    for (size_t i = 0; i < n; i++) {
        Ext acc = zero();
        for (size_t k = n; k-- > 0;) acc = ext_add(mul_base(acc, x), coeffs[k]);
        out[i] = acc;
        x = bb::mul(x, w);
    }
    return out;
}

// The unique polynomial of degree < N*d with the given values on the d cosets g^(j+1) * D, |D| = N
// = 2^l_skip (g = 31).  evals[z_idx * d + j] = P(g^(j+1) * w^z_idx).  (poly.rs:619-683 with
// shift = init = F::GENERATOR; interpolation is unique, so any exact method gives these coefficients.)
inline std::vector<Ext> interpolate_geometric_cosets(const std::vector<Ext>& evals, int l_skip, int d) {
    const size_t N = size_t(1) << l_skip;
    const uint32_t g = bb::to_mont(31);
    std::vector<std::vector<Ext>> rem(d);
    std::vector<uint32_t> pts(d);
    for (int j = 0; j < d; j++) {
        std::vector<Ext> col(N);
        for (size_t z = 0; z < N; z++) col[z] = evals[z * d + j];
        rem[j] = idft_small(col);
        const uint32_t s = bb::pow(g, (uint64_t)j + 1), s_inv = bb::inv(s);
        uint32_t p = bb::R1;
        for (size_t t = 0; t < N; t++) {
            rem[j][t] = mul_base(rem[j][t], p);
            p = bb::mul(p, s_inv);
        }
        pts[j] = bb::pow(s, N);
    }
    // Lagrange basis over the points a_j = (g^(j+1))^N, coefficient form
    std::vector<std::vector<uint32_t>> basis(d, std::vector<uint32_t>(d, 0));
    for (int i = 0; i < d; i++) {
        std::vector<uint32_t> poly{bb::R1};
        uint32_t denom = bb::R1;
        for (int j = 0; j < d; j++) {
            if (j == i) continue;
            poly.push_back(0);
            for (size_t k = poly.size() - 1; k >= 1; k--) poly[k] = bb::sub(poly[k - 1], bb::mul(pts[j], poly[k]));
            poly[0] = bb::neg(bb::mul(pts[j], poly[0]));
            denom = bb::mul(denom, bb::sub(pts[i], pts[j]));
        }
        const uint32_t inv = bb::inv(denom);
        for (int k = 0; k < d; k++) basis[i][k] = bb::mul(poly[k], inv);
    }
    std::vector<Ext> coeffs(N * d, zero());
    for (size_t t = 0; t < N; t++)
        for (int i = 0; i < d; i++)
            for (int k = 0; k < d; k++) coeffs[k * N + t] = ext_add(coeffs[k * N + t], mul_base(rem[i][t], basis[i][k]));
    return coeffs;
}

// Coefficients of the polynomial through (0, e0), (1, e1), ..., (len-1, e_{len-1})  (poly.rs:349-420)
inline std::vector<Ext> lagrange_interpolate_0n(const std::vector<Ext>& evals) {
    const size_t len = evals.size();
    std::vector<Ext> coeffs(len, zero());
    for (size_t i = 0; i < len; i++) {
        std::vector<uint32_t> poly{bb::R1};
        uint32_t denom = bb::R1;
        for (size_t j = 0; j < len; j++) {
            if (j == i) continue;
            const uint32_t pj = bb::to_mont((uint32_t)j);
            poly.push_back(0);
            for (size_t k = poly.size() - 1; k >= 1; k--) poly[k] = bb::sub(poly[k - 1], bb::mul(pj, poly[k]));
            poly[0] = bb::neg(bb::mul(pj, poly[0]));
            denom = bb::mul(denom, bb::sub(bb::to_mont((uint32_t)i), pj));
        }
        const uint32_t inv = bb::inv(denom);
        for (size_t k = 0; k < len; k++) coeffs[k] = ext_add(coeffs[k], mul_base(evals[i], bb::mul(poly[k], inv)));
    }
    return coeffs;
}

// Lagrange coefficients L_i(r) of the subgroup D = <w_N> at r:  (w^i / N) (r^N - 1) / (r - w^i)
// (what interpolate_coset_with_precomputation evaluates in fold_ple_evals, sumcheck.rs:204-251)
inline std::vector<Ext> lagrange_at(int l_skip, const Ext& r) {
    const size_t N = size_t(1) << l_skip;
    const uint32_t w = bb::two_adic_generator(l_skip);
    const Ext num = mul_base(ext_sub(exp_pow2(r, l_skip), one()), bb::inv(bb::to_mont((uint32_t)N)));
    std::vector<Ext> L(N);
    uint32_t wi = bb::R1;
    for (size_t i = 0; i < N; i++) {
        L[i] = mul_base(ext_mul(num, bb::ext_inv(ext_sub(r, from_base(wi)))), wi);
        wi = bb::mul(wi, w);
    }
    return L;
}

}  // namespace hp
}  // namespace swirl
