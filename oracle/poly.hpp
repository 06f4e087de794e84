// ORACLE — TEST INFRASTRUCTURE ONLY (see bb31.hpp header).
//
// Multilinear / univariate polynomial helpers of the SWIRL prover, plain CPU restatement of
//   crates/stark-backend/src/prover/poly.rs:24-131      Mle (evals<->coeffs, eval_at_point)
//   crates/stark-backend/src/prover/poly.rs:133-208     evals_eq_hypercube(s), evals_mobius_eq_hypercube
//   crates/stark-backend/src/poly_common.rs:7-140       eval_eq_mle, eval_mobius_eq_mle, eval_eq_uni(_at_one),
//                                                       eq_uni_poly, eval_in_uni, evals_eq_hypercube_serial
//   crates/stark-backend/src/poly_common.rs:141-260     eval_eq_sharp_uni, rot kernels, horner, interpolate_*
//   crates/stark-backend/src/prover/sumcheck.rs:395-440 fold_mle_evals
#pragma once
#include <stdexcept>
#include <vector>

#include "commit.hpp"

namespace orc {

inline EF ef_from_u64(uint64_t x) { return ef_from(from_canonical(x)); }
inline EF ef_dbl(EF a) { return a + a; }
inline EF ef_halve(EF a) { EF r; for (int i = 0; i < 4; i++) r.c[i] = halve(a.c[i]); return r; }
inline bool ef_is_zero(const EF& a) { return a == ef_zero(); }

// poly.rs:133-149
inline std::vector<EF> evals_eq_hypercube(const std::vector<EF>& x) {
    std::vector<EF> out(size_t(1) << x.size(), ef_zero());
    out[0] = ef_one();
    for (size_t i = 0; i < x.size(); i++) {
        const size_t half = size_t(1) << i;
        for (size_t j = 0; j < half; j++) {
            out[half + j] = out[j] * x[i];
            out[j] = out[j] * (ef_one() - x[i]);
        }
    }
    return out;
}
// poly.rs:160-178
inline std::vector<EF> evals_mobius_eq_hypercube(const std::vector<EF>& u) {
    std::vector<EF> out(size_t(1) << u.size(), ef_zero());
    out[0] = ef_one();
    for (size_t i = 0; i < u.size(); i++) {
        const EF w0 = ef_one() - ef_dbl(u[i]), w1 = u[i];
        const size_t half = size_t(1) << i;
        for (size_t j = 0; j < half; j++) {
            const EF prev = out[j];
            out[half + j] = prev * w1;
            out[j] = prev * w0;
        }
    }
    return out;
}
// poly.rs:182-193: concatenation of eq tables of every prefix length, masks in the other endianness
template <class It>
inline std::vector<EF> evals_eq_hypercubes(size_t n, It begin, It end) {
    std::vector<EF> out((size_t(2) << n) - 1, ef_zero());
    out[0] = ef_one();
    size_t i = 0;
    for (It it = begin; it != end; ++it, ++i) {
        const EF x_i = *it;
        for (size_t y = 0; y < (size_t(1) << i); y++) {
            out[(size_t(1) << (i + 1)) - 1 + (2 * y + 1)] = out[(size_t(1) << i) - 1 + y] * x_i;
            out[(size_t(1) << (i + 1)) - 1 + (2 * y)] = out[(size_t(1) << i) - 1 + y] * (ef_one() - x_i);
        }
    }
    return out;
}

// poly_common.rs:7-22
inline EF eval_eq_mle(const EF* x, const EF* y, size_t n) {
    EF acc = ef_one();
    for (size_t i = 0; i < n; i++) acc = acc * (ef_one() - y[i] - x[i] + ef_dbl(x[i] * y[i]));
    return acc;
}
inline EF eval_eq_mle1(EF x, bool b) {  // eq(x, b) for boolean b
    return b ? x : ef_one() - x;
}
// poly_common.rs:29-36
inline EF eval_mobius_eq_mle(const std::vector<EF>& u, const std::vector<EF>& x) {
    EF acc = ef_one();
    for (size_t i = 0; i < u.size(); i++) {
        const EF w0 = ef_one() - ef_dbl(u[i]);
        acc = acc * (w0 * (ef_one() - x[i]) + u[i] * x[i]);
    }
    return acc;
}
inline F f_half_pow(int l) { return f_pow(halve(f_one()), (uint64_t)l); }
// poly_common.rs:60-66
inline EF eval_eq_uni(int l_skip, EF x, EF y) {
    EF res = ef_one();
    for (int i = 0; i < l_skip; i++) {
        res = (x + y) * res + (x - ef_one()) * (y - ef_one());
        x = x * x;
        y = y * y;
    }
    return res * f_half_pow(l_skip);
}
// poly_common.rs:71-77
inline EF eval_eq_uni_at_one(int l_skip, EF x) {
    EF res = ef_one();
    for (int i = 0; i < l_skip; i++) {
        res = res * (x + ef_one());
        x = x * x;
    }
    return res * f_half_pow(l_skip);
}
inline EF ef_exp_power_of_2(EF x, int k) {
    for (int i = 0; i < k; i++) x = x * x;
    return x;
}
// poly_common.rs:104-114
inline EF eval_in_uni(int l_skip, int n, EF z) {
    if (n < 0) return eval_eq_uni_at_one(-n, ef_exp_power_of_2(z, l_skip + n));
    return ef_one();
}
// poly_common.rs:85-102: eq_D(x, Z) as coefficients in Z
inline std::vector<EF> eq_uni_poly(int l_skip, EF x) {
    const F n_inv = f_half_pow(l_skip);
    const size_t N = size_t(1) << l_skip;
    std::vector<EF> coeffs(N);
    EF xp = x;
    for (size_t i = 0; i < N; i++) {  // x^1 .. x^N scaled
        coeffs[i] = xp * n_inv;
        xp = xp * x;
    }
    std::vector<EF> rev(coeffs.rbegin(), coeffs.rend());
    rev[0] = ef_from(n_inv);
    return rev;
}
// poly_common.rs:134-176
inline EF eval_eq_sharp_uni(const std::vector<F>& omega_skip_pows, const EF* xi_1, int l_skip, EF z) {
    std::vector<EF> xi(xi_1, xi_1 + l_skip);
    std::vector<EF> eq_xi = evals_eq_hypercube(xi);
    EF res = ef_zero();
    for (size_t i = 0; i < omega_skip_pows.size(); i++)
        res += eval_eq_uni(l_skip, z, ef_from(omega_skip_pows[i])) * eq_xi[i];
    return res;
}
// poly_common.rs:196-207
inline void eval_eq_rot_cube(const EF* x, const EF* y, size_t n, EF* eq_out, EF* rot_out) {
    EF rot = ef_one(), eq = ef_one();
    for (size_t i = n; i-- > 0;) {
        rot = x[i] * (ef_one() - y[i]) * eq + (ef_one() - x[i]) * y[i] * rot;
        eq = eq * (x[i] * y[i] + (ef_one() - x[i]) * (ef_one() - y[i]));
    }
    *eq_out = eq;
    *rot_out = rot;
}
// poly_common.rs:181-193
inline EF eval_rot_kernel_prism(int l_skip, const EF* x, const EF* y, size_t n_plus_1) {
    const EF omega = ef_from(two_adic_generator(l_skip));
    EF eq_cube, rot_cube;
    eval_eq_rot_cube(x + 1, y + 1, n_plus_1 - 1, &eq_cube, &rot_cube);
    return eval_eq_uni(l_skip, x[0], y[0] * omega) * eq_cube +
           eval_eq_uni_at_one(l_skip, x[0]) * eval_eq_uni_at_one(l_skip, y[0] * omega) * (rot_cube - eq_cube);
}

// poly_common.rs:232-243
inline EF horner_eval(const std::vector<EF>& coeffs, EF x) {
    EF acc = ef_zero();
    for (size_t i = coeffs.size(); i-- > 0;) acc = acc * x + coeffs[i];
    return acc;
}
inline EF interpolate_linear_at_01(EF e0, EF e1, EF x) { return (e1 - e0) * x + e0; }
// poly_common.rs:256-263
inline EF interpolate_quadratic_at_012(const EF e[3], EF x) {
    const EF s1 = e[1] - e[0], s2 = e[2] - e[1];
    const EF p = ef_halve(s2 - s1), q = s1 - p;
    return (p * x + q) * x + e[0];
}
// poly_common.rs:268-283
inline EF interpolate_cubic_at_0123(const EF e[4], EF x) {
    const F inv6 = f_inv(from_canonical(6));
    const EF s1 = e[1] - e[0], s2 = e[2] - e[0], s3 = e[3] - e[0];
    const EF d3 = s3 - (s2 - s1) * from_canonical(3);
    const EF p = d3 * inv6;
    const EF q = ef_halve(s2 - d3) - s1;
    const EF r = s1 - p - q;
    return ((p * x + q) * x + r) * x + e[0];
}

// sumcheck.rs:395-414: fold the low variable of every column of a column-major EF matrix
inline void fold_mle_evals(std::vector<EF>& values, size_t& height, size_t width, EF r) {
    if (height <= 1) return;
    const size_t nh = height / 2;
    std::vector<EF> out(nh * width);
    const size_t h0 = height;
    parallel_for(nh * width, [&](size_t b, size_t e) {
        for (size_t i = b; i < e; i++) {
            const size_t c = i / nh, y = i % nh;
            const EF t0 = values[c * h0 + 2 * y], t1 = values[c * h0 + 2 * y + 1];
            out[i] = t0 + (t1 - t0) * r;
        }
    }, 4096);
    values.swap(out);
    height = nh;
}

// poly.rs:99-131 (generic over the element type through + and -)
template <class T>
inline void mle_evals_to_coeffs_inplace(std::vector<T>& a) {
    const size_t n = a.size();
    for (size_t step = 1; step < n; step <<= 1)
        for (size_t i = 0; i < n; i += 2 * step)
            for (size_t j = 0; j < step; j++) a[i + j + step] = a[i + j + step] - a[i + j];
}
template <class T>
inline void mle_coeffs_to_evals_inplace(T* a, size_t n) {
    for (size_t step = 1; step < n; step <<= 1)
        for (size_t i = 0; i < n; i += 2 * step)
            for (size_t j = 0; j < step; j++) a[i + j + step] = a[i + j + step] + a[i + j];
}
// poly.rs:60-76: naive evaluation of an MLE in coefficient form
inline EF mle_eval_at_point(const std::vector<EF>& coeffs, const std::vector<EF>& x) {
    EF res = ef_zero();
    for (size_t i = 0; i < coeffs.size(); i++) {
        EF term = coeffs[i];
        for (size_t j = 0; j < x.size(); j++)
            if ((i >> j) & 1) term = term * x[j];
        res += term;
    }
    return res;
}

}  // namespace orc
