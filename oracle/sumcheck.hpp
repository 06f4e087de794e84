// ORACLE — TEST INFRASTRUCTURE ONLY (see bb31.hpp header).
//
// Generic sumcheck building blocks of the SWIRL prover.  CPU restatement of
//   crates/stark-backend/src/prover/sumcheck.rs:49-187     sumcheck_uni_round0_poly (univariate skip round)
//   crates/stark-backend/src/prover/sumcheck.rs:204-251    fold_ple_evals
//   crates/stark-backend/src/prover/sumcheck.rs:271-393    sumcheck_round_poly_evals
//   crates/stark-backend/src/prover/poly.rs:619-750        UnivariatePoly::from_geometric_cosets_evals_idft,
//                                                          lagrange_basis_from_geometric_points
//   crates/stark-backend/src/prover/poly.rs:349-420        UnivariatePoly::lagrange_interpolate
// DFT contract (p3-dft 0.4.3 / dft/radix_2_bowers_serial.rs): natural order in and out,
// coset_dft(coeffs, shift)[i] = P(shift * w^i).
#pragma once
#include <array>
#include <functional>

#include "poly.hpp"

namespace orc {

// ---- tiny EF DFTs (sizes <= a few hundred): direct evaluation --------------------------------
inline std::vector<EF> ef_coset_dft(const std::vector<EF>& coeffs, F shift) {
    const size_t n = coeffs.size();
    const F w = two_adic_generator(log2_strict(n));
    std::vector<EF> out(n);
    F x = shift;
    for (size_t i = 0; i < n; i++) {
        EF acc = ef_zero();
        for (size_t k = n; k-- > 0;) acc = acc * x + coeffs[k];
        out[i] = acc;
        x *= w;
    }
    return out;
}
inline std::vector<EF> ef_dft(const std::vector<EF>& coeffs) { return ef_coset_dft(coeffs, f_one()); }
inline std::vector<EF> ef_idft(const std::vector<EF>& evals) {
    const size_t n = evals.size();
    const F w_inv = f_inv(two_adic_generator(log2_strict(n)));
    const F n_inv = f_inv(from_canonical(n));
    std::vector<EF> out(n);
    F x = f_one();
    for (size_t k = 0; k < n; k++) {
        EF acc = ef_zero();
        for (size_t i = n; i-- > 0;) acc = acc * x + evals[i];
        out[k] = acc * n_inv;
        x *= w_inv;
    }
    return out;
}
inline std::vector<F> f_idft_small(std::vector<F> v) {
    idft_inplace(v.data(), v.size());
    return v;
}
inline std::vector<F> f_coset_dft_small(std::vector<F> coeffs, F shift) {
    F s = f_one();
    for (auto& c : coeffs) {
        c *= s;
        s *= shift;
    }
    dft_inplace(coeffs.data(), coeffs.size());
    return coeffs;
}

// poly.rs:619-683: evals is row-major `height x width` (row = point index in D, col = coset),
// coset i is init * shift^i * D.  Returns height*width coefficients.
inline std::vector<EF> from_geometric_cosets_evals_idft(const std::vector<EF>& evals, size_t height, size_t width, F shift,
                                                        F init) {
    if (height == 0 || width == 0) return {};
    const int log_height = log2_strict(height);
    // iDFT within each coset, then unshift coefficient t by (init * shift^i)^-t
    std::vector<std::vector<EF>> rem(width);
    for (size_t i = 0; i < width; i++) {
        std::vector<EF> col(height);
        for (size_t r = 0; r < height; r++) col[r] = evals[r * width + i];
        rem[i] = ef_idft(col);
        const F s_inv = f_inv(init * f_pow(shift, i));
        F p = f_one();
        for (size_t t = 0; t < height; t++) {
            rem[i][t] = rem[i][t] * p;
            p *= s_inv;
        }
    }
    // interpolate across cosets at the points init^height * shift^(i*height)
    F base = shift, ib = init;
    for (int i = 0; i < log_height; i++) {
        base *= base;
        ib *= ib;
    }
    std::vector<F> pts(width);
    {
        F p = ib;
        for (size_t i = 0; i < width; i++) {
            pts[i] = p;
            p *= base;
        }
    }
    // Lagrange basis polynomials in coefficient form
    std::vector<std::vector<F>> basis(width, std::vector<F>(width, f_zero()));
    for (size_t i = 0; i < width; i++) {
        std::vector<F> poly{f_one()};
        F denom = f_one();
        for (size_t j = 0; j < width; j++) {
            if (j == i) continue;
            poly.push_back(f_zero());
            for (size_t k = poly.size() - 1; k >= 1; k--) poly[k] = poly[k - 1] - pts[j] * poly[k];
            poly[0] = -(pts[j] * poly[0]);
            denom *= pts[i] - pts[j];
        }
        const F inv = f_inv(denom);
        for (size_t k = 0; k < width; k++) basis[i][k] = poly[k] * inv;
    }
    std::vector<EF> coeffs(height * width, ef_zero());
    for (size_t t = 0; t < height; t++)
        for (size_t i = 0; i < width; i++)
            for (size_t k = 0; k < width; k++) coeffs[k * height + t] += rem[i][t] * basis[i][k];
    return coeffs;
}

// poly.rs:349-420 (points 0..len-1 in the callers here)
inline std::vector<EF> lagrange_interpolate(const std::vector<F>& points, const std::vector<EF>& evals) {
    const size_t len = points.size();
    std::vector<EF> coeffs(len, ef_zero());
    for (size_t i = 0; i < len; i++) {
        std::vector<F> poly{f_one()};
        F denom = f_one();
        for (size_t j = 0; j < len; j++) {
            if (j == i) continue;
            poly.push_back(f_zero());
            for (size_t k = poly.size() - 1; k >= 1; k--) poly[k] = poly[k - 1] - points[j] * poly[k];
            poly[0] = -(points[j] * poly[0]);
            denom *= points[i] - points[j];
        }
        const F inv = f_inv(denom);
        for (size_t k = 0; k < len; k++) coeffs[k] += evals[i] * (poly[k] * inv);
    }
    return coeffs;
}

// A (possibly rotated) base-field matrix part handed to the round-0 / fold routines
// (StridedColMajorMatrixView with stride 1 + the is_rot flag, sumcheck.rs:53,207).
struct MatPart {
    const F* values = nullptr;  // column-major, column stride = col_stride
    size_t height = 0, width = 0, col_stride = 0;
    bool is_rot = false;
    F at(size_t row, size_t col) const { return values[col * col_stride + row]; }
};

// sumcheck.rs:49-187.  w(z, x, rows) -> WD values, rows[m][col] = part m, column col at (z, x).
// Returns WD polynomials in coefficient form (d * 2^l_skip coefficients each).
template <size_t WD>
inline std::array<std::vector<EF>, WD> sumcheck_uni_round0_poly(
    int l_skip, int n, int d, const std::vector<MatPart>& mats,
    const std::function<std::array<EF, WD>(F, size_t, const std::vector<std::vector<F>>&)>& w) {
    std::array<std::vector<EF>, WD> out;
    if (d == 0) return out;
    const size_t N = size_t(1) << l_skip;
    const F g = f_generator(), omega_skip = two_adic_generator(l_skip);
    std::vector<F> shifts;
    {
        F s = g;
        for (int i = 0; i < d; i++) {
            shifts.push_back(s);
            s *= g;
        }
    }
    std::vector<std::array<EF, WD>> evals(N * d);
    for (auto& e : evals) e.fill(ef_zero());
    // the sum over x is split over host threads (per-worker partial sums; exact field additions commute)
    const size_t n_x = size_t(1) << n;
    const unsigned workers = par_workers(n_x, 4);
    std::vector<std::vector<std::array<EF, WD>>> partial(workers, evals);
    parallel_for_tid(n_x, workers, [&](unsigned wk, size_t x_begin, size_t x_end) {
    std::vector<std::array<EF, WD>>& evals = partial[wk];
    for (size_t x = x_begin; x < x_end; x++) {
        // mats_at_zs[m][col][coset * N + z_idx]
        std::vector<std::vector<std::vector<F>>> at(mats.size());
        for (size_t mi = 0; mi < mats.size(); mi++) {
            const MatPart& mat = mats[mi];
            const size_t off = mat.is_rot ? 1 : 0;
            at[mi].resize(mat.width);
            for (size_t c = 0; c < mat.width; c++) {
                std::vector<F> col_x(N);
                for (size_t i = 0; i < N; i++) col_x[i] = mat.at(((x << l_skip) + i + off) % mat.height, c);
                const std::vector<F> coeffs = f_idft_small(col_x);
                for (F sh : shifts) {
                    std::vector<F> ev = f_coset_dft_small(coeffs, sh);
                    at[mi][c].insert(at[mi][c].end(), ev.begin(), ev.end());
                }
            }
        }
        F z = f_one();
        for (size_t z_idx = 0; z_idx < N; z_idx++) {
            for (int coset = 0; coset < d; coset++) {
                const size_t z_int = ((size_t)coset << l_skip) + z_idx;
                std::vector<std::vector<F>> rows(mats.size());
                for (size_t mi = 0; mi < mats.size(); mi++) {
                    rows[mi].resize(mats[mi].width);
                    for (size_t c = 0; c < mats[mi].width; c++) rows[mi][c] = at[mi][c][z_int];
                }
                const std::array<EF, WD> v = w(shifts[coset] * z, x, rows);
                for (size_t k = 0; k < WD; k++) evals[z_idx * d + coset][k] += v[k];
            }
            z *= omega_skip;
        }
    }
    });
    for (const auto& pe : partial)
        for (size_t i = 0; i < N * d; i++)
            for (size_t k = 0; k < WD; k++) evals[i][k] += pe[i][k];
    for (size_t k = 0; k < WD; k++) {
        std::vector<EF> vals(N * d);
        for (size_t i = 0; i < N * d; i++) vals[i] = evals[i][k];
        out[k] = from_geometric_cosets_evals_idft(vals, N, d, g, g);
    }
    return out;
}

inline size_t sumcheck_round0_deg(int l_skip, size_t d) { return d * ((size_t(1) << l_skip) - 1); }

// sumcheck.rs:204-251: evaluate, per 2^l_skip chunk, the interpolant over D at r (barycentric
// result == the unique interpolant's value).  Returns column-major EF matrix of height
// max(height, 2^l_skip) >> l_skip.
inline std::vector<EF> fold_ple_evals(int l_skip, const MatPart& mat, EF r, size_t* new_height_out) {
    const size_t N = size_t(1) << l_skip;
    const size_t lifted = std::max(mat.height, N), new_height = lifted >> l_skip;
    const F omega = two_adic_generator(l_skip);
    // Lagrange coefficients L_i(r) = (w^i / N) * (r^N - 1) / (r - w^i)
    std::vector<EF> L(N);
    {
        EF rN = r;
        for (int i = 0; i < l_skip; i++) rN = rN * rN;
        const EF num = (rN - ef_one()) * f_inv(from_canonical(N));
        F wi = f_one();
        for (size_t i = 0; i < N; i++) {
            L[i] = num * ef_inv(r - ef_from(wi)) * wi;
            wi *= omega;
        }
    }
    const size_t off = mat.is_rot ? 1 : 0;
    std::vector<EF> out(new_height * mat.width);
    parallel_for(mat.width * new_height, [&](size_t b, size_t e) {
        for (size_t i = b; i < e; i++) {
            const size_t j = i / new_height, x = i % new_height;
            EF acc = ef_zero();
            for (size_t z = 0; z < N; z++) acc += L[z] * mat.at(((x << l_skip) + z + off) % mat.height, j);
            out[i] = acc;
        }
    }, 1024);
    *new_height_out = new_height;
    return out;
}

// A column-major EF matrix view
struct EfPart {
    const EF* values = nullptr;
    size_t height = 0, width = 0;
    EF at(size_t row, size_t col) const { return values[col * height + row]; }
};

// sumcheck.rs:271-393: s(X) for X = 1..d of sum_{y in H_{n-1}} W(T(X, y)).
template <size_t WD>
inline std::array<std::vector<EF>, WD> sumcheck_round_poly_evals(
    int n, int d, const std::vector<EfPart>& mats,
    const std::function<std::array<EF, WD>(EF, size_t, const std::vector<std::vector<EF>>&)>& w) {
    std::array<std::vector<EF>, WD> out;
    if (n == 0) {
        std::vector<std::vector<EF>> rows(mats.size());
        for (size_t mi = 0; mi < mats.size(); mi++)
            for (size_t c = 0; c < mats[mi].width; c++) rows[mi].push_back(mats[mi].at(0, c));
        const auto v = w(ef_one(), 0, rows);
        for (size_t k = 0; k < WD; k++) out[k].assign(d, v[k]);
        return out;
    }
    for (size_t k = 0; k < WD; k++) out[k].assign(d, ef_zero());
    const size_t n_y = size_t(1) << (n - 1);
    const unsigned workers = par_workers(n_y, 16);
    std::vector<std::array<std::vector<EF>, WD>> partial(workers, out);
    parallel_for_tid(n_y, workers, [&](unsigned wk, size_t y_begin, size_t y_end) {
        std::array<std::vector<EF>, WD>& acc = partial[wk];
        for (size_t y = y_begin; y < y_end; y++)
            for (int X = 1; X <= d; X++) {
                const EF xe = ef_from_u64((uint64_t)X);
                std::vector<std::vector<EF>> rows(mats.size());
                for (size_t mi = 0; mi < mats.size(); mi++)
                    for (size_t c = 0; c < mats[mi].width; c++) {
                        const EF t0 = mats[mi].at(2 * y, c), t1 = mats[mi].at(2 * y + 1, c);
                        rows[mi].push_back(t0 + (t1 - t0) * xe);
                    }
                const auto v = w(xe, y, rows);
                for (size_t k = 0; k < WD; k++) acc[k][X - 1] += v[k];
            }
    });
    for (const auto& pa : partial)
        for (size_t k = 0; k < WD; k++)
            for (int X = 0; X < d; X++) out[k][X] += pa[k][X];
    return out;
}

}  // namespace orc
