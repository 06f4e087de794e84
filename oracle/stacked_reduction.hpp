// ORACLE — TEST INFRASTRUCTURE ONLY (see bb31.hpp header).
//
// Stacked opening reduction (batch sumcheck from trace-column openings to stacked-column
// openings), prover and verifier.  CPU restatement of
//   crates/stark-backend/src/prover/stacked_reduction.rs:67-127     prove_stacked_opening_reduction
//   crates/stark-backend/src/prover/stacked_reduction.rs:129-506    StackedReductionCpu
//   crates/stark-backend/src/verifier/stacked_reduction.rs:27-239   verify_stacked_reduction
// PARITY UNPINNED against reference outputs; pinned by the verifier restatement accepting the
// prover restatement and by CUDA == oracle (tests/test_stacked_reduction.py).
#pragma once
#include <map>

#include "sumcheck.hpp"
#include "transcript.hpp"

namespace orc {

struct StackingProof {
    std::vector<EF> univariate_round_coeffs;
    std::vector<std::array<EF, 2>> sumcheck_round_polys;
    std::vector<std::vector<EF>> stacking_openings;  // per commit, per stacked column
};

inline size_t rot_prev(size_t x, int n) { return x == 0 ? (size_t(1) << n) - 1 : x - 1; }

// stacked_reduction.rs:67-127 with StackedReductionCpu inlined.  r has n_max + 1 entries
// (r[0] is the univariate point); returns u (n_stack + 1 entries) through u_out.
inline StackingProof prove_stacked_opening_reduction(DuplexSponge& ts, int l_skip, int n_stack,
                                                     const std::vector<const StackedPcsData*>& commits,
                                                     const std::vector<std::vector<bool>>& need_rot_per_commit,
                                                     const std::vector<EF>& r, std::vector<EF>* u_out) {
    const EF lambda = ts.sample_ext();
    const F omega_skip = two_adic_generator(l_skip);
    struct View {
        size_t com_idx;
        StackedSlice slice;
        size_t lambda_eq_idx;
        long lambda_rot_idx;  // -1: none
    };
    std::vector<View> views;
    size_t lambda_idx = 0;
    for (size_t ci = 0; ci < commits.size(); ci++)
        for (const SortedCol& sc : commits[ci]->layout.sorted_cols) {
            View v{ci, sc.slice, lambda_idx, -1};
            lambda_idx++;
            if (need_rot_per_commit[ci][sc.mat_idx]) v.lambda_rot_idx = (long)lambda_idx;
            lambda_idx++;
            views.push_back(v);
        }
    std::vector<EF> lambda_pows(lambda_idx);
    {
        EF a = ef_one();
        for (auto& x : lambda_pows) {
            x = a;
            a = a * lambda;
        }
    }
    std::vector<size_t> ht_diff_idxs;
    std::map<int, std::vector<EF>> eq_r_per_lht, k_rot_r_per_lht;
    int last_height = -1;
    for (size_t i = 0; i < views.size(); i++) {
        const int lh = views[i].slice.log_height;
        const int n_lift = std::max(lh - l_skip, 0);
        if (i == 0 || lh != last_height) {
            ht_diff_idxs.push_back(i);
            last_height = lh;
        }
        if (!eq_r_per_lht.count(lh)) eq_r_per_lht[lh] = evals_eq_hypercube(std::vector<EF>(r.begin() + 1, r.begin() + 1 + n_lift));
    }
    ht_diff_idxs.push_back(views.size());
    const EF r_0 = r[0];
    const EF eq_const = eval_eq_uni_at_one(l_skip, r_0 * omega_skip);
    std::vector<EF> eq_ub_per_trace(views.size(), ef_one());

    // ---- round 0 -----------------------------------------------------------------------------
    const size_t s_0_deg = sumcheck_round0_deg(l_skip, 2);
    std::vector<EF> s_0(s_0_deg + 1, ef_zero());
    for (size_t wi = 0; wi + 1 < ht_diff_idxs.size(); wi++) {
        const size_t w0 = ht_diff_idxs[wi], w1 = ht_diff_idxs[wi + 1];
        const int log_height = views[w0].slice.log_height;
        const int n = log_height - l_skip;
        const int n_lift = std::max(n, 0);
        const std::vector<EF>& eq_rs = eq_r_per_lht[log_height];
        std::vector<MatPart> parts;
        for (size_t t = w0; t < w1; t++) {
            const ColMajor& q = commits[views[t].com_idx]->matrix;
            const StackedSlice& s = views[t].slice;
            MatPart p;
            p.values = q.col(s.col_idx) + s.row_idx;
            p.height = s.len(l_skip);
            p.width = 1;
            p.col_stride = 0;
            parts.push_back(p);
        }
        auto polys = sumcheck_uni_round0_poly<2>(
            l_skip, n_lift, 2, parts, [&](F zf, size_t x, const std::vector<std::vector<F>>& evals) {
                const EF z = ef_from(zf);
                const EF eq_cube = eq_rs[x];
                int l = l_skip;
                F omega = omega_skip;
                EF r_uni = r_0;
                if (n < 0) {
                    l = l_skip + n;
                    for (int i = 0; i < -n; i++) omega *= omega;
                    r_uni = ef_exp_power_of_2(r_0, -n);
                }
                const EF ind = eval_in_uni(l_skip, n, z);
                const EF eq_uni_r0 = eval_eq_uni(l, z, r_uni);
                const EF eq_uni_r0_rot = eval_eq_uni(l, z, r_uni * omega);
                const EF eq_uni_1 = eval_eq_uni_at_one(l_skip, z);
                const EF k_rot_cube = eq_rs[rot_prev(x, n_lift)];
                const EF eq = eq_uni_r0 * eq_cube;
                const EF k_rot = eq_uni_r0_rot * eq_cube + eq_const * eq_uni_1 * (k_rot_cube - eq_cube);
                std::array<EF, 2> acc{ef_zero(), ef_zero()};
                for (size_t i = 0; i < evals.size(); i++) {
                    const View& tv = views[w0 + i];
                    const F q = evals[i][0];
                    acc[0] += lambda_pows[tv.lambda_eq_idx] * eq * q * ind;
                    if (tv.lambda_rot_idx >= 0) acc[1] += lambda_pows[tv.lambda_rot_idx] * k_rot * q * ind;
                }
                return acc;
            });
        for (size_t i = 0; i <= s_0_deg; i++)
            for (int k = 0; k < 2; k++) s_0[i] += polys[k][i];
    }
    for (const EF& c : s_0) ts.observe_ext(c);
    std::vector<EF> u_vec{ts.sample_ext()};
    const EF u_0 = u_vec[0];

    // ---- fold_ple_evals ------------------------------------------------------------------------
    std::vector<std::vector<EF>> q_evals(commits.size());
    std::vector<size_t> q_h(commits.size());
    for (size_t ci = 0; ci < commits.size(); ci++) {
        const ColMajor& m = commits[ci]->matrix;
        MatPart p;
        p.values = m.values.data();
        p.height = m.height;
        p.width = m.width;
        p.col_stride = m.height;
        q_evals[ci] = fold_ple_evals(l_skip, p, u_0, &q_h[ci]);
    }
    {
        const EF eq_uni_u0r0 = eval_eq_uni(l_skip, u_0, r_0);
        const EF eq_uni_u0r0_rot = eval_eq_uni(l_skip, u_0, r_0 * omega_skip);
        const EF eq_uni_u01 = eval_eq_uni_at_one(l_skip, u_0);
        for (auto& kv : eq_r_per_lht) {
            const int log_height = kv.first;
            std::vector<EF>& mat = kv.second;
            const int n = log_height - l_skip, n_lift = std::max(n, 0);
            const EF ind = eval_in_uni(l_skip, n, u_0);
            EF eq_uni = eq_uni_u0r0, eq_uni_rot = eq_uni_u0r0_rot;
            if (n < 0) {
                F omega = omega_skip;
                for (int i = 0; i < -n; i++) omega *= omega;
                const EF rr = ef_exp_power_of_2(r_0, -n);
                eq_uni = eval_eq_uni(l_skip + n, u_0, rr);
                eq_uni_rot = eval_eq_uni(l_skip + n, u_0, rr * omega);
            }
            std::vector<EF> k(mat.size());
            for (size_t x = 0; x < mat.size(); x++) {
                const EF eq_cube = mat[x], k_rot_cube = mat[rot_prev(x, n_lift)];
                k[x] = ind * (eq_uni_rot * eq_cube + eq_const * eq_uni_u01 * (k_rot_cube - eq_cube));
            }
            for (auto& v : mat) v = v * (ind * eq_uni);
            k_rot_r_per_lht[log_height] = k;
        }
    }

    // ---- MLE rounds ------------------------------------------------------------------------------
    StackingProof proof;
    proof.univariate_round_coeffs = s_0;
    for (int round = 1; round <= n_stack; round++) {
        std::array<EF, 2> s{ef_zero(), ef_zero()};
        for (size_t wi = 0; wi + 1 < ht_diff_idxs.size(); wi++) {
            const size_t w0 = ht_diff_idxs[wi], w1 = ht_diff_idxs[wi + 1];
            const int log_height = views[w0].slice.log_height;
            const int n_lift = std::max(log_height - l_skip, 0);
            const int hypercube_dim = std::max(n_lift - round, 0);
            const std::vector<EF>& eq_rs = eq_r_per_lht[log_height];
            const std::vector<EF>& k_rot_rs = k_rot_r_per_lht[log_height];
            std::vector<EfPart> cols;
            for (size_t t = w0; t < w1; t++) {
                const StackedSlice& sl = views[t].slice;
                const size_t row_start = round <= n_lift ? (sl.row_idx >> log_height) << (hypercube_dim + 1)
                                                         : (sl.row_idx >> (l_skip + round)) << 1;
                EfPart p;
                p.values = q_evals[views[t].com_idx].data() + sl.col_idx * q_h[views[t].com_idx] + row_start;
                p.height = size_t(2) << hypercube_dim;
                p.width = 1;
                cols.push_back(p);
            }
            auto ev = sumcheck_round_poly_evals<2>(
                hypercube_dim + 1, 2, cols, [&](EF x, size_t y, const std::vector<std::vector<EF>>& evals) {
                    std::array<EF, 2> acc{ef_zero(), ef_zero()};
                    for (size_t i = 0; i < evals.size(); i++) {
                        const size_t t_idx = w0 + i;
                        const View& tv = views[t_idx];
                        const EF q = evals[i][0];
                        EF eq_ub = eq_ub_per_trace[t_idx];
                        EF eq, k_rot;
                        if (round > n_lift) {
                            const bool b = (tv.slice.row_idx >> (l_skip + round - 1)) & 1;
                            eq_ub = eq_ub * eval_eq_mle1(x, b);
                            eq = eq_rs[0] * eq_ub;
                            k_rot = k_rot_rs[0] * eq_ub;
                        } else {
                            const EF eq_r = eq_rs[2 * y] * (ef_one() - x) + eq_rs[2 * y + 1] * x;
                            const EF k_rot_r = k_rot_rs[2 * y] * (ef_one() - x) + k_rot_rs[2 * y + 1] * x;
                            eq = eq_r * eq_ub;
                            k_rot = k_rot_r * eq_ub;
                        }
                        acc[0] += lambda_pows[tv.lambda_eq_idx] * q * eq;
                        if (tv.lambda_rot_idx >= 0) acc[1] += lambda_pows[tv.lambda_rot_idx] * q * k_rot;
                    }
                    return acc;
                });
            for (int X = 0; X < 2; X++) s[X] += ev[0][X] + ev[1][X];
        }
        ts.observe_ext(s[0]);
        ts.observe_ext(s[1]);
        proof.sumcheck_round_polys.push_back(s);
        const EF u_round = ts.sample_ext();
        u_vec.push_back(u_round);
        // fold
        for (size_t ci = 0; ci < commits.size(); ci++) fold_mle_evals(q_evals[ci], q_h[ci], commits[ci]->matrix.width, u_round);
        for (auto& kv : eq_r_per_lht) {
            size_t h = kv.second.size();
            fold_mle_evals(kv.second, h, 1, u_round);
        }
        for (auto& kv : k_rot_r_per_lht) {
            size_t h = kv.second.size();
            fold_mle_evals(kv.second, h, 1, u_round);
        }
        for (size_t t = 0; t < views.size(); t++) {
            const int n_lift = std::max(views[t].slice.log_height - l_skip, 0);
            if (round > n_lift) {
                const bool b = (views[t].slice.row_idx >> (l_skip + round - 1)) & 1;
                eq_ub_per_trace[t] = eq_ub_per_trace[t] * eval_eq_mle1(u_round, b);
            }
        }
    }
    proof.stacking_openings = q_evals;
    for (auto& v : proof.stacking_openings)
        for (const EF& c : v) ts.observe_ext(c);
    *u_out = u_vec;
    return proof;
}

// verifier/stacked_reduction.rs:27-239.  t_claims: the (claim, rot claim) pairs in prover order
// (per commit, per sorted column) — the verifier derives this order from column_openings; here the
// caller passes them already ordered.
inline bool verify_stacked_reduction(DuplexSponge& ts, const StackingProof& proof, const std::vector<const StackedLayout*>& layouts,
                                     const std::vector<std::vector<bool>>& need_rot_per_commit, int l_skip, int n_stack,
                                     const std::vector<std::pair<EF, EF>>& t_claims, const std::vector<EF>& r,
                                     std::vector<EF>* u_out) {
    const size_t omega_order = size_t(1) << l_skip;
    size_t t_claims_len = 0;
    for (auto* l : layouts) t_claims_len += l->sorted_cols.size();
    if (t_claims.size() != t_claims_len) return false;
    if (proof.univariate_round_coeffs.size() != 2 * (omega_order - 1) + 1) return false;
    if ((int)proof.sumcheck_round_polys.size() != n_stack) return false;
    const EF lambda = ts.sample_ext();
    std::vector<EF> lambda_sqr_powers(t_claims_len);
    {
        EF a = ef_one();
        const EF l2 = lambda * lambda;
        for (auto& x : lambda_sqr_powers) {
            x = a;
            a = a * l2;
        }
    }
    EF s_0 = ef_zero();
    for (size_t i = 0; i < t_claims_len; i++) s_0 += (t_claims[i].first + t_claims[i].second * lambda) * lambda_sqr_powers[i];
    EF s_0_sum = ef_zero();
    for (size_t i = 0; i < proof.univariate_round_coeffs.size(); i += omega_order) s_0_sum += proof.univariate_round_coeffs[i];
    s_0_sum = s_0_sum * from_canonical(omega_order);
    if (s_0 != s_0_sum) return false;
    for (const EF& c : proof.univariate_round_coeffs) ts.observe_ext(c);
    std::vector<EF> u(n_stack + 1);
    u[0] = ts.sample_ext();
    EF claim = horner_eval(proof.univariate_round_coeffs, u[0]);
    for (int j = 1; j <= n_stack; j++) {
        const EF s1 = proof.sumcheck_round_polys[j - 1][0], s2 = proof.sumcheck_round_polys[j - 1][1];
        ts.observe_ext(s1);
        ts.observe_ext(s2);
        u[j] = ts.sample_ext();
        const EF ev[3] = {claim - s1, s1, s2};
        claim = interpolate_quadratic_at_012(ev, u[j]);
    }
    if (proof.stacking_openings.size() != layouts.size()) return false;
    EF final_sum = ef_zero();
    size_t lambda_i = 0;
    std::vector<std::vector<EF>> q_coeffs;
    for (size_t ci = 0; ci < layouts.size(); ci++) {
        std::vector<EF> coeffs(proof.stacking_openings[ci].size(), ef_zero());
        for (const SortedCol& sc : layouts[ci]->sorted_cols) {
            const StackedSlice& s = sc.slice;
            const bool need_rot = need_rot_per_commit[ci][sc.mat_idx];
            const int n = s.log_height - l_skip, n_lift = std::max(n, 0);
            std::vector<EF> b;
            for (int j = l_skip + n_lift; j < l_skip + n_stack; j++) b.push_back(((s.row_idx >> j) & 1) ? ef_one() : ef_zero());
            const EF eq_mle = eval_eq_mle(u.data() + n_lift + 1, b.data(), b.size());
            const EF ind = eval_in_uni(l_skip, n, u[0]);
            int l = l_skip;
            std::vector<EF> rs_n(r.begin(), r.begin() + n_lift + 1);
            if (n < 0) {
                l = l_skip + n;
                rs_n.assign(1, ef_exp_power_of_2(r[0], -n));
            }
            const EF eq_prism = eval_eq_uni(l, u[0], rs_n[0]) * eval_eq_mle(u.data() + 1, rs_n.data() + 1, n_lift);
            EF batched = lambda_sqr_powers[lambda_i] * eq_prism;
            if (need_rot) batched += lambda_sqr_powers[lambda_i] * lambda * eval_rot_kernel_prism(l, u.data(), rs_n.data(), n_lift + 1);
            coeffs[s.col_idx] += eq_mle * batched * ind;
            lambda_i++;
        }
        q_coeffs.push_back(coeffs);
    }
    for (size_t ci = 0; ci < layouts.size(); ci++)
        for (size_t j = 0; j < q_coeffs[ci].size(); j++) {
            ts.observe_ext(proof.stacking_openings[ci][j]);
            final_sum += q_coeffs[ci][j] * proof.stacking_openings[ci][j];
        }
    if (claim != final_sum) return false;
    *u_out = u;
    return true;
}

}  // namespace orc
