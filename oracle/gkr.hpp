// ORACLE — TEST INFRASTRUCTURE ONLY (see bb31.hpp header).
//
// LogUp-GKR fractional sumcheck, prover and verifier.  CPU restatement of
//   crates/stark-backend/src/prover/logup_zerocheck/fractional_sumcheck_gkr.rs:29-213   prover
//   crates/stark-backend/src/prover/sumcheck.rs:271-393   sumcheck_round_poly_evals (as used here)
//   crates/stark-backend/src/verifier/fractional_sumcheck_gkr.rs:49-234                 verifier
// The verifier is restated so that the oracle prover can be checked for self-consistency here,
// where the reference verifier cannot run (no Rust toolchain): PARITY UNPINNED against reference
// outputs, pinned only by "oracle verifier accepts oracle prover" + "CUDA prover == oracle prover".
#pragma once
#include <array>

#include "poly.hpp"
#include "transcript.hpp"

namespace orc {

struct Frac {
    EF p, q;
};
inline Frac frac_add(const Frac& a, const Frac& b) { return Frac{a.p * b.q + a.q * b.p, a.q * b.q}; }

struct GkrLayerClaims {
    EF p_xi_0, q_xi_0, p_xi_1, q_xi_1;
};
struct FracSumcheckProof {
    EF frac_sum_p, frac_sum_q;
    std::vector<GkrLayerClaims> claims_per_layer;
    std::vector<std::vector<std::array<EF, 3>>> sumcheck_polys;
};
struct NonZeroRootSum : std::runtime_error {
    NonZeroRootSum() : std::runtime_error("LogupZerocheckError::NonZeroRootSum") {}
};

// fractional_sumcheck_gkr.rs:60-213
inline FracSumcheckProof fractional_sumcheck(DuplexSponge& ts, const std::vector<Frac>& evals, bool assert_zero,
                                             std::vector<EF>* xi_out) {
    FracSumcheckProof proof;
    if (evals.empty()) {
        proof.frac_sum_p = ef_zero();
        proof.frac_sum_q = ef_one();
        xi_out->clear();
        return proof;
    }
    const int total_rounds = log2_strict(evals.size());
    std::vector<Frac> tree(size_t(2) << total_rounds);
    for (size_t i = 0; i < evals.size(); i++) tree[(size_t(1) << total_rounds) + i] = evals[i];
    for (int level = total_rounds - 1; level >= 0; level--) {  // same nodes as `for node = 2^n - 1 .. 1`, level by level
        const size_t first = size_t(1) << level;
        parallel_for(first, [&](size_t b, size_t e) {
            for (size_t node = first + b; node < first + e; node++) tree[node] = frac_add(tree[2 * node], tree[2 * node + 1]);
        }, 1024);
    }
    const Frac frac_sum = tree[1];
    if (assert_zero) {
        if (!ef_is_zero(frac_sum.p)) throw NonZeroRootSum();
    } else {
        ts.observe_ext(frac_sum.p);
    }
    ts.observe_ext(frac_sum.q);
    proof.frac_sum_p = frac_sum.p;
    proof.frac_sum_q = frac_sum.q;

    auto push_claims = [&](const GkrLayerClaims& c) {
        proof.claims_per_layer.push_back(c);
        ts.observe_ext(c.p_xi_0);
        ts.observe_ext(c.q_xi_0);
        ts.observe_ext(c.p_xi_1);
        ts.observe_ext(c.q_xi_1);
    };
    if (total_rounds == 0) {  // a single leaf: no layers (log2_strict(1) == 0)
        xi_out->clear();
        return proof;
    }
    push_claims(GkrLayerClaims{tree[2].p, tree[2].q, tree[3].p, tree[3].q});
    std::vector<EF> xi_prev{ts.sample_ext()};

    for (int round = 1; round < total_rounds; round++) {
        const size_t eval_size = size_t(1) << round;
        const EF lambda = ts.sample_ext();
        // columns p_j0, q_j0, p_j1, q_j1
        std::vector<EF> pq(4 * eval_size);
        const Frac* seg = &tree[2 * eval_size];
        parallel_for(eval_size, [&](size_t b, size_t e) {
            for (size_t x = b; x < e; x++) {
                pq[x] = seg[2 * x].p;
                pq[eval_size + x] = seg[2 * x].q;
                pq[2 * eval_size + x] = seg[2 * x + 1].p;
                pq[3 * eval_size + x] = seg[2 * x + 1].q;
            }
        }, 4096);
        size_t pq_h = eval_size, eq_h = eval_size;
        std::vector<EF> eq_xis = evals_eq_hypercube(xi_prev);
        std::vector<std::array<EF, 3>> round_polys;
        std::vector<EF> rho;
        for (int sr = 0; sr < round; sr++) {
            // s(X) at X in {1,2,3} of eq * (p0 q1 + p1 q0 + lambda q0 q1), summed over y in H_{n-1}
            std::array<EF, 3> s{ef_zero(), ef_zero(), ef_zero()};
            const size_t ny = pq_h / 2;
            const unsigned workers = par_workers(ny, 512);
            std::vector<std::array<EF, 3>> partial(workers, s);
            parallel_for_tid(ny, workers, [&](unsigned wk, size_t y_begin, size_t y_end) {
                std::array<EF, 3> acc{ef_zero(), ef_zero(), ef_zero()};
                for (size_t y = y_begin; y < y_end; y++)
                    for (int X = 1; X <= 3; X++) {
                        const F xf = from_canonical((uint64_t)X);
                        auto at = [&](const std::vector<EF>& m, size_t h, size_t col) {
                            const EF t0 = m[col * h + 2 * y], t1 = m[col * h + 2 * y + 1];
                            return t0 + (t1 - t0) * xf;
                        };
                        const EF eq = at(eq_xis, eq_h, 0);
                        const EF p0 = at(pq, pq_h, 0), q0 = at(pq, pq_h, 1), p1 = at(pq, pq_h, 2), q1 = at(pq, pq_h, 3);
                        acc[X - 1] += eq * ((p0 * q1 + p1 * q0) + lambda * (q0 * q1));
                    }
                partial[wk] = acc;
            });
            for (const auto& pa : partial)
                for (int X = 0; X < 3; X++) s[X] += pa[X];
            for (const EF& e : s) ts.observe_ext(e);
            round_polys.push_back(s);
            const EF r = ts.sample_ext();
            fold_mle_evals(pq, pq_h, 4, r);
            fold_mle_evals(eq_xis, eq_h, 1, r);
            rho.push_back(r);
        }
        push_claims(GkrLayerClaims{pq[0], pq[1], pq[2], pq[3]});
        const EF mu = ts.sample_ext();
        xi_prev.assign(1, mu);
        xi_prev.insert(xi_prev.end(), rho.begin(), rho.end());
        proof.sumcheck_polys.push_back(round_polys);
    }
    *xi_out = xi_prev;
    return proof;
}

// verifier/fractional_sumcheck_gkr.rs:49-149 (q0_claim = frac_sum_q; numerator asserted zero).
// Returns false on any failed check; outputs (p(xi), q(xi), xi).
inline bool verify_gkr(const FracSumcheckProof& proof, DuplexSponge& ts, int total_rounds, EF* numer, EF* denom,
                       std::vector<EF>* xi) {
    if ((int)proof.claims_per_layer.size() != total_rounds) return false;
    if ((int)proof.sumcheck_polys.size() != (total_rounds > 0 ? total_rounds - 1 : 0)) return false;
    ts.observe_ext(proof.frac_sum_q);
    auto observe = [&](const GkrLayerClaims& c) {
        ts.observe_ext(c.p_xi_0);
        ts.observe_ext(c.q_xi_0);
        ts.observe_ext(c.p_xi_1);
        ts.observe_ext(c.q_xi_1);
    };
    const GkrLayerClaims& c0 = proof.claims_per_layer[0];
    observe(c0);
    if (!ef_is_zero(c0.p_xi_0 * c0.q_xi_1 + c0.p_xi_1 * c0.q_xi_0)) return false;
    if (c0.q_xi_0 * c0.q_xi_1 != proof.frac_sum_q) return false;
    EF mu = ts.sample_ext();
    EF numer_claim = interpolate_linear_at_01(c0.p_xi_0, c0.p_xi_1, mu);
    EF denom_claim = interpolate_linear_at_01(c0.q_xi_0, c0.q_xi_1, mu);
    std::vector<EF> gkr_r{mu};
    for (int round = 1; round < total_rounds; round++) {
        const EF lambda = ts.sample_ext();
        EF claim = numer_claim + lambda * denom_claim;
        const auto& polys = proof.sumcheck_polys[round - 1];
        if ((int)polys.size() != round) return false;
        std::vector<EF> r_prime;
        EF eq = ef_one();
        for (int sr = 0; sr < round; sr++) {
            for (const EF& e : polys[sr]) ts.observe_ext(e);
            const EF ri = ts.sample_ext();
            r_prime.push_back(ri);
            const EF ev[4] = {claim - polys[sr][0], polys[sr][0], polys[sr][1], polys[sr][2]};
            claim = interpolate_cubic_at_0123(ev, ri);
            const EF x = gkr_r[sr];
            eq = eq * (x * ri + (ef_one() - x) * (ef_one() - ri));
        }
        const GkrLayerClaims& c = proof.claims_per_layer[round];
        observe(c);
        const EF pc = c.p_xi_0 * c.q_xi_1 + c.p_xi_1 * c.q_xi_0, qc = c.q_xi_0 * c.q_xi_1;
        if ((pc + lambda * qc) * eq != claim) return false;
        mu = ts.sample_ext();
        numer_claim = interpolate_linear_at_01(c.p_xi_0, c.p_xi_1, mu);
        denom_claim = interpolate_linear_at_01(c.q_xi_0, c.q_xi_1, mu);
        gkr_r.assign(1, mu);
        gkr_r.insert(gkr_r.end(), r_prime.begin(), r_prime.end());
    }
    *numer = numer_claim;
    *denom = denom_claim;
    *xi = gkr_r;
    return true;
}

}  // namespace orc
