// ORACLE — TEST INFRASTRUCTURE ONLY (see bb31.hpp header).
//
// Batch constraint sumcheck: LogUp input layer + GKR, zerocheck / LogUp univariate round 0
// (constraint evaluation on cosets of the skip domain, quotient by the vanishing polynomial),
// and the front-loaded batched MLE rounds; prover and verifier.  CPU restatement of
//   crates/stark-backend/src/prover/logup_zerocheck/mod.rs:40-438     prove_zerocheck_and_logup
//   crates/stark-backend/src/prover/logup_zerocheck/cpu.rs:72-695     LogupZerocheckCpu
//   crates/stark-backend/src/prover/logup_zerocheck/single.rs:17-190  EvalHelper
//   crates/stark-backend/src/prover/logup_zerocheck/evaluator.rs      ProverConstraintEvaluator
//   crates/stark-backend/src/air_builders/symbolic/dag.rs:17-55       SymbolicExpressionNode / Dag
//   crates/stark-backend/src/air_builders/symbolic/symbolic_expression.rs:351-385  eval_nodes
//   crates/stark-backend/src/verifier/batch_constraints.rs:52-387     verify_zerocheck_and_logup
//   crates/stark-backend/src/verifier/evaluator.rs                    VerifierConstraintEvaluator
//   crates/stark-backend/src/lib.rs:82-93                             calculate_n_logup
// PARITY UNPINNED against reference outputs; pinned by the verifier restatement accepting the
// prover restatement and by CUDA == oracle (tests/test_batch_constraints.py).
#pragma once
#include <algorithm>

#include "gkr.hpp"
#include "sumcheck.hpp"

namespace orc {

enum : uint32_t {
    NODE_VAR_PREP = 0,   // a = column index, b = offset
    NODE_VAR_MAIN = 1,   // a = column index, b = offset, c = part index
    NODE_VAR_PUBLIC = 2, // a = index
    NODE_IS_FIRST = 3,
    NODE_IS_LAST = 4,
    NODE_IS_TRANSITION = 5,
    NODE_CONST = 6,  // a = Montgomery word
    NODE_ADD = 7,
    NODE_SUB = 8,
    NODE_NEG = 9,
    NODE_MUL = 10,
};
struct DagNode {
    uint32_t op, a, b, c;
};
struct Interaction {
    uint32_t count_node = 0, bus_index = 0;
    std::vector<uint32_t> message;
};
// One present AIR with its trace (AirProvingContext + the parts of the proving key the prover reads)
struct AirCtx {
    std::vector<DagNode> nodes;
    std::vector<uint32_t> constraint_idx;
    std::vector<Interaction> interactions;
    int constraint_degree = 0;
    bool need_rot = false;
    std::vector<F> public_values;
    const ColMajor* common_main = nullptr;
    std::vector<const ColMajor*> cached_mains;
    const ColMajor* preprocessed = nullptr;
    size_t height() const { return common_main->height; }
    // single.rs:38-70 view_mats: (prep[, prep_rot]), (cached_i[, rot])..., (common[, rot])
    std::vector<MatPart> view_mats() const {
        std::vector<MatPart> out;
        auto push = [&](const ColMajor* m) {
            MatPart p;
            p.values = m->values.data();
            p.height = m->height;
            p.width = m->width;
            p.col_stride = m->height;
            p.is_rot = false;
            out.push_back(p);
            if (need_rot) {
                p.is_rot = true;
                out.push_back(p);
            }
        };
        if (preprocessed) push(preprocessed);
        for (auto* c : cached_mains) push(c);
        push(common_main);
        return out;
    }
};

inline F lift_to(F x, F*) { return x; }
inline EF lift_to(F x, EF*) { return ef_from(x); }

// symbolic_expression.rs:351-385 with ProverConstraintEvaluator (evaluator.rs): row_parts[0] =
// [is_first, is_transition, is_last], then the view_mats parts.
template <class T>
inline std::vector<T> eval_nodes(const AirCtx& air, const std::vector<std::vector<T>>& row_parts) {
    const size_t stride = air.need_rot ? 2 : 1;
    const size_t prep_base = 1, main_base = 1 + (air.preprocessed ? stride : 0);
    std::vector<T> v;
    v.reserve(air.nodes.size());
    for (const DagNode& nd : air.nodes) {
        T x;
        switch (nd.op) {
            case NODE_VAR_PREP: x = row_parts[prep_base + nd.b][nd.a]; break;
            case NODE_VAR_MAIN: x = row_parts[main_base + nd.c * stride + nd.b][nd.a]; break;
            case NODE_VAR_PUBLIC: x = lift_to(air.public_values[nd.a], (T*)nullptr); break;
            case NODE_IS_FIRST: x = row_parts[0][0]; break;
            case NODE_IS_TRANSITION: x = row_parts[0][1]; break;
            case NODE_IS_LAST: x = row_parts[0][2]; break;
            case NODE_CONST: x = lift_to(F::raw(nd.a), (T*)nullptr); break;
            case NODE_ADD: x = v[nd.a] + v[nd.b]; break;
            case NODE_SUB: x = v[nd.a] - v[nd.b]; break;
            case NODE_NEG: x = -v[nd.a]; break;
            case NODE_MUL: x = v[nd.a] * v[nd.b]; break;
            default: throw std::runtime_error("bad node");
        }
        v.push_back(x);
    }
    return v;
}
inline EF ef_times(EF a, F b) { return a * b; }
inline EF ef_times(EF a, EF b) { return a * b; }
// single.rs:77-89
template <class T>
inline EF acc_constraints(const AirCtx& air, const std::vector<std::vector<T>>& row_parts, const std::vector<EF>& lambda_pows) {
    const std::vector<T> nodes = eval_nodes(air, row_parts);
    EF acc = ef_zero();
    for (size_t k = 0; k < air.constraint_idx.size() && k < lambda_pows.size(); k++)
        acc += ef_times(lambda_pows[k], nodes[air.constraint_idx[k]]);
    return acc;
}
// single.rs:121-146: (numer, denom without alpha) per interaction
template <class T>
inline std::vector<std::pair<T, EF>> eval_interactions(const AirCtx& air, const std::vector<std::vector<T>>& row_parts,
                                                       const std::vector<EF>& beta_pows) {
    const std::vector<T> nodes = eval_nodes(air, row_parts);
    std::vector<std::pair<T, EF>> out;
    for (const Interaction& it : air.interactions) {
        const size_t len = it.message.size();
        EF denom = beta_pows[len] * from_canonical((uint64_t)it.bus_index + 1);
        for (size_t j = 0; j < len; j++) denom += ef_times(beta_pows[j], nodes[it.message[j]]);
        out.push_back({nodes[it.count_node], denom});
    }
    return out;
}
// single.rs:96-119
template <class T>
inline std::array<EF, 2> acc_interactions(const AirCtx& air, const std::vector<std::vector<T>>& row_parts,
                                          const std::vector<EF>& beta_pows, const std::vector<EF>& eq_3bs) {
    auto ev = eval_interactions(air, row_parts, beta_pows);
    EF numer = ef_zero(), denom = ef_zero();
    for (size_t i = 0; i < ev.size() && i < eq_3bs.size(); i++) {
        numer += ef_times(eq_3bs[i], ev[i].first);
        denom += eq_3bs[i] * ev[i].second;
    }
    return {numer, denom};
}

inline int calculate_n_logup(int l_skip, uint64_t total_interactions) {
    if (total_interactions == 0) return 0;
    int bits = 0;
    while (total_interactions >> bits) bits++;
    return bits - l_skip;
}

struct GkrProof {
    F logup_pow_witness;
    EF q0_claim;
    std::vector<GkrLayerClaims> claims_per_layer;
    std::vector<std::vector<std::array<EF, 3>>> sumcheck_polys;
};
struct BatchConstraintProof {
    std::vector<EF> numerator_term_per_air, denominator_term_per_air, univariate_round_coeffs;
    std::vector<std::vector<EF>> sumcheck_round_polys;
    std::vector<std::vector<std::vector<EF>>> column_openings;  // per air, per part (common main first), flat
};

// mod.rs:40-438 with LogupZerocheckCpu (cpu.rs) inlined.  airs sorted by descending height.
inline void prove_zerocheck_and_logup(DuplexSponge& ts, int l_skip, int max_constraint_degree, int logup_pow_bits,
                                      const std::vector<AirCtx>& airs, GkrProof* gkr_out, BatchConstraintProof* bc_out,
                                      std::vector<EF>* r_out) {
    const int constraint_degree = max_constraint_degree;
    const size_t num_traces = airs.size();
    const size_t N = size_t(1) << l_skip;
    const int n_max = std::max(log2_strict(airs[0].height()) - l_skip, 0);
    std::vector<std::pair<size_t, int>> interactions_meta;
    uint64_t total_interactions = 0;
    for (const AirCtx& a : airs) {
        const int lh = log2_strict(a.height()), llh = std::max(lh, l_skip);
        total_interactions += (uint64_t)a.interactions.size() << llh;
        interactions_meta.push_back({a.interactions.size(), llh});
    }
    const int n_logup = calculate_n_logup(l_skip, total_interactions);
    const StackedLayout ilayout = make_stacked_layout(0, l_skip + n_logup, interactions_meta);

    const F logup_pow_witness = ts.grind(logup_pow_bits);
    const EF alpha = ts.sample_ext(), beta = ts.sample_ext();
    size_t max_len = 0;
    size_t max_num_constraints = 0;
    for (const AirCtx& a : airs) {
        for (auto& it : a.interactions) max_len = std::max(max_len, it.message.size());
        max_num_constraints = std::max(max_num_constraints, a.constraint_idx.size());
    }
    std::vector<EF> beta_pows(max_len + 1);
    {
        EF p = ef_one();
        for (auto& x : beta_pows) {
            x = p;
            p = p * beta;
        }
    }
    std::vector<int> n_per_trace;
    for (const AirCtx& a : airs) n_per_trace.push_back(log2_strict(a.height()) - l_skip);

    // ---- GKR input layer (mod.rs:103-168) ------------------------------------------------------
    std::vector<Frac> gkr_input;
    if (!ilayout.sorted_cols.empty()) {
        std::vector<std::vector<std::vector<std::pair<F, EF>>>> unstacked(num_traces);
        for (size_t t = 0; t < num_traces; t++) {
            const AirCtx& a = airs[t];
            const std::vector<MatPart> mats = a.view_mats();
            const size_t height = a.height();
            unstacked[t].resize(height);
            parallel_for(height, [&](size_t i_begin, size_t i_end) {
                for (size_t i = i_begin; i < i_end; i++) {
                    std::vector<std::vector<F>> row_parts;
                    row_parts.push_back({i == 0 ? f_one() : f_zero(), i != height - 1 ? f_one() : f_zero(),
                                         i == height - 1 ? f_one() : f_zero()});
                    for (const MatPart& m : mats) {
                        std::vector<F> row(m.width);
                        for (size_t j = 0; j < m.width; j++) row[j] = m.at((i + (m.is_rot ? 1 : 0)) % height, j);
                        row_parts.push_back(row);
                    }
                    unstacked[t][i] = eval_interactions(a, row_parts, beta_pows);
                }
            }, 16);
        }
        gkr_input.assign(size_t(1) << (l_skip + n_logup), Frac{ef_zero(), ef_zero()});
        for (const SortedCol& sc : ilayout.sorted_cols) {
            const auto& pq = unstacked[sc.mat_idx];
            const size_t height = pq.size(), len = sc.slice.len(0);
            const F norm = f_inv(from_canonical(len / height));
            for (size_t off = 0; off < len; off += height)
                for (size_t i = 0; i < height; i++) {
                    Frac& f = gkr_input[sc.slice.row_idx + off + i];
                    f.p = ef_from(pq[i][sc.col_in_mat].first * norm);
                    f.q = pq[i][sc.col_in_mat].second;
                }
        }
        for (auto& f : gkr_input) f.q += alpha;
    }
    std::vector<EF> xi;
    FracSumcheckProof fsp = fractional_sumcheck(ts, gkr_input, true, &xi);
    const int n_global = std::max(n_max, n_logup);
    while ((int)xi.size() != l_skip + n_global) xi.push_back(ts.sample_ext());

    // ---- batch sumcheck ---------------------------------------------------------------------------
    const EF lambda = ts.sample_ext();
    std::vector<EF> lambda_pows(max_num_constraints);
    {
        EF p = ef_one();
        for (auto& x : lambda_pows) {
            x = p;
            p = p * lambda;
        }
    }
    const F omega_skip = two_adic_generator(l_skip);
    std::vector<F> omega_skip_pows(N);
    {
        F p = f_one();
        for (auto& x : omega_skip_pows) {
            x = p;
            p *= omega_skip;
        }
    }
    // eq_3b per trace (cpu.rs:247-283)
    std::vector<std::vector<EF>> eq_3b_per_trace(num_traces);
    for (size_t t = 0; t < num_traces; t++) {
        const int n_lift = std::max(n_per_trace[t], 0);
        for (size_t i = 0; i < airs[t].interactions.size(); i++) {
            size_t stacked_idx = 0;
            bool found = false;
            for (const SortedCol& sc : ilayout.sorted_cols)
                if (sc.mat_idx == t && sc.col_in_mat == i) {
                    stacked_idx = sc.slice.row_idx;
                    found = true;
                }
            if (!found) throw std::runtime_error("InteractionsLayoutMissing");
            size_t b_int = stacked_idx >> (l_skip + n_lift);
            std::vector<EF> b(n_logup - n_lift);
            for (auto& x : b) {
                x = (b_int & 1) ? ef_one() : ef_zero();
                b_int >>= 1;
            }
            eq_3b_per_trace[t].push_back(eval_eq_mle(xi.data() + l_skip + n_lift, b.data(), b.size()));
        }
    }
    // eq_xi trees (cpu.rs:289-299) and selector matrices (cpu.rs:306-322)
    std::vector<std::vector<EF>> eq_xi_per_trace(num_traces);
    std::vector<ColMajor> sels_base(num_traces);
    for (size_t t = 0; t < num_traces; t++) {
        const int n_lift = std::max(n_per_trace[t], 0);
        std::vector<EF> rev(xi.begin() + l_skip, xi.begin() + l_skip + n_lift);
        std::reverse(rev.begin(), rev.end());
        eq_xi_per_trace[t] = evals_eq_hypercubes(n_lift, rev.begin(), rev.end());
        const size_t height = airs[t].height(), lifted = std::max(height, N);
        ColMajor m(lifted, 3);
        for (size_t i = 0; i < lifted; i++) m.values[lifted + i] = f_one();
        for (size_t i = 0; i < lifted; i += height) {
            m.values[i] = f_one();
            m.values[lifted + i + height - 1] = f_zero();
            m.values[2 * lifted + i + height - 1] = f_one();
        }
        sels_base[t] = m;
    }
    auto sels_part = [&](size_t t) {
        MatPart p;
        p.values = sels_base[t].values.data();
        p.height = sels_base[t].height;
        p.width = 3;
        p.col_stride = sels_base[t].height;
        return p;
    };
    // round 0 polynomials: logup (numer, denom) per trace, then zerocheck per trace (cpu.rs:324-424)
    std::vector<std::vector<EF>> sp_0_polys(3 * num_traces);
    for (size_t t = 0; t < num_traces; t++) {
        const AirCtx& a = airs[t];
        const int n_lift = std::max(n_per_trace[t], 0);
        const EF* eq_xi = eq_xi_per_trace[t].data() + ((size_t(1) << n_lift) - 1);
        std::vector<MatPart> parts{sels_part(t)};
        for (auto& m : a.view_mats()) parts.push_back(m);
        const int cd = a.constraint_degree;
        if (cd > 0) {
            auto q = sumcheck_uni_round0_poly<1>(l_skip, n_lift, cd - 1, parts,
                                                 [&](F z, size_t x, const std::vector<std::vector<F>>& rows) {
                                                     const EF ce = acc_constraints(a, rows, lambda_pows);
                                                     F zN = z;
                                                     for (int i = 0; i < l_skip; i++) zN *= zN;
                                                     return std::array<EF, 1>{eq_xi[x] * ce * f_inv(zN - f_one())};
                                                 })[0];
            const size_t deg = sumcheck_round0_deg(l_skip, cd);
            std::vector<EF> coeffs(deg + 1);
            for (size_t i = 0; i <= deg; i++) {
                EF c = i < q.size() ? -q[i] : ef_zero();
                if (i >= N) c += q[i - N];  // q has (cd-1)*N coefficients; i - N < q.size() by degree
                coeffs[i] = c;
            }
            sp_0_polys[2 * num_traces + t] = coeffs;
        }
        if (!a.interactions.empty()) {
            const F norm = f_inv(from_canonical(size_t(1) << std::max(l_skip - log2_strict(a.height()), 0)));
            auto nd = sumcheck_uni_round0_poly<2>(l_skip, n_lift, cd, parts,
                                                  [&](F, size_t x, const std::vector<std::vector<F>>& rows) {
                                                      auto v = acc_interactions(a, rows, beta_pows, eq_3b_per_trace[t]);
                                                      return std::array<EF, 2>{eq_xi[x] * v[0], eq_xi[x] * v[1]};
                                                  });
            for (auto& p : nd[0]) p = p * norm;
            sp_0_polys[2 * t] = nd[0];
            sp_0_polys[2 * t + 1] = nd[1];
        }
    }
    const size_t sp_0_deg = sumcheck_round0_deg(l_skip, constraint_degree);
    const int s_deg = constraint_degree + 1;
    const size_t s_0_deg = sumcheck_round0_deg(l_skip, s_deg);
    size_t large = 1;
    while (large < s_0_deg + 1) large <<= 1;
    auto poly_mul_trunc = [&](std::vector<EF> a, std::vector<EF> b) {
        // product via evaluations on the size-`large` domain (mod.rs:208-236): exact when deg(a*b) < large
        a.resize(large, ef_zero());
        b.resize(large, ef_zero());
        std::vector<EF> ea = ef_dft(a), eb = ef_dft(b);
        for (size_t i = 0; i < large; i++) ea[i] = ea[i] * eb[i];
        return ef_idft(ea);
    };
    // logup: s_0 = eq_sharp_uni * sp_0 per (trace, numer/denom)
    std::vector<EF> eq_sharp_coeffs;
    {
        std::vector<EF> x1(xi.begin(), xi.begin() + l_skip);
        eq_sharp_coeffs = ef_idft(evals_eq_hypercube(x1));
    }
    std::vector<std::vector<EF>> s_0_logup(2 * num_traces);
    for (size_t i = 0; i < 2 * num_traces; i++) {
        std::vector<EF> c = sp_0_polys[i];
        c.resize(std::min(c.size(), sp_0_deg + 1));
        s_0_logup[i] = poly_mul_trunc(eq_sharp_coeffs, c);
    }
    BatchConstraintProof bc;
    const F skip_domain_size = from_canonical(N);
    for (size_t t = 0; t < num_traces; t++) {
        EF sums[2];
        for (int d = 0; d < 2; d++) {
            EF s = ef_zero();
            for (size_t j = 0; j <= s_0_deg; j += N) s += s_0_logup[2 * t + d][j];
            sums[d] = s * skip_domain_size;
        }
        ts.observe_ext(sums[0]);
        ts.observe_ext(sums[1]);
        bc.numerator_term_per_air.push_back(sums[0]);
        bc.denominator_term_per_air.push_back(sums[1]);
    }
    const EF mu = ts.sample_ext();
    std::vector<EF> mu_pows(3 * num_traces);
    {
        EF p = ef_one();
        for (auto& x : mu_pows) {
            x = p;
            p = p * mu;
        }
    }
    std::vector<EF> s_0_zc;
    {
        std::vector<EF> sp(large, ef_zero());
        for (size_t j = 0; j <= sp_0_deg; j++)
            for (size_t t = 0; t < num_traces; t++) {
                const auto& poly = sp_0_polys[2 * num_traces + t];
                if (j < poly.size()) sp[j] += mu_pows[2 * num_traces + t] * poly[j];
            }
        s_0_zc = poly_mul_trunc(eq_uni_poly(l_skip, xi[0]), sp);
    }
    std::vector<EF> s_0_poly(s_0_deg + 1);
    for (size_t j = 0; j <= s_0_deg; j++) {
        EF c = s_0_zc[j];
        for (size_t i = 0; i < 2 * num_traces; i++) c += mu_pows[i] * s_0_logup[i][j];
        ts.observe_ext(c);
        s_0_poly[j] = c;
    }
    std::vector<EF> r{ts.sample_ext()};
    const EF r_0 = r[0];
    EF prev_s_eval = horner_eval(s_0_poly, r_0);

    // ---- fold_ple_evals (cpu.rs:430-458) -----------------------------------------------------------
    std::vector<std::vector<std::vector<EF>>> mat_evals(num_traces);  // per trace, per view mat: col-major EF
    std::vector<std::vector<size_t>> mat_w(num_traces);
    std::vector<size_t> mat_h(num_traces);
    std::vector<std::vector<EF>> sels(num_traces);
    for (size_t t = 0; t < num_traces; t++) {
        size_t h = 0;
        for (const MatPart& m : airs[t].view_mats()) {
            mat_evals[t].push_back(fold_ple_evals(l_skip, m, r_0, &h));
            mat_w[t].push_back(m.width);
        }
        sels[t] = fold_ple_evals(l_skip, sels_part(t), r_0, &h);
        mat_h[t] = h;
    }
    std::vector<EF> eq_ns{eval_eq_uni(l_skip, xi[0], r_0)};
    std::vector<EF> eq_sharp_ns{eval_eq_sharp_uni(omega_skip_pows, xi.data(), l_skip, r_0)};
    for (auto& eq : eq_xi_per_trace)
        if (eq.size() > 1) eq.resize(eq.size() / 2);

    std::vector<EF> zerocheck_tilde(num_traces, ef_zero());
    std::vector<std::array<EF, 2>> logup_tilde(num_traces, std::array<EF, 2>{ef_zero(), ef_zero()});
    // ---- MLE rounds (mod.rs:314-395, cpu.rs:463-642) --------------------------------------------------
    for (int round = 1; round <= n_max; round++) {
        const EF r_prev = r[round - 1];
        const int sp_deg = constraint_degree;
        const EF eq_r_acc = eq_ns.back(), eq_sharp_r_acc = eq_sharp_ns.back();
        std::vector<std::vector<EF>> sp_evals(3 * num_traces);
        for (size_t t = 0; t < num_traces; t++) {
            const AirCtx& a = airs[t];
            const int n_lift = std::max(n_per_trace[t], 0);
            std::vector<EfPart> parts;
            parts.push_back(EfPart{sels[t].data(), mat_h[t], 3});
            for (size_t mi = 0; mi < mat_evals[t].size(); mi++) parts.push_back(EfPart{mat_evals[t][mi].data(), mat_h[t], mat_w[t][mi]});
            auto row0 = [&]() {
                std::vector<std::vector<EF>> rows;
                for (const EfPart& p : parts) {
                    std::vector<EF> row(p.width);
                    for (size_t c = 0; c < p.width; c++) row[c] = p.at(0, c);
                    rows.push_back(row);
                }
                return rows;
            };
            const F norm = f_inv(from_canonical(size_t(1) << std::max(-n_per_trace[t], 0)));
            // zerocheck
            if (round > n_lift) {
                if (round == n_lift + 1)
                    zerocheck_tilde[t] = eq_r_acc * acc_constraints(a, row0(), lambda_pows);
                else
                    zerocheck_tilde[t] = zerocheck_tilde[t] * r_prev;
                sp_evals[2 * num_traces + t] = {zerocheck_tilde[t]};
            } else {
                const int log_num_y = n_lift - round;
                const EF* eq_xi = eq_xi_per_trace[t].data() + ((size_t(1) << log_num_y) - 1);
                sp_evals[2 * num_traces + t] = sumcheck_round_poly_evals<1>(
                    log_num_y + 1, sp_deg, parts, [&](EF, size_t y, const std::vector<std::vector<EF>>& rows) {
                        return std::array<EF, 1>{eq_xi[y] * acc_constraints(a, rows, lambda_pows)};
                    })[0];
            }
            // logup
            if (a.interactions.empty()) {
                sp_evals[2 * t].assign(sp_deg, ef_zero());
                sp_evals[2 * t + 1].assign(sp_deg, ef_zero());
            } else if (round > n_lift) {
                if (round == n_lift + 1) {
                    auto v = acc_interactions(a, row0(), beta_pows, eq_3b_per_trace[t]);
                    logup_tilde[t] = {eq_sharp_r_acc * v[0] * norm, eq_sharp_r_acc * v[1]};
                } else {
                    logup_tilde[t][0] = logup_tilde[t][0] * r_prev;
                    logup_tilde[t][1] = logup_tilde[t][1] * r_prev;
                }
                sp_evals[2 * t] = {logup_tilde[t][0]};
                sp_evals[2 * t + 1] = {logup_tilde[t][1]};
            } else {
                const int log_num_y = n_lift - round;
                const EF* eq_xi = eq_xi_per_trace[t].data() + ((size_t(1) << log_num_y) - 1);
                auto nd = sumcheck_round_poly_evals<2>(log_num_y + 1, sp_deg, parts,
                                                       [&](EF, size_t y, const std::vector<std::vector<EF>>& rows) {
                                                           auto v = acc_interactions(a, rows, beta_pows, eq_3b_per_trace[t]);
                                                           return std::array<EF, 2>{eq_xi[y] * v[0], eq_xi[y] * v[1]};
                                                       });
                for (auto& p : nd[0]) p = p * norm;
                sp_evals[2 * t] = nd[0];
                sp_evals[2 * t + 1] = nd[1];
            }
        }
        size_t tail_start = num_traces;
        for (size_t t = 0; t < num_traces; t++)
            if (round > n_per_trace[t]) {
                tail_start = t;
                break;
            }
        std::vector<EF> sp_head_zc(constraint_degree, ef_zero()), sp_head_logup(constraint_degree, ef_zero());
        EF sp_tail = ef_zero();
        for (size_t t = 0; t < num_traces; t++) {
            const size_t zc = 2 * num_traces + t, nu = 2 * t, de = nu + 1;
            if (t < tail_start) {
                for (int i = 0; i < constraint_degree; i++) {
                    sp_head_zc[i] += mu_pows[zc] * sp_evals[zc][i];
                    sp_head_logup[i] += mu_pows[nu] * sp_evals[nu][i] + mu_pows[de] * sp_evals[de][i];
                }
            } else {
                sp_tail += mu_pows[zc] * sp_evals[zc][0] + mu_pows[nu] * sp_evals[nu][0] + mu_pows[de] * sp_evals[de][0];
            }
        }
        std::vector<EF> sp_head_evals(s_deg, ef_zero());
        for (int i = 0; i < constraint_degree; i++)
            sp_head_evals[i + 1] = eq_ns[round - 1] * sp_head_zc[i] + eq_sharp_ns[round - 1] * sp_head_logup[i];
        const EF xi_cur = xi[l_skip + round - 1];
        sp_head_evals[0] = (prev_s_eval - xi_cur * sp_head_evals[1] - sp_tail) * ef_inv(ef_one() - xi_cur);
        std::vector<F> pts;
        for (int i = 0; i < s_deg; i++) pts.push_back(from_canonical((uint64_t)i));
        std::vector<EF> coeffs = lagrange_interpolate(pts, sp_head_evals);
        coeffs.push_back(ef_zero());
        {
            const EF b = ef_one() - xi_cur, a = xi_cur - b;
            for (int i = s_deg - 1; i >= 0; i--) coeffs[i + 1] = a * coeffs[i] + b * coeffs[i + 1];
            coeffs[0] = coeffs[0] * b;
            coeffs[1] += sp_tail;
        }
        std::vector<EF> batch_s_evals;
        for (int i = 1; i <= s_deg; i++) {
            const EF e = horner_eval(coeffs, ef_from_u64((uint64_t)i));
            ts.observe_ext(e);
            batch_s_evals.push_back(e);
        }
        bc.sumcheck_round_polys.push_back(batch_s_evals);
        const EF r_round = ts.sample_ext();
        r.push_back(r_round);
        prev_s_eval = horner_eval(coeffs, r_round);
        // fold (cpu.rs:582-597)
        for (size_t t = 0; t < num_traces; t++) {
            size_t h = mat_h[t];
            for (size_t mi = 0; mi < mat_evals[t].size(); mi++) {
                h = mat_h[t];
                fold_mle_evals(mat_evals[t][mi], h, mat_w[t][mi], r_round);
            }
            h = mat_h[t];
            fold_mle_evals(sels[t], h, 3, r_round);
            mat_h[t] = h;
            if (eq_xi_per_trace[t].size() > 1) eq_xi_per_trace[t].resize(eq_xi_per_trace[t].size() / 2);
        }
        const EF eq_r = eval_eq_mle(&xi_cur, &r_round, 1);
        eq_ns.push_back(eq_ns[round - 1] * eq_r);
        eq_sharp_ns.push_back(eq_sharp_ns[round - 1] * eq_r);
    }
    // ---- column openings (cpu.rs:644-694) + transcript order (mod.rs:404-421) -------------------------
    for (size_t t = 0; t < num_traces; t++) {
        const AirCtx& a = airs[t];
        std::vector<std::vector<EF>> m = mat_evals[t];  // each has height 1 now
        std::vector<size_t> w = mat_w[t];
        std::vector<std::vector<EF>> openings;
        auto take_pair = [&](size_t i0) {
            std::vector<EF> v;
            if (a.need_rot) {
                for (size_t c = 0; c < w[i0]; c++) {
                    v.push_back(m[i0][c]);
                    v.push_back(m[i0 + 1][c]);
                }
            } else {
                v = m[i0];
            }
            return v;
        };
        const size_t stride = a.need_rot ? 2 : 1;
        openings.push_back(take_pair(m.size() - stride));  // common main first
        for (size_t i = 0; i + stride < m.size(); i += stride) openings.push_back(take_pair(i));
        bc.column_openings.push_back(openings);
    }
    auto observe_part = [&](const std::vector<EF>& part, bool need_rot) {
        if (need_rot) {
            for (const EF& e : part) ts.observe_ext(e);  // (claim, claim_rot) interleaved already
        } else {
            for (const EF& e : part) {
                ts.observe_ext(e);
                ts.observe_ext(ef_zero());
            }
        }
    };
    for (size_t t = 0; t < num_traces; t++) observe_part(bc.column_openings[t][0], airs[t].need_rot);
    for (size_t t = 0; t < num_traces; t++)
        for (size_t p = 1; p < bc.column_openings[t].size(); p++) observe_part(bc.column_openings[t][p], airs[t].need_rot);
    bc.univariate_round_coeffs = s_0_poly;
    gkr_out->logup_pow_witness = logup_pow_witness;
    gkr_out->q0_claim = fsp.frac_sum_q;
    gkr_out->claims_per_layer = fsp.claims_per_layer;
    gkr_out->sumcheck_polys = fsp.sumcheck_polys;
    *bc_out = bc;
    *r_out = r;
}

}  // namespace orc

namespace orc {

// verifier/evaluator.rs:14-24
inline EF progression_exp_2(EF m, int l) {
    EF pow = m, sum = ef_one();
    for (int i = 0; i < l; i++) {
        sum = sum * (ef_one() + pow);
        pow = pow * pow;
    }
    return sum;
}

// verifier/batch_constraints.rs:52-387.  vk-side data is taken from `airs` (nodes, interactions,
// need_rot, preprocessed != nullptr, cached count, public values); traces themselves are not read.
inline bool verify_zerocheck_and_logup(DuplexSponge& ts, int l_skip, int max_constraint_degree, int logup_pow_bits,
                                       const std::vector<AirCtx>& airs, const std::vector<int>& n_per_trace,
                                       const GkrProof& gkr, const BatchConstraintProof& bc, std::vector<EF>* r_out) {
    const size_t num_traces = airs.size();
    const size_t N = size_t(1) << l_skip;
    if (bc.numerator_term_per_air.size() != num_traces || bc.denominator_term_per_air.size() != num_traces) return false;
    if (!ts.check_witness(logup_pow_bits, gkr.logup_pow_witness)) return false;
    const EF alpha = ts.sample_ext(), beta = ts.sample_ext();
    uint64_t total_interactions = 0;
    for (size_t t = 0; t < num_traces; t++)
        total_interactions += (uint64_t)airs[t].interactions.size() << (l_skip + std::max(n_per_trace[t], 0));
    const int n_logup = calculate_n_logup(l_skip, total_interactions);
    std::vector<EF> xi;
    EF p_xi_claim = ef_zero(), q_xi_claim = alpha;
    if (total_interactions > 0) {
        FracSumcheckProof fsp;
        fsp.frac_sum_p = ef_zero();
        fsp.frac_sum_q = gkr.q0_claim;
        fsp.claims_per_layer = gkr.claims_per_layer;
        fsp.sumcheck_polys = gkr.sumcheck_polys;
        if (!verify_gkr(fsp, ts, l_skip + n_logup, &p_xi_claim, &q_xi_claim, &xi)) return false;
    } else if (gkr.q0_claim != ef_one()) {
        return false;
    }
    int n_max = 0;
    for (int n : n_per_trace) n_max = std::max(n_max, n);
    const int n_global = std::max(n_max, n_logup);
    while ((int)xi.size() != l_skip + n_global) xi.push_back(ts.sample_ext());
    const EF lambda = ts.sample_ext();
    for (size_t t = 0; t < num_traces; t++) {
        p_xi_claim -= bc.numerator_term_per_air[t];
        q_xi_claim -= bc.denominator_term_per_air[t];
        ts.observe_ext(bc.numerator_term_per_air[t]);
        ts.observe_ext(bc.denominator_term_per_air[t]);
    }
    if (!ef_is_zero(p_xi_claim)) return false;
    if (q_xi_claim != alpha) return false;
    const EF mu = ts.sample_ext();
    EF sum_claim = ef_zero(), cur_mu = ef_one();
    for (size_t t = 0; t < num_traces; t++) {
        sum_claim += bc.numerator_term_per_air[t] * cur_mu;
        cur_mu = cur_mu * mu;
        sum_claim += bc.denominator_term_per_air[t] * cur_mu;
        cur_mu = cur_mu * mu;
    }
    for (const EF& c : bc.univariate_round_coeffs) ts.observe_ext(c);
    const int s_deg = max_constraint_degree + 1;
    const EF r_0 = ts.sample_ext();
    if (bc.univariate_round_coeffs.size() != (size_t)(max_constraint_degree + 1) * (N - 1) + 1) return false;
    EF sum_univ = ef_zero();
    for (size_t i = 0; i < bc.univariate_round_coeffs.size(); i += N) sum_univ += bc.univariate_round_coeffs[i];
    sum_univ = sum_univ * from_canonical(N);
    if (sum_claim != sum_univ) return false;
    EF cur_sum = horner_eval(bc.univariate_round_coeffs, r_0);
    std::vector<EF> rs{r_0};
    if ((int)bc.sumcheck_round_polys.size() != n_max) return false;
    for (int round = 0; round < n_max; round++) {
        const std::vector<EF>& evs = bc.sumcheck_round_polys[round];
        if ((int)evs.size() != s_deg) return false;
        for (const EF& e : evs) ts.observe_ext(e);
        std::vector<EF> all{cur_sum - evs[0]};
        all.insert(all.end(), evs.begin(), evs.end());
        std::vector<F> fact(s_deg + 1, f_one());
        for (int i = 1; i <= s_deg; i++) fact[i] = fact[i - 1] * from_canonical((uint64_t)i);
        const EF r = ts.sample_ext();
        std::vector<EF> pref(s_deg + 1, ef_one()), suf(s_deg + 1, ef_one());
        for (int i = 0; i < s_deg; i++) {
            pref[i + 1] = pref[i] * (r - ef_from_u64((uint64_t)i));
            suf[i + 1] = suf[i] * (ef_from_u64((uint64_t)(s_deg - i)) - r);
        }
        EF acc = ef_zero();
        for (int i = 0; i <= s_deg; i++) acc += all[i] * pref[i] * suf[s_deg - i] * f_inv(fact[i]) * f_inv(fact[s_deg - i]);
        cur_sum = acc;
        rs.push_back(r);
    }
    // eq_3b per trace
    size_t stacked_idx = 0;
    std::vector<std::vector<EF>> eq_3b(num_traces);
    for (size_t t = 0; t < num_traces; t++) {
        const int n_lift = std::max(n_per_trace[t], 0);
        for (size_t i = 0; i < airs[t].interactions.size(); i++) {
            size_t b_int = stacked_idx >> (l_skip + n_lift);
            std::vector<EF> b(n_logup - n_lift);
            for (auto& x : b) {
                x = (b_int & 1) ? ef_one() : ef_zero();
                b_int >>= 1;
            }
            stacked_idx += size_t(1) << (l_skip + n_lift);
            eq_3b[t].push_back(eval_eq_mle(xi.data() + l_skip + n_lift, b.data(), b.size()));
        }
    }
    std::vector<F> omega_pows(N);
    {
        F p = f_one();
        const F w = two_adic_generator(l_skip);
        for (auto& x : omega_pows) {
            x = p;
            p *= w;
        }
    }
    std::vector<EF> eq_ns(n_max + 1, ef_one()), eq_sharp_ns(n_max + 1, ef_one());
    eq_ns[0] = eval_eq_uni(l_skip, xi[0], r_0);
    eq_sharp_ns[0] = eval_eq_sharp_uni(omega_pows, xi.data(), l_skip, r_0);
    for (int i = 1; i <= n_max; i++) {
        const EF e = eval_eq_mle(&xi[l_skip + i - 1], &rs[i], 1);
        eq_ns[i] = eq_ns[i - 1] * e;
        eq_sharp_ns[i] = eq_sharp_ns[i - 1] * e;
    }
    EF r_rev_prod = rs[n_max];
    for (int i = n_max - 1; i >= 0; i--) {
        eq_ns[i] = eq_ns[i] * r_rev_prod;
        eq_sharp_ns[i] = eq_sharp_ns[i] * r_rev_prod;
        r_rev_prod = r_rev_prod * rs[i];
    }
    if (bc.column_openings.size() != num_traces) return false;
    auto pairs_of = [](const std::vector<EF>& v, bool need_rot) {
        std::vector<std::pair<EF, EF>> out;
        if (need_rot)
            for (size_t i = 0; i + 1 < v.size(); i += 2) out.push_back({v[i], v[i + 1]});
        else
            for (const EF& e : v) out.push_back({e, ef_zero()});
        return out;
    };
    for (size_t t = 0; t < num_traces; t++)
        for (auto& pr : pairs_of(bc.column_openings[t][0], airs[t].need_rot)) {
            ts.observe_ext(pr.first);
            ts.observe_ext(pr.second);
        }
    std::vector<EF> interactions_evals, constraints_evals;
    for (size_t t = 0; t < num_traces; t++) {
        const AirCtx& a = airs[t];
        const int n = n_per_trace[t], n_lift = std::max(n, 0);
        const auto& ao = bc.column_openings[t];
        for (size_t p = 1; p < ao.size(); p++)
            for (auto& pr : pairs_of(ao[p], a.need_rot)) {
                ts.observe_ext(pr.first);
                ts.observe_ext(pr.second);
            }
        // row_parts in EvalHelper order from the openings: sels, then prep, cached..., common (each local[, next])
        int l = l_skip;
        std::vector<EF> rs_n(rs.begin(), rs.begin() + n_lift + 1);
        F norm = f_one();
        if (n < 0) {
            l = l_skip + n;
            rs_n.assign(1, ef_exp_power_of_2(rs[0], -n));
            norm = f_inv(from_canonical(size_t(1) << (-n)));
        }
        const F omega = two_adic_generator(l);
        const EF inv = ef_from(f_inv(from_canonical(size_t(1) << l)));
        EF prod0 = ef_one(), prod1 = ef_one();
        for (size_t i = 1; i < rs_n.size(); i++) {
            prod0 = prod0 * (ef_one() - rs_n[i]);
            prod1 = prod1 * rs_n[i];
        }
        const EF is_first = inv * progression_exp_2(rs_n[0], l) * prod0;
        const EF is_last = inv * progression_exp_2(rs_n[0] * omega, l) * prod1;
        std::vector<std::vector<EF>> rows;
        rows.push_back({is_first, ef_one() - is_last, is_last});
        auto push_part = [&](const std::vector<EF>& flat) {
            std::vector<EF> loc, nxt;
            for (auto& pr : pairs_of(flat, a.need_rot)) {
                loc.push_back(pr.first);
                nxt.push_back(pr.second);
            }
            rows.push_back(loc);
            if (a.need_rot) rows.push_back(nxt);
        };
        for (size_t p = 1; p < ao.size(); p++) push_part(ao[p]);  // preprocessed (if any) then cached
        push_part(ao[0]);                                          // common main last
        std::vector<EF> lambda_pows(a.constraint_idx.size());
        {
            EF p = ef_one();
            for (auto& x : lambda_pows) {
                x = p;
                p = p * lambda;
            }
        }
        constraints_evals.push_back(eq_ns[n_lift] * acc_constraints(a, rows, lambda_pows));
        size_t max_len = 0;
        for (auto& it : a.interactions) max_len = std::max(max_len, it.message.size());
        std::vector<EF> beta_pows(max_len + 1);
        {
            EF p = ef_one();
            for (auto& x : beta_pows) {
                x = p;
                p = p * beta;
            }
        }
        auto v = acc_interactions(a, rows, beta_pows, eq_3b[t]);
        interactions_evals.push_back(v[0] * norm * eq_sharp_ns[n_lift]);
        interactions_evals.push_back(v[1] * eq_sharp_ns[n_lift]);
    }
    EF evaluated = ef_zero(), mp = ef_one();
    for (const EF& x : interactions_evals) {
        evaluated += x * mp;
        mp = mp * mu;
    }
    for (const EF& x : constraints_evals) {
        evaluated += x * mp;
        mp = mp * mu;
    }
    if (cur_sum != evaluated) return false;
    *r_out = rs;
    return true;
}

}  // namespace orc
