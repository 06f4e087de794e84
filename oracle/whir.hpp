// ORACLE — TEST INFRASTRUCTURE ONLY (see bb31.hpp header).
//
// WHIR opening proof of the stacked PCS, prover and verifier.  CPU restatement of
//   crates/stark-backend/src/prover/whir.rs:78-352            prove_whir_opening, w_evals_accumulate
//   crates/stark-backend/src/verifier/whir.rs:27-400          verify_whir, binary_k_fold, merkle_verify
//   crates/stark-backend/src/config.rs:172-197                WhirConfig / WhirRoundConfig
//   crates/stark-backend/src/hasher.rs:29-42                  tree_compress
// PARITY UNPINNED against reference outputs (no Rust toolchain here); pinned by "oracle verifier
// accepts oracle prover" and "CUDA prover == oracle prover" (tests/test_whir.py).
//
// Flat proof (uint32 words, field elements in Montgomery form), sections in this order:
//   mu_pow_witness[1] | whir_sumcheck_polys[R*k][2 EF] | codeword_commits[R-1][8] | ood_values[R-1][EF]
//   | folding_pow_witnesses[R*k] | query_phase_pow_witnesses[R]
//   | initial_round_opened_rows: per commit, per query: [2^k][width] | initial_round_merkle_proofs:
//   per commit, per query: [m + log_blowup - k][8] | codeword_opened_values: per round 1..R-1, per
//   query: [2^k][EF] | codeword_merkle_proofs: per round 1..R-1, per query: [m + log_blowup - r - k][8]
//   | final_poly[2^(m - R*k)][EF]
#pragma once
#include "poly.hpp"
#include "transcript.hpp"

namespace orc {

struct WhirConfig {
    int k = 0;
    std::vector<int> num_queries;  // per WHIR round
    int mu_pow_bits = 0, query_phase_pow_bits = 0, folding_pow_bits = 0;
    int num_rounds() const { return (int)num_queries.size(); }
};

inline void push_ef(std::vector<uint32_t>& out, const EF& e) {
    for (int i = 0; i < 4; i++) out.push_back(e.c[i].v);
}
inline void push_digest(std::vector<uint32_t>& out, const Digest& d) {
    for (int i = 0; i < 8; i++) out.push_back(d.w[i].v);
}

// whir.rs:343-352
inline void w_evals_accumulate(std::vector<EF>& w, EF z, EF gamma) {
    const int dim = log2_strict(w.size());
    std::vector<EF> z_pows;
    for (int i = 0; i < dim; i++) {
        z_pows.push_back(z);
        z = z * z;
    }
    std::vector<EF> ev = evals_eq_hypercube(z_pows);
    for (size_t i = 0; i < w.size(); i++) w[i] += gamma * ev[i];
}

// whir.rs:78-341.  `u` has m = log2(stacked height) entries.
inline std::vector<uint32_t> prove_whir_opening(DuplexSponge& ts, int l_skip, int log_blowup, const WhirConfig& cfg,
                                                const std::vector<const StackedPcsData*>& commits,
                                                const std::vector<EF>& u) {
    const F mu_pow_witness = ts.grind(cfg.mu_pow_bits);
    const EF mu = ts.sample_ext();
    const size_t height = commits[0]->matrix.height;
    int m = log2_strict(height);
    const int k = cfg.k, R = cfg.num_rounds();
    if ((int)u.size() != m) throw std::invalid_argument("u length");

    // f_evals = sum_j mu^j * hypercube evals of the MLE whose coefficients are the RS message of column j
    std::vector<EF> f_evals(height, ef_zero());
    {
        // columns are split over host threads, each with its own partial sum (exact additions commute)
        std::vector<const F*> cols;
        for (const StackedPcsData* d : commits) {
            if (d->matrix.height != height) throw std::invalid_argument("commit heights differ");
            for (size_t c = 0; c < d->matrix.width; c++) cols.push_back(d->matrix.col(c));
        }
        std::vector<EF> mu_pows(cols.size(), ef_one());
        for (size_t j = 1; j < cols.size(); j++) mu_pows[j] = mu_pows[j - 1] * mu;
        const unsigned workers = par_workers(cols.size(), 1);
        std::vector<std::vector<EF>> partial(workers);
        parallel_for_tid(cols.size(), workers, [&](unsigned wk, size_t c_begin, size_t c_end) {
            std::vector<EF>& acc = partial[wk];
            acc.assign(height, ef_zero());
            for (size_t j = c_begin; j < c_end; j++) {
                std::vector<F> x(cols[j], cols[j] + height);
                eval_to_coeff_rs_message_inplace(l_skip, x.data(), height);
                mle_coeffs_to_evals_inplace(x.data(), height);
                for (size_t i = 0; i < height; i++) acc[i] += mu_pows[j] * x[i];
            }
        });
        for (const auto& pa : partial)
            for (size_t i = 0; i < height; i++) f_evals[i] += pa[i];
    }
    std::vector<EF> w_evals = evals_mobius_eq_hypercube(u);

    std::vector<uint32_t> sec_polys, sec_commits, sec_ood, sec_fold_pow, sec_query_pow, sec_rows, sec_proofs0, sec_vals,
        sec_proofs, sec_final;
    std::vector<std::vector<uint32_t>> rows_per_commit(commits.size()), proofs_per_commit(commits.size());
    MerkleTree rs_tree;
    bool have_rs_tree = false;
    int log_rs_domain_size = m + log_blowup;
    for (int whir_round = 0; whir_round < R; whir_round++) {
        const bool is_last = whir_round == R - 1;
        for (int round = 0; round < k; round++) {
            EF s[2] = {ef_zero(), ef_zero()};
            const size_t ny = f_evals.size() / 2;
            {
                const unsigned workers = par_workers(ny, 2048);
                std::vector<std::array<EF, 2>> partial(workers, {ef_zero(), ef_zero()});
                parallel_for_tid(ny, workers, [&](unsigned wk, size_t y_begin, size_t y_end) {
                    for (int X = 1; X <= 2; X++) {
                        const F xf = from_canonical((uint64_t)X);
                        EF acc = ef_zero();
                        for (size_t y = y_begin; y < y_end; y++) {
                            const EF f_x = f_evals[2 * y] + (f_evals[2 * y + 1] - f_evals[2 * y]) * xf;
                            const EF w_x = w_evals[2 * y] + (w_evals[2 * y + 1] - w_evals[2 * y]) * xf;
                            acc += f_x * w_x;
                        }
                        partial[wk][X - 1] = acc;
                    }
                });
                for (const auto& pa : partial) {
                    s[0] += pa[0];
                    s[1] += pa[1];
                }
            }
            ts.observe_ext(s[0]);
            ts.observe_ext(s[1]);
            push_ef(sec_polys, s[0]);
            push_ef(sec_polys, s[1]);
            sec_fold_pow.push_back(ts.grind(cfg.folding_pow_bits).v);
            const EF alpha = ts.sample_ext();
            {
                std::vector<EF> f2(ny), w2(ny);
                parallel_for(ny, [&](size_t b, size_t e) {
                    for (size_t y = b; y < e; y++) {
                        f2[y] = f_evals[2 * y] + alpha * (f_evals[2 * y + 1] - f_evals[2 * y]);
                        w2[y] = w_evals[2 * y] + alpha * (w_evals[2 * y + 1] - w_evals[2 * y]);
                    }
                }, 4096);
                f_evals.swap(f2);
                w_evals.swap(w2);
            }
        }
        std::vector<EF> g_coeffs = f_evals;
        mle_evals_to_coeffs_inplace(g_coeffs);
        MerkleTree g_tree;
        EF z_0 = ef_zero();
        if (!is_last) {
            // RS codeword of g on the domain of size 2^(log_rs_domain_size - 1): component-wise DFT
            const size_t N = size_t(1) << (log_rs_domain_size - 1);
            ColMajor cw(N, 4);
            parallel_for(4, [&](size_t b, size_t e) {
                for (size_t comp = b; comp < e; comp++) {
                    F* col = cw.col(comp);
                    for (size_t i = 0; i < g_coeffs.size(); i++) col[i] = g_coeffs[i].c[comp];
                    dft_inplace(col, N);
                }
            });
            g_tree = merkle_tree_new(std::move(cw), size_t(1) << k);
            const Digest g_commit = g_tree.root();
            ts.observe_commit(g_commit);
            push_digest(sec_commits, g_commit);
            z_0 = ts.sample_ext();
            std::vector<EF> z0_vec;
            EF zp = z_0;
            for (int i = 0; i < m - k; i++) {
                z0_vec.push_back(zp);
                zp = zp * zp;
            }
            const EF g_opened = mle_eval_at_point(g_coeffs, z0_vec);
            ts.observe_ext(g_opened);
            push_ef(sec_ood, g_opened);
        } else {
            for (const EF& c : g_coeffs) {
                ts.observe_ext(c);
                push_ef(sec_final, c);
            }
        }
        const F omega = two_adic_generator(log_rs_domain_size - k);
        const int nq = cfg.num_queries[whir_round];
        sec_query_pow.push_back(ts.grind(cfg.query_phase_pow_bits).v);
        std::vector<size_t> idxs;
        for (int q = 0; q < nq; q++) idxs.push_back((size_t)ts.sample_bits(log_rs_domain_size - k));
        std::vector<F> zs;
        for (int q = 0; q < nq; q++) {
            const size_t index = idxs[q];
            zs.push_back(f_pow(omega, index));
            if (whir_round == 0) {
                for (size_t ci = 0; ci < commits.size(); ci++) {
                    const MerkleTree& tree = commits[ci]->tree;
                    if (tree.backing.height != (size_t(1) << log_rs_domain_size)) throw std::runtime_error("TreeHeightMismatch");
                    for (auto& row : tree.get_opened_rows(index))
                        for (F v : row) rows_per_commit[ci].push_back(v.v);
                    for (auto& d : tree.query_merkle_proof(index)) push_digest(proofs_per_commit[ci], d);
                }
            } else {
                if (!have_rs_tree) throw std::runtime_error("RsTreeNone");
                for (auto& row : rs_tree.get_opened_rows(index))
                    for (F v : row) sec_vals.push_back(v.v);
                for (auto& d : rs_tree.query_merkle_proof(index)) push_digest(sec_proofs, d);
            }
        }
        rs_tree = std::move(g_tree);
        have_rs_tree = !is_last;
        const EF gamma = ts.sample_ext();
        if (!is_last) {
            w_evals_accumulate(w_evals, z_0, gamma);
            EF gp = gamma * gamma;
            for (F z : zs) {
                w_evals_accumulate(w_evals, ef_from(z), gp);
                gp = gp * gamma;
            }
        }
        m -= k;
        log_rs_domain_size -= 1;
    }
    std::vector<uint32_t> out;
    out.push_back(mu_pow_witness.v);
    auto app = [&](const std::vector<uint32_t>& v) { out.insert(out.end(), v.begin(), v.end()); };
    app(sec_polys);
    app(sec_commits);
    app(sec_ood);
    app(sec_fold_pow);
    app(sec_query_pow);
    for (auto& v : rows_per_commit) app(v);
    for (auto& v : proofs_per_commit) app(v);
    app(sec_vals);
    app(sec_proofs);
    app(sec_final);
    return out;
}

inline size_t whir_proof_words(int m, int log_blowup, const WhirConfig& cfg, const std::vector<size_t>& widths) {
    const int k = cfg.k, R = cfg.num_rounds();
    size_t n = 1 + (size_t)R * k * 8 + (size_t)(R - 1) * 12 + (size_t)R * k + R;
    for (size_t w : widths) n += (size_t)cfg.num_queries[0] * ((w << k) + (size_t)(m + log_blowup - k) * 8);
    for (int r = 1; r < R; r++) n += (size_t)cfg.num_queries[r] * ((size_t(4) << k) + (size_t)(m + log_blowup - r - k) * 8);
    n += size_t(4) << (m - R * k);
    return n;
}

// verifier/whir.rs:352-389
inline EF binary_k_fold(std::vector<EF> values, const std::vector<EF>& alphas, F x) {
    const size_t n = values.size();
    const int k = (int)alphas.size();
    const F omega_k = two_adic_generator(k), omega_k_inv = f_inv(omega_k);
    std::vector<F> tw(size_t(1) << (k - 1)), inv_tw(tw.size());
    F a = f_one(), b = f_one();
    for (size_t i = 0; i < tw.size(); i++) {
        tw[i] = a;
        inv_tw[i] = b;
        a *= omega_k;
        b *= omega_k_inv;
    }
    F x_pow = x, x_inv_pow = f_inv(x);
    for (int j = 0; j < k; j++) {
        const size_t mm = n >> (j + 1);
        for (size_t i = 0; i < mm; i++) {
            const F t = tw[i << j] * x_pow, t_inv = inv_tw[i << j] * x_inv_pow;
            values[i] += (alphas[j] - ef_from(t)) * (values[i] - values[mm + i]) * halve(t_inv);
        }
        x_pow *= x_pow;
        x_inv_pow *= x_inv_pow;
    }
    return values[0];
}
inline Digest tree_compress(std::vector<Digest> h) {
    while (h.size() > 1) {
        std::vector<Digest> nx;
        for (size_t i = 0; i < h.size(); i += 2) nx.push_back(compress(h[i], h[i + 1]));
        h.swap(nx);
    }
    return h[0];
}
inline bool merkle_verify(const Digest& root, uint32_t idx, Digest cur, const uint32_t* proof, size_t depth) {
    for (size_t l = 0; l < depth; l++) {
        Digest sib;
        memcpy(&sib, proof + 8 * l, 32);
        cur = (idx & 1) == 0 ? compress(cur, sib) : compress(sib, cur);
        idx >>= 1;
    }
    return memcmp(&root, &cur, 32) == 0;
}

// verifier/whir.rs:27-318 over the flat proof.  stacking_openings: per commit, width EF values.
inline bool verify_whir(DuplexSponge& ts, int l_skip, int n_stack, int log_blowup, const WhirConfig& cfg,
                        const std::vector<uint32_t>& proof, const std::vector<std::vector<EF>>& stacking_openings,
                        const std::vector<Digest>& commitments, const std::vector<EF>& u) {
    const int m = l_skip + n_stack, k = cfg.k, R = cfg.num_rounds();
    std::vector<size_t> widths;
    for (auto& v : stacking_openings) widths.push_back(v.size());
    if (proof.size() != whir_proof_words(m, log_blowup, cfg, widths)) return false;
    const uint32_t* p = proof.data();
    auto rd_ef = [](const uint32_t* q) { EF e; memcpy(&e, q, 16); return e; };
    const uint32_t* mu_pow = p; p += 1;
    const uint32_t* polys = p; p += (size_t)R * k * 8;
    const uint32_t* commits = p; p += (size_t)(R - 1) * 8;
    const uint32_t* ood = p; p += (size_t)(R - 1) * 4;
    const uint32_t* fold_pow = p; p += (size_t)R * k;
    const uint32_t* query_pow = p; p += R;
    std::vector<const uint32_t*> rows0, proofs0;
    for (size_t w : widths) { rows0.push_back(p); p += (size_t)cfg.num_queries[0] * (w << k); }
    for (size_t i = 0; i < widths.size(); i++) { proofs0.push_back(p); p += (size_t)cfg.num_queries[0] * (m + log_blowup - k) * 8; }
    std::vector<const uint32_t*> vals_r(R, nullptr), proofs_r(R, nullptr);
    for (int r = 1; r < R; r++) { vals_r[r] = p; p += (size_t)cfg.num_queries[r] * (size_t(4) << k); }
    for (int r = 1; r < R; r++) { proofs_r[r] = p; p += (size_t)cfg.num_queries[r] * (m + log_blowup - r - k) * 8; }
    std::vector<EF> final_poly(size_t(1) << (m - R * k));
    for (size_t i = 0; i < final_poly.size(); i++) final_poly[i] = rd_ef(p + 4 * i);

    if (!ts.check_witness(cfg.mu_pow_bits, F::raw(mu_pow[0]))) return false;
    const EF mu = ts.sample_ext();
    size_t total_w = 0;
    for (size_t w : widths) total_w += w;
    std::vector<EF> mu_pows(total_w);
    {
        EF a = ef_one();
        for (auto& x : mu_pows) { x = a; a = a * mu; }
    }
    EF claim = ef_zero();
    {
        size_t j = 0;
        for (auto& v : stacking_openings)
            for (const EF& o : v) claim += mu_pows[j++] * o;
    }
    std::vector<EF> gammas, z0s, alphas;
    std::vector<std::vector<F>> zs;
    int log_rs = m + log_blowup;
    size_t sc_i = 0;
    for (int wr = 0; wr < R; wr++) {
        const bool is_initial = wr == 0, is_final = wr == R - 1;
        std::vector<EF> alphas_round;
        for (int i = 0; i < k; i++, sc_i++) {
            const EF ev1 = rd_ef(polys + sc_i * 8), ev2 = rd_ef(polys + sc_i * 8 + 4);
            ts.observe_ext(ev1);
            ts.observe_ext(ev2);
            if (!ts.check_witness(cfg.folding_pow_bits, F::raw(fold_pow[sc_i]))) return false;
            const EF alpha = ts.sample_ext();
            alphas_round.push_back(alpha);
            const EF ev[3] = {claim - ev1, ev1, ev2};
            claim = interpolate_quadratic_at_012(ev, alpha);
        }
        bool have_y0 = false;
        EF y0 = ef_zero();
        Digest round_commit{};
        if (is_final) {
            for (const EF& c : final_poly) ts.observe_ext(c);
        } else {
            memcpy(&round_commit, commits + 8 * wr, 32);
            ts.observe_commit(round_commit);
            z0s.push_back(ts.sample_ext());
            y0 = rd_ef(ood + 4 * wr);
            ts.observe_ext(y0);
            have_y0 = true;
        }
        if (!ts.check_witness(cfg.query_phase_pow_bits, F::raw(query_pow[wr]))) return false;
        const int nq = cfg.num_queries[wr];
        std::vector<uint64_t> idxs;
        for (int q = 0; q < nq; q++) idxs.push_back(ts.sample_bits(log_rs - k));
        std::vector<F> zs_round;
        std::vector<EF> ys_round;
        const F omega = two_adic_generator(log_rs);
        const size_t depth = (size_t)(log_rs - k);
        for (int q = 0; q < nq; q++) {
            const uint64_t index = idxs[q];
            const F zi_root = f_pow(omega, index);
            F zi = zi_root;
            for (int i = 0; i < k; i++) zi *= zi;
            EF yi;
            if (is_initial) {
                std::vector<EF> codeword_vals(size_t(1) << k, ef_zero());
                size_t mu_i = 0;
                for (size_t ci = 0; ci < widths.size(); ci++) {
                    const size_t w = widths[ci];
                    const uint32_t* rows = rows0[ci] + (size_t)q * (w << k);
                    std::vector<Digest> leaf;
                    for (size_t j = 0; j < (size_t(1) << k); j++)
                        leaf.push_back(hash_slice(reinterpret_cast<const F*>(rows + j * w), w));
                    if (!merkle_verify(commitments[ci], (uint32_t)index, tree_compress(leaf),
                                       proofs0[ci] + (size_t)q * depth * 8, depth))
                        return false;
                    for (size_t c = 0; c < w; c++, mu_i++)
                        for (size_t j = 0; j < (size_t(1) << k); j++)
                            codeword_vals[j] += mu_pows[mu_i] * F::raw(rows[j * w + c]);
                }
                yi = binary_k_fold(codeword_vals, alphas_round, zi_root);
            } else {
                const uint32_t* vals = vals_r[wr] + (size_t)q * (size_t(4) << k);
                std::vector<EF> opened(size_t(1) << k);
                std::vector<Digest> leaf;
                for (size_t j = 0; j < opened.size(); j++) {
                    opened[j] = rd_ef(vals + 4 * j);
                    leaf.push_back(hash_slice(reinterpret_cast<const F*>(vals + 4 * j), 4));
                }
                Digest prev_commit;
                memcpy(&prev_commit, commits + 8 * (wr - 1), 32);
                if (!merkle_verify(prev_commit, (uint32_t)index, tree_compress(leaf), proofs_r[wr] + (size_t)q * depth * 8, depth))
                    return false;
                yi = binary_k_fold(opened, alphas_round, zi_root);
            }
            zs_round.push_back(zi);
            ys_round.push_back(yi);
        }
        const EF gamma = ts.sample_ext();
        if (have_y0) claim += y0 * gamma;
        EF gp = gamma * gamma;
        for (const EF& yi : ys_round) {
            claim += yi * gp;
            gp = gp * gamma;
        }
        gammas.push_back(gamma);
        zs.push_back(zs_round);
        alphas.insert(alphas.end(), alphas_round.begin(), alphas_round.end());
        log_rs -= 1;
    }
    const int t = k * R;
    std::vector<EF> u_pre(u.begin(), u.begin() + t), a_pre(alphas.begin(), alphas.begin() + t);
    const EF prefix = eval_mobius_eq_mle(u_pre, a_pre);
    std::vector<EF> fp = final_poly;
    {
        size_t len = fp.size();
        for (int j = m - 1; j >= t; j--) {
            len >>= 1;
            for (size_t i = 0; i < len; i++) fp[i] = fp[i] * (ef_one() - u[j]) + fp[len + i] * u[j];
        }
    }
    EF acc = prefix * fp[0];
    int j = k;
    for (int i = 0; i < R; i++) {
        const EF gamma = gammas[i];
        const EF* alpha_slc = alphas.data() + j;
        const size_t slc_len = (size_t)(t - j) + 1;
        auto term = [&](EF z) {
            std::vector<EF> zp;
            for (size_t q = 0; q < slc_len; q++) { zp.push_back(z); z = z * z; }
            return eval_eq_mle(alpha_slc, zp.data(), slc_len - 1) * horner_eval(final_poly, zp.back());
        };
        if (i != R - 1) acc += gamma * term(z0s[i]);
        EF gp = gamma * gamma;
        for (F zi : zs[i]) {
            acc += gp * term(ef_from(zi));
            gp = gp * gamma;
        }
        j += k;
    }
    return acc == claim;
}

}  // namespace orc
