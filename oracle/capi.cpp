// ORACLE — TEST INFRASTRUCTURE ONLY (see bb31.hpp header).
// Flat C entry points over the C++ restatement so that tests/ (ctypes) and bench.py's
// cpu_baseline leg can drive it.  All field words cross this API in Montgomery form, exactly as
// they cross the product's C-ABI (include/swirl_b200.h).
#include <cstring>
#include <stdexcept>

#include "commit.hpp"
#include "gkr.hpp"
#include "logup_zerocheck.hpp"
#include "stacked_reduction.hpp"
#include "transcript.hpp"
#include "whir.hpp"

using namespace orc;

#define ORC_TRY(body)                 \
    try {                             \
        body;                         \
        return 0;                     \
    } catch (const std::exception&) { \
        return 1;                     \
    }

extern "C" {

// ---- field -------------------------------------------------------------------------------
uint32_t orc_from_canonical(uint32_t x) { return from_canonical(x).v; }
uint32_t orc_to_canonical(uint32_t m) { return to_canonical(F::raw(m)); }
uint32_t orc_f_add(uint32_t a, uint32_t b) { return (F::raw(a) + F::raw(b)).v; }
uint32_t orc_f_sub(uint32_t a, uint32_t b) { return (F::raw(a) - F::raw(b)).v; }
uint32_t orc_f_mul(uint32_t a, uint32_t b) { return (F::raw(a) * F::raw(b)).v; }
uint32_t orc_f_inv(uint32_t a) { return f_inv(F::raw(a)).v; }
uint32_t orc_two_adic_generator(int bits) { return two_adic_generator(bits).v; }
void orc_from_canonical_vec(uint32_t* a, size_t n) {
    for (size_t i = 0; i < n; i++) a[i] = from_canonical(a[i]).v;
}
void orc_to_canonical_vec(uint32_t* a, size_t n) {
    for (size_t i = 0; i < n; i++) a[i] = to_canonical(F::raw(a[i]));
}
void orc_f_mul_vec(uint32_t* out, const uint32_t* a, const uint32_t* b, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = (F::raw(a[i]) * F::raw(b[i])).v;
}
void orc_ef_mul(uint32_t out[4], const uint32_t a[4], const uint32_t b[4]) {
    EF x, y;
    memcpy(&x, a, 16);
    memcpy(&y, b, 16);
    EF z = x * y;
    memcpy(out, &z, 16);
}
// verifier/whir.rs:352-389 on n = 2^k values; pinned by the reference's fold_single / fold_double identities
// (crates/backend-tests/src/lib.rs:1191-1227) in tests/test_whir.py
void orc_binary_k_fold(uint32_t out[4], const uint32_t* values, int k, const uint32_t* alphas, uint32_t x) {
    std::vector<EF> v(size_t(1) << k), al(k);
    memcpy(v.data(), values, v.size() * 16);
    memcpy(al.data(), alphas, al.size() * 16);
    F fx;
    memcpy(&fx, &x, 4);
    const EF z = binary_k_fold(v, al, fx);
    memcpy(out, &z, 16);
}
void orc_ef_inv(uint32_t out[4], const uint32_t a[4]) {
    EF x;
    memcpy(&x, a, 16);
    EF z = ef_inv(x);
    memcpy(out, &z, 16);
}

// ---- Poseidon2 ---------------------------------------------------------------------------
void orc_poseidon2_permute(uint32_t* states, size_t n_states) {
    parallel_for(n_states, [&](size_t i0, size_t i1) {
        for (size_t i = i0; i < i1; i++) poseidon2_permute(reinterpret_cast<F*>(states + 16 * i));
    }, 256);
}
void orc_hash_slice(uint32_t out[8], const uint32_t* vals, size_t n) {
    Digest d = hash_slice(reinterpret_cast<const F*>(vals), n);
    memcpy(out, &d, 32);
}
void orc_compress(uint32_t out[8], const uint32_t l[8], const uint32_t r[8]) {
    Digest a, b;
    memcpy(&a, l, 32);
    memcpy(&b, r, 32);
    Digest d = compress(a, b);
    memcpy(out, &d, 32);
}

// ---- DFT ---------------------------------------------------------------------------------
int orc_dft(uint32_t* a, size_t n, int inverse) {
    ORC_TRY(if (inverse) idft_inplace(reinterpret_cast<F*>(a), n);
            else dft_inplace(reinterpret_cast<F*>(a), n));
}
// batch of `cols` columns of length n each (contiguous), forward natural-order DFT
int orc_dft_batch(uint32_t* a, size_t n, size_t cols, int inverse) {
    try {
        std::vector<F> tw = dft_twiddles(n, inverse != 0);
        F ninv = f_inv(from_canonical(n));
        parallel_for(cols, [&](size_t c0, size_t c1) {
            for (size_t c = c0; c < c1; c++) {
                F* p = reinterpret_cast<F*>(a) + c * n;
                dft_with_twiddles(p, n, tw);
                if (inverse)
                    for (size_t i = 0; i < n; i++) p[i] *= ninv;
            }
        });
        return 0;
    } catch (const std::exception&) {
        return 1;
    }
}
int orc_coset_dft(uint32_t* a, size_t n, uint32_t shift) {
    ORC_TRY(coset_dft_inplace(reinterpret_cast<F*>(a), n, F::raw(shift)));
}

// ---- stacking ----------------------------------------------------------------------------
// out_cols: per sorted column 5 x u64 = (mat_idx, col_in_mat, stacked col_idx, stacked row_idx,
// log_height).  Returns 0 ok; 1 on layout error.  *out_n receives the number of sorted columns
// (call with out_cols == NULL to size).
int orc_stacked_layout(int l_skip, int log_stacked_height, size_t n_mats, const uint64_t* widths,
                       const int32_t* log_heights, uint64_t* out_width, uint64_t* out_n,
                       uint64_t* out_cols) {
    try {
        std::vector<std::pair<size_t, int>> meta;
        for (size_t i = 0; i < n_mats; i++) meta.push_back({(size_t)widths[i], (int)log_heights[i]});
        StackedLayout lay = make_stacked_layout(l_skip, log_stacked_height, meta);
        *out_width = lay.width;
        *out_n = lay.sorted_cols.size();
        if (out_cols)
            for (size_t i = 0; i < lay.sorted_cols.size(); i++) {
                auto& s = lay.sorted_cols[i];
                out_cols[5 * i + 0] = s.mat_idx;
                out_cols[5 * i + 1] = s.col_in_mat;
                out_cols[5 * i + 2] = s.slice.col_idx;
                out_cols[5 * i + 3] = s.slice.row_idx;
                out_cols[5 * i + 4] = (uint64_t)s.slice.log_height;
            }
        return 0;
    } catch (const std::exception&) {
        return 1;
    }
}

static std::vector<ColMajor> wrap_traces(size_t n, const uint32_t* const* ptrs, const uint64_t* heights,
                                         const uint64_t* widths) {
    std::vector<ColMajor> v(n);
    for (size_t i = 0; i < n; i++) {
        v[i] = ColMajor(heights[i], widths[i]);
        memcpy(v[i].values.data(), ptrs[i], heights[i] * widths[i] * 4);
    }
    return v;
}

// out must hold 2^(l_skip+n_stack) * (*out_width) words; call with out == NULL to get width.
int orc_stacked_matrix(int l_skip, int n_stack, size_t n, const uint32_t* const* ptrs,
                       const uint64_t* heights, const uint64_t* widths, uint64_t* out_width, uint32_t* out) {
    try {
        std::vector<ColMajor> tr = wrap_traces(n, ptrs, heights, widths);
        std::vector<const ColMajor*> refs;
        for (auto& t : tr) refs.push_back(&t);
        StackedLayout lay;
        ColMajor q = stacked_matrix(l_skip, n_stack, refs, &lay);
        *out_width = q.width;
        if (out) memcpy(out, q.values.data(), q.values.size() * 4);
        return 0;
    } catch (const std::exception&) {
        return 1;
    }
}

// ---- RS encode ---------------------------------------------------------------------------
int orc_eval_to_coeff_rs_message(int l_skip, uint32_t* a, size_t n) {
    ORC_TRY(eval_to_coeff_rs_message_inplace(l_skip, reinterpret_cast<F*>(a), n));
}
int orc_rs_code_matrix(int l_skip, int log_blowup, const uint32_t* evals, size_t height, size_t width,
                       uint32_t* out) {
    try {
        ColMajor m(height, width);
        memcpy(m.values.data(), evals, height * width * 4);
        ColMajor r = rs_code_matrix(l_skip, log_blowup, m);
        memcpy(out, r.values.data(), r.values.size() * 4);
        return 0;
    } catch (const std::exception&) {
        return 1;
    }
}

// ---- Merkle ------------------------------------------------------------------------------
// out_layers: all digest layers concatenated, layer 0 (query_stride digests) first, root last;
// total digests = 2*query_stride - 1.
int orc_merkle_tree(const uint32_t* matrix, size_t height, size_t width, size_t rows_per_query,
                    uint32_t* out_layers) {
    try {
        ColMajor m(height, width);
        memcpy(m.values.data(), matrix, height * width * 4);
        MerkleTree t = merkle_tree_new(std::move(m), rows_per_query);
        uint32_t* o = out_layers;
        for (auto& l : t.layers) {
            memcpy(o, l.data(), l.size() * 32);
            o += l.size() * 8;
        }
        return 0;
    } catch (const std::exception&) {
        return 1;
    }
}

// Whole commitment.  out_codeword (optional) receives the (H << log_blowup) x W matrix,
// out_layers (optional) the concatenated digest layers, out_root the 8-word root.
int orc_stacked_commit(int l_skip, int n_stack, int log_blowup, int k_whir, size_t n,
                       const uint32_t* const* ptrs, const uint64_t* heights, const uint64_t* widths,
                       uint32_t out_root[8], uint64_t* out_width, uint32_t* out_codeword,
                       uint32_t* out_layers) {
    try {
        std::vector<ColMajor> tr = wrap_traces(n, ptrs, heights, widths);
        std::vector<const ColMajor*> refs;
        for (auto& t : tr) refs.push_back(&t);
        StackedPcsData d;
        Digest root = stacked_commit(l_skip, n_stack, log_blowup, k_whir, refs, &d);
        memcpy(out_root, &root, 32);
        if (out_width) *out_width = d.matrix.width;
        if (out_codeword) memcpy(out_codeword, d.tree.backing.values.data(), d.tree.backing.values.size() * 4);
        if (out_layers) {
            uint32_t* o = out_layers;
            for (auto& l : d.tree.layers) {
                memcpy(o, l.data(), l.size() * 32);
                o += l.size() * 8;
            }
        }
        return 0;
    } catch (const std::exception&) {
        return 1;
    }
}

// ---- transcript --------------------------------------------------------------------------
// state layout: 16 words + absorb_idx + sample_idx  (== DeviceSpongeState, sponge.cu:13-17)
static DuplexSponge load_sponge(const uint32_t* s) {
    DuplexSponge d;
    memcpy(d.state, s, 64);
    d.absorb_idx = s[16];
    d.sample_idx = s[17];
    return d;
}
static void store_sponge(const DuplexSponge& d, uint32_t* s) {
    memcpy(s, d.state, 64);
    s[16] = d.absorb_idx;
    s[17] = d.sample_idx;
}
void orc_sponge_observe(uint32_t* st, const uint32_t* vals, size_t n) {
    DuplexSponge d = load_sponge(st);
    for (size_t i = 0; i < n; i++) d.observe(F::raw(vals[i]));
    store_sponge(d, st);
}
void orc_sponge_sample(uint32_t* st, uint32_t* out, size_t n) {
    DuplexSponge d = load_sponge(st);
    for (size_t i = 0; i < n; i++) out[i] = d.sample().v;
    store_sponge(d, st);
}
uint64_t orc_sponge_sample_bits(uint32_t* st, int bits) {
    DuplexSponge d = load_sponge(st);
    uint64_t r = d.sample_bits(bits);
    store_sponge(d, st);
    return r;
}
int orc_sponge_check_witness(uint32_t* st, int bits, uint32_t w) {
    DuplexSponge d = load_sponge(st);
    bool ok = d.check_witness(bits, F::raw(w));
    store_sponge(d, st);
    return ok ? 1 : 0;
}
// returns the (Montgomery) witness; canonical value is the smallest valid one >= start
uint32_t orc_sponge_grind(uint32_t* st, int bits, uint32_t start) {
    DuplexSponge d = load_sponge(st);
    F w = d.grind(bits, start);
    store_sponge(d, st);
    return w.v;
}


// ---- LogUp-GKR fractional sumcheck ----------------------------------------------------------------
// Flat formats as in include/swirl_b200.h (swirl_gkr_fractional_sumcheck).  Returns 0 ok,
// 2 NonZeroRootSum, 1 other error.
int orc_gkr_prove(uint32_t* st, const uint32_t* leaves, int log_n, int assert_zero, uint32_t* frac_sum,
                  uint32_t* claims, uint32_t* polys, uint32_t* xi) {
    try {
        DuplexSponge ts = load_sponge(st);
        std::vector<Frac> ev(size_t(1) << log_n);
        memcpy(ev.data(), leaves, ev.size() * sizeof(Frac));
        std::vector<EF> xi_v;
        FracSumcheckProof pr = fractional_sumcheck(ts, ev, assert_zero != 0, &xi_v);
        memcpy(frac_sum, &pr.frac_sum_p, 16);
        memcpy(frac_sum + 4, &pr.frac_sum_q, 16);
        for (size_t i = 0; i < pr.claims_per_layer.size(); i++) memcpy(claims + 16 * i, &pr.claims_per_layer[i], 64);
        size_t off = 0;
        for (auto& layer : pr.sumcheck_polys)
            for (auto& s : layer) {
                memcpy(polys + 12 * off, s.data(), 48);
                off++;
            }
        for (size_t i = 0; i < xi_v.size(); i++) memcpy(xi + 4 * i, &xi_v[i], 16);
        store_sponge(ts, st);
        return 0;
    } catch (const NonZeroRootSum&) {
        return 2;
    } catch (const std::exception&) {
        return 1;
    }
}
// Returns 1 when the verifier accepts (then numer/denom/xi are filled), 0 otherwise.
int orc_gkr_verify(uint32_t* st, int total_rounds, const uint32_t* frac_sum, const uint32_t* claims,
                   const uint32_t* polys, uint32_t* numer, uint32_t* denom, uint32_t* xi) {
    DuplexSponge ts = load_sponge(st);
    FracSumcheckProof pr;
    memcpy(&pr.frac_sum_p, frac_sum, 16);
    memcpy(&pr.frac_sum_q, frac_sum + 4, 16);
    pr.claims_per_layer.resize(total_rounds);
    for (int i = 0; i < total_rounds; i++) memcpy(&pr.claims_per_layer[i], claims + 16 * i, 64);
    size_t off = 0;
    for (int round = 1; round < total_rounds; round++) {
        std::vector<std::array<EF, 3>> layer(round);
        for (int sr = 0; sr < round; sr++) memcpy(layer[sr].data(), polys + 12 * (off++), 48);
        pr.sumcheck_polys.push_back(layer);
    }
    EF n, d;
    std::vector<EF> xi_v;
    if (!verify_gkr(pr, ts, total_rounds, &n, &d, &xi_v)) return 0;
    memcpy(numer, &n, 16);
    memcpy(denom, &d, 16);
    for (size_t i = 0; i < xi_v.size(); i++) memcpy(xi + 4 * i, &xi_v[i], 16);
    store_sponge(ts, st);
    return 1;
}
// MLE evaluation of a table of 2^n EF values at an EF point (poly_common.rs:42-54)
void orc_eval_mle_evals_at_point(const uint32_t* evals, int n, const uint32_t* x, uint32_t* out) {
    std::vector<EF> e(size_t(1) << n);
    memcpy(e.data(), evals, e.size() * 16);
    size_t len = e.size();
    for (int j = n; j-- > 0;) {
        EF xj;
        memcpy(&xj, x + 4 * j, 16);
        len >>= 1;
        for (size_t i = 0; i < len; i++) e[i] = e[i] * (ef_one() - xj) + e[len + i] * xj;
    }
    memcpy(out, &e[0], 16);
}

// ---- WHIR ------------------------------------------------------------------------------------------
static WhirConfig mk_whir_cfg(int k, int rounds, const int32_t* num_queries, int mu_pow, int query_pow, int fold_pow) {
    WhirConfig c;
    c.k = k;
    c.num_queries.assign(num_queries, num_queries + rounds);
    c.mu_pow_bits = mu_pow;
    c.query_phase_pow_bits = query_pow;
    c.folding_pow_bits = fold_pow;
    return c;
}
size_t orc_whir_proof_words(int m, int log_blowup, int k, int rounds, const int32_t* num_queries, size_t n_commits,
                            const uint64_t* widths) {
    std::vector<size_t> w(widths, widths + n_commits);
    return whir_proof_words(m, log_blowup, mk_whir_cfg(k, rounds, num_queries, 0, 0, 0), w);
}
// Each commit is given by its stacked matrix (height x widths[i], column-major).  roots: n_commits x 8.
int orc_whir_prove(uint32_t* st, int l_skip, int log_blowup, int k, int rounds, const int32_t* num_queries, int mu_pow,
                   int query_pow, int fold_pow, size_t n_commits, const uint32_t* const* mats, const uint64_t* widths,
                   size_t height, const uint32_t* u, uint32_t* roots, uint32_t* proof_out) {
    try {
        DuplexSponge ts = load_sponge(st);
        WhirConfig cfg = mk_whir_cfg(k, rounds, num_queries, mu_pow, query_pow, fold_pow);
        std::vector<StackedPcsData> data(n_commits);
        std::vector<const StackedPcsData*> ptrs;
        for (size_t i = 0; i < n_commits; i++) {
            data[i].matrix = ColMajor(height, widths[i]);
            memcpy(data[i].matrix.values.data(), mats[i], height * widths[i] * 4);
            data[i].tree = merkle_tree_new(rs_code_matrix(l_skip, log_blowup, data[i].matrix), size_t(1) << k);
            Digest r = data[i].tree.root();
            memcpy(roots + 8 * i, &r, 32);
            ptrs.push_back(&data[i]);
        }
        const int m = log2_strict(height);
        std::vector<EF> uv(m);
        memcpy(uv.data(), u, (size_t)m * 16);
        std::vector<uint32_t> pr = prove_whir_opening(ts, l_skip, log_blowup, cfg, ptrs, uv);
        memcpy(proof_out, pr.data(), pr.size() * 4);
        store_sponge(ts, st);
        return 0;
    } catch (const std::exception&) {
        return 1;
    }
}
// openings[j] = MLE with coefficient table eval_to_coeff_rs_message(column j), evaluated at u
void orc_whir_stacking_openings(int l_skip, const uint32_t* mat, size_t height, size_t width, const uint32_t* u,
                                uint32_t* out) {
    const int m = log2_strict(height);
    for (size_t c = 0; c < width; c++) {
        std::vector<F> x(height);
        memcpy(x.data(), mat + c * height, height * 4);
        eval_to_coeff_rs_message_inplace(l_skip, x.data(), height);
        std::vector<EF> e(height);
        for (size_t i = 0; i < height; i++) e[i] = ef_from(x[i]);
        size_t len = height;
        for (int j = m; j-- > 0;) {
            EF xj;
            memcpy(&xj, u + 4 * j, 16);
            len >>= 1;
            for (size_t i = 0; i < len; i++) e[i] = e[i] * (ef_one() - xj) + e[len + i] * xj;
        }
        memcpy(out + 4 * c, &e[0], 16);
    }
}
int orc_whir_verify(uint32_t* st, int l_skip, int n_stack, int log_blowup, int k, int rounds, const int32_t* num_queries,
                    int mu_pow, int query_pow, int fold_pow, const uint32_t* proof, size_t proof_words, size_t n_commits,
                    const uint64_t* widths, const uint32_t* openings, const uint32_t* roots, const uint32_t* u) {
    try {
        DuplexSponge ts = load_sponge(st);
        WhirConfig cfg = mk_whir_cfg(k, rounds, num_queries, mu_pow, query_pow, fold_pow);
        std::vector<uint32_t> pr(proof, proof + proof_words);
        std::vector<std::vector<EF>> so(n_commits);
        std::vector<Digest> com(n_commits);
        size_t off = 0;
        for (size_t i = 0; i < n_commits; i++) {
            so[i].resize(widths[i]);
            memcpy(so[i].data(), openings + 4 * off, widths[i] * 16);
            off += widths[i];
            memcpy(&com[i], roots + 8 * i, 32);
        }
        std::vector<EF> uv(l_skip + n_stack);
        memcpy(uv.data(), u, uv.size() * 16);
        bool ok = verify_whir(ts, l_skip, n_stack, log_blowup, cfg, pr, so, com, uv);
        if (ok) store_sponge(ts, st);
        return ok ? 1 : 0;
    } catch (const std::exception&) {
        return 0;
    }
}

// ---- stacked opening reduction -------------------------------------------------------------------
// Commit c stacks traces [trace_off[c], trace_off[c+1]) (already height-sorted).  Flat proof:
// univariate_round_coeffs[2(2^l_skip-1)+1][4] | sumcheck_round_polys[n_stack][2][4] | per commit openings[width][4].
static void build_commits(int l_skip, int n_stack, size_t n_commits, const uint64_t* trace_off, const uint32_t* const* ptrs,
                          const uint64_t* heights, const uint64_t* widths, std::vector<std::vector<ColMajor>>& traces,
                          std::vector<StackedPcsData>& data) {
    traces.resize(n_commits);
    data.resize(n_commits);
    for (size_t c = 0; c < n_commits; c++) {
        std::vector<const ColMajor*> tp;
        for (uint64_t t = trace_off[c]; t < trace_off[c + 1]; t++) {
            ColMajor m(heights[t], widths[t]);
            if (ptrs) memcpy(m.values.data(), ptrs[t], heights[t] * widths[t] * 4);
            traces[c].push_back(std::move(m));
        }
        for (auto& m : traces[c]) tp.push_back(&m);
        data[c].matrix = stacked_matrix(l_skip, n_stack, tp, &data[c].layout);
    }
}
size_t orc_stacked_reduction_proof_words(int l_skip, int n_stack, size_t n_commits, const uint64_t* stacked_widths) {
    size_t n = (2 * ((size_t(1) << l_skip) - 1) + 1) * 4 + (size_t)n_stack * 8;
    for (size_t c = 0; c < n_commits; c++) n += stacked_widths[c] * 4;
    return n;
}
int orc_stacked_reduction_prove(uint32_t* st, int l_skip, int n_stack, size_t n_commits, const uint64_t* trace_off,
                                const uint32_t* const* ptrs, const uint64_t* heights, const uint64_t* widths,
                                const uint8_t* need_rot, const uint32_t* r, size_t r_len, uint64_t* stacked_widths,
                                uint32_t* proof_out, uint32_t* u_out) {
    try {
        DuplexSponge ts = load_sponge(st);
        std::vector<std::vector<ColMajor>> traces;
        std::vector<StackedPcsData> data;
        build_commits(l_skip, n_stack, n_commits, trace_off, ptrs, heights, widths, traces, data);
        std::vector<const StackedPcsData*> cp;
        std::vector<std::vector<bool>> rot(n_commits);
        for (size_t c = 0; c < n_commits; c++) {
            cp.push_back(&data[c]);
            stacked_widths[c] = data[c].matrix.width;
            for (uint64_t t = trace_off[c]; t < trace_off[c + 1]; t++) rot[c].push_back(need_rot[t] != 0);
        }
        std::vector<EF> rv(r_len), u;
        memcpy(rv.data(), r, r_len * 16);
        StackingProof pr = prove_stacked_opening_reduction(ts, l_skip, n_stack, cp, rot, rv, &u);
        uint32_t* p = proof_out;
        for (auto& c : pr.univariate_round_coeffs) { memcpy(p, &c, 16); p += 4; }
        for (auto& s : pr.sumcheck_round_polys) { memcpy(p, s.data(), 32); p += 8; }
        for (auto& v : pr.stacking_openings)
            for (auto& c : v) { memcpy(p, &c, 16); p += 4; }
        memcpy(u_out, u.data(), u.size() * 16);
        store_sponge(ts, st);
        return 0;
    } catch (const std::exception&) {
        return 1;
    }
}
// Opening claim of one (possibly rotated) trace column at r = (r_0, r_1..): fold_ple_evals at r_0
// then MLE folds (what prove_zerocheck_and_logup leaves as column_openings; cpu.rs:430-445,582-587,644-694).
void orc_column_opening(int l_skip, const uint32_t* col, size_t height, int is_rot, const uint32_t* r, uint32_t* out) {
    MatPart p;
    p.values = reinterpret_cast<const F*>(col);
    p.height = height;
    p.width = 1;
    p.col_stride = 0;
    p.is_rot = is_rot != 0;
    EF r0;
    memcpy(&r0, r, 16);
    size_t h;
    std::vector<EF> v = fold_ple_evals(l_skip, p, r0, &h);
    int j = 1;
    while (h > 1) {
        EF rj;
        memcpy(&rj, r + 4 * j, 16);
        fold_mle_evals(v, h, 1, rj);
        j++;
    }
    memcpy(out, &v[0], 16);
}
int orc_stacked_reduction_verify(uint32_t* st, int l_skip, int n_stack, size_t n_commits, const uint64_t* trace_off,
                                 const uint64_t* heights, const uint64_t* widths, const uint8_t* need_rot,
                                 const uint32_t* t_claims, const uint32_t* r, size_t r_len, const uint32_t* proof,
                                 uint32_t* u_out) {
    try {
        DuplexSponge ts = load_sponge(st);
        std::vector<std::vector<ColMajor>> traces;
        std::vector<StackedPcsData> data;
        build_commits(l_skip, n_stack, n_commits, trace_off, nullptr, heights, widths, traces, data);
        std::vector<const StackedLayout*> layouts;
        std::vector<std::vector<bool>> rot(n_commits);
        size_t n_claims = 0;
        for (size_t c = 0; c < n_commits; c++) {
            layouts.push_back(&data[c].layout);
            n_claims += data[c].layout.sorted_cols.size();
            for (uint64_t t = trace_off[c]; t < trace_off[c + 1]; t++) rot[c].push_back(need_rot[t] != 0);
        }
        std::vector<std::pair<EF, EF>> claims(n_claims);
        for (size_t i = 0; i < n_claims; i++) {
            memcpy(&claims[i].first, t_claims + 8 * i, 16);
            memcpy(&claims[i].second, t_claims + 8 * i + 4, 16);
        }
        StackingProof pr;
        const uint32_t* p = proof;
        pr.univariate_round_coeffs.resize(2 * ((size_t(1) << l_skip) - 1) + 1);
        for (auto& c : pr.univariate_round_coeffs) { memcpy(&c, p, 16); p += 4; }
        pr.sumcheck_round_polys.resize(n_stack);
        for (auto& s : pr.sumcheck_round_polys) { memcpy(s.data(), p, 32); p += 8; }
        for (size_t c = 0; c < n_commits; c++) {
            std::vector<EF> v(data[c].matrix.width);
            for (auto& x : v) { memcpy(&x, p, 16); p += 4; }
            pr.stacking_openings.push_back(v);
        }
        std::vector<EF> rv(r_len), u;
        memcpy(rv.data(), r, r_len * 16);
        if (!verify_stacked_reduction(ts, pr, layouts, rot, l_skip, n_stack, claims, rv, &u)) return 0;
        memcpy(u_out, u.data(), u.size() * 16);
        store_sponge(ts, st);
        return 1;
    } catch (const std::exception&) {
        return 0;
    }
}

// ---- batch constraints (LogUp-GKR + zerocheck) ---------------------------------------------------
// air_meta[i] = {n_nodes, n_constraints, n_interactions, constraint_degree, need_rot, n_public, n_cached, has_prep};
// nodes / constraint_idx / interactions {count_node, bus_index, msg_off, msg_len} / msg_nodes / public_values are the
// per-AIR arrays concatenated (msg_off is relative to the AIR's own msg_nodes block, whose length is
// the sum of its msg_len); matrices per AIR in the order common_main, cached..., preprocessed.
// Flat proof: logup_pow_witness[1] | q0_claim[4] | claims[L][16] | gkr polys[L(L-1)/2][12] | numer[n][4] | denom[n][4]
//   | univariate_round_coeffs[(D+1)(2^l-1)+1][4] | sumcheck_round_polys[n_max][D+1][4] | column_openings per air,
//   per part (common main, preprocessed, cached...) flat EF.   L = l_skip + n_logup (0 without interactions).
struct BcInputs {
    std::vector<AirCtx> airs;
    std::vector<ColMajor> mats;  // storage
    std::vector<int> n_per_trace;
};
static void bc_parse(int l_skip, size_t n_airs, const uint64_t* meta, const uint32_t* nodes, const uint32_t* cidx,
                     const uint32_t* inter, const uint32_t* msg, const uint32_t* pubs, const uint32_t* const* mat_ptrs,
                     const uint64_t* mat_h, const uint64_t* mat_w, BcInputs& in) {
    size_t n_mats = 0;
    for (size_t i = 0; i < n_airs; i++) n_mats += 1 + meta[8 * i + 6] + meta[8 * i + 7];
    in.mats.resize(n_mats);
    for (size_t k = 0; k < n_mats; k++) {
        in.mats[k] = ColMajor(mat_h[k], mat_w[k]);
        if (mat_ptrs) memcpy(in.mats[k].values.data(), mat_ptrs[k], mat_h[k] * mat_w[k] * 4);
    }
    size_t mi = 0;
    for (size_t i = 0; i < n_airs; i++) {
        const uint64_t* m = meta + 8 * i;
        AirCtx a;
        for (uint64_t k = 0; k < m[0]; k++, nodes += 4) a.nodes.push_back(DagNode{nodes[0], nodes[1], nodes[2], nodes[3]});
        a.constraint_idx.assign(cidx, cidx + m[1]);
        cidx += m[1];
        size_t msg_total = 0;
        for (uint64_t k = 0; k < m[2]; k++, inter += 4) {
            Interaction it;
            it.count_node = inter[0];
            it.bus_index = inter[1];
            it.message.assign(msg + inter[2], msg + inter[2] + inter[3]);
            msg_total += inter[3];
            a.interactions.push_back(it);
        }
        msg += msg_total;
        a.constraint_degree = (int)m[3];
        a.need_rot = m[4] != 0;
        for (uint64_t k = 0; k < m[5]; k++) a.public_values.push_back(F::raw(pubs[k]));
        pubs += m[5];
        a.common_main = &in.mats[mi++];
        for (uint64_t k = 0; k < m[6]; k++) a.cached_mains.push_back(&in.mats[mi++]);
        if (m[7]) a.preprocessed = &in.mats[mi++];
        in.n_per_trace.push_back(log2_strict(a.common_main->height) - l_skip);
        in.airs.push_back(a);
    }
}
static size_t bc_words(int l_skip, int D, const BcInputs& in, int* L_out, int* n_max_out) {
    uint64_t total = 0;
    int n_max = 0;
    for (size_t t = 0; t < in.airs.size(); t++) {
        total += (uint64_t)in.airs[t].interactions.size() << (l_skip + std::max(in.n_per_trace[t], 0));
        n_max = std::max(n_max, in.n_per_trace[t]);
    }
    const int L = total ? l_skip + calculate_n_logup(l_skip, total) : 0;
    size_t n = 1 + 4 + (size_t)L * 16 + (size_t)L * (L > 0 ? L - 1 : 0) / 2 * 12 + in.airs.size() * 8 +
               ((size_t)(D + 1) * ((size_t(1) << l_skip) - 1) + 1) * 4 + (size_t)n_max * (D + 1) * 4;
    for (const AirCtx& a : in.airs) {
        size_t w = a.common_main->width;
        for (auto* c : a.cached_mains) w += c->width;
        if (a.preprocessed) w += a.preprocessed->width;
        n += w * (a.need_rot ? 2 : 1) * 4;
    }
    if (L_out) *L_out = L;
    if (n_max_out) *n_max_out = n_max;
    return n;
}
size_t orc_bc_proof_words(int l_skip, int D, size_t n_airs, const uint64_t* meta, const uint32_t* nodes, const uint32_t* cidx,
                          const uint32_t* inter, const uint32_t* msg, const uint32_t* pubs, const uint64_t* mat_h,
                          const uint64_t* mat_w) {
    try {
        BcInputs in;
        bc_parse(l_skip, n_airs, meta, nodes, cidx, inter, msg, pubs, nullptr, mat_h, mat_w, in);
        return bc_words(l_skip, D, in, nullptr, nullptr);
    } catch (const std::exception&) {
        return 0;
    }
}
// returns 0 ok, 2 NonZeroRootSum (unbalanced LogUp), 1 other
int orc_bc_prove(uint32_t* st, int l_skip, int D, int logup_pow_bits, size_t n_airs, const uint64_t* meta,
                 const uint32_t* nodes, const uint32_t* cidx, const uint32_t* inter, const uint32_t* msg,
                 const uint32_t* pubs, const uint32_t* const* mat_ptrs, const uint64_t* mat_h, const uint64_t* mat_w,
                 uint32_t* proof, uint32_t* r_out) {
    try {
        DuplexSponge ts = load_sponge(st);
        BcInputs in;
        bc_parse(l_skip, n_airs, meta, nodes, cidx, inter, msg, pubs, mat_ptrs, mat_h, mat_w, in);
        GkrProof g;
        BatchConstraintProof bc;
        std::vector<EF> r;
        prove_zerocheck_and_logup(ts, l_skip, D, logup_pow_bits, in.airs, &g, &bc, &r);
        uint32_t* p = proof;
        *p++ = g.logup_pow_witness.v;
        memcpy(p, &g.q0_claim, 16); p += 4;
        for (auto& c : g.claims_per_layer) { memcpy(p, &c, 64); p += 16; }
        for (auto& layer : g.sumcheck_polys)
            for (auto& sp : layer) { memcpy(p, sp.data(), 48); p += 12; }
        for (auto& e : bc.numerator_term_per_air) { memcpy(p, &e, 16); p += 4; }
        for (auto& e : bc.denominator_term_per_air) { memcpy(p, &e, 16); p += 4; }
        for (auto& e : bc.univariate_round_coeffs) { memcpy(p, &e, 16); p += 4; }
        for (auto& rp : bc.sumcheck_round_polys)
            for (auto& e : rp) { memcpy(p, &e, 16); p += 4; }
        for (auto& ao : bc.column_openings)
            for (auto& part : ao)
                for (auto& e : part) { memcpy(p, &e, 16); p += 4; }
        memcpy(r_out, r.data(), r.size() * 16);
        store_sponge(ts, st);
        return 0;
    } catch (const NonZeroRootSum&) {
        return 2;
    } catch (const std::exception&) {
        return 1;
    }
}
int orc_bc_verify(uint32_t* st, int l_skip, int D, int logup_pow_bits, size_t n_airs, const uint64_t* meta,
                  const uint32_t* nodes, const uint32_t* cidx, const uint32_t* inter, const uint32_t* msg,
                  const uint32_t* pubs, const uint64_t* mat_h, const uint64_t* mat_w, const uint32_t* proof,
                  uint32_t* r_out) {
    try {
        DuplexSponge ts = load_sponge(st);
        BcInputs in;
        bc_parse(l_skip, n_airs, meta, nodes, cidx, inter, msg, pubs, nullptr, mat_h, mat_w, in);
        int L = 0, n_max = 0;
        bc_words(l_skip, D, in, &L, &n_max);
        GkrProof g;
        BatchConstraintProof bc;
        const uint32_t* p = proof;
        g.logup_pow_witness = F::raw(*p++);
        memcpy(&g.q0_claim, p, 16); p += 4;
        g.claims_per_layer.resize(L);
        for (auto& c : g.claims_per_layer) { memcpy(&c, p, 64); p += 16; }
        for (int round = 1; round < L; round++) {
            std::vector<std::array<EF, 3>> layer(round);
            for (auto& sp : layer) { memcpy(sp.data(), p, 48); p += 12; }
            g.sumcheck_polys.push_back(layer);
        }
        auto rd = [&](std::vector<EF>& v, size_t n) {
            v.resize(n);
            for (auto& e : v) { memcpy(&e, p, 16); p += 4; }
        };
        rd(bc.numerator_term_per_air, n_airs);
        rd(bc.denominator_term_per_air, n_airs);
        rd(bc.univariate_round_coeffs, (size_t)(D + 1) * ((size_t(1) << l_skip) - 1) + 1);
        bc.sumcheck_round_polys.resize(n_max);
        for (auto& rp : bc.sumcheck_round_polys) rd(rp, D + 1);
        for (const AirCtx& a : in.airs) {
            std::vector<std::vector<EF>> ao;
            const size_t mul = a.need_rot ? 2 : 1;
            std::vector<EF> part;
            rd(part, a.common_main->width * mul);
            ao.push_back(part);
            if (a.preprocessed) { rd(part, a.preprocessed->width * mul); ao.push_back(part); }
            for (auto* c : a.cached_mains) { rd(part, c->width * mul); ao.push_back(part); }
            bc.column_openings.push_back(ao);
        }
        std::vector<EF> r;
        if (!verify_zerocheck_and_logup(ts, l_skip, D, logup_pow_bits, in.airs, in.n_per_trace, g, bc, &r)) return 0;
        memcpy(r_out, r.data(), r.size() * 16);
        store_sponge(ts, st);
        return 1;
    } catch (const std::exception&) {
        return 0;
    }
}
}  // extern "C"
