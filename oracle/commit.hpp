// ORACLE — TEST INFRASTRUCTURE ONLY (see bb31.hpp header).
//
// The stacked-PCS commitment: stacking layout, Reed–Solomon encoding of the stacked matrix and
// the Poseidon2 Merkle tree with strided query layers.  CPU restatement of
//   crates/stark-backend/src/prover/stacked_pcs.rs:116-134   stacked_commit
//   crates/stark-backend/src/prover/stacked_pcs.rs:144-203   StackedLayout::new
//   crates/stark-backend/src/prover/stacked_pcs.rs:294-335   stacked_matrix
//   crates/stark-backend/src/prover/stacked_pcs.rs:341-367   rs_code_matrix
//   crates/stark-backend/src/prover/stacked_pcs.rs:388-405   query_merkle_proof
//   crates/stark-backend/src/prover/stacked_pcs.rs:413-485   MerkleTree::new
//   crates/stark-backend/src/prover/stacked_pcs.rs:516-540   get_opened_rows
//   crates/stark-backend/src/prover/poly.rs:117-131,325-348  coeffs_to_evals_inplace,
//                                                            eval_to_coeff_rs_message
// DFT contract (p3-dft 0.4.3, not vendored): natural order in and out,
//   dft(c)[i] = sum_j c[j] * w^(i*j),  w = two_adic_generator(log2 n);   idft is its inverse.
// Golden vectors pinned: stacking matrices of stacked_pcs.rs:556-619 (tests/golden).
#pragma once
#include <algorithm>
#include <stdexcept>
#include <vector>

#include "par.hpp"
#include "poseidon2.hpp"

namespace orc {

inline int log2_strict(size_t n) {
    if (n == 0 || (n & (n - 1))) throw std::invalid_argument("not a power of two");
    int l = 0;
    while ((size_t(1) << l) < n) l++;
    return l;
}

// ------------------------------------------------------------------------------------------
// DFT
// ------------------------------------------------------------------------------------------
inline void bit_reverse_inplace(F* a, size_t n) {
    int lg = log2_strict(n);
    for (size_t i = 0; i < n; i++) {
        size_t j = 0;
        for (int b = 0; b < lg; b++) j |= ((i >> b) & 1) << (lg - 1 - b);
        if (i < j) std::swap(a[i], a[j]);
    }
}

// Per-size twiddle table w^0 .. w^(n/2-1).
inline std::vector<F> dft_twiddles(size_t n, bool inverse) {
    std::vector<F> tw(n / 2 ? n / 2 : 1);
    F w = two_adic_generator(log2_strict(n));
    if (inverse) w = f_inv(w);
    F cur = f_one();
    for (size_t i = 0; i < n / 2; i++) {
        tw[i] = cur;
        cur *= w;
    }
    return tw;
}

// In-place decimation-in-time radix-2 on a natural-order vector.
inline void dft_with_twiddles(F* a, size_t n, const std::vector<F>& tw) {
    if (n <= 1) return;
    bit_reverse_inplace(a, n);
    for (size_t half = 1; half < n; half <<= 1) {
        size_t step = n / (2 * half);  // twiddle stride
        for (size_t base = 0; base < n; base += 2 * half) {
            for (size_t k = 0; k < half; k++) {
                F u = a[base + k];
                F v = a[base + k + half] * tw[k * step];
                a[base + k] = u + v;
                a[base + k + half] = u - v;
            }
        }
    }
}

inline void dft_inplace(F* a, size_t n) { dft_with_twiddles(a, n, dft_twiddles(n, false)); }

inline void idft_inplace(F* a, size_t n) {
    dft_with_twiddles(a, n, dft_twiddles(n, true));
    F ninv = f_inv(from_canonical(n));
    for (size_t i = 0; i < n; i++) a[i] *= ninv;
}

// coset_dft(coeffs, shift)[i] = poly(shift * w^i)
inline void coset_dft_inplace(F* a, size_t n, F shift) {
    F s = f_one();
    for (size_t i = 0; i < n; i++) {
        a[i] *= s;
        s *= shift;
    }
    dft_inplace(a, n);
}

// ------------------------------------------------------------------------------------------
// Column-major matrix
// ------------------------------------------------------------------------------------------
struct ColMajor {
    std::vector<F> values;  // values[col * height + row]
    size_t height = 0, width = 0;
    ColMajor() {}
    ColMajor(size_t h, size_t w) : values(h * w), height(h), width(w) {}
    F* col(size_t c) { return values.data() + c * height; }
    const F* col(size_t c) const { return values.data() + c * height; }
};

// ------------------------------------------------------------------------------------------
// Stacking
// ------------------------------------------------------------------------------------------
struct StackedSlice {
    size_t col_idx, row_idx;
    int log_height;
    size_t len(int l_skip) const { return size_t(1) << std::max(log_height, l_skip); }
    size_t stride(int l_skip) const { return size_t(1) << (l_skip > log_height ? l_skip - log_height : 0); }
};

struct SortedCol {
    size_t mat_idx, col_in_mat;
    StackedSlice slice;
};

struct StackedLayout {
    int l_skip = 0;
    size_t height = 0, width = 0;
    std::vector<SortedCol> sorted_cols;
    std::vector<size_t> mat_starts;
};

// `sorted` = (width, log_height) per matrix, already sorted by descending log_height.
inline StackedLayout make_stacked_layout(int l_skip, int log_stacked_height,
                                         const std::vector<std::pair<size_t, int>>& sorted) {
    StackedLayout lay;
    lay.l_skip = l_skip;
    lay.height = size_t(1) << log_stacked_height;
    size_t col = 0, row = 0;
    for (size_t m = 0; m < sorted.size(); m++) {
        lay.mat_starts.push_back(lay.sorted_cols.size());
        size_t w = sorted[m].first;
        int lh = sorted[m].second;
        if (w == 0) continue;
        if (lh > log_stacked_height) throw std::invalid_argument("LayoutHeightExceeded");
        size_t slen = size_t(1) << std::max(lh, l_skip);
        for (size_t j = 0; j < w; j++) {
            if (row + slen > lay.height) {
                if (row != lay.height) throw std::invalid_argument("LayoutRowOverflow");
                col++;
                row = 0;
            }
            lay.sorted_cols.push_back({m, j, {col, row, lh}});
            row += slen;
        }
    }
    lay.width = col + (row != 0 ? 1 : 0);
    return lay;
}

inline ColMajor stacked_matrix(int l_skip, int n_stack, const std::vector<const ColMajor*>& traces,
                               StackedLayout* out_layout) {
    std::vector<std::pair<size_t, int>> meta;
    size_t total_cells = 0;
    for (auto* t : traces) {
        meta.push_back({t->width, log2_strict(t->height)});
        total_cells += std::max(t->height, size_t(1) << l_skip) * t->width;
    }
    StackedLayout lay = make_stacked_layout(l_skip, l_skip + n_stack, meta);
    size_t H = size_t(1) << (l_skip + n_stack);
    size_t W = (total_cells + H - 1) / H;
    ColMajor q(H, W);
    for (auto& sc : lay.sorted_cols) {
        const ColMajor* t = traces[sc.mat_idx];
        const F* src = t->col(sc.col_in_mat);
        F* dst = q.values.data() + sc.slice.col_idx * H + sc.slice.row_idx;
        size_t st = sc.slice.stride(l_skip);
        for (size_t i = 0; i < t->height; i++) dst[i * st] = src[i];
    }
    if (out_layout) *out_layout = lay;
    return q;
}

// ------------------------------------------------------------------------------------------
// Reed–Solomon encoding
// ------------------------------------------------------------------------------------------
// coeffs -> evals over the boolean cube (subset-zeta), poly.rs:117-131
inline void coeffs_to_evals_inplace(F* a, size_t n) {
    for (size_t step = 1; step < n; step <<= 1)
        for (size_t i = 0; i < n; i += 2 * step)
            for (size_t j = 0; j < step; j++) a[i + j + step] += a[i + j];
}

// poly.rs:325-348 — per 2^l_skip chunk: iDFT in Z, then zeta over the Z-index bits.
inline void eval_to_coeff_rs_message_inplace(int l_skip, F* a, size_t n) {
    size_t chunk = size_t(1) << l_skip;
    if (n < chunk) throw std::invalid_argument("prism dim < l_skip");
    std::vector<F> tw = dft_twiddles(chunk, true);
    F cinv = f_inv(from_canonical(chunk));
    for (size_t off = 0; off < n; off += chunk) {
        dft_with_twiddles(a + off, chunk, tw);
        for (size_t i = 0; i < chunk; i++) a[off + i] *= cinv;
        coeffs_to_evals_inplace(a + off, chunk);
    }
}

inline ColMajor rs_code_matrix(int l_skip, int log_blowup, const ColMajor& evals) {
    size_t H = evals.height, N = H << log_blowup;
    ColMajor out(N, evals.width);
    std::vector<F> tw = dft_twiddles(N, false);
    parallel_for(evals.width, [&](size_t c0, size_t c1) {
        for (size_t c = c0; c < c1; c++) {
            F* dst = out.col(c);
            std::copy(evals.col(c), evals.col(c) + H, dst);
            eval_to_coeff_rs_message_inplace(l_skip, dst, H);
            // dst[H..N) already zero
            dft_with_twiddles(dst, N, tw);
        }
    });
    return out;
}

// ------------------------------------------------------------------------------------------
// Merkle tree
// ------------------------------------------------------------------------------------------
struct MerkleTree {
    ColMajor backing;
    std::vector<std::vector<Digest>> layers;  // layers[0] has query_stride entries
    size_t rows_per_query = 1;
    size_t query_stride() const { return layers[0].size(); }
    size_t proof_depth() const { return layers.size() - 1; }
    Digest root() const { return layers.back()[0]; }

    std::vector<Digest> query_merkle_proof(size_t query_idx) const {
        if (query_idx >= query_stride()) throw std::out_of_range("MerkleTreeQueryOutOfBounds");
        std::vector<Digest> proof;
        size_t idx = query_idx;
        for (size_t l = 0; l < proof_depth(); l++) {
            proof.push_back(layers[l][idx ^ 1]);
            idx >>= 1;
        }
        return proof;
    }
    // rows { index + t * query_stride }, t < rows_per_query; each row has `width` entries
    std::vector<std::vector<F>> get_opened_rows(size_t index) const {
        if (index >= query_stride()) throw std::out_of_range("MerkleTreeOpenedRowsOutOfBounds");
        std::vector<std::vector<F>> rows;
        for (size_t t = 0; t < rows_per_query; t++) {
            size_t r = t * query_stride() + index;
            std::vector<F> row(backing.width);
            for (size_t c = 0; c < backing.width; c++)
                row[c] = r < backing.height ? backing.values[c * backing.height + r] : f_zero();
            rows.push_back(row);
        }
        return rows;
    }
};

// `ext_degree` = 1 for a base-field matrix; 4 when `matrix` holds EF columns flattened so that
// the 4 basis coefficients of an element are 4 consecutive base columns in the hash input
// (stacked_pcs.rs:438-441).  The caller flattens; here a row is simply `width` base elements.
inline MerkleTree merkle_tree_new(ColMajor matrix, size_t rows_per_query) {
    size_t height = matrix.height;
    if (height == 0) throw std::invalid_argument("MerkleTreeEmptyMatrix");
    if (rows_per_query == 0 || (rows_per_query & (rows_per_query - 1)))
        throw std::invalid_argument("MerkleTreeRowsPerQueryNotPow2");
    size_t num_leaves = 1;
    while (num_leaves < height) num_leaves <<= 1;
    if (rows_per_query > num_leaves) throw std::invalid_argument("MerkleTreeRowsPerQueryExceeded");
    size_t W = matrix.width;
    std::vector<Digest> cur(num_leaves);
    parallel_for(num_leaves, [&](size_t r0, size_t r1) {
        // blocks of rows so that column-major reads stay cache friendly
        const size_t RB = 64;
        std::vector<F> rows(RB * W);
        for (size_t rb = r0; rb < r1; rb += RB) {
            size_t nb = std::min(RB, r1 - rb);
            for (size_t c = 0; c < W; c++)
                for (size_t i = 0; i < nb; i++)
                    rows[i * W + c] = rb + i < height ? matrix.values[c * height + rb + i] : f_zero();
            for (size_t i = 0; i < nb; i++) cur[rb + i] = hash_slice(rows.data() + i * W, W);
        }
    }, 64);
    size_t qs = num_leaves / rows_per_query;
    for (size_t lvl = 1; lvl < rows_per_query; lvl <<= 1) {
        std::vector<Digest> nxt(cur.size() / 2);
        parallel_for(nxt.size(), [&](size_t i0, size_t i1) {
            for (size_t i = i0; i < i1; i++) {
                size_t x = i / qs, y = i % qs;
                nxt[i] = compress(cur[2 * x * qs + y], cur[(2 * x + 1) * qs + y]);
            }
        }, 256);
        cur.swap(nxt);
    }
    MerkleTree t;
    t.rows_per_query = rows_per_query;
    t.layers.push_back(std::move(cur));
    while (t.layers.back().size() > 1) {
        const std::vector<Digest>& prev = t.layers.back();
        std::vector<Digest> nxt(prev.size() / 2);
        parallel_for(nxt.size(), [&](size_t i0, size_t i1) {
            for (size_t i = i0; i < i1; i++) nxt[i] = compress(prev[2 * i], prev[2 * i + 1]);
        }, 256);
        t.layers.push_back(std::move(nxt));
    }
    t.backing = std::move(matrix);
    return t;
}

struct StackedPcsData {
    StackedLayout layout;
    ColMajor matrix;  // stacked evaluations, height 2^(l_skip+n_stack)
    MerkleTree tree;
};

inline Digest stacked_commit(int l_skip, int n_stack, int log_blowup, int k_whir,
                             const std::vector<const ColMajor*>& traces, StackedPcsData* out) {
    StackedPcsData d;
    d.matrix = stacked_matrix(l_skip, n_stack, traces, &d.layout);
    ColMajor rs = rs_code_matrix(l_skip, log_blowup, d.matrix);
    d.tree = merkle_tree_new(std::move(rs), size_t(1) << k_whir);
    Digest root = d.tree.root();
    if (out) *out = std::move(d);
    return root;
}

}  // namespace orc
