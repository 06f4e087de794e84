//! Flat proof sections of libswirl_b200 (Montgomery words, layouts documented in include/swirl_b200.h) -> the reference's
//! proof structs (crates/stark-backend/src/proof.rs:70-199).  stark-backend_b200/codec.py is the executable description
//! of the same offsets (its encoder is checked byte for byte against `Proof::encode_to_vec()` as restated from
//! proof.rs:204-420 in tests/test_codec.py).
use openvm_stark_backend::{
    proof::{BatchConstraintProof, GkrLayerClaims, GkrProof, StackingProof, WhirProof},
    prover::{DeviceMultiStarkProvingKey, MatrixDimensions, ProvingContext},
    SystemParams,
};
use openvm_stark_sdk::config::baby_bear_poseidon2::{BabyBearPoseidon2Config as SC, Digest, EF, F};
use p3_field::BasedVectorSpace;

use crate::{f_from_word, f_to_word, B200Backend};

pub fn ef(w: &[u32]) -> EF {
    EF::from_basis_coefficients_fn(|i| f_from_word(w[i]))
}
pub fn ef_vec(w: &[u32]) -> Vec<EF> {
    w.chunks_exact(4).map(ef).collect()
}
pub fn ef_words(e: &EF) -> [u32; 4] {
    let c: &[F] = e.as_basis_coefficients_slice();
    [f_to_word(c[0]), f_to_word(c[1]), f_to_word(c[2]), f_to_word(c[3])]
}
fn digest(w: &[u32]) -> Digest {
    std::array::from_fn(|i| f_from_word(w[i]))
}

/// Cursor over a flat section.
struct Rd<'a>(&'a [u32]);
impl<'a> Rd<'a> {
    fn take(&mut self, n: usize) -> &'a [u32] {
        let (a, b) = self.0.split_at(n);
        self.0 = b;
        a
    }
    fn f(&mut self) -> F {
        f_from_word(self.take(1)[0])
    }
    fn ef(&mut self) -> EF {
        ef(self.take(4))
    }
    fn efs(&mut self, n: usize) -> Vec<EF> {
        ef_vec(self.take(4 * n))
    }
    fn digest(&mut self) -> Digest {
        digest(self.take(8))
    }
}

/// What the batch-constraint section's lengths depend on (swirl_batch_constraints_proof_words).
pub struct BatchShape {
    pub gkr_layers: usize,       // L = l_skip + n_logup, 0 without interactions (calculate_n_logup, lib.rs:82-93)
    pub n_airs: usize,
    pub uni_coeffs: usize,       // (D + 1)(2^l_skip - 1) + 1
    pub n_max: usize,
    pub round_evals: usize,      // D + 1
    pub part_openings: Vec<Vec<usize>>, // per AIR (sorted order), per part: width * (need_rot ? 2 : 1)
}

impl BatchShape {
    pub fn new(params: &SystemParams, mpk: &DeviceMultiStarkProvingKey<B200Backend>, ctx: &ProvingContext<B200Backend>) -> Self {
        let l_skip = params.l_skip;
        let d = mpk.max_constraint_degree;
        let mut total_interactions = 0u64;
        let mut n_max = 0usize;
        let mut part_openings = Vec::new();
        for (air_id, air) in &ctx.per_trace {
            let pk = &mpk.per_air[*air_id];
            let log_h = air.common_main.height().trailing_zeros() as usize;
            total_interactions += (pk.vk.symbolic_constraints.interactions.len() as u64) << log_h.max(l_skip);
            n_max = n_max.max(log_h.saturating_sub(l_skip));
            let rot = if pk.vk.params.need_rot { 2 } else { 1 };
            // parts: common main, preprocessed (if any), cached.. (proof.rs:128-130)
            let mut parts = vec![air.common_main.width() * rot];
            if let Some(p) = &pk.preprocessed_data {
                parts.push(p.trace.width() * rot);
            }
            parts.extend(air.cached_mains.iter().map(|c| c.trace.width() * rot));
            part_openings.push(parts);
        }
        let gkr_layers = if total_interactions == 0 { 0 } else { (64 - total_interactions.leading_zeros()) as usize };
        Self {
            gkr_layers,
            n_airs: ctx.per_trace.len(),
            uni_coeffs: (d + 1) * ((1 << l_skip) - 1) + 1,
            n_max,
            round_evals: d + 1,
            part_openings,
        }
    }
}

/// logup_pow_witness[1] | q0_claim[4] | claims_per_layer[L][16] | sumcheck_polys[L(L-1)/2][12] | numerator_term_per_air[n][4]
/// | denominator_term_per_air[n][4] | univariate_round_coeffs | sumcheck_round_polys[n_max][D+1][4] | column_openings
pub fn split_gkr_and_batch(flat: &[u32], s: &BatchShape) -> (GkrProof<SC>, BatchConstraintProof<SC>) {
    let mut r = Rd(flat);
    let logup_pow_witness = r.f();
    let q0_claim = r.ef();
    let claims_per_layer = (0..s.gkr_layers)
        .map(|_| {
            // the library stores a layer's claims in transcript order p(xi,0), q(xi,0), p(xi,1), q(xi,1)
            // (fractional_sumcheck_gkr.rs observes p_xi_0, q_xi_0, p_xi_1, q_xi_1)
            let (p0, q0, p1, q1) = (r.ef(), r.ef(), r.ef(), r.ef());
            GkrLayerClaims { p_xi_0: p0, p_xi_1: p1, q_xi_0: q0, q_xi_1: q1 }
        })
        .collect();
    // layer j (1-based) has j - 1... layers 1..L-1 have 1..L-1 rounds: round r of the section belongs to the layer with r rounds
    let sumcheck_polys = (1..s.gkr_layers)
        .map(|rounds| (0..rounds).map(|_| [r.ef(), r.ef(), r.ef()]).collect())
        .collect();
    let gkr = GkrProof { logup_pow_witness, q0_claim, claims_per_layer, sumcheck_polys };
    let numerator_term_per_air = r.efs(s.n_airs);
    let denominator_term_per_air = r.efs(s.n_airs);
    let univariate_round_coeffs = r.efs(s.uni_coeffs);
    let sumcheck_round_polys = (0..s.n_max).map(|_| r.efs(s.round_evals)).collect();
    let column_openings = s.part_openings.iter().map(|parts| parts.iter().map(|&n| r.efs(n)).collect()).collect();
    debug_assert!(r.0.is_empty());
    (gkr, BatchConstraintProof { numerator_term_per_air, denominator_term_per_air, univariate_round_coeffs, sumcheck_round_polys, column_openings })
}

/// univariate_round_coeffs[2(2^l_skip - 1) + 1][4] | sumcheck_round_polys[n_stack][2][4] | stacking_openings per commit
pub fn split_stacking(flat: &[u32], params: &SystemParams, widths: &[usize]) -> StackingProof<SC> {
    let mut r = Rd(flat);
    let univariate_round_coeffs = r.efs(2 * ((1 << params.l_skip) - 1) + 1);
    let sumcheck_round_polys = (0..params.n_stack).map(|_| [r.ef(), r.ef()]).collect();
    let stacking_openings = widths.iter().map(|&w| r.efs(w)).collect();
    debug_assert!(r.0.is_empty());
    StackingProof { univariate_round_coeffs, sumcheck_round_polys, stacking_openings }
}

/// Layout of swirl_whir_proof_words (include/swirl_b200.h).
pub fn split_whir(flat: &[u32], params: &SystemParams, widths: &[usize]) -> WhirProof<SC> {
    let w = &params.whir;
    let (k, rounds) = (w.k, w.rounds.len());
    let m = params.l_skip + params.n_stack;
    let mut r = Rd(flat);
    let mu_pow_witness = r.f();
    let whir_sumcheck_polys = (0..rounds * k).map(|_| [r.ef(), r.ef()]).collect();
    let codeword_commits = (0..rounds - 1).map(|_| r.digest()).collect();
    let ood_values = r.efs(rounds - 1);
    let folding_pow_witnesses = (0..rounds * k).map(|_| r.f()).collect();
    let query_phase_pow_witnesses = (0..rounds).map(|_| r.f()).collect();
    let q0 = w.rounds[0].num_queries;
    let initial_round_opened_rows = widths
        .iter()
        .map(|&width| (0..q0).map(|_| (0..1usize << k).map(|_| r.take(width).iter().map(|&x| f_from_word(x)).collect()).collect()).collect())
        .collect();
    let depth0 = m + params.log_blowup - k;
    let initial_round_merkle_proofs = widths.iter().map(|_| (0..q0).map(|_| (0..depth0).map(|_| r.digest()).collect()).collect()).collect();
    let codeword_opened_values = (1..rounds)
        .map(|i| (0..w.rounds[i].num_queries).map(|_| r.efs(1 << k)).collect())
        .collect();
    let codeword_merkle_proofs = (1..rounds)
        .map(|i| (0..w.rounds[i].num_queries).map(|_| (0..depth0 - i).map(|_| r.digest()).collect()).collect())
        .collect();
    let final_poly = r.efs(1 << (m - rounds * k));
    debug_assert!(r.0.is_empty());
    WhirProof {
        mu_pow_witness,
        whir_sumcheck_polys,
        codeword_commits,
        ood_values,
        folding_pow_witnesses,
        query_phase_pow_witnesses,
        initial_round_opened_rows,
        initial_round_merkle_proofs,
        codeword_opened_values,
        codeword_merkle_proofs,
        final_poly,
    }
}
