// OpeningProver::prove_openings = stacked opening reduction, then WHIR at u_cube.
// Replaces (reference): crates/cuda-backend/src/gpu_backend.rs:150-212 (prove_openings);
// semantics = crates/stark-backend/src/prover/cpu_backend.rs:139-220.
#include <vector>

#include "hostpoly.hpp"
#include "kernels.cuh"
#include "pcs.cuh"

using namespace swirl;

extern "C" int swirl_prove_openings(swirl_ctx* ctx, swirl_transcript* ts, const swirl_whir_config* cfg,
                                    const swirl_pcs* const* pcs, size_t n_commits, const uint8_t* const* need_rot,
                                    const uint32_t* h_r, size_t r_len, uint32_t* h_stacking_proof, size_t stacking_words,
                                    uint32_t* h_whir_proof, size_t whir_words) {
    SWIRL_REQUIRE(ctx && ts && cfg && pcs && n_commits >= 1 && pcs[0], "null argument");
    const int l_skip = pcs[0]->params.l_skip, n_stack = pcs[0]->params.n_stack;
    std::vector<uint32_t> u((size_t)(n_stack + 1) * 4);
    SWIRL_TRY(swirl_stacked_reduction(ctx, ts, pcs, n_commits, need_rot, h_r, r_len, h_stacking_proof, stacking_words, u.data()));
    // u_cube = (u_0, u_0^2, .., u_0^(2^(l_skip-1)), u_1, .., u_n_stack)  (cpu_backend.rs:203-210)
    std::vector<uint32_t> u_cube((size_t)(l_skip + n_stack) * 4);
    bb::Ext p = hp::from_words(u.data());
    for (int i = 0; i < l_skip; i++) {
        memcpy(&u_cube[4 * i], p.c, 16);
        p = bb::ext_sqr(p);
    }
    memcpy(&u_cube[4 * (size_t)l_skip], &u[4], (size_t)n_stack * 16);
    return swirl_whir_open(ctx, ts, cfg, pcs, n_commits, u_cube.data(), h_whir_proof, whir_words);
}
