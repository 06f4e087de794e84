// Host side of the Fiat–Shamir transcript used by the phase-level entry points: an overwrite-mode
// duplex sponge (rate 8 / width 16) over the same `p2::permute` the kernels run (it is
// __host__ __device__), so challenges are derived next to the launches that need them.
//
// Replaces (reference, relative to /root/reference):
//   crates/stark-backend/src/transcript/duplex_sponge.rs:60-83      absorb / squeeze
//   crates/stark-backend/src/transcript/traits.rs:11-90             observe_ext / sample_ext / sample_bits /
//                                                                   check_witness / grind
//   crates/cuda-backend/src/sponge.rs:267-300                       DuplexSpongeGpu (state = 16 words + 2 indices)
// The POD layout (16 state words, absorb_idx, sample_idx) is the reference's DeviceSpongeState
// (cuda-backend/cuda/src/sponge.cu:13-17) and is what crosses the C ABI as `swirl_transcript`.
#pragma once
#include "../../include/swirl_b200.h"
#include "poseidon2.cuh"
#include "poseidon2_host.hpp"

namespace swirl {

struct Transcript {
    swirl_transcript* t;
    explicit Transcript(swirl_transcript* p) : t(p) {}

    void permute() {
        p2host::permute(t->state);  // AVX2 on the host (poseidon2_host.hpp), same function as p2::permute
        t->absorb_idx = 0;
        t->sample_idx = 8;
    }
    void observe(uint32_t v) {
        t->state[t->absorb_idx++] = v;
        if (t->absorb_idx == 8) permute();
    }
    uint32_t sample() {
        if (t->absorb_idx != 0 || t->sample_idx == 0) permute();
        return t->state[--t->sample_idx];
    }
    void observe_ext(const bb::Ext& e) {
        for (int i = 0; i < 4; i++) observe(e.c[i]);
    }
    void observe_digest(const uint32_t d[8]) {
        for (int i = 0; i < 8; i++) observe(d[i]);
    }
    bb::Ext sample_ext() {
        bb::Ext e;
        for (int i = 0; i < 4; i++) e.c[i] = sample();
        return e;
    }
    uint32_t sample_bits(int bits) { return bb::from_mont(sample()) & (uint32_t)((uint64_t(1) << bits) - 1); }
    bool check_witness(int bits, uint32_t w_mont) {
        if (bits == 0) return true;
        observe(w_mont);
        return sample_bits(bits) == 0;
    }
};

// grind(bits): smallest valid canonical witness; searched on the host for small `bits` (a launch +
// sync costs more than 2^bits scalar permutations) and by grind_kernel otherwise.  Mutates the
// transcript with the witness and returns it in Montgomery form through *w_mont.
int transcript_grind(swirl_ctx* ctx, swirl_transcript* t, int bits, uint32_t* w_mont);

}  // namespace swirl
