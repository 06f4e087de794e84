// ORACLE — TEST INFRASTRUCTURE ONLY (see bb31.hpp header).
//
// Fiat–Shamir transcript: overwrite-mode duplex sponge (rate 8 / width 16) over Poseidon2, plus
// bit sampling and proof-of-work grinding.  CPU restatement of
//   crates/stark-backend/src/transcript/duplex_sponge.rs:60-115   absorb / squeeze
//   crates/stark-backend/src/transcript/traits.rs:11-90           observe_ext, sample_ext,
//                                                                 sample_bits, check_witness, grind
// (must equal p3-challenger 0.4.3 DuplexChallenger — pinned in the reference by
// crates/stark-backend/tests/transcript.rs:20-42, which cannot run here; PARITY UNPINNED for
// challenge values, see poseidon2.hpp).
//
// grind(): the reference takes *any* witness (`par_iter().find_any`); the oracle is
// deterministic and returns the smallest valid witness, which is one of the reference's
// admissible outputs.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <vector>

#include "par.hpp"
#include "poseidon2.hpp"

namespace orc {

struct DuplexSponge {
    F state[16] = {};
    uint32_t absorb_idx = 0;  // 0 <= absorb_idx < 8
    uint32_t sample_idx = 0;  // 0 <= sample_idx <= 8

    void observe(F v) {
        state[absorb_idx++] = v;
        if (absorb_idx == P2_RATE) {
            poseidon2_permute(state);
            absorb_idx = 0;
            sample_idx = P2_RATE;
        }
    }
    F sample() {
        if (absorb_idx != 0 || sample_idx == 0) {
            poseidon2_permute(state);
            absorb_idx = 0;
            sample_idx = P2_RATE;
        }
        return state[--sample_idx];
    }
    void observe_commit(const Digest& d) {
        for (int i = 0; i < 8; i++) observe(d.w[i]);
    }
    void observe_ext(const EF& e) {
        for (int i = 0; i < 4; i++) observe(e.c[i]);
    }
    EF sample_ext() {
        EF e;
        for (int i = 0; i < 4; i++) e.c[i] = sample();
        return e;
    }
    uint64_t sample_bits(int bits) { return (uint64_t)to_canonical(sample()) & ((uint64_t(1) << bits) - 1); }
    bool check_witness(int bits, F w) {
        if (bits == 0) return true;
        observe(w);
        return sample_bits(bits) == 0;
    }
    // smallest canonical witness >= start that passes; mutates the transcript with it
    F grind(int bits, uint32_t start = 0) {
        if (bits == 0) return f_zero();
        // windows of candidates searched by all host threads (the reference searches with rayon); the smallest
        // witness of the first window that holds one is the smallest witness overall
        const uint64_t window = uint64_t(4096) * par_threads();
        for (uint64_t base = start; base < P; base += window) {
            const uint64_t n = std::min<uint64_t>(window, P - base);
            std::atomic<uint64_t> found{UINT64_MAX};
            const DuplexSponge snapshot = *this;
            parallel_for(n, [&](size_t i0, size_t i1) {
                for (size_t i = i0; i < i1; i++) {
                    DuplexSponge probe = snapshot;
                    if (probe.check_witness(bits, from_canonical(base + i))) {
                        uint64_t cur = found.load();
                        while (base + i < cur && !found.compare_exchange_weak(cur, base + i)) {}
                        return;
                    }
                }
            }, 256);
            const uint64_t best = found.load();
            if (best != UINT64_MAX) {
                const F w = from_canonical(best);
                const bool ok = check_witness(bits, w);
                (void)ok;
                return w;
            }
        }
        throw std::runtime_error("failed to find PoW witness");
    }
};

}  // namespace orc
