// ORACLE — TEST INFRASTRUCTURE ONLY (see bb31.hpp header).
// Minimal fork-join helper (this image has no libgomp): static block partition over std::thread.
// Thread count: ORC_THREADS env var, else hardware_concurrency.
#pragma once
#include <cstdlib>
#include <thread>
#include <vector>

namespace orc {

inline unsigned par_threads() {
    static unsigned n = [] {
        const char* e = std::getenv("ORC_THREADS");
        unsigned v = e ? (unsigned)std::atoi(e) : std::thread::hardware_concurrency();
        return v ? v : 1u;
    }();
    return n;
}

// body(begin, end) over a partition of [0, n)
template <class Body>
inline void parallel_for(size_t n, Body body, size_t min_chunk = 1) {
    unsigned nt = par_threads();
    if (nt <= 1 || n <= min_chunk) {
        body(size_t(0), n);
        return;
    }
    size_t chunks = (n + min_chunk - 1) / min_chunk;
    if (chunks < nt) nt = (unsigned)chunks;
    std::vector<std::thread> th;
    th.reserve(nt);
    for (unsigned t = 0; t < nt; t++) {
        size_t b = n * t / nt, e = n * (t + 1) / nt;
        th.emplace_back([=] { body(b, e); });
    }
    for (auto& x : th) x.join();
}

// how many workers parallel_for / parallel_for_tid would use for n items
inline unsigned par_workers(size_t n, size_t min_chunk = 1) {
    unsigned nt = par_threads();
    if (nt <= 1 || n <= min_chunk) return 1;
    size_t chunks = (n + min_chunk - 1) / min_chunk;
    return chunks < nt ? (unsigned)chunks : nt;
}

// body(worker, begin, end) over a partition of [0, n) into `nt` = par_workers(n, min_chunk) blocks: for reductions
// with one partial accumulator per worker (field addition is exact and associative, so the result does not depend
// on the partition).
template <class Body>
inline void parallel_for_tid(size_t n, unsigned nt, Body body) {
    if (nt <= 1) {
        body(0u, size_t(0), n);
        return;
    }
    std::vector<std::thread> th;
    th.reserve(nt);
    for (unsigned t = 0; t < nt; t++) {
        size_t b = n * t / nt, e = n * (t + 1) / nt;
        th.emplace_back([=] { body(t, b, e); });
    }
    for (auto& x : th) x.join();
}

}  // namespace orc
