// Proof-of-work grinding for the duplex-sponge transcript.
//
// Replaces (reference, relative to /root/reference):
//   crates/cuda-backend/cuda/src/sponge.cu:65-117     grind_kernel / _sponge_grind
//   crates/cuda-backend/src/sponge.rs:267-300          DuplexSpongeGpu::grind_gpu
// Semantics = crates/stark-backend/src/transcript/traits.rs:63-86 (check_witness / grind) over
// transcript/duplex_sponge.rs:60-83.  Whatever the absorb position, observing the witness and then
// sampling costs exactly one permutation and the sampled word is state[7] of the permuted state
// (observe fills slot absorb_idx; either that completes the rate block and permutes, or the
// following sample permutes; both leave sample_idx = 8 -> returns state[7]).
//
// Unlike the reference (first finder wins, non-deterministic) this search is deterministic: it
// returns the smallest valid witness, scanning ascending windows and taking an atomicMin inside
// the first window that contains a hit.
#include "kernels.cuh"
#include "poseidon2.cuh"
#include "poseidon2_v2.cuh"

namespace swirl {

struct GrindState {
    uint32_t s[16];
    uint32_t slot;
};

__global__ void __launch_bounds__(256)
grind_kernel(GrindState st, uint32_t mask, uint32_t w0, uint32_t w_end, uint32_t* __restrict__ result) {
    const uint32_t w = w0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= w_end) return;
    // blocks are scheduled in ascending order: once a smaller witness is known, the rest of the window is skipped
    if (__ldcg(result) < w0 + blockIdx.x * blockDim.x) return;
    uint32_t s[16];
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = st.s[i];
    const uint32_t wm = bb::to_mont(w);
#pragma unroll
    for (int i = 0; i < 8; i++)
        if ((uint32_t)i == st.slot) s[i] = wm;
    p2v2::permute(s);
    if ((bb::from_mont(s[7]) & mask) == 0) atomicMin(result, w);
}

}  // namespace swirl

using namespace swirl;

extern "C" int swirl_sponge_grind(swirl_ctx* ctx, const uint32_t h_state[18], int bits, uint32_t min_w,
                                  uint32_t max_w, uint32_t* h_witness) {
    SWIRL_REQUIRE(ctx && h_state && h_witness, "null argument");
    SWIRL_REQUIRE(bits >= 0 && bits < 31, "bits");
    SWIRL_REQUIRE(h_state[16] < 8 && h_state[17] <= 8, "sponge indices");
    if (max_w > bb::P) max_w = bb::P;
    *h_witness = 0xffffffffu;
    if (bits == 0) {
        *h_witness = 0;  // grind(0) returns ZERO without touching the transcript (traits.rs:78-80)
        return 0;
    }
    SWIRL_CUDA(cudaSetDevice(ctx->device));
    GrindState st;
    for (int i = 0; i < 16; i++) st.s[i] = h_state[i];
    st.slot = h_state[16];
    const uint32_t mask = (1u << bits) - 1;
    uint32_t* d_res = nullptr;
    SWIRL_CUDA(dev_alloc(ctx, &d_res, 1));
    // ascending windows of 2^bits candidates: each holds a witness with probability 1 - 1/e, so the
    // expected work is ~1.6 * 2^bits permutations (the smallest witness is in the first non-empty window)
    uint64_t window = uint64_t(1) << bits;
    if (window < (1u << 16)) window = 1u << 16;
    if (window > (1u << 24)) window = 1u << 24;
    int rc = 0;
    for (uint64_t w0 = min_w; w0 < max_w; w0 += window) {
        const uint32_t w_end = (uint32_t)(w0 + window < max_w ? w0 + window : max_w);
        cudaError_t e = cudaMemsetAsync(d_res, 0xff, 4, ctx->stream);
        if (e != cudaSuccess) {
            rc = cuda_fail(e, "memset", __FILE__, __LINE__);
            break;
        }
        const uint32_t cnt = w_end - (uint32_t)w0;
        grind_kernel<<<(cnt + 255) / 256, 256, 0, ctx->stream>>>(st, mask, (uint32_t)w0, w_end, d_res);
        ctx->launches++;
        uint32_t found = 0xffffffffu;
        e = cudaMemcpyAsync(&found, d_res, 4, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            rc = cuda_fail(e, "grind", __FILE__, __LINE__);
            break;
        }
        if (found != 0xffffffffu) {
            *h_witness = found;
            break;
        }
    }
    dev_free(ctx, d_res);
    return rc;
}

// ---- host transcript entry points (transcript.hpp) --------------------------------------------
#include <emmintrin.h>

#include "ext.cuh"
#include "transcript.hpp"

namespace swirl {

// Waits up to ~25 ms for the host's word: used once per context to find out whether launches are asynchronous at all.
__global__ void link_probe_kernel(const uint32_t* mail, uint32_t want, uint32_t* got) {
    const long long t0 = clock64();
    uint4 v;
    uint32_t tag;
    do {
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(mail));
        tag = (v.x >> 31) | ((v.y >> 31) << 1) | ((v.z >> 31) << 2) | ((v.w >> 31) << 3);
    } while (tag != want && clock64() - t0 < 50000000ll);
    *got = tag == want ? 1u : 0u;
}

// The round link needs the host to run WHILE an enqueued kernel waits for it.  Under a profiler or a sanitizer that
// serialises launches (ncu, compute-sanitizer, CUDA_LAUNCH_BLOCKING=1) the launch call itself blocks until the kernel
// ends, and every linked round would sit out its device-side time-out.  One probe per context: a kernel that waits for
// a word the host sends only after the launch call has returned.
static void link_probe(swirl_ctx* ctx, RoundScratch* rs) {
    if (!ctx->round_link) return;
    const RoundLink l = link_make(rs, true);
    uint32_t* d_got = rs->d_gate + 6;
    const auto t0 = std::chrono::steady_clock::now();
    link_probe_kernel<<<1, 1, 0, ctx->stream>>>(rs->d_link, link_mail_tag(l.seq), d_got);
    const double launch_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    link_send(rs, l.seq, bb::ext_zero());
    uint32_t got = 0;
    cudaError_t e = cudaMemcpyAsync(&got, d_got, 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess || !got || launch_ms > 10.0) {
        ctx->round_link = false;
        ctx->round_link_ok = false;
        if (getenv("SWIRL_TRACE"))
            fprintf(stderr, "[swirl] round link off: launches are serialised here (probe launch took %.1f ms, answered %u)\n", launch_ms, got);
    }
}

int round_scratch_get(swirl_ctx* ctx, RoundScratch** out) {
    if (!ctx->round_scratch) {
        RoundScratch* rs = new RoundScratch();
        rs->max_blocks = 8192;
        SWIRL_CUDA(cudaMalloc((void**)&rs->d_partials, (size_t)rs->max_blocks * 64 * sizeof(uint32_t)));
        SWIRL_CUDA(cudaMalloc((void**)&rs->d_ticket, 1024 * sizeof(unsigned int)));
        SWIRL_CUDA(cudaMemset(rs->d_ticket, 0, 1024 * sizeof(unsigned int)));
        SWIRL_CUDA(cudaHostAlloc((void**)&rs->h_result, 65536, cudaHostAllocMapped));
        memset(rs->h_result, 0, 65536);
        SWIRL_CUDA(cudaHostGetDevicePointer((void**)&rs->d_result, rs->h_result, 0));
        SWIRL_CUDA(cudaHostAlloc((void**)&rs->h_link, 4096, cudaHostAllocMapped));
        memset(rs->h_link, 0, 4096);
        SWIRL_CUDA(cudaHostGetDevicePointer((void**)&rs->d_link, rs->h_link, 0));
        SWIRL_CUDA(cudaMalloc((void**)&rs->d_gate, 8 * sizeof(uint32_t)));
        SWIRL_CUDA(cudaMemset(rs->d_gate, 0, 8 * sizeof(uint32_t)));
        ctx->round_scratch = rs;
        link_probe(ctx, rs);
    }
    *out = (RoundScratch*)ctx->round_scratch;
    return 0;
}

void link_begin(swirl_ctx* ctx, RoundScratch* rs, size_t offset, size_t nv) {
    if (rs->link_seq >= 0xfffffff0u) rs->link_seq = 0;
    const uint32_t not_ready = link_result_tag(rs->link_seq + 1) ^ 0x80000000u;
    volatile uint32_t* r = rs->h_result + offset;
    for (size_t i = 0; i < nv; i++) r[i] = not_ready;
    // idle: tag 14 = bits 1, 2, 3
    _mm_store_si128(reinterpret_cast<__m128i*>(rs->h_link), _mm_set_epi32((int)0x80000000u, (int)0x80000000u, (int)0x80000000u, 0));
    __atomic_thread_fence(__ATOMIC_SEQ_CST);
    // relay = idle as well, abort flag cleared (every earlier linked kernel has finished: its round was consumed)
    static const uint32_t idle_gate[8] = {0, 0x80000000u, 0x80000000u, 0x80000000u, 0, 0, 0, 0};
    cudaMemcpyAsync(rs->d_gate, idle_gate, sizeof(idle_gate), cudaMemcpyHostToDevice, ctx->stream);
}

void link_send(RoundScratch* rs, uint32_t seq, const Ext& r) {
    const uint32_t t = link_mail_tag(seq);
    alignas(16) uint32_t w[4];
    for (int k = 0; k < 4; k++) w[k] = r.c[k] | (((t >> k) & 1u) << 31);
    _mm_store_si128(reinterpret_cast<__m128i*>(rs->h_link), _mm_load_si128(reinterpret_cast<const __m128i*>(w)));  // one 16-byte store
    __atomic_thread_fence(__ATOMIC_SEQ_CST);
}
LinkAbortGuard::~LinkAbortGuard() {
    if (!armed) return;
    link_abort(rs);
    cudaStreamSynchronize(ctx->stream);
}
void link_abort(RoundScratch* rs) {
    _mm_store_si128(reinterpret_cast<__m128i*>(rs->h_link), _mm_set1_epi32((int)0x80000000u));  // tag 15
    __atomic_thread_fence(__ATOMIC_SEQ_CST);
}

int link_recv(swirl_ctx* ctx, RoundScratch* rs, uint32_t seq, size_t offset, int nv, size_t groups, size_t stride, uint32_t* out) {
    const volatile uint32_t* res = rs->h_result + offset;
    const uint32_t tag = link_result_tag(seq);
    const auto t0 = std::chrono::steady_clock::now();
    auto next_query = t0 + std::chrono::milliseconds(250);
    ctx->link_count++;
    auto ready = [&]() {
        for (size_t g = 0; g < groups; g++)
            for (int i = 0; i < nv; i++)
                if ((res[g * stride + i] & 0x80000000u) != tag) return false;
        __atomic_thread_fence(__ATOMIC_ACQUIRE);
        for (size_t g = 0; g < groups; g++)
            for (int i = 0; i < nv; i++) out[g * stride + i] = res[g * stride + i] & 0x7fffffffu;
        return true;
    };
    for (;;) {
        for (int spin = 0; spin < 256; spin++)
            if (ready()) return 0;
        const auto now = std::chrono::steady_clock::now();
        if (now < next_query) continue;
        next_query = now + std::chrono::milliseconds(250);
        const cudaError_t e = cudaStreamQuery(ctx->stream);
        if (e != cudaSuccess && e != cudaErrorNotReady) return cuda_fail(e, "round link: kernel failed", __FILE__, __LINE__);
        if (e == cudaSuccess) {  // the stream drained: either the result is there by now or the kernel gave up
            if (ready()) return 0;
            set_error("round link: the stream finished without publishing the round (kernel time-out or abort)");
            return SWIRL_ERR_INVALID;
        }
        if (now - t0 > std::chrono::seconds(20)) {
            link_abort(rs);
            cudaStreamSynchronize(ctx->stream);
            set_error("round link: no answer from the device within 20 s");
            return SWIRL_ERR_INVALID;
        }
    }
}

void round_scratch_free(swirl_ctx* ctx) {
    RoundScratch* rs = (RoundScratch*)ctx->round_scratch;
    if (!rs) return;
    cudaFree(rs->d_partials);
    cudaFree(rs->d_ticket);
    cudaFreeHost(rs->h_result);
    cudaFreeHost(rs->h_link);
    cudaFree(rs->d_gate);
    delete rs;
    ctx->round_scratch = nullptr;
}

int transcript_grind(swirl_ctx* ctx, swirl_transcript* t, int bits, uint32_t* w_mont) {
    *w_mont = 0;
    if (bits == 0) return 0;  // traits.rs:78-80: no transcript interaction
    SWIRL_REQUIRE(bits > 0 && bits < 31, "pow bits");
    uint32_t w = 0xffffffffu;
    if (bits <= 10) {
        // expected 2^bits scalar permutations (~1 us each) beat a launch + sync
        for (uint32_t c = 0; c < bb::P; c++) {
            swirl_transcript probe = *t;
            if (Transcript(&probe).check_witness(bits, bb::to_mont(c))) {
                w = c;
                break;
            }
        }
    } else {
        uint32_t st[18];
        for (int i = 0; i < 16; i++) st[i] = t->state[i];
        st[16] = t->absorb_idx;
        st[17] = t->sample_idx;
        SWIRL_TRY(swirl_sponge_grind(ctx, st, bits, 0, bb::P, &w));
    }
    if (w == 0xffffffffu) {
        set_error("failed to find proof-of-work witness");
        return SWIRL_ERR_POW;
    }
    *w_mont = bb::to_mont(w);
    if (!Transcript(t).check_witness(bits, *w_mont)) {
        set_error("internal error: grind witness rejected");
        return SWIRL_ERR_POW;
    }
    return 0;
}

}  // namespace swirl

extern "C" int swirl_transcript_observe(swirl_transcript* ts, const uint32_t* words, size_t n) {
    SWIRL_REQUIRE(ts && (words || n == 0), "null argument");
    SWIRL_REQUIRE(ts->absorb_idx < 8 && ts->sample_idx <= 8, "sponge indices");
    Transcript tr(ts);
    for (size_t i = 0; i < n; i++) {
        SWIRL_REQUIRE(words[i] < bb::P, "non-canonical Montgomery word");
        tr.observe(words[i]);
    }
    return 0;
}
extern "C" int swirl_transcript_sample(swirl_transcript* ts, uint32_t* out, size_t n) {
    SWIRL_REQUIRE(ts && (out || n == 0), "null argument");
    SWIRL_REQUIRE(ts->absorb_idx < 8 && ts->sample_idx <= 8, "sponge indices");
    Transcript tr(ts);
    for (size_t i = 0; i < n; i++) out[i] = tr.sample();
    return 0;
}
extern "C" int swirl_transcript_sample_bits(swirl_transcript* ts, int bits, uint32_t* out) {
    SWIRL_REQUIRE(ts && out, "null argument");
    SWIRL_REQUIRE(bits >= 0 && bits < 31, "bits");
    SWIRL_REQUIRE(ts->absorb_idx < 8 && ts->sample_idx <= 8, "sponge indices");
    *out = Transcript(ts).sample_bits(bits);
    return 0;
}
extern "C" int swirl_transcript_check_witness(swirl_transcript* ts, int bits, uint32_t witness, int* ok) {
    SWIRL_REQUIRE(ts && ok, "null argument");
    SWIRL_REQUIRE(bits >= 0 && bits < 31 && witness < bb::P, "bits / witness");
    SWIRL_REQUIRE(ts->absorb_idx < 8 && ts->sample_idx <= 8, "sponge indices");
    *ok = Transcript(ts).check_witness(bits, bb::to_mont(witness)) ? 1 : 0;
    return 0;
}
extern "C" int swirl_transcript_grind(swirl_ctx* ctx, swirl_transcript* ts, int bits, uint32_t* witness) {
    SWIRL_REQUIRE(ctx && ts && witness, "null argument");
    SWIRL_REQUIRE(ts->absorb_idx < 8 && ts->sample_idx <= 8, "sponge indices");
    uint32_t wm = 0;
    SWIRL_TRY(transcript_grind(ctx, ts, bits, &wm));
    *witness = bb::from_mont(wm);
    return 0;
}
