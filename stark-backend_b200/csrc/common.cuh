// Shared host-side plumbing for the swirl_b200 library: the context object behind the C ABI,
// error capture, stream-ordered scratch allocation.
//
// Reference counterparts (relative to /root/reference): cuda-common/src/stream.rs:132-151
// (one explicit non-blocking stream per device ctx), cuda-common/src/d_buffer.rs (DeviceBuffer),
// cuda-common/include/launcher.cuh:43-55 (CHECK_KERNEL).  We use the driver's stream-ordered pool
// (cudaMallocAsync with an unbounded release threshold) instead of the reference's VPMM pool.
#pragma once
#include <algorithm>
#include <chrono>
#include <cstring>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/swirl_b200.h"

struct swirl_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    cudaStream_t copy_stream = nullptr;  // H2D transport overlapping the compute stream (swirl_commit_host)
    int sm_count = 148;
    // twiddle tables: W = two_adic_generator(27) in Montgomery form
    //   tw_lo[i] = W^i            (i < 2^14)
    //   tw_hi[i] = W^(i * 2^14)   (i < 2^13)   == powers of the 2^13-th root of unity
    uint32_t* tw_lo = nullptr;
    uint32_t* tw_hi = nullptr;
    uint32_t* tw_lo_scaled[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // tw_lo * 2^-l, built on demand
    void* round_scratch = nullptr;  // swirl::RoundScratch (ext.cuh), created on first use
    uint64_t launches = 0;  // kernels launched through this ctx (bench.py reports it)
    int ntt_max_log_radix = 11;               // largest single-pass radix (log2)
    // GpuProverConfig::cache_rs_code_matrix (reference cuda-backend/src/device.rs:102-121): keep the RS codeword of a
    // commitment for the WHIR openings (true: 2x the trace in HBM) or stream it through a column-group scratch at commit
    // time and recompute the opened rows' columns in the openings (false: the large-trace mode, BASELINE configs[3])
    bool cache_codeword = true;
    bool round_link_ok = true;  // false once the probe found launches serialised (profiler, sanitizer): sponge.cu link_probe
    bool round_link = true;  // sumcheck rounds exchange results/challenges with the host through a mapped mailbox (ext.cuh: RoundLink)
    int jit_mode = 1;  // run-time compiled constraint kernels: 0 = never, 1 = traces of 2^17 rows and more, 2 = always (tests)
    bool jit_mle = true;  // the MLE rounds run compiled kernels too (batch.cu: generate_mle_source); SWIRL_JIT_MLE=0: interpreter
    uint64_t jit_stats[4] = {0, 0, 0, 0};  // round-0 kernels built / launched, MLE-round kernels built / launched (swirl_ctx_jit_stats)
    size_t ntt_scratch_bytes = size_t(4) << 30;  // inter-pass scratch per column group (measured: one big launch beats L2-sized groups)
    // optional per-kernel-family CUDA-event timing (bench.py's roofline numbers)
    bool timing = false;
    struct TimedSpan {
        int slot;
        cudaEvent_t a, b;
        uint64_t bytes;  // algorithmic bytes of the launch (0 = not accounted)
    };
    std::vector<TimedSpan> spans;
    // Arena of large device blocks (>= ARENA_MIN bytes), see dev_alloc below.
    std::multimap<size_t, void*> arena_free;        // size -> idle block
    std::unordered_map<void*, size_t> arena_live;   // block handed out -> size
    size_t arena_bytes = 0;                         // idle + live
    size_t arena_live_bytes = 0, arena_live_peak = 0;  // handed out now / high-water mark (swirl_ctx_mem_stats)
    // host-side synchronisation statistics (swirl_ctx_sync_stats): how much of a proof is spent waiting on the stream
    uint64_t sync_count = 0;
    double sync_ms = 0;
    uint64_t link_count = 0;  // round results received through the mapped mailbox instead of a stream synchronisation (ext.cuh: RoundLink)
    // pinned staging area for device-to-host copies into caller (pageable) memory, see d2h_staged
    void* program_cache = nullptr;  // compiled constraint programs (batch.cu: ProgramCache)
    void* h_stage = nullptr;
    size_t h_stage_bytes = 0;
};

// kernel families for swirl_ctx_timing_read
enum {
    SWIRL_T_LEAF = 0,      // fused row sponge + strided tree levels (leaf_tree_kernel)
    SWIRL_T_TREE = 1,      // upper adjacent compression layers
    SWIRL_T_CHUNK = 2,     // 2^l_skip chunk iDFT + zeta
    SWIRL_T_NTT_PASS = 3,  // strided NTT passes
    SWIRL_T_NTT_FINAL = 4, // final NTT pass (natural-order store)
    SWIRL_T_STACK = 5,     // stacking copies
    SWIRL_T_GKR = 6,       // GKR round kernels (gkr_round_kernel)
    SWIRL_T_BC_ROUND0 = 7, // batch constraints: round-0 coset evaluation
    SWIRL_T_BC_MLE = 8,    // batch constraints: MLE round kernels
    SWIRL_T_SLOTS = 9
};

struct SwirlTimed {  // RAII: records an event pair around a launch when ctx->timing is on
    swirl_ctx* ctx;
    cudaEvent_t b = nullptr;
    SwirlTimed(swirl_ctx* c, int slot, uint64_t bytes = 0) : ctx(c) {
        if (!c->timing) return;
        cudaEvent_t a;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a, c->stream);
        c->spans.push_back({slot, a, b, bytes});
    }
    ~SwirlTimed() {
        if (b) cudaEventRecord(b, ctx->stream);
    }
};

namespace swirl {

constexpr int TW_LO_BITS = 14;
constexpr int TW_HI_BITS = 13;

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

// SWIRL_STALL_DEBUG=<ms>: host calls (stream-ordered allocations, stream synchronisations) that take longer
// than <ms> of wall time are reported on stderr with their call site (box / allocator stall hunting).
void program_cache_clear(swirl_ctx* ctx);  // batch.cu
double stall_debug_ms();
void stall_report(const char* what, const char* file, int line, double ms, size_t bytes);
inline cudaError_t stream_sync(swirl_ctx* ctx, const char* file, int line) {
    const double lim = stall_debug_ms();
    const auto t0 = std::chrono::steady_clock::now();
    const cudaError_t e = cudaStreamSynchronize(ctx->stream);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    ctx->sync_count++;
    ctx->sync_ms += ms;
    if (lim > 0 && ms > lim) stall_report("cudaStreamSynchronize", file, line, ms, 0);
    return e;
}

// SWIRL_TRACE=1: wall-clock per phase on stderr; every mark synchronises the stream first.
inline void trace_mark(swirl_ctx* ctx, const char* phase, const char* name, std::chrono::steady_clock::time_point* prev) {
    static const bool on = getenv("SWIRL_TRACE") != nullptr;
    if (!on) return;
    cudaStreamSynchronize(ctx->stream);
    const auto now = std::chrono::steady_clock::now();
    if (name) fprintf(stderr, "[swirl trace] %s %-18s %8.3f ms\n", phase, name, std::chrono::duration<double, std::milli>(now - *prev).count());
    *prev = now;
}

// Device -> caller memory.  Proof buffers belong to the caller and are usually pageable; a direct cudaMemcpyAsync into
// them is staged by the driver in small chunks and blocks (~1 ms for the 3 MB of opened rows).  Copies of 64 KiB and more
// go through one pinned buffer of the context instead: one DMA at full PCIe rate, one synchronisation, one memcpy.
// Returns with the data in `dst` (the stream is synchronised).
inline cudaError_t d2h_staged(swirl_ctx* ctx, void* dst, const void* d_src, size_t bytes) {
    if (bytes < (size_t(64) << 10)) {
        cudaError_t e = cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream);
        return e != cudaSuccess ? e : cudaStreamSynchronize(ctx->stream);
    }
    if (ctx->h_stage_bytes < bytes) {
        if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
        ctx->h_stage = nullptr;
        ctx->h_stage_bytes = 0;
        const size_t want = std::max(bytes, size_t(8) << 20);
        cudaError_t e = cudaHostAlloc(&ctx->h_stage, want, cudaHostAllocDefault);
        if (e != cudaSuccess) return e;
        ctx->h_stage_bytes = want;
    }
    cudaError_t e = cudaMemcpyAsync(ctx->h_stage, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) memcpy(dst, ctx->h_stage, bytes);
    return e;
}

// Scratch allocation.  Small blocks come from the driver's stream-ordered pool.  Blocks of ARENA_MIN bytes
// and more are kept by the context and handed out again by size (exact size first, then the smallest idle
// block within +25 %): a proof allocates the same multi-GiB tables in the same order every time, and carving
// them out of the driver pool again and again was measured to stall for 0.1-0.9 s at random when the pool
// has to re-map physical memory behind a new virtual range (tools: SWIRL_STALL_DEBUG).  Reuse is ordered by
// ctx->stream like cudaFreeAsync was.  swirl_ctx_trim / an out-of-memory allocation return idle blocks to
// the driver.
constexpr size_t ARENA_MIN = size_t(1) << 20;
cudaError_t arena_alloc(swirl_ctx* ctx, void** p, size_t bytes);
void arena_free_block(swirl_ctx* ctx, void* p);
void arena_trim(swirl_ctx* ctx);

template <class T>
inline cudaError_t dev_alloc(swirl_ctx* ctx, T** p, size_t count) {
    return arena_alloc(ctx, (void**)p, count * sizeof(T));
}
template <class T>
inline void dev_free(swirl_ctx* ctx, T* p) {
    if (p) arena_free_block(ctx, (void*)p);
}

// Scratch blocks of one library call: handed back to the arena when the call returns, on every path.
struct ArenaGuard {
    swirl_ctx* ctx;
    std::vector<void*> blocks;
    explicit ArenaGuard(swirl_ctx* c) : ctx(c) {}
    ArenaGuard(const ArenaGuard&) = delete;
    ArenaGuard& operator=(const ArenaGuard&) = delete;
    template <class T>
    void add(T* p) {
        if (p) blocks.push_back((void*)p);
    }
    ~ArenaGuard() {
        for (void* p : blocks) arena_free_block(ctx, p);
    }
};

inline int ilog2(size_t n) {
    int l = 0;
    while ((size_t(1) << l) < n) l++;
    return l;
}
inline bool is_pow2(size_t n) { return n && !(n & (n - 1)); }

}  // namespace swirl

#define SWIRL_CUDA(expr)                                                         \
    do {                                                                         \
        cudaError_t _e = (expr);                                                 \
        if (_e != cudaSuccess) return swirl::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

#define SWIRL_LAUNCH_CHECK(ctx)                                                  \
    do {                                                                         \
        (ctx)->launches++;                                                       \
        cudaError_t _e = cudaGetLastError();                                     \
        if (_e != cudaSuccess) return swirl::cuda_fail(_e, "kernel launch", __FILE__, __LINE__); \
    } while (0)

#define SWIRL_REQUIRE(cond, msg)                       \
    do {                                               \
        if (!(cond)) {                                 \
            swirl::set_error(std::string("invalid argument: ") + (msg)); \
            return SWIRL_ERR_INVALID;                  \
        }                                              \
    } while (0)

#define SWIRL_TRY(expr)            \
    do {                           \
        int _rc = (expr);          \
        if (_rc != 0) return _rc;  \
    } while (0)
