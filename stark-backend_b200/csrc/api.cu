// C-ABI glue: context lifetime, error text, and the thin extern "C" wrappers over the kernel-level
// primitives (include/swirl_b200.h).  Phase-level entry points live next to their orchestration
// (commit.cu, sponge.cu).
#include <cstdlib>
#include <string>

#include "kernels.cuh"

namespace swirl {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    g_last_error = std::string("CUDA error ") + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ") in " + what +
                   " at " + file + ":" + std::to_string(line);
    return (int)e;
}

double stall_debug_ms() {
    static const double lim = [] {
        const char* e = getenv("SWIRL_STALL_DEBUG");
        return e ? atof(e) : 0.0;
    }();
    return lim;
}
static cudaError_t timed_malloc_async(swirl_ctx* ctx, void** p, size_t bytes) {
    const double lim = stall_debug_ms();
    if (lim <= 0) return cudaMallocAsync(p, bytes ? bytes : 4, ctx->stream);
    const auto t0 = std::chrono::steady_clock::now();
    const cudaError_t e = cudaMallocAsync(p, bytes ? bytes : 4, ctx->stream);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (ms > lim) stall_report("cudaMallocAsync", __FILE__, __LINE__, ms, bytes);
    return e;
}

cudaError_t arena_alloc(swirl_ctx* ctx, void** p, size_t bytes) {
    if (bytes < ARENA_MIN) return timed_malloc_async(ctx, p, bytes);
    bytes = (bytes + 511) & ~size_t(511);
    auto it = ctx->arena_free.lower_bound(bytes);
    if (it != ctx->arena_free.end() && it->first <= bytes + bytes / 4) {
        *p = it->second;
        ctx->arena_live[*p] = it->first;
        ctx->arena_live_bytes += it->first;
        ctx->arena_live_peak = std::max(ctx->arena_live_peak, ctx->arena_live_bytes);
        ctx->arena_free.erase(it);
        return cudaSuccess;
    }
    cudaError_t e = timed_malloc_async(ctx, p, bytes);
    if (e == cudaErrorMemoryAllocation) {  // give the idle blocks back and try once more
        cudaGetLastError();
        arena_trim(ctx);
        cudaStreamSynchronize(ctx->stream);
        e = timed_malloc_async(ctx, p, bytes);
    }
    if (e == cudaSuccess) {
        ctx->arena_live[*p] = bytes;
        ctx->arena_bytes += bytes;
        ctx->arena_live_bytes += bytes;
        ctx->arena_live_peak = std::max(ctx->arena_live_peak, ctx->arena_live_bytes);
    }
    return e;
}

void arena_free_block(swirl_ctx* ctx, void* p) {
    auto it = ctx->arena_live.find(p);
    if (it == ctx->arena_live.end()) {
        cudaFreeAsync(p, ctx->stream);
        return;
    }
    ctx->arena_free.emplace(it->second, p);
    ctx->arena_live_bytes -= it->second;
    ctx->arena_live.erase(it);
}

void arena_trim(swirl_ctx* ctx) {
    for (auto& kv : ctx->arena_free) {
        cudaFreeAsync(kv.second, ctx->stream);
        ctx->arena_bytes -= kv.first;
    }
    ctx->arena_free.clear();
}

void stall_report(const char* what, const char* file, int line, double ms, size_t bytes) {
    fprintf(stderr, "[swirl stall] %s %.1f ms (%zu bytes) at %s:%d\n", what, ms, bytes, file, line);
}

}  // namespace swirl

using namespace swirl;

extern "C" {
static void timing_clear(swirl_ctx* ctx);
}

static int ctx_init(int device, cudaStream_t stream, bool owns, swirl_ctx** out) {
    SWIRL_REQUIRE(out, "null out");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("no CUDA device available: libswirl_b200 has no CPU fallback");
        return SWIRL_ERR_NO_DEVICE;
    }
    SWIRL_REQUIRE(device >= 0 && device < count, "device index");
    SWIRL_CUDA(cudaSetDevice(device));
    swirl_ctx* ctx = new swirl_ctx();
    ctx->device = device;
    ctx->stream = stream;
    ctx->owns_stream = owns;
    // The round link needs asynchronous launches (ext.cuh).  Tools that serialise them announce themselves in the
    // environment of the process they start: Nsight Compute, Nsight Systems' CUDA injection, compute-sanitizer; a profiler
    // that only instruments SOME kernels (ncu -k / --launch-skip) would get past the start-up probe of sponge.cu.
    for (const char* marker : {"NV_COMPUTE_PROFILER_PERFWORKS_DIR", "NV_NSIGHT_INJECTION_TRANSPORT_TYPE", "NV_SANITIZER_INJECTION_TRANSPORT_TYPE",
                               "CUDA_INJECTION64_PATH"})
        if (getenv(marker)) ctx->round_link = ctx->round_link_ok = false;
    if (const char* env = getenv("CUDA_LAUNCH_BLOCKING"))
        if (atoi(env) != 0) ctx->round_link = ctx->round_link_ok = false;
    if (const char* env = getenv("SWIRL_JIT_MLE")) ctx->jit_mle = atoi(env) != 0;
    if (const char* env = getenv("SWIRL_JIT_MODE")) ctx->jit_mode = std::max(0, std::min(2, atoi(env)));  // initial swirl_ctx_set_jit mode (A/B runs)
    if (const char* env = getenv("SWIRL_ROUND_LINK")) ctx->round_link = atoi(env) != 0 && ctx->round_link_ok;  // A/B knob, see swirl_ctx_set_round_link
    if (owns) {
        e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete ctx;
            return cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__);
        }
    }
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    // keep freed blocks in the stream-ordered pool instead of returning them to the driver
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thresh = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh);
    }
    int rc = ntt_init_twiddles(ctx);
    if (rc == 0) {
        e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = cuda_fail(e, "twiddle init", __FILE__, __LINE__);
    }
    if (rc != 0) {
        swirl_ctx_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return 0;
}

extern "C" {

int swirl_ctx_create(int device, swirl_ctx** out) { return ctx_init(device, nullptr, true, out); }

int swirl_ctx_create_on_stream(int device, void* cuda_stream, swirl_ctx** out) {
    return ctx_init(device, (cudaStream_t)cuda_stream, false, out);
}

int swirl_ctx_destroy(swirl_ctx* ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    if (ctx->stream || !ctx->owns_stream) cudaStreamSynchronize(ctx->stream);
    timing_clear(ctx);
    swirl::program_cache_clear(ctx);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    arena_trim(ctx);
    for (auto& kv : ctx->arena_live) cudaFreeAsync(kv.first, ctx->stream);  // blocks the caller never released
    ctx->arena_live.clear();
    cudaStreamSynchronize(ctx->stream);
    if (ctx->tw_lo) cudaFree(ctx->tw_lo);
    if (ctx->tw_hi) cudaFree(ctx->tw_hi);
    for (uint32_t* t : ctx->tw_lo_scaled)
        if (t) cudaFree(t);
    round_scratch_free(ctx);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->owns_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return 0;
}

int swirl_ctx_synchronize(swirl_ctx* ctx) {
    SWIRL_REQUIRE(ctx, "null ctx");
    SWIRL_CUDA(swirl::stream_sync(ctx, __FILE__, __LINE__));
    return 0;
}

void* swirl_ctx_stream(swirl_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
uint64_t swirl_ctx_launch_count(swirl_ctx* ctx) { return ctx ? ctx->launches : 0; }

int swirl_ctx_set_ntt_plan(swirl_ctx* ctx, int max_log_radix, size_t scratch_bytes) {
    SWIRL_REQUIRE(ctx, "null ctx");
    SWIRL_REQUIRE(max_log_radix >= 1 && max_log_radix <= 13, "max_log_radix must be in [1, 13]");
    ctx->ntt_max_log_radix = max_log_radix;
    if (scratch_bytes) ctx->ntt_scratch_bytes = scratch_bytes;
    return 0;
}

int swirl_ctx_set_cache_rs_code_matrix(swirl_ctx* ctx, int on) {
    SWIRL_REQUIRE(ctx, "null ctx");
    ctx->cache_codeword = on != 0;
    return 0;
}

int swirl_ctx_set_jit(swirl_ctx* ctx, int mode) {
    SWIRL_REQUIRE(ctx && mode >= 0 && (mode & 3) <= 2 && mode < 8, "mode must be 0, 1 or 2, plus 4 for compiled MLE rounds");
    ctx->jit_mode = mode & 3;  // 0 switches every compiled kernel off, the MLE rounds' included
    if (mode & 4) ctx->jit_mle = true;  // otherwise the context keeps its setting (default, or SWIRL_JIT_MLE in the environment)
    return 0;
}

int swirl_ctx_jit_stats(swirl_ctx* ctx, uint64_t out[4]) {
    SWIRL_REQUIRE(ctx && out, "null argument");
    for (int i = 0; i < 4; i++) out[i] = ctx->jit_stats[i];
    return 0;
}

int swirl_ctx_mem_stats(swirl_ctx* ctx, int reset_peak, uint64_t out[4]) {
    SWIRL_REQUIRE(ctx && out, "null argument");
    size_t free_b = 0, total_b = 0;
    SWIRL_CUDA(cudaSetDevice(ctx->device));
    SWIRL_CUDA(cudaMemGetInfo(&free_b, &total_b));
    out[0] = ctx->arena_live_bytes;
    out[1] = ctx->arena_live_peak;
    out[2] = ctx->arena_bytes;
    out[3] = free_b;
    if (reset_peak) ctx->arena_live_peak = ctx->arena_live_bytes;
    return 0;
}

const char* swirl_last_error(void) { return g_last_error.c_str(); }

static void timing_clear(swirl_ctx* ctx) {
    for (auto& s : ctx->spans) {
        cudaEventDestroy(s.a);
        cudaEventDestroy(s.b);
    }
    ctx->spans.clear();
}

int swirl_ctx_timing_enable(swirl_ctx* ctx, int on) {
    SWIRL_REQUIRE(ctx, "null ctx");
    SWIRL_CUDA(swirl::stream_sync(ctx, __FILE__, __LINE__));
    timing_clear(ctx);
    ctx->timing = on != 0;
    return 0;
}

int swirl_ctx_timing_read(swirl_ctx* ctx, int slot, double* total_ms, uint64_t* count) {
    SWIRL_REQUIRE(ctx && total_ms && count, "null argument");
    SWIRL_REQUIRE(slot >= 0 && slot < SWIRL_T_SLOTS, "slot");
    SWIRL_CUDA(swirl::stream_sync(ctx, __FILE__, __LINE__));
    double tot = 0;
    uint64_t n = 0;
    for (auto& s : ctx->spans)
        if (s.slot == slot) {
            float ms = 0;
            SWIRL_CUDA(cudaEventElapsedTime(&ms, s.a, s.b));
            tot += ms;
            n++;
        }
    *total_ms = tot;
    *count = n;
    return 0;
}

int swirl_ctx_sync_stats(swirl_ctx* ctx, uint64_t* count, double* wait_ms) {
    SWIRL_REQUIRE(ctx && count && wait_ms, "null argument");
    *count = ctx->sync_count;
    *wait_ms = ctx->sync_ms;
    return 0;
}

int swirl_ctx_set_round_link(swirl_ctx* ctx, int on) {
    SWIRL_REQUIRE(ctx, "null ctx");
    ctx->round_link = on != 0 && ctx->round_link_ok;
    return 0;
}

int swirl_ctx_link_stats(swirl_ctx* ctx, uint64_t* count) {
    SWIRL_REQUIRE(ctx && count, "null argument");
    *count = ctx->link_count;
    return 0;
}

int swirl_ctx_timing_bytes(swirl_ctx* ctx, int slot, uint64_t* bytes) {
    SWIRL_REQUIRE(ctx && bytes, "null argument");
    SWIRL_REQUIRE(slot >= 0 && slot < SWIRL_T_SLOTS, "slot");
    uint64_t n = 0;
    for (auto& s : ctx->spans)
        if (s.slot == slot) n += s.bytes;
    *bytes = n;
    return 0;
}

int swirl_malloc(swirl_ctx* ctx, size_t bytes, void** d_out) {
    SWIRL_REQUIRE(ctx && d_out, "null argument");
    SWIRL_CUDA(cudaSetDevice(ctx->device));
    SWIRL_CUDA(arena_alloc(ctx, d_out, bytes ? bytes : 1));
    return 0;
}

int swirl_free(swirl_ctx* ctx, void* d_ptr) {
    SWIRL_REQUIRE(ctx, "null ctx");
    if (d_ptr) arena_free_block(ctx, d_ptr);
    return 0;
}

int swirl_ctx_trim(swirl_ctx* ctx) {
    SWIRL_REQUIRE(ctx, "null ctx");
    SWIRL_CUDA(cudaSetDevice(ctx->device));
    arena_trim(ctx);
    return 0;
}

int swirl_memcpy_h2d(swirl_ctx* ctx, void* d_dst, const void* h_src, size_t bytes) {
    SWIRL_REQUIRE(ctx && ((d_dst && h_src) || bytes == 0), "null argument");
    if (bytes) SWIRL_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

int swirl_memcpy_d2h(swirl_ctx* ctx, void* h_dst, const void* d_src, size_t bytes) {
    SWIRL_REQUIRE(ctx && ((h_dst && d_src) || bytes == 0), "null argument");
    if (bytes) SWIRL_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    SWIRL_CUDA(swirl::stream_sync(ctx, __FILE__, __LINE__));
    return 0;
}

int swirl_poseidon2_permute(swirl_ctx* ctx, uint32_t* d_states, size_t n) {
    SWIRL_REQUIRE(ctx && (d_states || n == 0), "null argument");
    SWIRL_CUDA(cudaSetDevice(ctx->device));
    return poseidon2_permute_batch(ctx, d_states, n);
}

int swirl_poseidon2_compress(swirl_ctx* ctx, const uint32_t* d_pairs, uint32_t* d_out, size_t n) {
    SWIRL_REQUIRE(ctx && ((d_pairs && d_out) || n == 0), "null argument");
    SWIRL_CUDA(cudaSetDevice(ctx->device));
    return poseidon2_compress_batch(ctx, d_pairs, d_out, n);
}

int swirl_ntt_batch(swirl_ctx* ctx, uint32_t* d_data, int log_n, size_t cols, int inverse) {
    SWIRL_REQUIRE(ctx && (d_data || cols == 0), "null argument");
    SWIRL_CUDA(cudaSetDevice(ctx->device));
    return ntt_batch(ctx, d_data, log_n, cols, inverse != 0);
}

int swirl_rs_encode(swirl_ctx* ctx, const uint32_t* d_in, size_t height, size_t width, int l_skip, int log_blowup,
                    uint32_t* d_out) {
    SWIRL_REQUIRE(ctx && ((d_in && d_out) || width == 0), "null argument");
    SWIRL_CUDA(cudaSetDevice(ctx->device));
    return rs_encode(ctx, d_in, height, height, width, l_skip, log_blowup, d_out);
}

int swirl_merkle_tree(swirl_ctx* ctx, const uint32_t* d_matrix, size_t height, size_t width, int log_rows_per_query,
                      uint32_t* d_layers) {
    SWIRL_REQUIRE(ctx && d_layers && (d_matrix || width == 0), "null argument");
    SWIRL_CUDA(cudaSetDevice(ctx->device));
    return merkle_commit(ctx, d_matrix, height, width, log_rows_per_query, d_layers);
}

int swirl_merkle_query_proofs(swirl_ctx* ctx, const uint32_t* d_layers, size_t query_stride, const uint32_t* d_indices,
                              size_t num_queries, uint32_t* d_out) {
    SWIRL_REQUIRE(ctx && d_layers && (num_queries == 0 || (d_indices && d_out)), "null argument");
    SWIRL_REQUIRE(is_pow2(query_stride), "query_stride must be a power of two");
    SWIRL_CUDA(cudaSetDevice(ctx->device));
    return merkle_query_proofs(ctx, d_layers, query_stride, d_indices, num_queries, d_out);
}

int swirl_matrix_open_rows(swirl_ctx* ctx, const uint32_t* d_matrix, size_t height, size_t width, size_t query_stride,
                           int log_rows_per_query, const uint32_t* d_indices, size_t num_queries, uint32_t* d_out) {
    SWIRL_REQUIRE(ctx && (num_queries == 0 || width == 0 || (d_matrix && d_indices && d_out)), "null argument");
    SWIRL_CUDA(cudaSetDevice(ctx->device));
    return matrix_open_rows(ctx, d_matrix, height, width, query_stride, log_rows_per_query, d_indices, num_queries,
                            d_out);
}

}  // extern "C"
