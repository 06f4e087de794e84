// NVRTC plumbing for jit.hpp: dlopen libnvrtc, compile to a cubin for sm_100a, load it through the runtime's library API.
#include "jit.hpp"

#include <dlfcn.h>
#include <nvrtc.h>

#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace swirl {

namespace {

struct Nvrtc {
    void* handle = nullptr;
    decltype(&nvrtcCreateProgram) create = nullptr;
    decltype(&nvrtcCompileProgram) compile = nullptr;
    decltype(&nvrtcGetCUBINSize) cubin_size = nullptr;
    decltype(&nvrtcGetCUBIN) cubin = nullptr;
    decltype(&nvrtcGetProgramLogSize) log_size = nullptr;
    decltype(&nvrtcGetProgramLog) log = nullptr;
    decltype(&nvrtcDestroyProgram) destroy = nullptr;
    decltype(&nvrtcGetErrorString) err_string = nullptr;
    bool ok = false;
};

Nvrtc& nvrtc() {
    static Nvrtc n;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* env = getenv("SWIRL_JIT");
        if (env && env[0] == '0') return;
        for (const char* name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"}) {
            n.handle = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (n.handle) break;
        }
        if (!n.handle) return;
#define SWIRL_SYM(field, sym) n.field = reinterpret_cast<decltype(n.field)>(dlsym(n.handle, #sym))
        SWIRL_SYM(create, nvrtcCreateProgram);
        SWIRL_SYM(compile, nvrtcCompileProgram);
        SWIRL_SYM(cubin_size, nvrtcGetCUBINSize);
        SWIRL_SYM(cubin, nvrtcGetCUBIN);
        SWIRL_SYM(log_size, nvrtcGetProgramLogSize);
        SWIRL_SYM(log, nvrtcGetProgramLog);
        SWIRL_SYM(destroy, nvrtcDestroyProgram);
        SWIRL_SYM(err_string, nvrtcGetErrorString);
#undef SWIRL_SYM
        n.ok = n.create && n.compile && n.cubin_size && n.cubin && n.log_size && n.log && n.destroy && n.err_string;
    });
    return n;
}

}  // namespace

bool jit_available() { return nvrtc().ok; }

const char* jit_prelude() {
    static const char* src =
#include "jit_prelude_str.inc"
        ;
    return src;
}

static int jit_load(swirl_ctx* ctx, const std::vector<char>& cubin, const char* name, JitKernel* out) {
    SWIRL_CUDA(cudaSetDevice(ctx->device));
    cudaLibrary_t lib = nullptr;
    SWIRL_CUDA(cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
    cudaKernel_t k = nullptr;
    cudaError_t e = cudaLibraryGetKernel(&k, lib, name);
    if (e != cudaSuccess) {
        cudaLibraryUnload(lib);
        return cuda_fail(e, "cudaLibraryGetKernel", __FILE__, __LINE__);
    }
    out->kernel = (void*)k;
    out->library = (void*)lib;
    return 0;
}

int jit_compile(swirl_ctx* ctx, const std::string& source, const char* name, JitKernel* out) {
    Nvrtc& n = nvrtc();
    if (!n.ok) {
        set_error("NVRTC is not available");
        return SWIRL_ERR_INVALID;
    }
    // cubins are shared by all contexts of the process (a prover per thread compiles each program once, not once per
    // context); loading a cubin into a context is cheap
    static std::mutex cache_mutex;
    static std::map<std::string, std::shared_ptr<const std::vector<char>>> cubin_cache;
    std::shared_ptr<const std::vector<char>> cached;
    {
        std::lock_guard<std::mutex> lock(cache_mutex);
        auto it = cubin_cache.find(source);
        if (it != cubin_cache.end()) cached = it->second;
    }
    if (cached) return jit_load(ctx, *cached, name, out);
    nvrtcProgram prog = nullptr;
    nvrtcResult r = n.create(&prog, source.c_str(), "swirl_jit.cu", 0, nullptr, nullptr);
    if (r != NVRTC_SUCCESS) {
        set_error(std::string("nvrtcCreateProgram: ") + n.err_string(r));
        return SWIRL_ERR_INVALID;
    }
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--extra-device-vectorization"};
    r = n.compile(prog, 4, opts);
    if (r != NVRTC_SUCCESS) {
        size_t ls = 0;
        n.log_size(prog, &ls);
        std::string log(ls, '\0');
        if (ls) n.log(prog, &log[0]);
        n.destroy(&prog);
        set_error(std::string("nvrtcCompileProgram: ") + n.err_string(r) + "\n" + log.substr(0, 4000));
        return SWIRL_ERR_INVALID;
    }
    size_t cs = 0;
    n.cubin_size(prog, &cs);
    auto cubin = std::make_shared<std::vector<char>>(cs);
    n.cubin(prog, cubin->data());
    n.destroy(&prog);
    {
        std::lock_guard<std::mutex> lock(cache_mutex);
        if (cubin_cache.size() >= 256) cubin_cache.clear();  // bound the cache; entries in use are kept alive by their users
        cubin_cache.emplace(source, cubin);
    }
    return jit_load(ctx, *cubin, name, out);
}

void jit_release(JitKernel* k) {
    if (k->library) cudaLibraryUnload((cudaLibrary_t)k->library);
    k->kernel = k->library = nullptr;
}

}  // namespace swirl
