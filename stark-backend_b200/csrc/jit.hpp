// Run-time compilation of per-AIR kernels with NVRTC.
//
// The reference turns every AIR's SymbolicConstraintsDag into three-address rules that a generic kernel interprets
// (crates/cuda-backend/src/logup_zerocheck/rules/mod.rs:27-130, cuda/include/{codec.cuh,dag_entry.cuh}); here the same
// program (batch.cu: compile_program) is additionally emitted as straight-line CUDA C++ -- one statement per
// instruction, value slots as registers -- compiled for sm_100a at first use and cached per program.  libnvrtc is
// dlopen'ed: without it (or with SWIRL_JIT=0) the interpreter kernels run, which is a GPU path as well.
#pragma once
#include <string>

#include "common.cuh"

namespace swirl {

struct JitKernel {
    void* kernel = nullptr;   // cudaKernel_t
    void* library = nullptr;  // cudaLibrary_t
};

// true when libnvrtc could be loaded and SWIRL_JIT is not 0
bool jit_available();
// Compiles `source` and resolves the extern "C" kernel `name`.  Returns 0, or an error code with set_error text
// (compile log included).  Blocking (seconds for a large program): callers cache the result.
int jit_compile(swirl_ctx* ctx, const std::string& source, const char* name, JitKernel* out);
void jit_release(JitKernel* k);

// the device-side prelude every generated translation unit starts with (csrc/jit_prelude.cuh, embedded at build time)
const char* jit_prelude();

}  // namespace swirl
