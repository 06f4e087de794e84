// Poseidon2 permutation throughput of the candidate implementations (perms/s on the whole chip),
// checked for equality against p2::permute on the same states.
//   nvcc -O3 -std=c++17 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -I../stark-backend_b200/csrc -o p2_bench.bin p2_bench.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "poseidon2.cuh"
#include "poseidon2_v2.cuh"

template <int V>
__device__ __forceinline__ void perm(uint32_t s[16]) {
    if (V == 0) p2::permute_v1(s);
    if (V == 1) p2v2::permute(s);
    if (V >= 100) p2v2::permute_pol<V - 100>(s);  // pin-policy sweep
}

// every thread iterates the permutation `iters` times on its own state (issue-bound measurement)
template <int V>
__global__ void __launch_bounds__(256) iterate(uint32_t* states, int iters, long long* cyc) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const long long t0 = clock64();
    uint4* p = reinterpret_cast<uint4*>(states + i * 16);
    uint4 v0 = p[0], v1 = p[1], v2 = p[2], v3 = p[3];
    uint32_t s[16] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w, v3.x, v3.y, v3.z, v3.w};
    for (int it = 0; it < iters; it++) perm<V>(s);
    p[0] = make_uint4(s[0], s[1], s[2], s[3]);
    p[1] = make_uint4(s[4], s[5], s[6], s[7]);
    p[2] = make_uint4(s[8], s[9], s[10], s[11]);
    p[3] = make_uint4(s[12], s[13], s[14], s[15]);
    if (cyc && threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}

template <int V>
double run(const char* name, const std::vector<uint32_t>& init, std::vector<uint32_t>& out, int threads_per_sm, int sms, int iters) {
    const size_t n = (size_t)threads_per_sm * sms;
    uint32_t* d;
    cudaMalloc(&d, n * 64);
    cudaMemcpy(d, init.data(), n * 64, cudaMemcpyHostToDevice);
    long long* cyc;
    cudaMalloc(&cyc, (n / 256) * 8);
    iterate<V><<<n / 256, 256>>>(d, 4 * iters, nullptr);  // warm the clocks
    cudaMemcpy(d, init.data(), n * 64, cudaMemcpyHostToDevice);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    iterate<V><<<n / 256, 256>>>(d, iters, cyc);
    cudaEventRecord(b);
    cudaError_t e = cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    out.resize(n * 16);
    cudaMemcpy(out.data(), d, n * 64, cudaMemcpyDeviceToHost);
    cudaFree(d);
    double gps = (double)n * iters / (ms * 1e-3) / 1e9;
    std::vector<long long> hc(n / 256);
    cudaMemcpy(hc.data(), cyc, (n / 256) * 8, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (auto c : hc) mx = c > mx ? c : mx;
    cudaFree(cyc);
    printf("{\"impl\": \"%s\", \"threads_per_sm\": %d, \"gperm_per_s\": %.3f, \"ms\": %.3f, \"clk_per_perm_per_sm\": %.2f, \"eff_mhz\": %.0f, \"err\": \"%s\"}\n", name, threads_per_sm, gps, ms,
           (double)mx / ((double)threads_per_sm * iters), (double)mx / (ms * 1e3), cudaGetErrorString(e));
    return gps;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    if (sms <= 0) return 1;
    const int iters = 256;
    for (int tps : {512, 1024, 2048}) {
        const size_t n = (size_t)tps * sms;
        std::vector<uint32_t> init(n * 16), o0, o1;
        uint64_t x = 88172645463325252ull;
        for (auto& v : init) {
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            v = (uint32_t)(x % bb::P);
        }
        run<0>("p2 (v1, unrolled, canonical)", init, o0, tps, sms, iters);
        run<1>("p2v2", init, o1, tps, sms, iters);
        size_t bad = 0;
        for (size_t i = 0; i < o0.size(); i++) bad += o0[i] != o1[i];
        printf("{\"check\": \"p2v2 == p2\", \"mismatches\": %zu}\n", bad);
        if (tps == 2048) {
            std::vector<uint32_t> o2;
#define POLRUN(P)                                                                     \
    {                                                                                 \
        run<100 + P>("p2v2 pin policy " #P, init, o2, tps, sms, iters);               \
        size_t b2 = 0;                                                                \
        for (size_t i = 0; i < o0.size(); i++) b2 += o0[i] != o2[i];                  \
        printf("{\"check\": \"policy " #P " == p2\", \"mismatches\": %zu}\n", b2); \
    }
            POLRUN(0) POLRUN(1) POLRUN(8) POLRUN(16) POLRUN(7) POLRUN(24) POLRUN(29) POLRUN(31) POLRUN(23) POLRUN(15) POLRUN(63) POLRUN(61) POLRUN(55) POLRUN(47) POLRUN(39) POLRUN(56)
        }
    }
    return 0;
}
