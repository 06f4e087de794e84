// Is the FP64 pipe of B200 a usable second engine for Poseidon2?  Measures, on the whole chip:
//   1. raw issue rates: DFMA chains alone, IMAD chains alone, and both kinds of warps together on every SM
//      (do the fp64 and fma-heavy pipes run concurrently, and at what rate each?);
//   2. the FP64 permutation (poseidon2_f64.cuh) against the integer one (poseidon2_v2.cuh): equality of outputs on
//      random states, and perms/s of each alone;
//   3. mixed CTAs: F of the 8 warps of every CTA run the FP64 permutation, the others the integer one, all
//      drawing batches of permutations from one shared counter (the dynamic hand-out the leaf kernel would use).
//   nvcc -O3 -std=c++17 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -I../stark-backend_b200/csrc -o p2_fp64_bench.bin p2_fp64_bench.cu
#include <cstdint>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "poseidon2_v2.cuh"
#include "poseidon2_f64.cuh"

// ---- 1. raw pipes ------------------------------------------------------------------------------------------
// mode bit 0: warps with odd index run DFMA, bit 1: warps with even index run IMAD (3 = both)
__global__ void __launch_bounds__(256) pipes(int mode, int iters, double* dout, uint32_t* iout) {
    const int warp = threadIdx.x >> 5;
    const bool fp = (warp & 1) != 0;
    if (fp && (mode & 1)) {
        double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
        const double m = 1.0000001, c = 0.5;
        for (int i = 0; i < iters; i++) {
            a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
            a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
        }
        dout[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    } else if (!fp && (mode & 2)) {
        uint32_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
        const uint32_t m = 0x9E3779B1u + blockIdx.x, c = 12345u;
        for (int i = 0; i < iters; i++) {
            a0 = a0 * m + c; a1 = a1 * m + c; a2 = a2 * m + c; a3 = a3 * m + c;
            a4 = a4 * m + c; a5 = a5 * m + c; a6 = a6 * m + c; a7 = a7 * m + c;
        }
        iout[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
    }
}

// ---- 2./3. permutations -------------------------------------------------------------------------------------
__device__ __forceinline__ void load16(const uint32_t* p, uint32_t s[16]) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 v0 = q[0], v1 = q[1], v2 = q[2], v3 = q[3];
    s[0] = v0.x; s[1] = v0.y; s[2] = v0.z; s[3] = v0.w; s[4] = v1.x; s[5] = v1.y; s[6] = v1.z; s[7] = v1.w;
    s[8] = v2.x; s[9] = v2.y; s[10] = v2.z; s[11] = v2.w; s[12] = v3.x; s[13] = v3.y; s[14] = v3.z; s[15] = v3.w;
}
__device__ __forceinline__ void store16(uint32_t* p, const uint32_t s[16]) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(s[0], s[1], s[2], s[3]); q[1] = make_uint4(s[4], s[5], s[6], s[7]);
    q[2] = make_uint4(s[8], s[9], s[10], s[11]); q[3] = make_uint4(s[12], s[13], s[14], s[15]);
}

// fixed iteration count; fp_warps of the 8 warps run the FP64 code.  Every iteration goes through canonical words
// only at the ends (like a sponge: the state stays in registers between permutations).
__global__ void __launch_bounds__(256) iterate(uint32_t* states, int iters, int fp_warps) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t s[16];
    load16(states + i * 16, s);
    if ((int)(threadIdx.x >> 5) < fp_warps) {
        double d[16];
#pragma unroll
        for (int k = 0; k < 16; k++) d[k] = p2f::from_word(s[k]);
        for (int it = 0; it < iters; it++) p2f::permute(d);
#pragma unroll
        for (int k = 0; k < 16; k++) s[k] = p2f::to_word(d[k]);
    } else {
        for (int it = 0; it < iters; it++) p2v2::permute(s);
    }
    store16(states + i * 16, s);
}

// dynamic hand-out: every warp draws batches of `batch` permutations from the CTA's counter until `total` are done
__global__ void __launch_bounds__(256) dynamic(uint32_t* states, int total, int batch, int fp_warps, unsigned* done_fp) {
    __shared__ int next;
    if (threadIdx.x == 0) next = 0;
    __syncthreads();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t s[16];
    load16(states + i * 16, s);
    const bool fp = (int)(threadIdx.x >> 5) < fp_warps;
    double d[16];
    if (fp) {
#pragma unroll
        for (int k = 0; k < 16; k++) d[k] = p2f::from_word(s[k]);
    }
    unsigned mine = 0;
    while (true) {
        int t = 0;
        if ((threadIdx.x & 31) == 0) t = atomicAdd(&next, batch);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= total) break;
        if (fp) {
            for (int it = 0; it < batch; it++) p2f::permute(d);
        } else {
            for (int it = 0; it < batch; it++) p2v2::permute(s);
        }
        mine += batch;
    }
    if (fp) {
#pragma unroll
        for (int k = 0; k < 16; k++) s[k] = p2f::to_word(d[k]);
        if ((threadIdx.x & 31) == 0) atomicAdd(done_fp, mine);
    }
    store16(states + i * 16, s);
}

static float timed(void (*launch)(void*), void* arg) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    launch(arg);  // warm
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    launch(arg);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    if (sms <= 0) return 1;
    // 1. raw pipes: 8 CTAs of 256 threads per SM
    {
        const int ctas = sms * 8, iters = 1 << 15;
        double* dout;
        uint32_t* iout;
        cudaMalloc(&dout, (size_t)ctas * 256 * 8);
        cudaMalloc(&iout, (size_t)ctas * 256 * 4);
        for (int mode = 1; mode <= 3; mode++) {
            struct A { int mode, iters, ctas; double* d; uint32_t* i; } a{mode, iters, ctas, dout, iout};
            float ms = timed([](void* p) { A* a = (A*)p; pipes<<<a->ctas, 256>>>(a->mode, a->iters, a->d, a->i); }, &a);
            const double ops = (double)ctas * 128 * 8.0 * iters;  // per kind: half the warps, 8 chains
            printf("{\"test\": \"pipes\", \"mode\": \"%s\", \"ms\": %.3f, \"dfma_per_clk_per_sm_at_1965\": %.1f, \"imad_per_clk_per_sm_at_1965\": %.1f}\n",
                   mode == 1 ? "dfma only" : mode == 2 ? "imad only" : "dfma + imad warps together", ms,
                   (mode & 1) ? ops / (ms * 1e-3) / sms / 1.965e9 : 0.0, (mode & 2) ? ops / (ms * 1e-3) / sms / 1.965e9 : 0.0);
        }
        cudaFree(dout);
        cudaFree(iout);
    }
    // 2. equality + throughput
    for (int tps : {1024, 2048}) {
        const size_t n = (size_t)tps * sms;
        std::vector<uint32_t> init(n * 16), o_int(n * 16), o_fp(n * 16);
        uint64_t x = 88172645463325252ull;
        for (auto& v : init) {
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            v = (uint32_t)(x % bb::P);
        }
        // a few edge states: zeros, p-1, iota
        for (int k = 0; k < 16; k++) { init[k] = 0; init[16 + k] = bb::P - 1; init[32 + k] = bb::mont(k); }
        uint32_t* d;
        cudaMalloc(&d, n * 64);
        struct A { uint32_t* d; int n, iters, fpw; } a{d, (int)(n / 256), 3, 0};
        auto launch = [](void* p) { A* a = (A*)p; iterate<<<a->n, 256>>>(a->d, a->iters, a->fpw); };
        cudaMemcpy(d, init.data(), n * 64, cudaMemcpyHostToDevice);
        launch(&a);
        cudaMemcpy(o_int.data(), d, n * 64, cudaMemcpyDeviceToHost);
        a.fpw = 8;
        cudaMemcpy(d, init.data(), n * 64, cudaMemcpyHostToDevice);
        launch(&a);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(o_fp.data(), d, n * 64, cudaMemcpyDeviceToHost);
        size_t bad = 0;
        for (size_t i = 0; i < n * 16; i++) bad += o_int[i] != o_fp[i];
        printf("{\"test\": \"fp64 permutation == integer permutation (3 chained perms, %zu states)\", \"mismatches\": %zu, \"err\": \"%s\"}\n", n, bad, cudaGetErrorString(e));
        a.iters = 128;
        for (int fpw : {0, 8, 2, 3, 4}) {
            a.fpw = fpw;
            float ms = timed(launch, &a);
            printf("{\"test\": \"static split\", \"threads_per_sm\": %d, \"fp_warps_of_8\": %d, \"ms\": %.3f, \"gperm_per_s\": %.3f}\n", tps, fpw, ms,
                   (double)n * a.iters / (ms * 1e-3) / 1e9);
        }
        unsigned* done_fp;
        cudaMalloc(&done_fp, 4);
        for (int fpw : {0, 2, 3, 4, 5, 8}) {
            struct B { uint32_t* d; int n, total, batch, fpw; unsigned* done; } b{d, (int)(n / 256), 8 * 128, 4, fpw, done_fp};
            cudaMemset(done_fp, 0, 4);
            float ms = timed([](void* p) { B* b = (B*)p; dynamic<<<b->n, 256>>>(b->d, b->total, b->batch, b->fpw, b->done); }, &b);
            unsigned h = 0;
            cudaMemcpy(&h, done_fp, 4, cudaMemcpyDeviceToHost);
            // every CTA does `total` warp-permutations = total * 32 thread permutations; two launches accumulated done_fp
            const double perms = (double)(n / 256) * b.total * 32;
            printf("{\"test\": \"dynamic hand-out\", \"threads_per_sm\": %d, \"fp_warps_of_8\": %d, \"ms\": %.3f, \"gperm_per_s\": %.3f, \"share_done_by_fp64_warps\": %.3f}\n",
                   tps, fpw, ms, perms / (ms * 1e-3) / 1e9, (double)h * 32 / (2.0 * perms));
        }
        cudaFree(done_fp);
        cudaFree(d);
    }
    return 0;
}
