// Round-trip latency of "launch a small kernel, get 48 bytes back on the host", the pattern of every
// sumcheck round: (a) cudaStreamSynchronize, (b) the kernel's last thread raises a flag in mapped pinned
// memory and the host spins on it, (c) cuStreamWriteValue32 after the kernel + host spin.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o sync_latency.bin sync_latency.cu
#include <chrono>
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__global__ void round_kernel(const uint32_t* in, uint32_t* result, volatile uint32_t* flag, uint32_t seq, int n) {
    __shared__ uint32_t s[256];
    uint32_t acc = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) acc += in[i];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x < 12) result[threadIdx.x] = s[0] + threadIdx.x;
    if (flag && blockIdx.x == 0) {
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) *flag = seq;
    }
}

typedef CUresult (*WriteValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

int main() {
    cudaStream_t st;
    cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    uint32_t *in, *h_res, *d_res;
    const int n = 1 << 12;
    cudaMalloc(&in, n * 4);
    cudaMemset(in, 0, n * 4);
    cudaHostAlloc(&h_res, 4096, cudaHostAllocMapped);
    cudaHostGetDevicePointer(&d_res, h_res, 0);
    volatile uint32_t* h_flag = h_res + 512;
    uint32_t* d_flag = d_res + 512;
    WriteValue32 wv = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuStreamWriteValue32", (void**)&wv, cudaEnableDefault, &qr);
    const int iters = 2000;
    for (int mode = 0; mode < 3; mode++) {
        uint32_t seq = 0;
        *h_flag = 0;
        uint64_t sink = 0;
        auto run = [&](int k) {
            for (int i = 0; i < k; i++) {
                seq++;
                if (mode == 0) {
                    round_kernel<<<4, 256, 0, st>>>(in, d_res, nullptr, seq, n);
                    cudaStreamSynchronize(st);
                } else if (mode == 1) {
                    round_kernel<<<4, 256, 0, st>>>(in, d_res, d_flag, seq, n);
                    while (*h_flag != seq) {}
                } else {
                    round_kernel<<<4, 256, 0, st>>>(in, d_res, nullptr, seq, n);
                    wv(st, (CUdeviceptr)d_flag, seq, 0);
                    while (*h_flag != seq) {}
                }
                sink += h_res[3];
            }
        };
        run(200);
        cudaStreamSynchronize(st);
        auto t0 = std::chrono::steady_clock::now();
        run(iters);
        cudaStreamSynchronize(st);
        auto t1 = std::chrono::steady_clock::now();
        const char* names[3] = {"cudaStreamSynchronize", "kernel flag in mapped memory + host spin", "cuStreamWriteValue32 + host spin"};
        printf("{\"mode\": \"%s\", \"us_per_round_trip\": %.2f, \"sink\": %llu}\n", names[mode],
               std::chrono::duration<double, std::micro>(t1 - t0).count() / iters, (unsigned long long)sink);
    }
    return 0;
}
