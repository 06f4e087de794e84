// Latency micro-benchmarks behind the design of the sumcheck tails ("host round trips"):
//  1. one Poseidon2 permutation running ALONE on the device: thread-per-state (p2v2::permute) against the
//     warp-cooperative form (poseidon2_warp.cuh, 16 lanes), chained so that only latency counts; outputs are
//     checked against the host permutation;
//  2. a persistent kernel that exchanges one result/challenge pair per round with the host through mapped pinned
//     memory (no launch, no stream synchronisation), against launch + cudaStreamSynchronize per round.
//   nvcc -O3 -std=c++17 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -I../stark-backend_b200/csrc -o latency_bench.bin latency_bench.cu \
//        -L../stark-backend_b200 -lswirl_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../stark-backend_b200'
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#include "poseidon2_v2.cuh"
#include "poseidon2_warp.cuh"
#include "ext.cuh"

__global__ void chain_thread(uint32_t* state, int n) {
    uint32_t s[16];
    for (int i = 0; i < 16; i++) s[i] = state[i];
    for (int it = 0; it < n; it++) p2v2::permute(s);
    for (int i = 0; i < 16; i++) state[i] = s[i];
}
__global__ void chain_warp(uint32_t* state, int n) {
    const int lane = threadIdx.x & 15;
    if (threadIdx.x >= 16) return;
    uint32_t x = state[lane];
    for (int it = 0; it < n; it++) x = p2w::permute(x, lane, 0xffffu);
    state[lane] = x;
}

// persistent kernel: round k publishes 8 words + seq in mapped memory, then waits for the host's reply k
__global__ void handshake_kernel(volatile uint32_t* to_host, volatile uint32_t* from_host, int rounds, uint32_t* sink) {
    uint32_t acc = 1;
    for (int k = 1; k <= rounds; k++) {
        if (threadIdx.x < 8) to_host[threadIdx.x] = acc + threadIdx.x;
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            to_host[16] = (uint32_t)k;
            long long t0 = clock64();
            while (from_host[16] != (uint32_t)k) {
                if (clock64() - t0 > 4000000000ll) break;  // ~2 s: never hang the GPU
            }
        }
        __syncthreads();
        acc = acc * 3 + from_host[0];
    }
    if (threadIdx.x == 0) *sink = acc;
}
__global__ void tiny_kernel(uint32_t* out, uint32_t v) {
    if (threadIdx.x < 8) out[threadIdx.x] = v + threadIdx.x;
}

// the library's round link: one pre-enqueued kernel per round, challenge in / result out through the mapped mailbox
__global__ void linked_round_kernel(swirl::RoundLink link, uint32_t* partials, unsigned int* ticket, uint32_t* result) {
    swirl::Ext r = bb::ext_zero();
    if (!swirl::link_wait(link, r)) return;
    uint32_t v[8];
    for (int k = 0; k < 8; k++) v[k] = bb::add(r.c[k & 3], threadIdx.x == 0 && blockIdx.x == 0 ? 1u : 0u);
    swirl::grid_sum<8>(v, partials, ticket, result, swirl::link_result_tag(link.seq));
}

int main() {
    uint32_t init[16], ref[16];
    for (int i = 0; i < 16; i++) init[i] = ref[i] = (uint32_t)(i * 0x1234567u + 99u) % bb::P;
    const int n = 2000;
    for (int it = 0; it < n; it++) p2v2::permute(ref);
    uint32_t* d;
    cudaMalloc(&d, 64);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int mode = 0; mode < 2; mode++) {
        float best = 1e9f;
        uint32_t out[16];
        for (int rep = 0; rep < 3; rep++) {
            cudaMemcpy(d, init, 64, cudaMemcpyHostToDevice);
            cudaEventRecord(a);
            if (mode == 0)
                chain_thread<<<1, 1>>>(d, n);
            else
                chain_warp<<<1, 32>>>(d, n);
            cudaEventRecord(b);
            cudaDeviceSynchronize();
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            if (ms < best) best = ms;
        }
        cudaMemcpy(out, d, 64, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int i = 0; i < 16; i++) bad += out[i] != ref[i];
        printf("{\"bench\": \"poseidon2 latency\", \"impl\": \"%s\", \"us_per_permutation\": %.3f, \"mismatches_vs_host\": %d, \"err\": \"%s\"}\n",
               mode == 0 ? "one thread (p2v2::permute)" : "16 lanes (p2w::permute)", best * 1e3 / n, bad, cudaGetErrorString(cudaGetLastError()));
    }
    // ---- host <-> persistent kernel handshake ------------------------------------------------------
    uint32_t *h_to, *h_from, *d_to, *d_from, *d_sink;
    cudaHostAlloc(&h_to, 4096, cudaHostAllocMapped);
    cudaHostAlloc(&h_from, 4096, cudaHostAllocMapped);
    cudaHostGetDevicePointer(&d_to, h_to, 0);
    cudaHostGetDevicePointer(&d_from, h_from, 0);
    cudaMalloc(&d_sink, 4);
    for (int threads : {32, 1024}) {
        const int rounds = 20000;
        h_to[16] = 0;
        h_from[16] = 0;
        cudaDeviceSynchronize();
        auto t0 = std::chrono::steady_clock::now();
        handshake_kernel<<<1, threads>>>(d_to, d_from, rounds, d_sink);
        volatile uint32_t* vt = h_to;
        volatile uint32_t* vf = h_from;
        for (int k = 1; k <= rounds; k++) {
            while (vt[16] != (uint32_t)k) {
            }
            vf[0] = vt[0] ^ 0x5u;  // "challenge"
            __sync_synchronize();
            vf[16] = (uint32_t)k;
        }
        cudaDeviceSynchronize();
        const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
        printf("{\"bench\": \"round trip\", \"mode\": \"persistent kernel <-> host through mapped memory, %d threads\", \"us_per_round_trip\": %.3f, \"err\": \"%s\"}\n",
               threads, us / rounds, cudaGetErrorString(cudaGetLastError()));
    }
    {
        const int rounds = 5000;
        cudaStream_t st;
        cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        auto t0 = std::chrono::steady_clock::now();
        for (int k = 1; k <= rounds; k++) {
            tiny_kernel<<<1, 32, 0, st>>>(d_to, (uint32_t)k);
            cudaStreamSynchronize(st);
        }
        const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
        printf("{\"bench\": \"round trip\", \"mode\": \"launch + cudaStreamSynchronize per round\", \"us_per_round_trip\": %.3f}\n", us / rounds);
    }
    {
        uint32_t *h_link, *d_link, *d_gate, *h_res, *d_res, *d_part;
        unsigned int* d_ticket;
        cudaHostAlloc(&h_link, 4096, cudaHostAllocMapped);
        memset(h_link, 0, 4096);
        cudaHostGetDevicePointer(&d_link, h_link, 0);
        cudaHostAlloc(&h_res, 4096, cudaHostAllocMapped);
        memset(h_res, 0, 4096);
        cudaHostGetDevicePointer(&d_res, h_res, 0);
        cudaMalloc(&d_gate, 32);
        cudaMemset(d_gate, 0, 32);
        cudaMalloc(&d_part, 8192 * 64 * 4);
        cudaMalloc(&d_ticket, 4096);
        cudaMemset(d_ticket, 0, 4096);
        cudaStream_t st;
        cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        swirl::RoundScratch rs;
        rs.h_result = h_res;
        rs.d_result = d_res;
        rs.h_link = h_link;
        rs.d_link = d_link;
        rs.d_gate = d_gate;
        swirl_ctx ctx;
        ctx.stream = st;
        for (int blocks : {1, 8, 592}) {
            for (int batch : {1, 16}) {
                const int rounds = 2000;
                cudaStreamSynchronize(st);
                auto t0 = std::chrono::steady_clock::now();
                int bad = 0;
                for (int k0 = 0; k0 < rounds; k0 += batch) {
                    swirl::link_begin(&ctx, &rs, 0, 8);
                    uint32_t seqs[16];
                    for (int b = 0; b < batch; b++) {
                        const swirl::RoundLink l = swirl::link_make(&rs, true);
                        seqs[b] = l.seq;
                        linked_round_kernel<<<blocks, 256, 0, st>>>(l, d_part, d_ticket, d_res);
                    }
                    for (int b = 0; b < batch; b++) {
                        swirl::Ext r{{(uint32_t)(k0 + b), 1u, 2u, 3u}};
                        swirl::link_send(&rs, seqs[b], r);
                        uint32_t out[8];
                        if (swirl::link_recv(&ctx, &rs, seqs[b], 0, 8, out) != 0) bad++;
                        // every thread adds r (blocks * 256 of them), one adds 1 more
                        const uint64_t n = (uint64_t)blocks * 256;
                        if (out[1] != (uint32_t)((n * 1 + 1) % bb::P) || out[0] != (uint32_t)((n * (k0 + b) + 1) % bb::P)) bad++;
                    }
                }
                cudaStreamSynchronize(st);
                const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
                printf("{\"bench\": \"round trip\", \"mode\": \"round link: %d kernels of %d blocks enqueued ahead, mailbox exchange per round\", \"us_per_round_trip\": %.3f, \"wrong_results\": %d, \"err\": \"%s\"}\n",
                       batch, blocks, us / rounds, bad, cudaGetErrorString(cudaGetLastError()));
            }
        }
    }
    return 0;
}
