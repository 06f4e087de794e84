// INT32 issue-rate micro-benchmark for sm_100a: measures lane-ops per clock per SM of the integer
// instructions the BabyBear kernels are built from (the INT32 roofline is not in
// MEASURED_PEAKS.json, so it is measured here).  Each thread runs CH independent dependency chains
// of one instruction (or a fixed mix), long enough to be issue-bound; clocks come from clock64().
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o int_roofline.bin int_roofline.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int CH = 8;
constexpr int ITERS = 2048;

enum Op { IADD, VIADDMN, IMADLO, IMADWIDE, IMADHI, LOP, MIX_MUL, MIX_ADD, MIX_HALF, SHF, MONT_MUL, MONT_SIGNED, MONT_SIGNED_HI, MONT_HI_ONLY };

template <int OP>
__device__ __forceinline__ void step(uint32_t& a, uint64_t& w, uint32_t p, uint32_t b, uint32_t c) {
    if (OP == IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(p));
    if (OP == VIADDMN) {  // min(a + b, a)  -> VIADDMNMX.U32
        asm volatile("{.reg .u32 t; add.u32 t, %0, %1; min.u32 %0, t, %0;}" : "+r"(a) : "r"(b));
    }
    if (OP == IMADLO) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == IMADWIDE) {  // 64-bit accumulator chain: acc = lo(acc) * b + acc
        asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0;}" : "+l"(w) : "r"(b));
    }
    if (OP == IMADHI) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a) : "r"(b));
    if (OP == LOP) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == MIX_MUL) {  // 1 IMAD + 1 IADD per step (dual pipe)
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
        asm volatile("{.reg .u32 t; add.u32 t, %0, %1; min.u32 %0, t, %0;}" : "+r"(a) : "r"(b));
    }
    if (OP == MIX_ADD) {  // canonical add as the compiler emits it: add + viaddmnmx
        uint32_t s = a + b;
        uint32_t t = s - 0x78000001u;
        a = s < t ? s : t;
    }
    if (OP == MIX_HALF) {  // 1 IMAD : 2 ALU
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
        asm volatile("{.reg .u32 t; add.u32 t, %0, %1; min.u32 %0, t, %0;}" : "+r"(a) : "r"(b));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
    }
    if (OP == MONT_MUL) {  // unsigned Montgomery product, canonical output
        uint64_t x = (uint64_t)a * p;
        uint32_t q = (uint32_t)x * 0x77ffffffu;
        uint64_t t = x + (uint64_t)q * 0x78000001u;
        uint32_t r = (uint32_t)(t >> 32);
        uint32_t u = r - 0x78000001u;
        a = r < u ? r : u;
    }
    if (OP == MONT_SIGNED) {  // signed Montgomery product, closed on int32, no correction
        asm volatile("{.reg .s64 x; .reg .s32 lo, hi, q; mul.wide.s32 x, %0, %1; mov.b64 {lo, hi}, x;"
                     " mul.lo.s32 q, lo, 0x77ffffff; mad.wide.s32 x, q, 0x78000001, x; mov.b64 {lo, %0}, x;}" : "+r"(a) : "r"(p));
    }
}

template <>
__device__ __forceinline__ void step<MONT_SIGNED_HI>(uint32_t& a, uint64_t& w, uint32_t p, uint32_t b, uint32_t c) {
    // signed Montgomery: r = hi(a*p) - hi(q*P), q = lo(a*p) * P^-1; closed on int32, no correction
    asm volatile("{.reg .s64 x; .reg .s32 lo, hi, q, h; mul.wide.s32 x, %0, %1; mov.b64 {lo, hi}, x;"
                 " mul.lo.s32 q, lo, 0x88000001; mul.hi.s32 h, q, 0x78000001; sub.s32 %0, hi, h;}" : "+r"(a) : "r"(p));
}
template <>
__device__ __forceinline__ void step<MONT_HI_ONLY>(uint32_t& a, uint64_t& w, uint32_t p, uint32_t b, uint32_t c) {
    // same without IMAD.WIDE: separate lo / hi products
    asm volatile("{.reg .s32 lo, hi, q, h; mul.lo.s32 lo, %0, %1; mul.hi.s32 hi, %0, %1;"
                 " mul.lo.s32 q, lo, 0x88000001; mul.hi.s32 h, q, 0x78000001; sub.s32 %0, hi, h;}" : "+r"(a) : "r"(p));
}

template <int OP>
__global__ void __launch_bounds__(256) bench(uint32_t* out, uint32_t b, uint32_t c, long long* cycles) {
    uint32_t v[CH];
    uint64_t w[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) {
        v[i] = threadIdx.x * 2654435761u + i * 40503u + blockIdx.x;
        w[i] = v[i] * 0x100000001ull;
    }
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it += 4) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < CH; i++) step<OP>(v[i], w[i], v[(i + 1) % CH], b, c);
    }
    long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) acc ^= v[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int ops_per_step, int sms, int blocks_per_sm) {
    const int grid = sms * blocks_per_sm;
    uint32_t* out;
    long long* cyc;
    cudaMalloc(&out, grid * 256 * 4);
    cudaMalloc(&cyc, grid * 8);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    bench<OP><<<grid, 256>>>(out, 0x12345677u, 0x9abcdef1u, cyc);
    cudaEventRecord(a);
    bench<OP><<<grid, 256>>>(out, 0x12345677u, 0x9abcdef1u, cyc);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    long long* h = new long long[grid];
    cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < grid; i++) avg += h[i];
    avg /= grid;
    // all blocks_per_sm CTAs of an SM run concurrently (256 thr, few regs).  ONE rate is reported: lane-ops per second on the
    // whole chip from the CUDA-event time, and the same per clock per SM at the EFFECTIVE clock (cycles a CTA counted with
    // clock64 / the kernel's duration) -- round 1's clock64-only column over-counted because the CTAs of a wave do not all
    // start at the same time.
    const double lane_ops_per_sm = (double)blocks_per_sm * 256 * CH * ITERS * ops_per_step;
    double max_cyc = 0;
    for (int i = 0; i < grid; i++) max_cyc = h[i] > max_cyc ? (double)h[i] : max_cyc;
    const double eff_mhz = max_cyc / (ms * 1e3);
    const double tops = lane_ops_per_sm * sms / (ms * 1e-3) / 1e12;
    printf("{\"op\": \"%s\", \"chip_Tops_per_s\": %.3f, \"effective_mhz\": %.0f, \"lane_ops_per_clk_per_sm\": %.2f, \"ms\": %.4f}\n", name,
           tops, eff_mhz, tops * 1e12 / sms / (eff_mhz * 1e6), ms);
    (void)avg;
    cudaFree(out);
    cudaFree(cyc);
    delete[] h;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    if (sms <= 0) {
        fprintf(stderr, "no CUDA device\n");
        return 1;
    }
    const int bps = 8;
    run<IADD>("IADD3", 1, sms, bps);
    run<VIADDMN>("VIADDMNMX.U32", 1, sms, bps);
    run<LOP>("LOP3", 1, sms, bps);
    run<SHF>("SHF", 1, sms, bps);
    run<IMADLO>("IMAD", 1, sms, bps);
    run<IMADWIDE>("IMAD.WIDE.U32", 1, sms, bps);
    run<IMADHI>("IMAD.HI.U32", 1, sms, bps);
    run<MIX_MUL>("IMAD+VIADDMNMX (1:1)", 2, sms, bps);
    run<MIX_HALF>("IMAD+VIADDMNMX+LOP3 (1:2)", 3, sms, bps);
    run<MIX_ADD>("bb::add (canonical)", 1, sms, bps);
    run<MONT_MUL>("bb::mul (unsigned, canonical)", 1, sms, bps);
    run<MONT_SIGNED>("bb::mul (signed, mad.wide accumulate)", 1, sms, bps);
    run<MONT_SIGNED_HI>("bb::mul (signed, wide+lo+hi+sub)", 1, sms, bps);
    run<MONT_HI_ONLY>("bb::mul (signed, lo+hi+lo+hi+sub)", 1, sms, bps);
    return 0;
}
