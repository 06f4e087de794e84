// Device-side prelude of the run-time compiled constraint kernels (jit.hpp).  Self-contained CUDA C++ for NVRTC: no
// #include, every type spelled out.  It restates, for the generated straight-line programs, exactly what the interpreter
// kernel `batch_round0_kernel<NS, 4>` of batch.cu does around `run_program_on`: same thread -> (hypercube point, coset
// point) mapping, same 16-term Lagrange dot product per column load, same reduction into `partials`, so the two are
// interchangeable launch for launch.  All field functions return canonical representatives, hence bit-identical results.
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
typedef unsigned short uint16_t;

#define SW_P 0x78000001u
#define SW_NEG_PINV 0x77ffffffu
#define SW_R1 0x0ffffffeu
#define SW_BETA 0x37ffffe9u /* Montgomery form of 11 */

__device__ __forceinline__ uint32_t f_add(uint32_t a, uint32_t b) {
    const uint32_t s = a + b, t = s - SW_P;
    return s < t ? s : t;
}
__device__ __forceinline__ uint32_t f_sub(uint32_t a, uint32_t b) {
    const uint32_t d = a - b, t = d + SW_P;
    return d < t ? d : t;
}
__device__ __forceinline__ uint32_t f_neg(uint32_t a) { return a ? SW_P - a : 0u; }
__device__ __forceinline__ uint32_t f_reduce(uint64_t x) {  // x < p * 2^32 -> x / 2^32 mod p, canonical
    const uint32_t q = (uint32_t)x * SW_NEG_PINV;
    const uint64_t t = x + (uint64_t)q * SW_P;
    const uint32_t r = (uint32_t)(t >> 32), u = r - SW_P;
    return r < u ? r : u;
}
__device__ __forceinline__ uint32_t f_mul(uint32_t a, uint32_t b) { return f_reduce((uint64_t)a * b); }
__device__ __forceinline__ uint32_t f_dot4(uint32_t a0, uint32_t b0, uint32_t a1, uint32_t b1, uint32_t a2, uint32_t b2, uint32_t a3,
                                           uint32_t b3) {
    uint64_t s = ((uint64_t)a0 * b0 + (uint64_t)a1 * b1) + ((uint64_t)a2 * b2 + (uint64_t)a3 * b3);
    const uint32_t hi = (uint32_t)(s >> 32), hm = hi - SW_P;  // conditional - p*2^32: one unsigned min on the high word
    return f_reduce(((uint64_t)(hi < hm ? hi : hm) << 32) | (uint32_t)s);
}

struct Ext {
    uint32_t c[4];
};
__device__ __forceinline__ Ext ext_zero() { return Ext{{0u, 0u, 0u, 0u}}; }
__device__ __forceinline__ Ext ext_add(const Ext& a, const Ext& b) {
    return Ext{{f_add(a.c[0], b.c[0]), f_add(a.c[1], b.c[1]), f_add(a.c[2], b.c[2]), f_add(a.c[3], b.c[3])}};
}
__device__ __forceinline__ Ext ext_mul_base(const Ext& a, uint32_t s) {
    return Ext{{f_mul(a.c[0], s), f_mul(a.c[1], s), f_mul(a.c[2], s), f_mul(a.c[3], s)}};
}
__device__ __forceinline__ Ext ext_mul(const Ext& a, const Ext& b) {
    const uint32_t w1 = f_mul(b.c[1], SW_BETA), w2 = f_mul(b.c[2], SW_BETA), w3 = f_mul(b.c[3], SW_BETA);
    Ext r;
    r.c[0] = f_dot4(a.c[0], b.c[0], a.c[1], w3, a.c[2], w2, a.c[3], w1);
    r.c[1] = f_dot4(a.c[0], b.c[1], a.c[1], b.c[0], a.c[2], w3, a.c[3], w2);
    r.c[2] = f_dot4(a.c[0], b.c[2], a.c[1], b.c[1], a.c[2], b.c[0], a.c[3], w3);
    r.c[3] = f_dot4(a.c[0], b.c[3], a.c[1], b.c[2], a.c[2], b.c[1], a.c[3], b.c[0]);
    return r;
}
__device__ __forceinline__ Ext ldg_ext(const uint32_t* p) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    return Ext{{v.x, v.y, v.z, v.w}};
}

// ---- the structures of batch.cu, member for member -----------------------------------------------------------
struct BasePart {
    const uint32_t* ptr;  // column-major, column stride = height
    uint32_t height;      // power of two
    uint32_t rot;
};
struct R0Args {
    const void* code;
    uint32_t n_instr;
    const BasePart* parts;
    const uint32_t* weights;
    const uint32_t* lde;
    const uint32_t* eq_xi;
    int l_skip, n_lift, P, x_per_block;
    uint32_t first_block, n_blocks;
    uint32_t n_parts;
    uint32_t* partials;
    uint32_t* result;
};
#define SW_R0_SMEM_PARTS 24
#define SW_BLOCK 256

__device__ __forceinline__ uint32_t chunk_dot16(const uint32_t (&l)[16], const uint32_t (&c)[16]) {
    const uint32_t a0 = f_dot4(l[0], c[0], l[1], c[1], l[2], c[2], l[3], c[3]);
    const uint32_t a1 = f_dot4(l[4], c[4], l[5], c[5], l[6], c[6], l[7], c[7]);
    const uint32_t a2 = f_dot4(l[8], c[8], l[9], c[9], l[10], c[10], l[11], c[11]);
    const uint32_t a3 = f_dot4(l[12], c[12], l[13], c[13], l[14], c[14], l[15], c[15]);
    return f_add(f_add(a0, a1), f_add(a2, a3));
}

// value of column `col` of part `bp` at this thread's coset point for the hypercube point x: Lagrange combination of the
// 16-row chunk (l_skip = 4)
SW_LOAD_ATTR uint32_t load_point(const BasePart bp, uint32_t col, size_t x, const uint32_t (&lreg)[16], const uint32_t* lde) {
    const uint32_t* c = bp.ptr + (size_t)col * bp.height;
    const size_t r0 = x << 4;
    if (bp.height >= 16) {
        const uint4* q = reinterpret_cast<const uint4*>(c + r0);
        const uint4 v0 = __ldg(q), v1 = __ldg(q + 1), v2 = __ldg(q + 2), v3 = __ldg(q + 3);
        uint32_t ch[16] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w, v3.x, v3.y, v3.z, v3.w};
        if (bp.rot) {  // rows r0+1 .. r0+16 (cyclic)
            const uint32_t nxt = __ldg(c + ((r0 + 16) & (bp.height - 1)));
#pragma unroll
            for (int i = 0; i < 15; i++) ch[i] = ch[i + 1];
            ch[15] = nxt;
        }
        return chunk_dot16(lreg, ch);
    }
    uint32_t v = 0;
    for (int i = 0; i < 16; i++) v = f_add(v, f_mul(__ldg(lde + i), __ldg(c + ((r0 + bp.rot + i) & (bp.height - 1)))));
    return v;
}

// statements the generator emits (two hypercube points per thread in lockstep, lanes _0 and _1)
#define LD(part, col, o0, o1)                                                        \
    {                                                                                \
        const BasePart bp_ = parts_in_smem ? sparts[part] : a.parts[part];           \
        o0 = load_point(bp_, col, xs0, lreg, lde);                                   \
        o1 = load_point(bp_, col, xs1, lreg, lde);                                   \
    }
#define PF(part, col)                                                                                               \
    {                                                                                                               \
        const BasePart bp_ = parts_in_smem ? sparts[part] : a.parts[part];                                          \
        const uint32_t* c_ = bp_.ptr + (size_t)(col) * bp_.height;                                                  \
        asm volatile("prefetch.global.L1 [%0];" ::"l"(c_ + (((xs0 << 4) + bp_.rot) & (bp_.height - 1))));           \
        asm volatile("prefetch.global.L1 [%0];" ::"l"(c_ + (((xs1 << 4) + bp_.rot) & (bp_.height - 1))));           \
    }
// acc[lane][k] += weights[w] * v
#define ACC(k, w, v0, v1)                                   \
    {                                                       \
        const Ext w_ = ldg_ext(a.weights + 4 * (w));        \
        acc0[k] = ext_add(acc0[k], ext_mul_base(w_, v0));   \
        acc1[k] = ext_add(acc1[k], ext_mul_base(w_, v1));   \
    }

// A generated kernel is:   SW_R0_SIGNATURE(name) { SW_R0_PROLOGUE  <statements>  SW_R0_EPILOGUE }
#define SW_R0_SIGNATURE(NAME)                                                                      \
    extern "C" __global__ void __launch_bounds__(SW_BLOCK, SW_MIN_BLOCKS) NAME(const R0Args* __restrict__ descs, \
                                                                               const uint16_t* __restrict__ block_air)
#define SW_R0_PROLOGUE                                                                                                \
    __shared__ uint32_t sm[SW_BLOCK * 13];                                                                            \
    __shared__ BasePart sparts[SW_R0_SMEM_PARTS];                                                                     \
    const R0Args a = descs[block_air[blockIdx.x]];                                                                    \
    const uint32_t bidx = blockIdx.x - a.first_block;                                                                 \
    if (threadIdx.x < a.n_parts && threadIdx.x < SW_R0_SMEM_PARTS) sparts[threadIdx.x] = a.parts[threadIdx.x];        \
    __syncthreads();                                                                                                  \
    const bool parts_in_smem = a.n_parts <= SW_R0_SMEM_PARTS;                                                         \
    const int P = a.P;                                                                                                \
    const int G = blockDim.x / P;                                                                                     \
    const bool idle = (int)threadIdx.x >= P * G;                                                                      \
    const int p = idle ? 0 : threadIdx.x % P, g = idle ? 0 : threadIdx.x / P;                                         \
    const size_t nx = size_t(1) << a.n_lift;                                                                          \
    const size_t x0 = (size_t)bidx * a.x_per_block, x1 = idle ? 0 : min(x0 + (size_t)a.x_per_block, nx);              \
    const uint32_t* lde = a.lde + (size_t)p * 16;                                                                     \
    uint32_t lreg[16];                                                                                                \
    _Pragma("unroll") for (int i = 0; i < 16; i++) lreg[i] = __ldg(lde + i);                                          \
    Ext tot[3] = {ext_zero(), ext_zero(), ext_zero()};                                                                \
    for (size_t xb = x0 + g; xb < x1; xb += (size_t)2 * G) {                                                          \
        const bool live1 = xb + (size_t)G < x1;                                                                       \
        const size_t xs0 = xb, xs1 = live1 ? xb + (size_t)G : xb;                                                     \
        Ext acc0[3] = {ext_zero(), ext_zero(), ext_zero()}, acc1[3] = {ext_zero(), ext_zero(), ext_zero()};
#define SW_R0_EPILOGUE                                                                                                \
        {                                                                                                             \
            const Ext e = ldg_ext(a.eq_xi + 4 * xs0);                                                                 \
            _Pragma("unroll") for (int k = 0; k < 3; k++) tot[k] = ext_add(tot[k], ext_mul(e, acc0[k]));              \
        }                                                                                                             \
        if (live1) {                                                                                                  \
            const Ext e = ldg_ext(a.eq_xi + 4 * xs1);                                                                 \
            _Pragma("unroll") for (int k = 0; k < 3; k++) tot[k] = ext_add(tot[k], ext_mul(e, acc1[k]));              \
        }                                                                                                             \
    }                                                                                                                 \
    _Pragma("unroll") for (int k = 0; k < 3; k++)                                                                     \
        _Pragma("unroll") for (int c = 0; c < 4; c++) sm[threadIdx.x * 13 + 4 * k + c] = tot[k].c[c];                 \
    __syncthreads();                                                                                                  \
    for (int o = threadIdx.x; o < P * 12; o += blockDim.x) {                                                          \
        const int pi = o / 12, k = o % 12;                                                                            \
        uint32_t s = 0;                                                                                               \
        for (int gg = 0; gg < G; gg++) s = f_add(s, sm[(gg * P + pi) * 13 + k]);                                      \
        a.partials[(size_t)bidx * (P * 12) + o] = s;                                                                  \
    }

// ---- MLE rounds: what `batch_mle_kernel<NS, D>` of batch.cu does around `run_program`, for generated programs ----------
// One launch serves a run of AIRs that share the kernel; a descriptor is one (AIR, sub-program) pair and names its
// `case` of the generated switch.  Sub-programs that differ only by a constant shift of their columns and weights (the
// same constraint applied to the next 16 columns) share a case: the shifts travel in the descriptor.  Value slots are
// arrays of SW_D extension-field lanes (X = 1..D in lockstep), all indices static after unrolling.
#ifdef SW_D
__device__ __forceinline__ Ext ext_one() { return Ext{{SW_R1, 0u, 0u, 0u}}; }
__device__ __forceinline__ Ext ext_from(uint32_t a) { return Ext{{a, 0u, 0u, 0u}}; }
__device__ __forceinline__ Ext ext_sub(const Ext& a, const Ext& b) {
    return Ext{{f_sub(a.c[0], b.c[0]), f_sub(a.c[1], b.c[1]), f_sub(a.c[2], b.c[2]), f_sub(a.c[3], b.c[3])}};
}
__device__ __forceinline__ Ext ext_neg(const Ext& a) { return Ext{{f_neg(a.c[0]), f_neg(a.c[1]), f_neg(a.c[2]), f_neg(a.c[3])}}; }

struct MleArgs {  // batch.cu, member for member
    const void* code;
    uint32_t n_instr;
    const uint32_t* base;
    size_t h;
    const uint32_t* weights;
    const uint32_t* eq_xi;
    size_t ny;
    int single;
    uint32_t first_block, n_blocks;
    uint32_t* partials;
    unsigned int* ticket;
    uint32_t* result;
    uint32_t sub, col_shift, w_shift;
};

#ifndef SW_HOST_EMU
__device__ __forceinline__ uint32_t warp_sum_f(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = f_add(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <int NV>
__device__ __forceinline__ uint32_t block_sum_f(uint32_t (&v)[NV]) {
    __shared__ uint32_t sm_part[32][NV];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; i++) {
        const uint32_t s = warp_sum_f(v[i]);
        if (lane == 0) sm_part[warp][i] = s;
    }
    __syncthreads();
    uint32_t s = 0;
    if (threadIdx.x < NV)
        for (int w = 0; w < nwarps; w++) s = f_add(s, sm_part[w][threadIdx.x]);
    return s;
}
// ext.cuh: group_sum -- the last block of the group to finish adds the partials and publishes the tagged totals
template <int NV>
__device__ __forceinline__ void group_sum_f(uint32_t (&v)[NV], uint32_t* __restrict__ partials, unsigned int* __restrict__ ticket,
                                            uint32_t* __restrict__ result, unsigned nblocks, unsigned bidx, uint32_t tag) {
    __shared__ bool sm_last;
    const uint32_t tot = block_sum_f<NV>(v);
    if (nblocks == 1) {
        if (threadIdx.x < NV) result[threadIdx.x] = tot | tag;
        return;
    }
    if (threadIdx.x < NV) partials[(size_t)bidx * NV + threadIdx.x] = tot;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        sm_last = (atomicAdd(ticket, 1u) == nblocks - 1);
        if (sm_last) *ticket = 0;
    }
    __syncthreads();
    if (sm_last) {
        __threadfence();
        uint32_t acc[NV];
#pragma unroll
        for (int i = 0; i < NV; i++) acc[i] = 0;
        for (unsigned b = threadIdx.x; b < nblocks; b += blockDim.x) {
#pragma unroll
            for (int i = 0; i < NV; i++) acc[i] = f_add(acc[i], __ldcg(partials + (size_t)b * NV + i));
        }
        const uint32_t t2 = block_sum_f<NV>(acc);
        if (threadIdx.x < NV) result[threadIdx.x] = t2 | tag;
    }
}
#else
// host emulation (tests): the harness runs the threads one after the other; totals are accumulated in place
template <int NV>
void group_sum_f(uint32_t (&v)[NV], uint32_t*, unsigned int*, uint32_t* result, unsigned, unsigned, uint32_t tag) {
    for (int i = 0; i < NV; i++) result[i] = f_add(result[i] & 0x7fffffffu, v[i]) | tag;
}
#endif

// rows 2y, 2y+1 of column `gcol` -> its values at X = 1..D: t1, t1 + d, t1 + 2d, ...; a single-row table gives lane 0
__device__ __forceinline__ void mle_load(const MleArgs& a, size_t y, uint32_t gcol, Ext (&out)[SW_D]) {
    const uint32_t* c = a.base + ((size_t)gcol * a.h) * 4;
    if (a.single) {
        const Ext t = ldg_ext(c);
#pragma unroll
        for (int l = 0; l < SW_D; l++) out[l] = t;
        return;
    }
    const Ext t0 = ldg_ext(c + 8 * y), t1 = ldg_ext(c + 8 * y + 4);
    const Ext d = ext_sub(t1, t0);
    out[0] = t1;
#pragma unroll
    for (int l = 1; l < SW_D; l++) out[l] = ext_add(out[l - 1], d);
}
#ifndef SW_HOST_EMU
__device__ __forceinline__ void mle_prefetch(const MleArgs& a, size_t y, uint32_t gcol) {
    const uint32_t* c = a.base + ((size_t)gcol * a.h) * 4;
    asm volatile("prefetch.global.L1 [%0];" ::"l"(c + (a.single ? 0 : 8 * y)));
}
#else
inline void mle_prefetch(const MleArgs&, size_t, uint32_t) {}
#endif

// statements the generator emits; cs / ws are the descriptor's column / weight shifts
#define SW_LANES _Pragma("unroll") for (int l_ = 0; l_ < SW_D; l_++)
#define LDV(dst, gcol) mle_load(a, y, (uint32_t)(gcol) + cs, dst);
#define PFV(gcol) mle_prefetch(a, y, (uint32_t)(gcol) + cs);
#define CST(dst, val) { SW_LANES dst[l_] = ext_from(val); }
#define ADD(dst, p, q) { SW_LANES dst[l_] = ext_add(p[l_], q[l_]); }
#define SUB(dst, p, q) { SW_LANES dst[l_] = ext_sub(p[l_], q[l_]); }
#define MUL(dst, p, q) { SW_LANES dst[l_] = ext_mul(p[l_], q[l_]); }
#define NEG(dst, p) { SW_LANES dst[l_] = ext_neg(p[l_]); }
// acc[lane][k] += weights[w] * s   /   += weights[w] * (p * q)
#define ACCV(k, w, s)                                                      \
    {                                                                      \
        const Ext w_ = ldg_ext(a.weights + 4 * ((uint32_t)(w) + ws));      \
        SW_LANES acc[l_][k] = ext_add(acc[l_][k], ext_mul(w_, s[l_]));     \
    }
#define MACV(k, w, p, q)                                                               \
    {                                                                                  \
        const Ext w_ = ldg_ext(a.weights + 4 * ((uint32_t)(w) + ws));                  \
        SW_LANES acc[l_][k] = ext_add(acc[l_][k], ext_mul(w_, ext_mul(p[l_], q[l_]))); \
    }
#define SLOT(name) Ext name[SW_D];

// A generated kernel is:  SW_MLE_SIGNATURE(name) { SW_MLE_PROLOGUE  case 0: { ... } break; ...  SW_MLE_EPILOGUE }
// `block_air` points at the launch's first block; block_base is that block's index in the round's numbering (the
// descriptors' first_block counts in it).
#define SW_MLE_SIGNATURE(NAME)                                                                                          \
    extern "C" __global__ void __launch_bounds__(128, SW_MLE_MIN_BLOCKS) NAME(const MleArgs* __restrict__ descs,        \
                                                                              const uint16_t* __restrict__ block_air,   \
                                                                              uint32_t block_base, uint32_t result_tag)
#define SW_MLE_PROLOGUE                                                                                                 \
    const MleArgs a = descs[block_air[blockIdx.x]];                                                                     \
    const uint32_t bidx = blockIdx.x + block_base - a.first_block;                                                      \
    const uint32_t cs = a.col_shift, ws = a.w_shift;                                                                    \
    uint32_t v[SW_D * 12];                                                                                              \
    _Pragma("unroll") for (int i = 0; i < SW_D * 12; i++) v[i] = 0;                                                     \
    for (size_t y = (size_t)bidx * blockDim.x + threadIdx.x; y < a.ny; y += (size_t)a.n_blocks * blockDim.x) {          \
        const Ext e = a.single ? ext_one() : ldg_ext(a.eq_xi + 4 * y);                                                  \
        Ext acc[SW_D][3];                                                                                               \
        _Pragma("unroll") for (int X = 0; X < SW_D; X++)                                                                \
            _Pragma("unroll") for (int k = 0; k < 3; k++) acc[X][k] = ext_zero();                                       \
        switch (a.sub) {
#define SW_MLE_EPILOGUE                                                                                                 \
            default: break;                                                                                             \
        }                                                                                                               \
        _Pragma("unroll") for (int X = 0; X < SW_D; X++)                                                                \
            _Pragma("unroll") for (int k = 0; k < 3; k++) {                                                             \
                const Ext t = ext_mul(e, acc[X][k]);                                                                    \
                _Pragma("unroll") for (int c = 0; c < 4; c++) v[X * 12 + 4 * k + c] = f_add(v[X * 12 + 4 * k + c], t.c[c]); \
            }                                                                                                           \
    }                                                                                                                   \
    group_sum_f<SW_D * 12>(v, a.partials, a.ticket, a.result, a.n_blocks, bidx, result_tag);
#endif  // SW_D
