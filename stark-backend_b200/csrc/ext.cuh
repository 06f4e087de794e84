// Device helpers shared by the sumcheck-type kernels (GKR, stacked reduction, WHIR, batch
// constraints): EF loads/stores as 16-byte vectors, block-wide EF sums, and the "last block
// finishes the reduction" epilogue that leaves a round's few field elements in mapped host memory.
#pragma once
#include "bb31.cuh"
#include "common.cuh"

namespace swirl {

using bb::Ext;

__device__ __forceinline__ Ext ld_ext(const uint32_t* p) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    return Ext{{v.x, v.y, v.z, v.w}};
}
__device__ __forceinline__ Ext ldg_ext(const uint32_t* p) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    return Ext{{v.x, v.y, v.z, v.w}};
}
__device__ __forceinline__ void st_ext(uint32_t* p, const Ext& e) {
    *reinterpret_cast<uint4*>(p) = make_uint4(e.c[0], e.c[1], e.c[2], e.c[3]);
}
__host__ __device__ __forceinline__ Ext ext_one_minus(const Ext& a) { return bb::ext_sub(bb::ext_one(), a); }
// t0 + (t1 - t0) * r
__host__ __device__ __forceinline__ Ext ext_lerp(const Ext& t0, const Ext& t1, const Ext& r) {
    return bb::ext_add(t0, bb::ext_mul(bb::ext_sub(t1, t0), r));
}

__device__ __forceinline__ uint32_t warp_sum_bb(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = bb::add(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum of NV words per thread; thread i < NV returns total i (others return garbage).
template <int NV>
__device__ __forceinline__ uint32_t block_sum(uint32_t (&v)[NV]) {
    __shared__ uint32_t sm_part[32][NV];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    __syncthreads();  // sm_part may still be read by a previous call
#pragma unroll
    for (int i = 0; i < NV; i++) {
        const uint32_t s = warp_sum_bb(v[i]);
        if (lane == 0) sm_part[warp][i] = s;
    }
    __syncthreads();
    uint32_t s = 0;
    if (threadIdx.x < NV)
        for (int w = 0; w < nwarps; w++) s = bb::add(s, sm_part[w][threadIdx.x]);
    return s;
}

// Sums NV base-field words per thread over the whole grid.  Every block writes its partial to
// `partials[blockIdx.x * NV ..]`; the last block to finish (atomic ticket) adds the partials and
// writes the NV totals to `result` (mapped pinned host memory or device).  Field addition is
// exact, so the summation order does not affect the value.  `ticket` must be zero on entry and is
// reset for the next launch.  blockDim.x must be a multiple of 32, <= 1024, >= NV.
// Group form: the blocks [0, nblocks) of one logical group (bidx = this block's index in the group)
// reduce into `result`; several groups may share a launch, each with its own ticket and partials.
template <int NV>
__device__ __forceinline__ void group_sum(uint32_t (&v)[NV], uint32_t* __restrict__ partials,
                                          unsigned int* __restrict__ ticket, uint32_t* __restrict__ result,
                                          unsigned nblocks, unsigned bidx) {
    __shared__ bool sm_last;
    const uint32_t tot = block_sum<NV>(v);
    if (nblocks == 1) {
        if (threadIdx.x < NV) result[threadIdx.x] = tot;
        return;
    }
    if (threadIdx.x < NV) partials[(size_t)bidx * NV + threadIdx.x] = tot;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) sm_last = (atomicAdd(ticket, 1u) == nblocks - 1);
    __syncthreads();
    if (sm_last) {
        __threadfence();
        uint32_t acc[NV];
#pragma unroll
        for (int i = 0; i < NV; i++) acc[i] = 0;
        for (unsigned b = threadIdx.x; b < nblocks; b += blockDim.x) {
#pragma unroll
            for (int i = 0; i < NV; i++) acc[i] = bb::add(acc[i], __ldcg(partials + (size_t)b * NV + i));
        }
        const uint32_t t2 = block_sum<NV>(acc);
        if (threadIdx.x < NV) result[threadIdx.x] = t2;
        if (threadIdx.x == 0) *ticket = 0;
    }
}
template <int NV>
__device__ __forceinline__ void grid_sum(uint32_t (&v)[NV], uint32_t* __restrict__ partials,
                                         unsigned int* __restrict__ ticket, uint32_t* __restrict__ result) {
    group_sum<NV>(v, partials, ticket, result, gridDim.x, blockIdx.x);
}

// Scratch every sumcheck-type phase needs: block partials, the ticket, and a mapped pinned result
// area the host reads after synchronising the stream.
struct RoundScratch {
    uint32_t* d_partials = nullptr;  // max_blocks * max_nv words
    unsigned int* d_ticket = nullptr;  // 1024 zero-initialised tickets (one per group of a launch)
    uint32_t* h_result = nullptr;  // pinned, mapped
    uint32_t* d_result = nullptr;  // device alias of h_result
    int max_blocks = 0;
};
int round_scratch_get(swirl_ctx* ctx, RoundScratch** out);

}  // namespace swirl
