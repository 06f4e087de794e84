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
// Returns true (block-uniform) in the block that wrote `result`.
template <int NV>
__device__ __forceinline__ bool group_sum(uint32_t (&v)[NV], uint32_t* __restrict__ partials,
                                          unsigned int* __restrict__ ticket, uint32_t* __restrict__ result,
                                          unsigned nblocks, unsigned bidx, uint32_t tag = 0) {
    __shared__ bool sm_last;
    const uint32_t tot = block_sum<NV>(v);
    if (nblocks == 1) {
        if (threadIdx.x < NV) result[threadIdx.x] = tot | tag;
        return true;
    }
    if (threadIdx.x < NV) partials[(size_t)bidx * NV + threadIdx.x] = tot;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        sm_last = (atomicAdd(ticket, 1u) == nblocks - 1);
        if (sm_last) *ticket = 0;  // before the result is stored: a linked successor may start as soon as the host has seen it
    }
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
        if (threadIdx.x < NV) result[threadIdx.x] = t2 | tag;
    }
    return sm_last;
}
template <int NV>
__device__ __forceinline__ bool grid_sum(uint32_t (&v)[NV], uint32_t* __restrict__ partials,
                                         unsigned int* __restrict__ ticket, uint32_t* __restrict__ result, uint32_t tag = 0) {
    return group_sum<NV>(v, partials, ticket, result, gridDim.x, blockIdx.x, tag);
}

// ---- round link: host <-> kernel mailbox through mapped pinned memory -----------------------------------------
// A sumcheck round is "sweep a table, send a few field elements to the transcript, get a challenge back".  With one
// launch + cudaStreamSynchronize per round that exchange costs ~10 us (tools/latency_bench.cu); a device-resident
// sponge is no way out, because a lone Poseidon2 permutation takes 4.2 us on the GPU (0.86 us on the host).  So the
// transcript stays on the host and the *launches and stream synchronisations* leave the critical path instead: the
// kernels of a round are enqueued while the previous round still runs; the kernel that needs the previous round's
// challenge waits for it on the device (block 0 polls the mapped mailbox, the other blocks poll block 0's relay in
// device memory), and the host polls the round's result words in mapped memory.  (Starting the next kernel early with programmatic
// dependent launch on top of this was measured and changes nothing: profiles/r2m_*.)
// Every access that crosses PCIe is a full round trip (~1.5 us), so each direction is ONE transaction and carries its
// own "ready" mark instead of a separate flag + fence: field words are < 2^31, their top bits are free.
//  * host -> device: the challenge is one aligned 16-byte store; the top bits of its four words spell seq % 14
//    (14 = idle, written when a sumcheck starts; 15 = abort), the kernel polls with one 16-byte volatile load until
//    they spell its own sequence number;
//  * device -> host: the result words carry (seq & 1) in their top bit; the host waits until all of them do.
// Sequence numbers are consecutive inside one sumcheck, so the previous content never looks ready.
// Failure handling: a kernel that sees the abort mark, or waits ~2 s in vain (a dead host must not hang the GPU),
// raises a sticky flag next to the relay and returns WITHOUT publishing or folding; every later waiting kernel returns
// at once, the stream drains, and the host reports the missing round (link_recv) or the raised flag (link_aborted)
// instead of using a made-up challenge.
// Rule for the host: NO CUDA call while an enqueued kernel may be waiting for a mail that has not been sent.  Another
// thread's device-synchronising call (cudaFree, a context being destroyed, a module being loaded) holds the driver's lock
// until all enqueued work has finished; a launch of ours queued behind that lock, with a kernel of ours waiting for us,
// is a deadlock that only the device time-out breaks (seen with three concurrent provers).  Hence
//  * the rounds are enqueued exactly ONE ahead: kernel k + 1 is launched right after the challenge of kernel k has been
//    sent, while k runs;
//  * at most one kernel waits per challenge and it is the LAST launch of its round (a round whose challenge has several
//    consumers folds them into one launch: sr_fold_multi_kernel, ef_fold_multi_kernel);
//  * link_recv only looks at the stream after 250 ms of silence.
struct RoundLink {
    const uint32_t* mail;  // mapped pinned, written by the host: 4 tagged challenge words (16-byte aligned)
    uint32_t* gate;        // device: [0,4) block 0's relay of the mailbox (same format), [4] sticky abort flag
    uint32_t seq;          // this launch (0 = no link: plain stream-ordered kernel)
    uint32_t wait;         // 1: the launch needs a challenge before it starts; 2: same, but a kernel earlier in the stream has
                           // already relayed it (several kernels consume one challenge): every block reads the relay
};
constexpr uint32_t LINK_TAG_IDLE = 14u, LINK_TAG_ABORT = 15u;
__host__ __device__ __forceinline__ uint32_t link_mail_tag(uint32_t seq) { return seq % 14u; }
__host__ __device__ __forceinline__ uint32_t link_result_tag(uint32_t seq) { return (seq & 1u) << 31; }  // 0 when seq == 0

__device__ __forceinline__ bool link_poll(const uint32_t* p, uint32_t want, uint4& v) {
    const long long t0 = clock64();
    for (unsigned it = 1;; it++) {
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
        const uint32_t tag = (v.x >> 31) | ((v.y >> 31) << 1) | ((v.z >> 31) << 2) | ((v.w >> 31) << 3);
        if (tag == want) return true;
        if (tag == LINK_TAG_ABORT) return false;
        if ((it & 255u) == 0 && clock64() - t0 > 4000000000ll) return false;
    }
}
// All threads of every block call this first.  Returns false (block-uniform) when the round was aborted: the kernel
// must return without publishing a result.
__device__ __forceinline__ bool link_wait(const RoundLink& l, Ext& r) {
    __shared__ uint4 sm_link;
    __shared__ bool sm_ok;
    if (l.seq == 0 || !l.wait) return true;
    if (threadIdx.x == 0) {
        volatile uint32_t* g = l.gate;
        uint4 v = make_uint4(0, 0, 0, 0);
        bool ok = g[4] == 0;
        if (ok) {
            const uint32_t want = link_mail_tag(l.seq);
            if (l.wait == 1 && blockIdx.x == 0 && blockIdx.y == 0) {
                ok = link_poll(l.mail, want, v);
                if (!ok) {
                    g[4] = 1;
                    v = make_uint4(0x80000000u, 0x80000000u, 0x80000000u, 0x80000000u);  // abort mark for the other blocks
                }
                asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(l.gate), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                             : "memory");  // one 16-byte store: the relay is consistent as well
            } else {
                ok = link_poll(l.gate, want, v);
            }
        }
        sm_link = v;
        sm_ok = ok;
    }
    __syncthreads();
    const uint4 v = sm_link;
    r = Ext{{v.x & 0x7fffffffu, v.y & 0x7fffffffu, v.z & 0x7fffffffu, v.w & 0x7fffffffu}};
    return sm_ok;
}

// Scratch every sumcheck-type phase needs: block partials, the ticket, and a mapped pinned result
// area the host reads after synchronising the stream.
struct RoundScratch {
    uint32_t* d_partials = nullptr;  // max_blocks * max_nv words
    unsigned int* d_ticket = nullptr;  // 1024 zero-initialised tickets (one per group of a launch)
    uint32_t* h_result = nullptr;  // pinned, mapped
    uint32_t* d_result = nullptr;  // device alias of h_result
    int max_blocks = 0;
    // round link (see RoundLink)
    uint32_t* h_link = nullptr;  // mapped pinned mailbox (4 words used)
    uint32_t* d_link = nullptr;  // device alias
    uint32_t* d_gate = nullptr;  // 8 words of device memory
    uint32_t link_seq = 0;
};
int round_scratch_get(swirl_ctx* ctx, RoundScratch** out);

// ---- host side of the round link ----
// Start of a sumcheck whose rounds leave `nv` result words at h_result + offset: marks them "not ready" for the first
// sequence number (earlier, unlinked kernels store untagged words there).  All earlier rounds must have been consumed.
// Also resets the mailbox to "idle" and (stream-ordered) the relay and its abort flag.
void link_begin(swirl_ctx* ctx, RoundScratch* rs, size_t offset, size_t nv);
// Marks `nv` result words "not ready" for the launch `seq`.  Needed between rounds only when a round may read result
// words that the previous round did not write (their mark would be two rounds old, i.e. look ready); call it after the
// previous round has been received and before its challenge is sent.
inline void link_expect(RoundScratch* rs, size_t offset, size_t nv, uint32_t seq) {
    const uint32_t not_ready = link_result_tag(seq) ^ 0x80000000u;
    volatile uint32_t* r = rs->h_result + offset;
    for (size_t i = 0; i < nv; i++) r[i] = not_ready;
    __atomic_thread_fence(__ATOMIC_SEQ_CST);
}
inline RoundLink link_make(RoundScratch* rs, bool wait) {
    ++rs->link_seq;
    return RoundLink{rs->d_link, rs->d_gate, rs->link_seq, wait ? 1u : 0u};
}
// Kernels that wait for a challenge but publish nothing (folds) cannot report a time-out through a missing result: the
// sticky flag they raise is fetched with the data the phase copies back anyway (link_flag_fetch before that copy's stream
// synchronisation, link_aborted after it), so a kernel that gave up waiting ends the proof with an error, never with
// tables that were silently left unfolded.
inline cudaError_t link_flag_fetch(swirl_ctx* ctx, RoundScratch* rs) {
    return cudaMemcpyAsync(rs->h_link + 64, rs->d_gate + 4, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
}
inline bool link_aborted(RoundScratch* rs) { return ((volatile uint32_t*)rs->h_link)[64] != 0; }
// An early return between link_begin and the last link_send must release the kernels that still wait.
struct LinkAbortGuard {
    swirl_ctx* ctx;
    RoundScratch* rs;
    bool armed = false;
    ~LinkAbortGuard();
};
void link_send(RoundScratch* rs, uint32_t seq, const Ext& r);  // challenge for the launch `seq`
void link_abort(RoundScratch* rs);                             // releases every launch that still waits
// Waits until the launch `seq` has left its result words in host memory: `groups` runs of `nv` words, `stride` words
// apart, starting at h_result + offset; copies them (untagged) to out at the same spacing.
// Polls the stream every ~0.1 ms so that a faulted kernel surfaces as its CUDA error instead of a hang.
int link_recv(swirl_ctx* ctx, RoundScratch* rs, uint32_t seq, size_t offset, int nv, size_t groups, size_t stride, uint32_t* out);
inline int link_recv(swirl_ctx* ctx, RoundScratch* rs, uint32_t seq, size_t offset, int nv, uint32_t* out) {
    return link_recv(ctx, rs, seq, offset, nv, 1, 0, out);
}

}  // namespace swirl
