// ks_internal.cuh — shared device-side views, PTX helpers and launch plumbing (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/ksched.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libksched is written for sm_100a (B200) only"
#endif

namespace ks {

constexpr int TILE_N = 1024;        // nodes per shared-memory tile of the direct kernel
constexpr int DIRECT_THREADS = 256; // 8 warps

// Device-resident node table (SoA, padded to a multiple of TILE_N with never-feasible sentinels:
// free = INT64_MIN, labels = 0).  labels are word-major: labels[w*Npad + n].
struct NodeTable {
    const int64_t* free_cpu;
    const int64_t* free_mem;
    const int64_t* alloc_cpu;
    const int64_t* alloc_mem;
    const uint64_t* labels;
    uint32_t N, Npad, W;
};

struct PodView {
    const int64_t* req_cpu;
    const int64_t* req_mem;
    const uint64_t* sel; // row-major [P*W]
    uint32_t P;
};

struct OutView {
    int32_t* node_idx;
    int64_t* score;
    uint32_t* cnt;
    uint32_t* mask;          // may be nullptr
    uint64_t mask_row_words; // row pitch in 32-bit words
    uint32_t mask_valid_words; // words per row that may be written: 8*ceil(N/256)
};

// Fused all-gather of the bindings (ks_exchange, include/ksched.h): peer-mapped destinations of this rank's shard
struct PeerOut {
    uint32_t n = 0; // peers; 0 = no exchange
    uint32_t world = 1, rank = 0;
    int32_t* idx[KS_MAX_PEERS];
    int64_t* score[KS_MAX_PEERS];
    uint32_t* flag[KS_MAX_PEERS];
    uint32_t* local_flags = nullptr;
    uint32_t* state = nullptr; // [0] = step sequence number, [1] = CTA completion counter, [2..11] = stamps (below)
};

// Diagnostic trace of the last exchange step: %globaltimer (ns) of 0 = first argmax CTA started, 1 = flags published,
// 2 = wait kernel started, 3 = wait kernel saw every peer, 4 = first CTA of the tail kernel started.  Five 64-bit words
// after the two state words.
__device__ __forceinline__ void exchange_stamp(const PeerOut& po, int which) {
    if (po.n == 0) return;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    reinterpret_cast<volatile unsigned long long*>(po.state + 2)[which] = t;
}

// Last CTA of the kernel that completes this rank's bindings: publish a new sequence number to every peer.
// Call at the very end of the kernel, by all threads of the CTA.
__device__ __forceinline__ void exchange_signal(const PeerOut& po) {
    if (po.n == 0) return;
    __syncthreads(); // every store of this CTA has been issued
    if (threadIdx.x == 0) {
        __threadfence_system(); // ... and is visible system-wide before the counter moves
        const uint32_t ctas = gridDim.x * gridDim.y * gridDim.z;
        if (atomicAdd(po.state + 1, 1u) == ctas - 1) {
            po.state[1] = 0; // every CTA has arrived: ready for the next launch
            const uint32_t seq = po.state[0] + 1;
            po.state[0] = seq;
            exchange_stamp(po, 1);
            __threadfence_system();
            for (uint32_t k = 0; k < po.n; k++)
                asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(po.flag[k]), "r"(seq) : "memory");
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(po.local_flags + po.rank), "r"(seq) : "memory");
        }
    }
}

// Per-(chunk,pod) partial results when the node dimension is split across CTAs (small P).
struct PartialView {
    int64_t* key;  // best policy key (LEFTOVER: node priority; LEAST_ALLOCATED: score)
    int32_t* idx;
    uint32_t* cnt;
};

extern std::atomic<uint64_t> g_launches;

// ---- PTX helpers: mbarrier + 1-D TMA bulk copy (cp.async.bulk -> SASS UBLKCP) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy, completion signalled on an mbarrier (bytes multiple of 16, 16B-aligned)
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- launch-side plumbing shared by the API and the kernel files ----
struct SelectLaunch {
    NodeTable nt;
    PodView pv;
    OutView ov;
    int policy;
    cudaStream_t stream;
    // host-space outputs: when set, the launcher may copy these results back as soon as they exist (the bindings
    // are ready long before the mask kernel ends) and clears the pointer it has served
    int32_t* host_node_idx = nullptr;
    int64_t* host_score = nullptr;
    cudaEvent_t ready_event = nullptr; // caller's "bindings are final" event (ks_bindings.bindings_ready_event)
    PeerOut po;                        // fused all-gather of the bindings (po.n == 0: none)
};

} // namespace ks
