// ks_launch.h — host-callable launchers implemented in ks_direct.cu / ks_bitpar.cu
#pragma once
#include "ks_internal.cuh"

namespace ks {

cudaError_t launch_free_reduce(int64_t* free_cpu, int64_t* free_mem, const int32_t* bnode, const int64_t* bcpu,
                               const int64_t* bmem, uint64_t B, cudaStream_t st);
cudaError_t launch_node_prio(const int64_t* free_cpu, const int64_t* free_mem, int64_t* prio, uint32_t N,
                             uint32_t Npad, int* range_flag, cudaStream_t st);
cudaError_t launch_check_cells(const NodeTable& nt, const PodView& pv, uint8_t* codes, uint32_t node_begin,
                               uint32_t node_count, cudaStream_t st);
cudaError_t launch_select_sampling(const NodeTable& nt, const PodView& pv, uint32_t attempts, uint64_t seed,
                                   uint64_t pod_offset, int32_t* node_idx, uint32_t* n_attempts, int32_t* draw_node,
                                   uint8_t* draw_code, cudaStream_t st);
// multi-GPU exchange helpers (ks_direct.cu): push finished bindings to the peers (per-cell path), wait for all ranks
cudaError_t launch_exchange_push(const PeerOut& po, const int32_t* node_idx, const int64_t* score, uint32_t P, cudaStream_t st);
cudaError_t launch_exchange_wait(const PeerOut& po, int* error_flag, cudaStream_t st);
cudaError_t launch_fill256(void* dst, uint64_t bytes, uint32_t v, int sms, cudaStream_t st);
uint32_t direct_pods_per_cta(uint32_t W);
cudaError_t prepare_select_direct(uint32_t W);
cudaError_t launch_select_direct(const SelectLaunch& L, const PartialView& part, uint32_t n_chunks,
                                 uint32_t tiles_per_chunk);

} // namespace ks

namespace ks {
cudaError_t launch_stream_resolve(int64_t* free_cpu, int64_t* free_mem, const int32_t* claim_node, const int64_t* req_cpu,
                                  const int64_t* req_mem, uint32_t n, uint32_t N, uint8_t* accepted, cudaStream_t st);
uint32_t stream_max_claims();

// device-side micro-batch loop (k_stream_batch, ks_stream.cu)
constexpr uint32_t STREAM_BATCH_MAX = 1024; // pods per cooperative launch
struct StreamBatchArgs {
    uint32_t N, Npad;
    const int64_t *alloc_cpu, *alloc_mem;
    const uint64_t* labels;
    int64_t *free_cpu, *free_mem;
    int policy;
    uint32_t m;
    const int64_t *req_cpu, *req_mem;
    const uint64_t* sel;
    int64_t* pkey;  // [STREAM_BATCH_MAX * grid]
    int32_t* pidx;  // [STREAM_BATCH_MAX * grid]
    uint32_t* pend; // [2 * STREAM_BATCH_MAX]
    uint32_t* ctl;  // [2]: pending count, rounds
    int32_t* out_idx;
    int64_t* out_score;
    uint32_t max_rounds;
    uint32_t grid;
};
cudaError_t launch_stream_batch(const StreamBatchArgs& a, uint32_t W, cudaStream_t st);
} // namespace ks
