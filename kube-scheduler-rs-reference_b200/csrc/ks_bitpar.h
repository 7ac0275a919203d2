// ks_bitpar.h — bit-parallel fast path (ks_bitpar.cu).
//
// Idea: `req <= free` depends only on ORDER.  Nodes are ranked once per snapshot (global position of each
// node in ascending free_cpu / free_mem order); a pod's request becomes a global rank threshold (binary
// search once per pod).  For a tile of 256 nodes the feasible bits of one resource are then a row of a
// prefix table indexed by the tile-local rank of the threshold, and the tile-local rank is
//   base[bucket][tile] + popc(member[bucket][tile] & lowmask)
// (bucket = 64 consecutive global positions).  Label selectors are ANDs of per-(key,value) node columns.
// One thread produces 256 cells (8 mask words) with ~100 instructions; the design target is the HBM write of the
// mask (measured: 53-60 % of it).  KS_SCORE_LEFTOVER is separable (node part - pod part), so argmax-score is the first
// feasible node in a static priority order: a short early-exit scan per pod (k_first_fit).
#pragma once
#include "ks_internal.cuh"

namespace ks {

constexpr int BP_TILE = 256;          // nodes per tile = 8 mask words = one 32-byte row
constexpr int BP_ROWS = BP_TILE + 1;  // prefix-table rows per tile and resource (local rank 0..256)
constexpr int BP_TABLE_BYTES = BP_ROWS * 32;
constexpr int BP_THREADS = 1024;
constexpr int BP_SMEM_MAX = 232448;   // 227 KB opt-in limit per CTA

struct BitparLayout { // byte offsets inside one column-block blob
    uint32_t nt;       // tiles per column block
    uint32_t nb;       // buckets of 64 global positions: (N >> 6) + 1
    uint32_t ncb;      // column blocks
    uint32_t off_baseC, off_membC, off_baseM, off_membM, off_tabC, off_tabM, off_pairs;
    uint32_t pstride;  // bytes between two label-pair columns
    uint32_t blob_bytes;
};

struct BitparIndex {
    // per-snapshot index (device)
    int64_t* sortedC = nullptr;  // [N] free_cpu ascending
    int64_t* sortedM = nullptr;  // [N] free_mem ascending
    uint32_t* gposC = nullptr;   // [N] position of node n in sortedC (ties by node index)
    uint32_t* gposM = nullptr;
    int64_t* ord_prio = nullptr; // nodes in descending priority order (ties by node index), padded to a tile
    int32_t* ord_idx = nullptr;
    int64_t* splC = nullptr;     // every `spl_stride`-th element of sortedC / sortedM (<= 1024 splitters)
    int64_t* splM = nullptr;
    uint8_t* blob = nullptr;     // ncb blobs of lay.blob_bytes: node-index order, staged in shared memory
    uint8_t* blobP = nullptr;    // one blob of layP.blob_bytes: priority order (tile k = priority ranks 256k..),
                                 // read through L1/L2 by k_first_fit_bp
    uint2* pod_ranks = nullptr;  // per-call scratch [P]
    uint32_t* tail_list = nullptr; // per-call scratch [cap_pods + 1]: pods left for k_first_fit_tail, then the count
    uint32_t* pod_bin = nullptr;   // per-call scratch: threshold bucket of each pod, slot inside the bucket
    uint32_t* pod_loc = nullptr;
    uint2* rk_s = nullptr;         // pods in bucket order: thresholds, original pod index, selector words
    uint32_t* pid_s = nullptr;
    unsigned long long* sel_s = nullptr;
    uint2* plist_s = nullptr;      // experimental variant 2 only: <= 4 label-pair column offsets per sorted pod
    uint32_t* hist = nullptr;      // [65536] bucket histogram -> exclusive scan
    uint32_t* rk_hist = nullptr;   // node sample sort scratch: [3][256] bucket counts, splitters, per-node bucket / slot, lists
    int64_t* rk_spl_v = nullptr;
    uint32_t* rk_spl_i = nullptr;
    uint8_t* rk_bkt = nullptr;
    uint32_t* rk_loc = nullptr;
    uint32_t* rk_perm = nullptr;
    size_t cap_nodes = 0, cap_blob = 0, cap_blobP = 0, cap_pods = 0, cap_sel = 0;
    uint32_t N = 0, Nord = 0, W = 0, spl_stride = 1, n_spl = 0;
    BitparLayout lay{}, layP{};
    bool valid = false;
    int sms = 0;
    cudaStream_t aux = nullptr; // k_first_fit_bp runs here, overlapped with k_mask_bitpar
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

cudaError_t bitpar_build(BitparIndex& ix, const NodeTable& nt, const int64_t* prio, cudaStream_t st);
bool bitpar_profitable(const BitparIndex& ix, uint32_t P);
cudaError_t bitpar_prepare(BitparIndex& ix, uint32_t P);
cudaError_t bitpar_select(BitparIndex& ix, SelectLaunch& L, cudaEvent_t before_mask, cudaEvent_t after_mask);
void bitpar_release(BitparIndex& ix);

} // namespace ks
