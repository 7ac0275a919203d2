// ks_bitpar.h — bit-parallel fast path (ks_bitpar.cu): per-tile threshold tables + label-pair columns
// turn 32 cells into one LOP3; argmax = first feasible node in static priority order.
#pragma once
#include "ks_internal.cuh"

namespace ks {

struct BitparIndex {
    void* blob = nullptr;      // per-tile index blobs, contiguous (see ks_bitpar.cu for the layout)
    size_t blob_cap = 0;
    void* order = nullptr;     // nodes in descending priority order: SoA {free_cpu, free_mem, labels[W], node_idx}
    size_t order_cap = 0;
    uint32_t N = 0, Npad = 0, W = 0, n_tiles = 0;
    bool valid = false;
};

cudaError_t bitpar_build(BitparIndex& ix, const NodeTable& nt, const int64_t* prio, cudaStream_t st);
bool bitpar_profitable(const BitparIndex& ix, uint32_t P);
cudaError_t bitpar_select(BitparIndex& ix, const SelectLaunch& L, const int64_t* prio, cudaEvent_t after_mask);
void bitpar_release(BitparIndex& ix);

} // namespace ks
