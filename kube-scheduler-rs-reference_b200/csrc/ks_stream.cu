// ks_stream.cu — K3: in-batch conflict resolution + capacity commit for streaming reconcile (config C5).
// Mirrors what the reference gets from re-LISTing bound pods per cell (/root/reference/src/predicates.rs:34-38):
// a later pod of the batch sees the capacity taken by earlier pods of the same batch.
#include "ks_internal.cuh"
#include "ks_launch.h"

namespace ks {

constexpr int STREAM_MAX = 4096; // claims per launch (one CTA sorts them in shared memory)

// One CTA.  (1) bitonic sort of (node, arrival) keys in shared memory; (2) one thread per node segment walks
// its claimants in arrival order: accept iff the request still fits, then decrement free[] (single writer
// per node: no atomics, deterministic).
__global__ void __launch_bounds__(1024)
    k_stream_resolve(int64_t* __restrict__ free_cpu, int64_t* __restrict__ free_mem, const int32_t* __restrict__ claim_node,
                     const int64_t* __restrict__ req_cpu, const int64_t* __restrict__ req_mem, uint32_t n, uint32_t N,
                     uint8_t* __restrict__ accepted) {
    __shared__ unsigned long long key[STREAM_MAX];
    uint32_t m = 1;
    while (m < n) m <<= 1;
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
        unsigned long long k = ~0ull; // padding and "no claim" sort last
        if (i < n) {
            const int32_t nd = claim_node[i];
            if (nd >= 0 && (uint32_t)nd < N) k = ((unsigned long long)(uint32_t)nd << 32) | i;
            else accepted[i] = 0;
        }
        key[i] = k;
    }
    __syncthreads();
    for (uint32_t size = 2; size <= m; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t i = threadIdx.x; i < m / 2; i += blockDim.x) {
                const uint32_t lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long a = key[lo], b = key[hi];
                if ((a > b) == up) {
                    key[lo] = b;
                    key[hi] = a;
                }
            }
            __syncthreads();
        }
    }
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
        const unsigned long long k = key[i];
        if (k == ~0ull) continue;
        const uint32_t node = (uint32_t)(k >> 32);
        if (i > 0 && (uint32_t)(key[i - 1] >> 32) == node) continue; // not a segment head
        int64_t fc = free_cpu[node], fm = free_mem[node];
        for (uint32_t j = i; j < m && (uint32_t)(key[j] >> 32) == node && key[j] != ~0ull; j++) {
            const uint32_t c = (uint32_t)key[j];
            const int64_t rc = req_cpu[c], rm = req_mem[c];
            const bool ok = rc <= fc && rm <= fm; // predicates.rs:42 against what is left
            accepted[c] = ok;
            if (ok) { // util.rs:31-36
                fc -= rc;
                fm -= rm;
            }
        }
        free_cpu[node] = fc;
        free_mem[node] = fm;
    }
}

cudaError_t launch_stream_resolve(int64_t* free_cpu, int64_t* free_mem, const int32_t* claim_node, const int64_t* req_cpu,
                                  const int64_t* req_mem, uint32_t n, uint32_t N, uint8_t* accepted, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    if (n > STREAM_MAX) return cudaErrorInvalidValue;
    k_stream_resolve<<<1, 1024, 0, st>>>(free_cpu, free_mem, claim_node, req_cpu, req_mem, n, N, accepted);
    g_launches++;
    return cudaGetLastError();
}

uint32_t stream_max_claims() { return STREAM_MAX; }

} // namespace ks
