// write_bw.cu — what a WRITE-ONLY kernel can reach on this B200 (the mask kernel writes 6.25 GB and reads ~0.1 GB).
// MEASURED_PEAKS.json's hbm_gbs is a copy (read + write bytes); this measures the store side alone, several ways.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/build/write_bw scripts/write_bw.cu
//   scripts/build/write_bw [GiB]        -> one JSON line
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            fprintf(stderr, "%s failed: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__); \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

__global__ void k_st128(uint4* __restrict__ dst, size_t n16, uint32_t v) {
    const uint4 x = make_uint4(v, v + 1, v + 2, v + 3);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = x;
}

template <int HINT> // 0 none, 1 L2 evict_first, 2 .cs (streaming)
__global__ void k_st256(uint8_t* __restrict__ dst, size_t n32, uint32_t v) {
    uint64_t pol = 0;
    if (HINT == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n32; i += (size_t)gridDim.x * blockDim.x) {
        uint8_t* p = dst + i * 32;
        if (HINT == 1)
            asm volatile("st.global.L2::cache_hint.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1}, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
        else
            asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
    }
}

__global__ void k_st128_cs(uint4* __restrict__ dst, size_t n16, uint32_t v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
        asm volatile("st.global.cs.v4.b32 [%0], {%1,%1,%1,%1};" ::"l"(dst + i), "r"(v) : "memory");
}

// the mask kernel's pattern: 8 lanes write 256 contiguous bytes of one row; rows (pods) in a scrambled order; the
// column block (256-byte column of every row) is the slow dimension
__global__ void k_rows_pattern(uint8_t* __restrict__ dst, uint32_t n_rows, uint32_t row_bytes, uint32_t v) {
    const uint32_t lane8 = threadIdx.x & 7, sub = (threadIdx.x >> 3) & 3, warp = threadIdx.x >> 5;
    const uint32_t n_cb = row_bytes / 256, n_groups = n_rows / 4;
    const uint64_t total = (uint64_t)n_cb * n_groups;
    const uint64_t per = (total + gridDim.x - 1) / gridDim.x;
    const uint64_t f0 = per * blockIdx.x, f1 = f0 + per < total ? f0 + per : total;
    for (uint64_t f = f0 + warp; f < f1; f += blockDim.x / 32) {
        const uint32_t cb = (uint32_t)(f / n_groups), g = (uint32_t)(f % n_groups);
        const uint32_t row = (uint32_t)(((uint64_t)(g * 4 + sub) * 2654435761ull) % n_rows); // scrambled pod order
        uint8_t* p = dst + (size_t)row * row_bytes + (size_t)cb * 256 + lane8 * 32;
        asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
    }
}

// generalised row pattern: `chunk` contiguous bytes per (row, column block) written by chunk/32 lanes.
//   mode 0: column-block-major (a CTA stays in one column block and walks the rows: the mask kernel's order)
//   mode 1: row-major "lockstep": CTA i owns column block i % n_cb and every (gridDim/n_cb)-th row group; all CTAs walk the
//           rows in the same order, so the pieces of one row are written at about the same time by n_cb different CTAs
//   scramble: rows visited in a pseudo-random order (sorted pods) or in index order
__global__ void k_rows_general(uint8_t* __restrict__ dst, uint32_t n_rows, uint32_t row_bytes, uint32_t chunk, int mode, int scramble,
                               uint32_t v) {
    const uint32_t lpc = chunk / 32;              // lanes per chunk
    const uint32_t cpw = 32 / lpc;                // chunks (rows) per warp instruction
    const uint32_t l = threadIdx.x & 31, sub = l / lpc, lane = l % lpc, warp = threadIdx.x >> 5, wpc = blockDim.x / 32;
    const uint32_t n_cb = row_bytes / chunk, n_groups = n_rows / cpw;
    if (mode == 0) {
        const uint64_t total = (uint64_t)n_cb * n_groups, per = (total + gridDim.x - 1) / gridDim.x;
        const uint64_t f0 = per * blockIdx.x, f1 = f0 + per < total ? f0 + per : total;
        for (uint64_t f = f0 + warp; f < f1; f += wpc) {
            const uint32_t cb = (uint32_t)(f / n_groups), g = (uint32_t)(f % n_groups);
            uint32_t row = g * cpw + sub;
            if (scramble) row = (uint32_t)(((uint64_t)row * 2654435761ull) % n_rows);
            asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(dst + (size_t)row * row_bytes + (size_t)cb * chunk + lane * 32), "r"(v)
                         : "memory");
        }
    } else {
        const uint32_t per_cb = gridDim.x / n_cb; // CTAs per column block; the rest of the grid idles
        if (blockIdx.x >= per_cb * n_cb) return;
        const uint32_t cb = blockIdx.x % n_cb, j = blockIdx.x / n_cb;
        for (uint32_t g = j * wpc + warp; g < n_groups; g += per_cb * wpc) {
            uint32_t row = g * cpw + sub;
            if (scramble) row = (uint32_t)(((uint64_t)row * 2654435761ull) % n_rows);
            asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(dst + (size_t)row * row_bytes + (size_t)cb * chunk + lane * 32), "r"(v)
                         : "memory");
        }
    }
}

// TMA bulk stores: shared -> global through the copy engine of the SM instead of the LSU
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int CHUNK>
__global__ void __launch_bounds__(256) k_tma_store(uint8_t* __restrict__ dst, size_t bytes, uint32_t v) {
    extern __shared__ __align__(128) uint8_t sm[];
    for (uint32_t i = threadIdx.x; i < CHUNK / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = v + i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const size_t n_chunks = bytes / CHUNK;
        uint32_t in_flight = 0;
        for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + c * CHUNK), "r"(smem_u32(sm)), "r"(CHUNK)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (++in_flight >= 8) asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

__global__ void k_copy128(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n16) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
__global__ void k_read128(const uint4* __restrict__ src, size_t n16, uint32_t* out) {
    uint32_t acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 x = src[i];
        acc ^= x.x ^ x.y ^ x.z ^ x.w;
    }
    if (acc == 0x12345) *out = acc;
}

int main(int argc, char** argv) {
    const double gib = argc > 1 ? atof(argv[1]) : 5.84; // ~ the C3 mask (1M rows x 6272 B)
    const uint32_t row_bytes = 6272;
    const uint32_t n_rows = (uint32_t)(gib * (1ull << 30) / row_bytes) / 4 * 4;
    const size_t bytes = (size_t)n_rows * row_bytes;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    uint8_t *a, *b;
    uint32_t* sink;
    CK(cudaMalloc(&a, bytes));
    CK(cudaMalloc(&b, bytes));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(b, 1, bytes));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    std::vector<std::pair<std::string, double>> res;
    auto timeit = [&](const char* name, double moved_bytes, auto launch) {
        float best = 1e30f;
        for (int it = 0; it < 6; it++) {
            CK(cudaEventRecord(e0));
            launch();
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (it > 0) best = std::min(best, ms);
        }
        res.push_back({name, moved_bytes / (best * 1e-3) / 1e9});
        fprintf(stderr, "%s %.1f GB/s\n", name, res.back().second);
    };
    const size_t n16 = bytes / 16, n32 = bytes / 32;
    timeit("cudaMemset", (double)bytes, [&] { CK(cudaMemsetAsync(a, 0x5a, bytes)); });
    timeit("st128_grid_sms_x8", (double)bytes, [&] { k_st128<<<sms * 8, 256>>>((uint4*)a, n16, 7); });
    timeit("st128_grid_sms_x32", (double)bytes, [&] { k_st128<<<sms * 32, 256>>>((uint4*)a, n16, 7); });
    timeit("st128_cs", (double)bytes, [&] { k_st128_cs<<<sms * 8, 256>>>((uint4*)a, n16, 7); });
    timeit("st256", (double)bytes, [&] { k_st256<0><<<sms * 8, 256>>>(a, n32, 7); });
    timeit("st256_evict_first", (double)bytes, [&] { k_st256<1><<<sms * 8, 256>>>(a, n32, 7); });
    timeit("st256_persistent_1024thr", (double)bytes, [&] { k_st256<0><<<sms, 1024>>>(a, n32, 7); });
    timeit("rows_pattern_256B_chunks", (double)bytes, [&] { k_rows_pattern<<<sms, 1024>>>(a, n_rows, row_bytes, 7); });
    {
        const uint32_t rb = 6144; // row pitch divisible by every chunk size below
        const uint32_t nr = (uint32_t)(bytes / rb) / 32 * 32;
        static char names[32][64];
        int ni = 0;
        for (uint32_t chunk : {128u, 256u, 512u, 1024u}) // chunk / 32 lanes <= one warp
            for (int mode = 0; mode < 2; mode++)
                for (int scr = 0; scr < 2; scr++) {
                    snprintf(names[ni], 64, "rows_chunk%u_%s_%s", chunk, mode ? "rowmajor_lockstep" : "cbmajor", scr ? "scrambled" : "inorder");
                    timeit(names[ni++], (double)nr * rb, [&] { k_rows_general<<<sms, 1024>>>(a, nr, rb, chunk, mode, scr, 7); });
                }
    }
    CK(cudaFuncSetAttribute(k_tma_store<32768>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    timeit("tma_bulk_store_32KB", (double)(bytes / 32768 * 32768), [&] { k_tma_store<32768><<<sms * 2, 256, 32768>>>(a, bytes, 7); });
    timeit("tma_bulk_store_2KB", (double)(bytes / 2048 * 2048), [&] { k_tma_store<2048><<<sms * 8, 256, 2048>>>(a, bytes, 7); });
    timeit("copy128_read_plus_write", 2.0 * bytes, [&] { k_copy128<<<sms * 8, 256>>>((uint4*)a, (const uint4*)b, n16); });
    timeit("read128", (double)bytes, [&] { k_read128<<<sms * 8, 256>>>((const uint4*)b, n16, sink); });
    printf("{\"bytes\": %zu, \"sms\": %d, \"unit\": \"GB/s\"", bytes, sms);
    for (auto& r : res) printf(", \"%s\": %.1f", r.first.c_str(), r.second);
    printf("}\n");
    return 0;
}
