// ks_bitpar.cu — bit-parallel fused feasibility pass for sm_100a (design notes in ks_bitpar.h).
//
//   per snapshot   k_rank_nodes     global rank of every node in free_cpu / free_mem / priority order
//                  k_build_tile     per-tile prefix tables, bucket base+membership, label-pair columns
//   per call       k_pod_ranks      request -> global rank threshold (2 binary searches per pod)
//                  k_mask_bitpar    persistent; column-block index blob staged in shared memory by TMA bulk
//                                   copies (cp.async.bulk + mbarrier); 1 thread = 1 (pod, 256-node tile):
//                                   2x(base + popc(member & low)) -> 2 table rows -> AND label columns ->
//                                   two 128-bit stores of the mask row segment; counts reduced in smem
//                  k_first_fit      argmax KS_SCORE_LEFTOVER = first feasible node in priority order
// Semantics per cell are exactly predicates.rs:42 / :45-61 (see include/ksched.h); only the evaluation
// order differs, and every output is compared bit-for-bit with the oracle in tests/.
#include "ks_bitpar.h"

#include <algorithm>

namespace ks {

// ------------------------------------------------------------------------------------------------ build
__global__ void __launch_bounds__(256)
    k_rank_nodes(NodeTable nt, const int64_t* __restrict__ prio, int64_t* __restrict__ sortedC,
                 int64_t* __restrict__ sortedM, uint32_t* __restrict__ gposC, uint32_t* __restrict__ gposM,
                 int64_t* __restrict__ ord_fc, int64_t* __restrict__ ord_fm, int64_t* __restrict__ ord_prio,
                 uint64_t* __restrict__ ord_lab, int32_t* __restrict__ ord_idx, uint32_t Nord) {
    __shared__ int64_t s_fc[1024], s_fm[1024], s_pr[1024];
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t N = nt.N;
    const bool real = n < N;
    const int64_t fc = real ? nt.free_cpu[n] : 0, fm = real ? nt.free_mem[n] : 0, pr = real ? prio[n] : 0;
    uint32_t cC = 0, cM = 0, cP = 0;
    for (uint32_t j0 = 0; j0 < N; j0 += 1024) {
        for (uint32_t k = threadIdx.x; k < 1024; k += blockDim.x) {
            const uint32_t j = j0 + k;
            s_fc[k] = j < N ? nt.free_cpu[j] : 0;
            s_fm[k] = j < N ? nt.free_mem[j] : 0;
            s_pr[k] = j < N ? prio[j] : 0;
        }
        __syncthreads();
        const uint32_t lim = min(1024u, N - j0);
        for (uint32_t k = 0; k < lim; k++) {
            const uint32_t j = j0 + k;
            const bool before = j < n;
            const int64_t vc = s_fc[k], vm = s_fm[k], vp = s_pr[k];
            cC += (vc < fc) || (vc == fc && before);
            cM += (vm < fm) || (vm == fm && before);
            cP += (vp > pr) || (vp == pr && before); // descending priority, ties -> lower node index first
        }
        __syncthreads();
    }
    if (real) {
        gposC[n] = cC;
        gposM[n] = cM;
        sortedC[cC] = fc;
        sortedM[cM] = fm;
        ord_fc[cP] = fc;
        ord_fm[cP] = fm;
        ord_prio[cP] = pr;
        ord_idx[cP] = (int32_t)n;
        for (uint32_t w = 0; w < nt.W; w++) ord_lab[(size_t)w * Nord + cP] = nt.labels[(size_t)w * nt.Npad + n];
    } else if (n < Nord) { // padding of the priority order: never feasible
        ord_fc[n] = INT64_MIN;
        ord_fm[n] = INT64_MIN;
        ord_prio[n] = INT64_MIN;
        ord_idx[n] = -1;
        for (uint32_t w = 0; w < nt.W; w++) ord_lab[(size_t)w * Nord + n] = 0;
    }
}

__device__ __forceinline__ uint32_t table_chunk(uint32_t row, uint32_t half) {
    // 16-byte chunk index of (row, half) inside a tile table; the XOR spreads the first halves of
    // consecutive rows over all eight 16-byte bank groups
    return 2u * row + (half ^ ((row >> 2) & 1u));
}

__global__ void __launch_bounds__(288)
    k_build_tile(NodeTable nt, const uint32_t* __restrict__ gposC, const uint32_t* __restrict__ gposM,
                 uint8_t* __restrict__ blob, BitparLayout lay) {
    __shared__ uint32_t s_g[2][BP_TILE];
    __shared__ uint16_t s_lr[2][BP_TILE];
    __shared__ uint8_t s_valid[BP_TILE];
    __shared__ uint64_t s_lab[KS_MAX_LABEL_WORDS][BP_TILE];
    const uint32_t tile_g = blockIdx.x, cb = tile_g / lay.nt, t = tile_g % lay.nt, s = threadIdx.x;
    uint8_t* B = blob + (size_t)cb * lay.blob_bytes;
    if (s < BP_TILE) {
        const uint32_t n = tile_g * BP_TILE + s;
        const bool v = n < nt.N;
        s_valid[s] = v;
        s_g[0][s] = v ? gposC[n] : 0xFFFFFFFFu;
        s_g[1][s] = v ? gposM[n] : 0xFFFFFFFFu;
        for (uint32_t w = 0; w < nt.W; w++) s_lab[w][s] = v ? nt.labels[(size_t)w * nt.Npad + n] : 0ull;
    }
    __syncthreads();
    if (s < BP_TILE) {
        for (int r = 0; r < 2; r++) {
            const uint32_t g = s_g[r][s];
            uint32_t c = 0;
            for (int k = 0; k < BP_TILE; k++) c += s_g[r][k] < g;
            s_lr[r][s] = (uint16_t)c;
        }
    }
    __syncthreads();
    // prefix tables: row r = nodes of the tile whose tile-local rank is >= r  (i.e. free >= threshold)
    if (s < BP_ROWS) {
        for (int r = 0; r < 2; r++) {
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                uint32_t acc = 0;
                for (int b = 0; b < 32; b++) {
                    const int k = j * 32 + b;
                    acc |= (uint32_t)(s_valid[k] && s_lr[r][k] >= s) << b;
                }
                w[j] = acc;
            }
            uint4* tab = reinterpret_cast<uint4*>(B + (r ? lay.off_tabM : lay.off_tabC) + (size_t)t * BP_TABLE_BYTES);
            tab[table_chunk(s, 0)] = make_uint4(w[0], w[1], w[2], w[3]);
            tab[table_chunk(s, 1)] = make_uint4(w[4], w[5], w[6], w[7]);
        }
    }
    // bucket membership + base counts, layout [bucket][tile]
    unsigned long long* membC = reinterpret_cast<unsigned long long*>(B + lay.off_membC);
    unsigned long long* membM = reinterpret_cast<unsigned long long*>(B + lay.off_membM);
    uint16_t* baseC = reinterpret_cast<uint16_t*>(B + lay.off_baseC);
    uint16_t* baseM = reinterpret_cast<uint16_t*>(B + lay.off_baseM);
    for (uint32_t hi = s; hi < lay.nb; hi += blockDim.x) {
        membC[(size_t)hi * lay.nt + t] = 0ull;
        membM[(size_t)hi * lay.nt + t] = 0ull;
    }
    __syncthreads();
    if (s < BP_TILE && s_valid[s]) {
        atomicOr(&membC[(size_t)(s_g[0][s] >> 6) * lay.nt + t], 1ull << (s_g[0][s] & 63));
        atomicOr(&membM[(size_t)(s_g[1][s] >> 6) * lay.nt + t], 1ull << (s_g[1][s] & 63));
    }
    __threadfence();
    __syncthreads();
    const uint32_t warp = s >> 5, lane = s & 31;
    if (warp < 2) {
        const unsigned long long* memb = warp ? membM : membC;
        uint16_t* base = warp ? baseM : baseC;
        uint32_t running = 0;
        for (uint32_t h0 = 0; h0 < lay.nb; h0 += 32) {
            const uint32_t hi = h0 + lane;
            const uint32_t c = hi < lay.nb ? __popcll(__ldcg(&memb[(size_t)hi * lay.nt + t])) : 0;
            uint32_t inc = c;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, inc, off);
                if (lane >= (uint32_t)off) inc += o;
            }
            if (hi < lay.nb) base[(size_t)hi * lay.nt + t] = (uint16_t)(running + inc - c);
            running += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
    // label-pair columns, layout [bit][tile][8 words]
    uint4* pairs = reinterpret_cast<uint4*>(B + lay.off_pairs);
    for (uint32_t bit = s; bit < 64u * nt.W; bit += blockDim.x) {
        const uint32_t w_ = bit >> 6, sh = bit & 63;
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            uint32_t acc = 0;
            for (int b = 0; b < 32; b++) acc |= (uint32_t)((s_lab[w_][j * 32 + b] >> sh) & 1ull) << b;
            w[j] = acc;
        }
        pairs[((size_t)bit * lay.nt + t) * 2 + 0] = make_uint4(w[0], w[1], w[2], w[3]);
        pairs[((size_t)bit * lay.nt + t) * 2 + 1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
}

// ------------------------------------------------------------------------------------------------ per call
__device__ __forceinline__ uint32_t lower_bound_i64(const int64_t* __restrict__ a, uint32_t n, int64_t x) {
    uint32_t lo = 0, len = n; // number of elements < x
    while (len > 0) {
        const uint32_t half = len >> 1;
        if (__ldg(a + lo + half) < x) {
            lo += half + 1;
            len -= half + 1;
        } else {
            len = half;
        }
    }
    return lo;
}

__global__ void k_pod_ranks(PodView pv, const int64_t* __restrict__ sortedC, const int64_t* __restrict__ sortedM,
                            uint32_t N, uint2* __restrict__ rk) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= pv.P) return;
    // nodes at sorted positions >= rank have free >= request  <=>  request <= free (predicates.rs:42)
    rk[p] = make_uint2(lower_bound_i64(sortedC, N, pv.req_cpu[p]), lower_bound_i64(sortedM, N, pv.req_mem[p]));
}

template <int W>
__global__ void __launch_bounds__(BP_THREADS, 1)
    k_mask_bitpar(const uint8_t* __restrict__ blob, BitparLayout lay, PodView pv, const uint2* __restrict__ rk,
                  OutView ov) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    uint32_t* cnt_s = reinterpret_cast<uint32_t*>(smem + lay.blob_bytes);

    const uint32_t tid = threadIdx.x, nt = lay.nt, P = pv.P;
    // static split of the (column block, pod) units over the persistent CTAs
    const uint64_t units = (uint64_t)lay.ncb * P;
    uint64_t u0 = units * blockIdx.x / gridDim.x;
    const uint64_t u1 = units * (blockIdx.x + 1) / gridDim.x;
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t phase = 0;
    const uint32_t div_magic = ((1u << 20) + nt - 1) / nt; // exact i/nt for i < 2^15 (BP_POD_CHUNK*nt <= 2^15)

    const uint16_t* baseC = reinterpret_cast<const uint16_t*>(smem + lay.off_baseC);
    const uint16_t* baseM = reinterpret_cast<const uint16_t*>(smem + lay.off_baseM);
    const unsigned long long* membC = reinterpret_cast<const unsigned long long*>(smem + lay.off_membC);
    const unsigned long long* membM = reinterpret_cast<const unsigned long long*>(smem + lay.off_membM);
    const uint4* pairs = reinterpret_cast<const uint4*>(smem + lay.off_pairs);

    while (u0 < u1) {
        const uint32_t cb = (uint32_t)(u0 / P);
        const uint32_t pa = (uint32_t)(u0 - (uint64_t)cb * P);
        const uint32_t pb = (uint32_t)min((uint64_t)P, u1 - (uint64_t)cb * P);
        u0 = (uint64_t)cb * P + pb;

        // ---- stage this column block's index blob: TMA bulk copies signalled on one mbarrier ----
        __syncthreads(); // all generic-proxy reads of the previous blob are done
        if (tid == 0) {
            fence_proxy_async();
            mbar_arrive_expect_tx(&bar, lay.blob_bytes);
            const uint8_t* src = blob + (size_t)cb * lay.blob_bytes;
            for (uint32_t off = 0; off < lay.blob_bytes; off += 32768u)
                tma_bulk_g2s(smem + off, src + off, min(32768u, lay.blob_bytes - off), &bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1;

        for (uint32_t pc0 = pa; pc0 < pb; pc0 += BP_POD_CHUNK) {
            const uint32_t n_p = min((uint32_t)BP_POD_CHUNK, pb - pc0);
            if (ov.cnt) {
                for (uint32_t i = tid; i < n_p; i += BP_THREADS) cnt_s[i] = 0;
                __syncthreads();
            }
            const uint32_t items = n_p * nt;
            for (uint32_t i = tid; i < items; i += BP_THREADS) {
                const uint32_t pl = (i * div_magic) >> 20;
                const uint32_t t = i - pl * nt;
                const uint32_t p = pc0 + pl;
                const uint2 r = __ldg(rk + p);
                // tile-local rank of each threshold: nodes of the tile at global positions < threshold
                const uint32_t hc = r.x >> 6, hm = r.y >> 6;
                const uint32_t rankC = baseC[hc * nt + t] + __popcll(membC[hc * nt + t] & ((1ull << (r.x & 63)) - 1ull));
                const uint32_t rankM = baseM[hm * nt + t] + __popcll(membM[hm * nt + t] & ((1ull << (r.y & 63)) - 1ull));
                const uint4* tc = reinterpret_cast<const uint4*>(smem + lay.off_tabC + t * BP_TABLE_BYTES);
                const uint4* tm = reinterpret_cast<const uint4*>(smem + lay.off_tabM + t * BP_TABLE_BYTES);
                const uint4 c0 = tc[table_chunk(rankC, 0)], c1 = tc[table_chunk(rankC, 1)];
                const uint4 m0 = tm[table_chunk(rankM, 0)], m1 = tm[table_chunk(rankM, 1)];
                uint4 a = make_uint4(c0.x & m0.x, c0.y & m0.y, c0.z & m0.z, c0.w & m0.w);
                uint4 b = make_uint4(c1.x & m1.x, c1.y & m1.y, c1.z & m1.z, c1.w & m1.w);
#pragma unroll
                for (int w = 0; w < W; w++) {
                    unsigned long long bits = __ldg(pv.sel + (size_t)p * W + w);
                    while (bits) { // AND the node column of every required (key,value) pair (predicates.rs:48-53)
                        const uint32_t bit = w * 64 + __ffsll((long long)bits) - 1;
                        bits &= bits - 1;
                        const uint4 q0 = pairs[(bit * nt + t) * 2], q1 = pairs[(bit * nt + t) * 2 + 1];
                        a.x &= q0.x; a.y &= q0.y; a.z &= q0.z; a.w &= q0.w;
                        b.x &= q1.x; b.y &= q1.y; b.z &= q1.z; b.w &= q1.w;
                    }
                }
                if (ov.cnt) {
                    const uint32_t c = __popc(a.x) + __popc(a.y) + __popc(a.z) + __popc(a.w) + __popc(b.x) +
                                       __popc(b.y) + __popc(b.z) + __popc(b.w);
                    if (c) atomicAdd(&cnt_s[pl], c);
                }
                if (ov.mask) {
                    const uint32_t word = (cb * nt + t) * 8;
                    if (word < ov.mask_valid_words) {
                        uint4* dst = reinterpret_cast<uint4*>(ov.mask + (size_t)p * ov.mask_row_words + word);
                        __stcs(dst, a);
                        __stcs(dst + 1, b);
                    }
                }
            }
            if (ov.cnt) {
                __syncthreads();
                for (uint32_t i = tid; i < n_p; i += BP_THREADS) {
                    const uint32_t c = cnt_s[i];
                    if (lay.ncb == 1) ov.cnt[pc0 + i] = c;
                    else if (c) atomicAdd(&ov.cnt[pc0 + i], c);
                }
                __syncthreads();
            }
        }
    }
}

// argmax of the separable score = first feasible node in descending priority order.  One warp per pod,
// 32 candidates per step, early exit.  Pods whose feasible count is already known to be 0 are skipped.
template <int W>
__global__ void __launch_bounds__(256)
    k_first_fit(const int64_t* __restrict__ ord_fc, const int64_t* __restrict__ ord_fm,
                const int64_t* __restrict__ ord_prio, const uint64_t* __restrict__ ord_lab,
                const int32_t* __restrict__ ord_idx, uint32_t Nord, PodView pv, OutView ov) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= pv.P) return;
    const int64_t rc = __ldg(pv.req_cpu + p), rm = __ldg(pv.req_mem + p);
    uint64_t sel[W];
#pragma unroll
    for (int w = 0; w < W; w++) sel[w] = __ldg(pv.sel + (size_t)p * W + w);
    int32_t best = -1;
    int64_t score = 0;
    if (!(ov.cnt && ov.cnt[p] == 0)) {
        for (uint32_t base = 0; base < Nord; base += 32) {
            const uint32_t i = base + lane;
            bool ok = (rc <= __ldg(ord_fc + i)) & (rm <= __ldg(ord_fm + i));
            uint64_t miss = 0;
#pragma unroll
            for (int w = 0; w < W; w++) miss |= sel[w] & ~__ldg(ord_lab + (size_t)w * Nord + i);
            ok = ok && miss == 0 && __ldg(ord_idx + i) >= 0;
            const uint32_t b = __ballot_sync(0xffffffffu, ok);
            if (b) {
                const uint32_t f = base + __ffs(b) - 1;
                best = __ldg(ord_idx + f);
                score = __ldg(ord_prio + f) - (int64_t)(((uint64_t)rc << 22) + (uint64_t)rm);
                break;
            }
        }
    }
    if (lane == 0) {
        if (ov.node_idx) ov.node_idx[p] = best;
        if (ov.score) ov.score[p] = score;
    }
}

// ------------------------------------------------------------------------------------------------ host side
static uint32_t round16(uint32_t x) { return (x + 15u) & ~15u; }

static bool make_layout(uint32_t N, uint32_t W, BitparLayout* lay) {
    const uint32_t n_tiles = (N + BP_TILE - 1) / BP_TILE;
    const uint32_t nb = (N >> 6) + 1;
    const uint64_t per_tile = (uint64_t)nb * 20 + 2ull * BP_TABLE_BYTES + 64ull * W * 32;
    const uint64_t avail = BP_SMEM_MAX - BP_POD_CHUNK * 4 - 1024;
    uint32_t nt_max = (uint32_t)std::min<uint64_t>(32, avail / per_tile);
    if (n_tiles == 0 || nt_max == 0) return false;
    const uint32_t ncb = (n_tiles + nt_max - 1) / nt_max;
    const uint32_t nt = (n_tiles + ncb - 1) / ncb;
    lay->nt = nt;
    lay->nb = nb;
    lay->ncb = ncb;
    uint32_t off = 0;
    lay->off_baseC = off;
    off += round16(nb * nt * 2);
    lay->off_baseM = off;
    off += round16(nb * nt * 2);
    lay->off_membC = off;
    off += round16(nb * nt * 8);
    lay->off_membM = off;
    off += round16(nb * nt * 8);
    lay->off_tabC = off;
    off += nt * BP_TABLE_BYTES;
    lay->off_tabM = off;
    off += nt * BP_TABLE_BYTES;
    lay->off_pairs = off;
    off += 64 * W * nt * 32;
    lay->blob_bytes = (off + 127u) & ~127u;
    return lay->blob_bytes + BP_POD_CHUNK * 4 <= (uint32_t)BP_SMEM_MAX;
}

template <class T>
static cudaError_t regrow(T*& p, size_t count) {
    if (p) cudaFree(p);
    p = nullptr;
    return cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T));
}

void bitpar_release(BitparIndex& ix) {
    void* ptrs[] = {ix.sortedC, ix.sortedM, ix.gposC, ix.gposM, ix.ord_fc, ix.ord_fm, ix.ord_prio, ix.ord_lab,
                    ix.ord_idx, ix.blob, ix.pod_ranks};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    ix = BitparIndex();
}

cudaError_t bitpar_build(BitparIndex& ix, const NodeTable& nt, const int64_t* prio, cudaStream_t st) {
    ix.valid = false;
    ix.N = nt.N;
    ix.W = nt.W;
    if (nt.N == 0) return cudaSuccess;
    const uint32_t Nord = (nt.N + 31u) & ~31u;
    ix.Nord = Nord;
    cudaError_t e;
    if (Nord > ix.cap_nodes || (size_t)Nord * nt.W > ix.cap_lab) {
        const size_t cap = Nord + Nord / 8 + 32;
        if ((e = regrow(ix.sortedC, cap)) != cudaSuccess) return e;
        if ((e = regrow(ix.sortedM, cap)) != cudaSuccess) return e;
        if ((e = regrow(ix.gposC, cap)) != cudaSuccess) return e;
        if ((e = regrow(ix.gposM, cap)) != cudaSuccess) return e;
        if ((e = regrow(ix.ord_fc, cap)) != cudaSuccess) return e;
        if ((e = regrow(ix.ord_fm, cap)) != cudaSuccess) return e;
        if ((e = regrow(ix.ord_prio, cap)) != cudaSuccess) return e;
        if ((e = regrow(ix.ord_idx, cap)) != cudaSuccess) return e;
        if ((e = regrow(ix.ord_lab, cap * nt.W)) != cudaSuccess) return e;
        ix.cap_nodes = cap;
        ix.cap_lab = cap * nt.W;
    }
    k_rank_nodes<<<(Nord + 255) / 256, 256, 0, st>>>(nt, prio, ix.sortedC, ix.sortedM, ix.gposC, ix.gposM, ix.ord_fc,
                                                     ix.ord_fm, ix.ord_prio, ix.ord_lab, ix.ord_idx, Nord);
    g_launches++;
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    BitparLayout lay{};
    if (!make_layout(nt.N, nt.W, &lay)) return cudaSuccess; // index does not fit: only k_first_fit is usable
    const size_t need = (size_t)lay.ncb * lay.blob_bytes;
    if (need > ix.cap_blob) {
        if ((e = regrow(ix.blob, need + need / 8)) != cudaSuccess) return e;
        ix.cap_blob = need + need / 8;
    }
    if ((e = cudaMemsetAsync(ix.blob, 0, need, st)) != cudaSuccess) return e;
    k_build_tile<<<lay.ncb * lay.nt, 288, 0, st>>>(nt, ix.gposC, ix.gposM, ix.blob, lay);
    g_launches++;
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    ix.lay = lay;
    ix.valid = true;
    return cudaSuccess;
}

bool bitpar_profitable(const BitparIndex& ix, uint32_t P) {
    // below ~16M cells the per-call rank pass and blob staging outweigh the per-cell kernel
    return ix.valid && (uint64_t)P * ix.N >= (1ull << 24);
}

template <int W>
static cudaError_t select_w(BitparIndex& ix, const SelectLaunch& L, cudaEvent_t after_mask) {
    cudaError_t e;
    const uint32_t P = L.pv.P;
    const bool need_mask_pass = L.ov.mask || L.ov.cnt;
    if (need_mask_pass) {
        if (P > ix.cap_pods) {
            const size_t cap = (size_t)P + P / 8 + 64;
            if ((e = regrow(ix.pod_ranks, cap)) != cudaSuccess) return e;
            ix.cap_pods = cap;
        }
        k_pod_ranks<<<(P + 255) / 256, 256, 0, L.stream>>>(L.pv, ix.sortedC, ix.sortedM, ix.N, ix.pod_ranks);
        g_launches++;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        if (L.ov.cnt && ix.lay.ncb > 1)
            if ((e = cudaMemsetAsync(L.ov.cnt, 0, (size_t)P * 4, L.stream)) != cudaSuccess) return e;
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const uint64_t units = (uint64_t)ix.lay.ncb * P;
        const uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)sms, (units + 63) / 64);
        const uint32_t smem = ix.lay.blob_bytes + BP_POD_CHUNK * 4;
        auto kern = k_mask_bitpar<W>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
            return e;
        kern<<<grid, BP_THREADS, smem, L.stream>>>(ix.blob, ix.lay, L.pv, ix.pod_ranks, L.ov);
        g_launches++;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    if (after_mask && need_mask_pass)
        if ((e = cudaEventRecord(after_mask, L.stream)) != cudaSuccess) return e;
    if (L.ov.node_idx || L.ov.score) {
        k_first_fit<W><<<(P + 7) / 8, 256, 0, L.stream>>>(ix.ord_fc, ix.ord_fm, ix.ord_prio, ix.ord_lab, ix.ord_idx,
                                                          ix.Nord, L.pv, L.ov);
        g_launches++;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    if (after_mask && !need_mask_pass)
        if ((e = cudaEventRecord(after_mask, L.stream)) != cudaSuccess) return e;
    return cudaSuccess;
}

cudaError_t bitpar_select(BitparIndex& ix, const SelectLaunch& L, const int64_t*, cudaEvent_t after_mask) {
    if (!ix.valid) return cudaErrorNotSupported;
    switch (ix.W) {
        case 1: return select_w<1>(ix, L, after_mask);
        case 2: return select_w<2>(ix, L, after_mask);
        case 4: return select_w<4>(ix, L, after_mask);
        case 8: return select_w<8>(ix, L, after_mask);
        default: return cudaErrorInvalidValue;
    }
}

} // namespace ks
