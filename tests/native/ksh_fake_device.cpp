// ksh_fake_device.cpp — TEST INFRASTRUCTURE.  A stand-in for the CUDA core of libksched.so, answering the ks_* calls of
// include/ksched.h on the host with the ORACLE's packed flavour (oracle/oracle.h).  tests/test_host_layer.py links it with
// csrc/host/ksh_host.cpp into a scratch "libksched.so" and runs the object-level (-m gpu) host-layer tests against it on
// machines without a GPU: that checks everything the host layer does around the device calls (packing, upload order,
// index <-> name mapping, capacity commit, Binding JSON).  It is never built into, shipped with or loaded by the product.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

#include "ksched.h"
extern "C" {
#include "oracle.h"
}

struct ks_snapshot {
    uint32_t N = 0, W = 1;
    std::vector<int64_t> alloc_cpu, alloc_mem, free_cpu, free_mem;
    std::vector<uint64_t> labels; // node-major [N*W]
};

static thread_local char g_err[512] = "";
static int fail(int code, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}

extern "C" {
void ks__set_error(const char* msg) { snprintf(g_err, sizeof(g_err), "%s", msg ? msg : ""); }
const char* ks_last_error(void) { return g_err; }
int ks_version(void) { return 100; }
int ks_device_count(void) { return 1; }
uint64_t ks_launch_count(void) { return 0; }
uint64_t ks_mask_row_bytes(uint32_t n) { return 32ull * ((n + 255ull) / 256ull); }

int ks_snapshot_create(int device, ks_snapshot** out) {
    if (!out || device != 0) return fail(KS_ERR_INVALID, "fake device: only device 0");
    *out = new ks_snapshot();
    return KS_OK;
}
void ks_snapshot_destroy(ks_snapshot* s) { delete s; }
uint32_t ks_snapshot_num_nodes(const ks_snapshot* s) { return s ? s->N : 0; }
uint32_t ks_snapshot_label_words(const ks_snapshot* s) { return s ? s->W : 0; }

int ks_snapshot_set_nodes(ks_snapshot* s, uint32_t n, uint32_t w, const int64_t* ac, const int64_t* am, const uint64_t* lab) {
    if (!s || (w != 1 && w != 2 && w != 4 && w != 8)) return fail(KS_ERR_INVALID, "bad set_nodes");
    s->N = n;
    s->W = w;
    s->alloc_cpu.assign(ac, ac + n);
    s->alloc_mem.assign(am, am + n);
    s->free_cpu = s->alloc_cpu;
    s->free_mem = s->alloc_mem;
    s->labels.assign(lab, lab + (size_t)n * w);
    return KS_OK;
}
int ks_snapshot_set_bound(ks_snapshot* s, uint64_t b, const int32_t* node, const int64_t* cpu, const int64_t* mem) {
    if (!s) return fail(KS_ERR_INVALID, "NULL");
    for (uint64_t i = 0; i < b; i++)
        if (node[i] < 0 || (uint32_t)node[i] >= s->N) return fail(KS_ERR_INVALID, "bound pod: node index out of range");
    if (s->N == 0) return KS_OK;
    return orc_free_reduce(s->N, s->alloc_cpu.data(), s->alloc_mem.data(), b, node, cpu, mem, s->free_cpu.data(), s->free_mem.data());
}
int ks_snapshot_apply_bind(ks_snapshot* s, int32_t node, int64_t cpu, int64_t mem) {
    if (!s || node < 0 || (uint32_t)node >= s->N) return fail(KS_ERR_INVALID, "node index out of range");
    s->free_cpu[node] -= cpu;
    s->free_mem[node] -= mem;
    return KS_OK;
}
int ks_snapshot_get_free(ks_snapshot* s, int64_t* fc, int64_t* fm) {
    if (!s) return fail(KS_ERR_INVALID, "NULL");
    if (s->N) {
        memcpy(fc, s->free_cpu.data(), (size_t)s->N * 8);
        memcpy(fm, s->free_mem.data(), (size_t)s->N * 8);
    }
    return KS_OK;
}

static int run(ks_snapshot* s, const ks_pods* p, int policy, int32_t* idx, int64_t* score, uint32_t* cnt, uint8_t* mask,
               uint64_t row, uint8_t* codes) {
    if (p->mem_space != KS_MEM_HOST) return fail(KS_ERR_INVALID, "fake device: host buffers only");
    if (s->N == 0) {
        for (uint64_t i = 0; i < p->n; i++) {
            if (idx) idx[i] = -1;
            if (score) score[i] = 0;
            if (cnt) cnt[i] = 0;
        }
        return KS_OK;
    }
    return orc_run_packed(s->N, s->W, s->free_cpu.data(), s->free_mem.data(), s->alloc_cpu.data(), s->alloc_mem.data(),
                          s->labels.data(), p->n, p->req_cpu, p->req_mem, p->sel, policy, idx, score, cnt, mask, row, codes, 1);
}

int ks_check_cells(ks_snapshot* s, const ks_pods* pods, uint8_t* out_codes) {
    if (!s || !pods || !out_codes) return fail(KS_ERR_INVALID, "NULL");
    if (pods->n == 0 || s->N == 0) return KS_OK;
    return run(s, pods, 0, nullptr, nullptr, nullptr, nullptr, 0, out_codes);
}
int ks_check_cell(ks_snapshot* s, int64_t rc, int64_t rm, const uint64_t* sel, uint32_t node) {
    if (!s || !sel || node >= s->N) return fail(KS_ERR_INVALID, "node index out of range");
    ks_pods p{1, &rc, &rm, sel, KS_MEM_HOST};
    std::vector<uint8_t> codes(s->N);
    const int e = run(s, &p, 0, nullptr, nullptr, nullptr, nullptr, 0, codes.data());
    return e ? e : (int)codes[node];
}
int ks_select(ks_snapshot* s, const ks_pods* pods, int policy, uint32_t, ks_bindings* out, void*) {
    if (!s || !pods || !out) return fail(KS_ERR_INVALID, "NULL");
    if (out->mem_space != KS_MEM_HOST || (out->mask && out->mask_space != KS_MEM_HOST)) return fail(KS_ERR_INVALID, "host only");
    if (pods->n == 0) return KS_OK;
    return run(s, pods, policy, out->node_idx, out->score, out->feasible_cnt, out->mask, out->mask_row_bytes, nullptr);
}
int ks_last_timings(ks_snapshot*, float* ms) {
    if (ms) ms[0] = ms[1] = ms[2] = 0.f;
    return KS_OK;
}
const char* ks_last_path(const ks_snapshot*) { return "fake"; }
int ks_last_trace(ks_snapshot*, uint64_t*) { return KS_ERR_INVALID; }

int ks_select_sampling(ks_snapshot* s, const ks_pods* pods, uint32_t attempts, uint64_t seed, uint64_t first, int32_t* idx,
                       uint32_t* used, int32_t* dn, uint8_t* dc) {
    if (!s || !pods || (pods->n && !idx)) return fail(KS_ERR_INVALID, "NULL");
    std::vector<uint64_t> st(pods->n);
    for (uint64_t p = 0; p < pods->n; p++) st[p] = KS_SAMPLING_STREAM(seed, first + p);
    return orc_select_sampling_packed(s->N, s->W, s->free_cpu.data(), s->free_mem.data(), s->labels.data(), pods->n, pods->req_cpu,
                                      pods->req_mem, pods->sel, attempts, st.data(), idx, used, dn, dc);
}
int ks_snapshot_commit_claims(ks_snapshot* s, uint64_t n, const int32_t* node, const int64_t* rc, const int64_t* rm, uint8_t* acc) {
    if (!s) return fail(KS_ERR_INVALID, "NULL");
    return orc_commit_claims(s->N, s->free_cpu.data(), s->free_mem.data(), n, node, rc, rm, acc);
}
int ks_stream_bind(ks_snapshot* s, const ks_pods* pods, int policy, int32_t* idx, int64_t* score, uint32_t* rounds) {
    if (!s || !pods) return fail(KS_ERR_INVALID, "NULL");
    if (pods->n == 0) return KS_OK;
    if (s->N == 0) {
        for (uint64_t i = 0; i < pods->n; i++) idx[i] = -1;
        if (rounds) *rounds = 0;
        return KS_OK;
    }
    return orc_stream_bind_packed(s->N, s->W, s->free_cpu.data(), s->free_mem.data(), s->alloc_cpu.data(), s->alloc_mem.data(),
                                  s->labels.data(), pods->n, pods->req_cpu, pods->req_mem, pods->sel, policy, idx, score, rounds);
}
// multi-GPU exchange / IPC entry points: nothing to fake on a CPU box (the ctypes binding only needs the symbols)
int ks_stream_open(ks_snapshot*, int, uint32_t, ks_stream**) { return fail(KS_ERR_NO_DEVICE, "fake device"); }
int ks_stream_submit(ks_stream*, uint64_t, const int64_t*, const int64_t*, const uint64_t*, const uint64_t*) { return KS_ERR_NO_DEVICE; }
int ks_stream_poll(ks_stream*, uint64_t, uint64_t*, int32_t*, int64_t*, uint64_t*) { return KS_ERR_NO_DEVICE; }
int ks_stream_flush(ks_stream*) { return KS_ERR_NO_DEVICE; }
int ks_stream_stats(ks_stream*, uint64_t*, uint64_t*, uint64_t*) { return KS_ERR_NO_DEVICE; }
void ks_stream_close(ks_stream*) {}
uint64_t ks_mask_row_bytes_aligned(uint32_t n) { return 256ull * ((n + 2047ull) / 2048ull); }
int ks_exchange_check(ks_snapshot*) { return KS_OK; }
int ks_ipc_alloc(int, uint64_t, void**, uint8_t*) { return fail(KS_ERR_NO_DEVICE, "fake device"); }
int ks_ipc_open(int, const uint8_t*, void**) { return fail(KS_ERR_NO_DEVICE, "fake device"); }
int ks_ipc_close(int, void*) { return KS_OK; }
int ks_ipc_free(int, void*) { return KS_OK; }
int ks_measure_write_bandwidth(int, void*, uint64_t, int, double*) { return fail(KS_ERR_NO_DEVICE, "fake device"); }
int ks_device_read(int, const void*, void*, uint64_t) { return fail(KS_ERR_NO_DEVICE, "fake device"); }
}
