// pack_bench.cpp — host packer throughput on natively built objects: Pod/Node OBJECTS with
// quantity strings and label maps -> the SoA int64 / label-bitmask arrays the device consumes, through a
// packing-only context (KSH_DEVICE_NONE).  No GPU needed.  Shapes follow BASELINE.json configs[1]/[2]:
//   pack_bench [nodes=10000] [pods=100000] [bound=100000] [device=-1]
// With a device ordinal (a B200) it also times the object-level calls end to end: ksh_select_nodes (objects in, bindings out:
// pack + upload + kernels + copy back) and ksh_reconcile_batch on the first 10000 pods (micro-batch loop + commits + Binding bodies).
// Prints one JSON line.  KSH_THREADS sets the host thread count (default: all cores, at most 32).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "ksched_host.h"

static uint64_t splitmix(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

int main(int argc, char** argv) {
    const uint32_t N = argc > 1 ? (uint32_t)atoi(argv[1]) : 10000;
    const uint64_t P = argc > 2 ? (uint64_t)atoll(argv[2]) : 100000, B = argc > 3 ? (uint64_t)atoll(argv[3]) : 100000;
    const int device = argc > 4 ? atoi(argv[4]) : KSH_DEVICE_NONE;
    uint64_t rng = 0xB2000002;
    std::vector<std::string> strs;
    strs.reserve((P + B) * 10 + (size_t)N * 24);
    auto S = [&](std::string s) {
        strs.push_back(std::move(s));
        return strs.back().c_str();
    };
    std::vector<ks_kv> kvs;
    kvs.reserve((P + B) * 10 + (size_t)N * 12);
    std::vector<ks_container_obj> cts;
    cts.reserve((P + B) * 3);
    std::vector<ks_node_obj> nodes(N);
    std::vector<ks_pod_obj> pods(P), bound(B);
    static const int cores[] = {4, 8, 16, 32, 64, 96};
    for (uint32_t n = 0; n < N; n++) {
        const size_t l0 = kvs.size();
        for (int k = 0; k < 8; k++) kvs.push_back({S("key" + std::to_string(k)), S("v" + std::to_string(splitmix(rng) % 8))});
        const size_t a0 = kvs.size();
        kvs.push_back({"cpu", S(std::to_string(cores[splitmix(rng) % 6]))});
        kvs.push_back({"memory", S(std::to_string((16ll << (splitmix(rng) % 5)) << 30))});
        nodes[n] = {S("node-" + std::to_string(n)), 1, 8, &kvs[l0], 1, 2, &kvs[a0]};
    }
    auto make_pod = [&](ks_pod_obj& pod, const char* prefix, uint64_t i, const char* node_name, bool selectors) {
        const uint32_t nc = 1 + (uint32_t)(splitmix(rng) % 3);
        const size_t c0 = cts.size();
        for (uint32_t c = 0; c < nc; c++) {
            const size_t r0 = kvs.size();
            kvs.push_back({"cpu", S(std::to_string(50 * (1 + splitmix(rng) % 27)) + "m")});
            kvs.push_back({"memory", S(std::to_string(67108864ll * (1 + splitmix(rng) % 85)))});
            cts.push_back({1, 2, &kvs[r0]});
        }
        const uint64_t r = splitmix(rng) % 100;
        const uint32_t ns = !selectors ? 0 : (r < 50 ? 0 : r < 80 ? 1 : r < 95 ? 2 : 3);
        const size_t s0 = kvs.size();
        for (uint32_t k = 0; k < ns; k++)
            kvs.push_back({S("key" + std::to_string((splitmix(rng) % 8))), S("v" + std::to_string(splitmix(rng) % (r == 99 ? 9 : 8)))});
        pod = {"default", S(prefix + std::to_string(i)), 1, node_name, nc, &cts[c0], ns > 0, ns, ns ? &kvs[s0] : nullptr};
    };
    for (uint64_t b = 0; b < B; b++) make_pod(bound[b], "bound-", b, nodes[splitmix(rng) % N].name, false);
    for (uint64_t p = 0; p < P; p++) make_pod(pods[p], "pod-", p, nullptr, true);

    ksh_context* ctx = nullptr;
    if (ksh_context_create(device, &ctx)) {
        fprintf(stderr, "ksh_context_create failed: %s\n", ks_last_error());
        return 1;
    }
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::milli>(b - a).count();
    };
    std::vector<int64_t> rc(P), rm(P);
    std::vector<uint64_t> sel(P * 8);
    double best[4] = {1e30, 1e30, 1e30, 1e30};
    int W = 0;
    for (int it = 0; it < 5; it++) {
        auto t0 = now();
        if (ksh_context_set_nodes(ctx, nodes.data(), N)) return 2;
        auto t1 = now();
        if (ksh_context_set_cluster_pods(ctx, bound.data(), B)) return 3;
        auto t2 = now();
        W = ksh_pack_pods(ctx, pods.data(), P, rc.data(), rm.data(), sel.data(), 8);
        if (W < 0) return 4;
        auto t3 = now();
        uint32_t idx;
        if (ksh_context_upsert_node(ctx, &nodes[N / 2], &idx)) return 5; // one informer event at full scale
        if (ksh_context_pod_deleted(ctx, &bound[B / 2]) || ksh_context_pod_bound(ctx, &bound[B / 2])) return 6;
        auto t4 = now();
        const double v[4] = {ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, t4)};
        for (int k = 0; k < 4; k++) best[k] = v[k] < best[k] ? v[k] : best[k];
    }
    double select_ms = -1, batch_ms = -1;
    unsigned batch_rounds = 0;
    uint64_t batch_bound = 0;
    if (device != KSH_DEVICE_NONE) {
        if (ksh_context_set_nodes(ctx, nodes.data(), N) || ksh_context_set_cluster_pods(ctx, bound.data(), B)) return 7;
        std::vector<int32_t> idx(P);
        std::vector<int64_t> score(P);
        std::vector<uint32_t> cnt(P);
        for (int it = 0; it < 4; it++) { // first call: uploads, index build, module load
            auto t0 = now();
            if (ksh_select_nodes(ctx, pods.data(), P, KS_SCORE_LEFTOVER, idx.data(), score.data(), cnt.data())) {
                fprintf(stderr, "ksh_select_nodes failed: %s\n", ks_last_error());
                return 8;
            }
            const double v = ms(t0, now());
            if (it > 0 && (select_ms < 0 || v < select_ms)) select_ms = v;
        }
        const uint64_t Q = P < 10000 ? P : 10000;
        std::vector<int32_t> st(Q), nd(Q);
        std::vector<int64_t> off(Q);
        std::vector<char> bodies(Q * 256);
        auto t0 = now();
        if (ksh_reconcile_batch(ctx, pods.data(), Q, KS_SCORE_LEFTOVER, st.data(), nd.data(), bodies.data(), bodies.size(), off.data(), &batch_rounds)) {
            fprintf(stderr, "ksh_reconcile_batch failed: %s\n", ks_last_error());
            return 9;
        }
        batch_ms = ms(t0, now());
        for (uint64_t i = 0; i < Q; i++) batch_bound += nd[i] >= 0;
    }
    const char* th = getenv("KSH_THREADS");
    printf("{\"metric\": \"host_packer_objects_per_sec\", \"objects\": \"native\", \"threads\": \"%s\", \"hardware_concurrency\": %u, "
           "\"nodes\": %u, \"nodes_per_s\": %.0f, \"bound_pods\": %llu, \"bound_pods_per_s\": %.0f, \"pods\": %llu, "
           "\"pods_per_s\": %.0f, \"label_words\": %d, \"ms\": {\"set_nodes\": %.3f, \"set_cluster_pods\": %.3f, "
           "\"pack_pods\": %.3f, \"three_events\": %.4f}, \"device\": %d, \"select_nodes_objects_ms\": %.3f, "
           "\"select_nodes_objects_cells_per_s\": %.4g, \"reconcile_batch_10k_ms\": %.3f, \"reconcile_batch_rounds\": %u, "
           "\"reconcile_batch_bound\": %llu}\n",
           th ? th : "default", std::thread::hardware_concurrency(), N, N / best[0] * 1e3, (unsigned long long)B, B / best[1] * 1e3,
           (unsigned long long)P, P / best[2] * 1e3, W, best[0], best[1], best[2], best[3], device, select_ms,
           select_ms > 0 ? (double)P * N / (select_ms * 1e-3) : 0.0, batch_ms, batch_rounds, (unsigned long long)batch_bound);
    ksh_context_destroy(ctx);
    return 0;
}
