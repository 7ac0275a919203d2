// Four threads share one packing-only context (node events, pod events, packs, reads) — the per-context lock must make
// that legal (Arc<Context> shared by concurrent reconciles).  Run under ThreadSanitizer.  Test infrastructure.
#include <atomic>
#include <cstdio>
#include <string>
#include <thread>
#include <vector>
#include "ksched_host.h"

int main() {
    ksh_context* ctx;
    if (ksh_context_create(KSH_DEVICE_NONE, &ctx)) return 1;
    ks_kv alloc[2] = {{"cpu", "64"}, {"memory", "274877906944"}};
    std::atomic<int> failures{0};
    auto worker = [&](int t) {
        uint64_t rng = 1000 + t;
        auto rnd = [&] { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
        for (int step = 0; step < 20000; step++) {
            const int ev = rnd() % 6;
            std::string node = "n" + std::to_string(rnd() % 50);
            if (ev == 0) {
                std::string v = "v" + std::to_string(rnd() % 200);
                ks_kv lab[1] = {{"rev", v.c_str()}};
                ks_node_obj nd = {node.c_str(), 1, 1, lab, 1, 2, alloc};
                uint32_t idx;
                if (ksh_context_upsert_node(ctx, &nd, &idx)) failures++;
            } else if (ev == 1) {
                if (ksh_context_remove_node(ctx, node.c_str())) failures++;
            } else if (ev == 2 || ev == 3) {
                std::string pn = "p" + std::to_string(rnd() % 500);
                ks_kv req[2] = {{"cpu", "10m"}, {"memory", "1048576"}};
                ks_container_obj ct = {1, 2, req};
                ks_pod_obj pod = {"ns", pn.c_str(), 1, node.c_str(), 1, &ct, 0, 0, nullptr};
                if (ev == 2 ? ksh_context_pod_bound(ctx, &pod) : ksh_context_pod_deleted(ctx, &pod)) failures++;
            } else if (ev == 4) {
                std::string v = "v" + std::to_string(rnd() % 200);
                ks_kv sel[1] = {{"rev", v.c_str()}};
                ks_pod_obj pod = {"ns", "q", 1, nullptr, 0, nullptr, 1, 1, sel};
                int64_t c, m;
                uint64_t s[8];
                if (ksh_pack_pods(ctx, &pod, 1, &c, &m, s, 8) < 0) failures++;
            } else {
                const uint32_t n = ksh_context_num_nodes(ctx);
                const uint64_t b = ksh_context_num_bound(ctx);
                std::vector<int64_t> ac(n + 64), am(n + 64), bc(b + 4096), bm(b + 4096);
                std::vector<uint64_t> lab((size_t)(n + 64) * 8);
                std::vector<int32_t> bn(b + 4096);
                (void)ksh_context_label_words(ctx); // sizes may move under us; only exercising the lock here
                (void)ac; (void)am; (void)bc; (void)bm; (void)lab; (void)bn;
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < 4; t++) th.emplace_back(worker, t);
    for (auto& x : th) x.join();
    printf("concurrent ok: %u nodes, %llu bound pods, failures %d\n", ksh_context_num_nodes(ctx),
           (unsigned long long)ksh_context_num_bound(ctx), failures.load());
    ksh_context_destroy(ctx);
    return failures.load() ? 2 : 0;
}
