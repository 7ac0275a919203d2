// Random informer events (node upsert/remove, pod bound/re-bound/deleted, selector packs) on a packing-only context,
// checked against a std::map model; run under ASan+UBSan by tests/test_host_layer.py.  Test infrastructure.
#include <cstdio>
#include <string>
#include <vector>
#include <map>
#include <cstdlib>
#include "ksched_host.h"
static uint64_t rng = 12345;
static uint64_t rnd() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; }
int main() {
    ksh_context* ctx; if (ksh_context_create(KSH_DEVICE_NONE, &ctx)) return 1;
    ks_kv alloc[2] = {{"cpu", "64"}, {"memory", "274877906944"}};
    std::vector<std::string> names; std::map<std::string, std::string> podnode; // model: pod -> node
    std::vector<std::string> live_nodes;
    for (int step = 0; step < 300000; step++) {
        int ev = rnd() % 100;
        if (ev < 15 || live_nodes.size() < 4) {
            std::string name = "n" + std::to_string(rnd() % 400);
            std::string v = "v" + std::to_string(rnd() % 5000);
            ks_kv lab[2] = {{"rev", v.c_str()}, {"zone", "a"}};
            ks_node_obj nd = {name.c_str(), 1, 2, lab, 1, 2, alloc};
            uint32_t idx; if (ksh_context_upsert_node(ctx, &nd, &idx)) return 2;
            bool known = false; for (auto& s : live_nodes) known |= s == name;
            if (!known) live_nodes.push_back(name);
        } else if (ev < 22) {
            size_t i = rnd() % live_nodes.size();
            std::string name = live_nodes[i];
            if (ksh_context_remove_node(ctx, name.c_str())) return 3;
            live_nodes.erase(live_nodes.begin() + i);
            for (auto it = podnode.begin(); it != podnode.end();) it = it->second == name ? podnode.erase(it) : std::next(it);
        } else if (ev < 70) {
            std::string pn = "pod-" + std::to_string(rnd() % 3000) + std::string(rnd() % 30, 'y');
            std::string node = live_nodes[rnd() % live_nodes.size()];
            ks_kv req[2] = {{"cpu", "10m"}, {"memory", "1048576"}};
            ks_container_obj ct = {1, 2, req};
            ks_pod_obj pod = {"ns", pn.c_str(), 1, node.c_str(), 1, &ct, 0, 0, nullptr};
            if (ksh_context_pod_bound(ctx, &pod)) return 4;
            podnode[pn] = node;
        } else if (!podnode.empty()) {
            auto it = podnode.begin(); std::advance(it, rnd() % podnode.size());
            ks_pod_obj pod = {"ns", it->first.c_str(), 1, nullptr, 0, nullptr, 0, 0, nullptr};
            if (ksh_context_pod_deleted(ctx, &pod)) return 5;
            podnode.erase(it);
        }
        if (step % 1000 == 0) {
            std::string v = "v" + std::to_string(rnd() % 5000);
            ks_kv sel[1] = {{"rev", v.c_str()}};
            ks_pod_obj pod = {"ns", "q", 1, nullptr, 0, nullptr, 1, 1, sel};
            int64_t c, m; uint64_t s[8];
            if (ksh_pack_pods(ctx, &pod, 1, &c, &m, s, 8) < 0) return 6;
        }
        if (ksh_context_num_nodes(ctx) != live_nodes.size() || ksh_context_num_bound(ctx) != podnode.size()) {
            printf("MISMATCH at step %d: nodes %u vs %zu, bound %llu vs %zu\n", step, ksh_context_num_nodes(ctx), live_nodes.size(),
                   (unsigned long long)ksh_context_num_bound(ctx), podnode.size());
            return 7;
        }
    }
    for (uint32_t i = 0; i < ksh_context_num_nodes(ctx); i++)
        if (live_nodes[i] != ksh_context_node_name(ctx, i)) { printf("order mismatch at %u\n", i); return 8; }
    printf("churn ok: %u nodes, %llu bound pods, W=%u\n", ksh_context_num_nodes(ctx), (unsigned long long)ksh_context_num_bound(ctx), ksh_context_label_words(ctx));
    ksh_context_destroy(ctx);
    return 0;
}
