// ksh_host.cpp — host layer over Pod/Node objects (include/ksched_host.h): quantity parsing, packing into
// SoA int64 + label bit columns, and the mirror of check_node_validity / select_node_for_pod / reconcile.
// It only packs and dispatches: every predicate is evaluated by the CUDA core (ks_* in ksched.h).
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../../include/ksched_host.h"

extern "C" void ks__set_error(const char* msg); // defined in ks_api.cu (not part of the public ABI)

namespace {

int fail(int code, const std::string& msg) {
    ks__set_error(msg.c_str());
    return code;
}

typedef unsigned __int128 u128;

// Kubernetes quantity: sign? digits [. digits] (binarySI | decimalSI | decimalExponent)?
// Result = value * 10^unit_shift as an exact integer (unit_shift 3 -> milli-units, 0 -> units).
int parse_quantity(const char* s, int unit_shift, int64_t* out) {
    if (!s) return KS_ERR_PARSE;
    size_t i = 0;
    bool neg = false;
    if (s[i] == '+' || s[i] == '-') neg = s[i++] == '-';
    u128 mant = 0;
    int int_digits = 0, frac_digits = 0;
    const u128 cap = (u128)1 << 96;
    while (s[i] >= '0' && s[i] <= '9') {
        if (mant >= cap) return KS_ERR_RANGE;
        mant = mant * 10 + (unsigned)(s[i++] - '0');
        int_digits++;
    }
    if (s[i] == '.') {
        i++;
        while (s[i] >= '0' && s[i] <= '9') {
            if (mant >= cap) return KS_ERR_RANGE;
            mant = mant * 10 + (unsigned)(s[i++] - '0');
            frac_digits++;
        }
    }
    if (int_digits + frac_digits == 0) return KS_ERR_PARSE;
    int p10 = -frac_digits, p2 = 0;
    const char c = s[i];
    if (c == '\0') {
    } else if (s[i + 1] == 'i' && s[i + 2] == '\0' && std::strchr("KMGTPE", c)) {
        p2 = 10 * (int)(std::strchr("KMGTPE", c) - "KMGTPE" + 1);
        i += 2;
    } else if ((c == 'e' || c == 'E') && s[i + 1] != '\0' &&
               ((s[i + 1] >= '0' && s[i + 1] <= '9') ||
                ((s[i + 1] == '+' || s[i + 1] == '-') && s[i + 2] >= '0' && s[i + 2] <= '9'))) {
        i++;
        bool eneg = false;
        if (s[i] == '+' || s[i] == '-') eneg = s[i++] == '-';
        int e = 0;
        while (s[i] >= '0' && s[i] <= '9') {
            e = e * 10 + (s[i++] - '0');
            if (e > 4000) return KS_ERR_RANGE;
        }
        p10 += eneg ? -e : e;
    } else {
        static const char suf[] = "numkMGTPE";
        static const int exp[] = {-9, -6, -3, 3, 6, 9, 12, 15, 18};
        const char* f = c ? std::strchr(suf, c) : nullptr;
        if (!f) return KS_ERR_PARSE;
        p10 += exp[f - suf];
        i++;
    }
    if (s[i] != '\0') return KS_ERR_PARSE;
    p10 += unit_shift;
    u128 v = mant;
    if (v != 0) {
        if (p2) {
            if (v >= (cap >> p2)) return KS_ERR_RANGE;
            v <<= p2;
        }
        while (p10 > 0) {
            if (v >= cap) return KS_ERR_RANGE;
            v *= 10;
            p10--;
        }
        while (p10 < 0) {
            if (v % 10 != 0) return KS_ERR_INEXACT;
            v /= 10;
            p10++;
        }
    }
    if (v > (u128)INT64_MAX) return KS_ERR_RANGE;
    *out = neg ? -(int64_t)v : (int64_t)v;
    return KS_OK;
}

const char* kv_find(const ks_kv* kv, uint32_t n, const char* key) {
    for (uint32_t i = 0; i < n; i++)
        if (kv[i].key && std::strcmp(kv[i].key, key) == 0) return kv[i].val;
    return nullptr;
}

std::string pair_key(const char* k, const char* v) {
    std::string s(k ? k : "");
    s.push_back('\0');
    s.append(v ? v : "");
    return s;
}

void json_escape(std::string& out, const char* s) {
    for (; s && *s; s++) {
        const unsigned char c = (unsigned char)*s;
        if (c == '"' || c == '\\') {
            out.push_back('\\');
            out.push_back((char)c);
        } else if (c < 0x20) {
            char buf[8];
            std::snprintf(buf, sizeof(buf), "\\u%04x", c);
            out += buf;
        } else {
            out.push_back((char)c);
        }
    }
}

} // namespace

struct ksh_context {
    ks_snapshot* snap = nullptr;
    uint32_t N = 0;
    std::vector<std::string> node_names;
    std::unordered_map<std::string, uint32_t> name2idx;
    std::vector<std::vector<std::string>> node_pairs; // per node: pair keys "k\0v"
    std::unordered_set<std::string> all_pairs;        // pairs carried by at least one node
    std::unordered_map<std::string, uint32_t> dict;   // pair -> bit id, only pairs some selector has named
    uint32_t W = 1;
    std::vector<int64_t> alloc_cpu, alloc_mem;
    std::vector<int32_t> bnode;
    std::vector<int64_t> bcpu, bmem;
    std::vector<std::string> bkey;                    // "namespace/name" of each bound pod ("" = unknown)
    std::unordered_map<std::string, size_t> bkey2pos; // bound pod -> position in the four lists above
    bool dirty = true; // device snapshot must be re-uploaded
};

static std::string pod_key(const ks_pod_obj* pod) {
    std::string k(pod->ns ? pod->ns : "");
    k.push_back('/');
    k.append(pod->name ? pod->name : "");
    return k;
}

static void bound_push(ksh_context* c, const std::string& key, int32_t node, int64_t cpu, int64_t mem) {
    if (!key.empty() && key != "/") c->bkey2pos[key] = c->bnode.size();
    c->bnode.push_back(node);
    c->bcpu.push_back(cpu);
    c->bmem.push_back(mem);
    c->bkey.push_back(key);
}

static void bound_erase(ksh_context* c, size_t pos) { // swap-remove
    const size_t last = c->bnode.size() - 1;
    c->bkey2pos.erase(c->bkey[pos]);
    if (pos != last) {
        c->bnode[pos] = c->bnode[last];
        c->bcpu[pos] = c->bcpu[last];
        c->bmem[pos] = c->bmem[last];
        c->bkey[pos] = c->bkey[last];
        if (!c->bkey[pos].empty() && c->bkey[pos] != "/") c->bkey2pos[c->bkey[pos]] = pos;
    }
    c->bnode.pop_back();
    c->bcpu.pop_back();
    c->bmem.pop_back();
    c->bkey.pop_back();
}

// parse one node object into (allocatable, label pairs); mirrors src/predicates.rs:27-32
static int parse_node(const ks_node_obj& nd, std::string* name, int64_t* ac, int64_t* am, std::vector<std::string>* pairs) {
    *name = nd.name ? nd.name : "";
    *ac = 0;
    *am = 0; // status/allocatable None => PodResources::new() = (0,0)   (predicates.rs:27-28)
    if (nd.has_allocatable) {
        const char* q = kv_find(nd.allocatable, nd.n_allocatable, "cpu");
        if (!q) return fail(KS_ERR_MISSING, "node '" + *name + "': allocatable has no cpu (reference panics, predicates.rs:29)");
        int rc = ksh_parse_cpu_millicores(q, ac);
        if (rc) return rc;
        q = kv_find(nd.allocatable, nd.n_allocatable, "memory");
        if (!q) return fail(KS_ERR_MISSING, "node '" + *name + "': allocatable has no memory (reference panics, predicates.rs:30)");
        rc = ksh_parse_memory_bytes(q, am);
        if (rc) return rc;
        if (*ac > KS_MAX_CPU_MILLI || *ac < -KS_MAX_CPU_MILLI || *am > KS_MAX_MEM_BYTES || *am < -KS_MAX_MEM_BYTES)
            return fail(KS_ERR_RANGE, "node '" + *name + "': allocatable out of range");
    }
    pairs->clear();
    if (nd.has_labels)
        for (uint32_t l = 0; l < nd.n_labels; l++) pairs->push_back(pair_key(nd.labels[l].key, nd.labels[l].val));
    return KS_OK;
}

// pairs carried by at least one node, and dictionary bits restricted to them
static void rebuild_pair_universe(ksh_context* c) {
    c->all_pairs.clear();
    for (const auto& v : c->node_pairs)
        for (const auto& p : v) c->all_pairs.insert(p);
    std::unordered_map<std::string, uint32_t> kept;
    for (auto& kv : c->dict)
        if (c->all_pairs.count(kv.first)) kept.emplace(kv.first, (uint32_t)kept.size());
    c->dict.swap(kept);
    c->W = 1;
    while ((uint64_t)c->W * 64 < (uint64_t)c->dict.size() + 1) c->W <<= 1;
}

static uint32_t words_for_bits(uint32_t real_bits) {
    // +1: the last bit of the last word is the shared "no node carries this pair" bit
    uint32_t w = 1;
    while ((uint64_t)w * 64 < (uint64_t)real_bits + 1) w <<= 1;
    return w;
}

// node label words [N*W] under the current dictionary
static void fill_label_words(const ksh_context* c, uint64_t* lab) {
    std::memset(lab, 0, (size_t)c->N * c->W * sizeof(uint64_t));
    for (uint32_t n = 0; n < c->N; n++)
        for (const std::string& pk : c->node_pairs[n]) {
            auto it = c->dict.find(pk);
            if (it != c->dict.end()) lab[(size_t)n * c->W + (it->second >> 6)] |= 1ull << (it->second & 63);
        }
}

static int upload(ksh_context* c) {
    if (!c->snap) return fail(KS_ERR_NO_DEVICE, "packing-only context (KSH_DEVICE_NONE): no device snapshot, nothing is computed on the host");
    if (!c->dirty) return KS_OK;
    std::vector<uint64_t> lab((size_t)c->N * c->W, 0);
    fill_label_words(c, lab.data());
    int rc = ks_snapshot_set_nodes(c->snap, c->N, c->W, c->alloc_cpu.data(), c->alloc_mem.data(), lab.data());
    if (rc) return rc;
    rc = ks_snapshot_set_bound(c->snap, c->bnode.size(), c->bnode.data(), c->bcpu.data(), c->bmem.data());
    if (rc) return rc;
    c->dirty = false;
    return KS_OK;
}

// register every selector pair of the batch; returns <0 on error, else 0
static int grow_dictionary(ksh_context* c, const ks_pod_obj* pods, uint64_t n) {
    for (uint64_t p = 0; p < n; p++) {
        const ks_pod_obj& pod = pods[p];
        if (!(pod.has_spec && pod.has_node_selector)) continue;
        for (uint32_t i = 0; i < pod.n_selector; i++) {
            std::string pk = pair_key(pod.selector[i].key, pod.selector[i].val);
            if (c->dict.count(pk) || !c->all_pairs.count(pk)) continue; // known, or absent everywhere
            const uint32_t bit = (uint32_t)c->dict.size();
            if (words_for_bits(bit + 1) > KS_MAX_LABEL_WORDS)
                return fail(KS_ERR_RANGE, "more than 511 distinct (key,value) pairs referenced by selectors");
            c->dict.emplace(std::move(pk), bit);
            c->dirty = true;
        }
    }
    const uint32_t w = words_for_bits((uint32_t)c->dict.size());
    if (w != c->W) {
        c->W = w;
        c->dirty = true;
    }
    return KS_OK;
}

extern "C" {

int ksh_parse_cpu_millicores(const char* q, int64_t* out) {
    if (!out) return fail(KS_ERR_INVALID, "out is NULL");
    int rc = parse_quantity(q, 3, out);
    if (rc) return fail(rc, std::string("cannot convert cpu quantity '") + (q ? q : "(null)") + "' to integer millicores");
    return KS_OK;
}

int ksh_parse_memory_bytes(const char* q, int64_t* out) {
    if (!out) return fail(KS_ERR_INVALID, "out is NULL");
    int rc = parse_quantity(q, 0, out);
    if (rc) return fail(rc, std::string("cannot convert memory quantity '") + (q ? q : "(null)") + "' to integer bytes");
    return KS_OK;
}

// src/util.rs:54-75: sum of resources.requests over spec.containers only (no initContainers, no limits)
int ksh_total_pod_resources(const ks_pod_obj* pod, int64_t* cpu, int64_t* mem) {
    if (!pod || !cpu || !mem) return fail(KS_ERR_INVALID, "NULL argument");
    int64_t c = 0, m = 0;
    if (pod->has_spec) {
        for (uint32_t i = 0; i < pod->n_containers; i++) {
            const ks_container_obj& ct = pod->containers[i];
            if (!ct.has_requests) continue;
            if (const char* q = kv_find(ct.requests, ct.n_requests, "cpu")) {
                int64_t v;
                int rc = ksh_parse_cpu_millicores(q, &v); // reference: .expect("invalid pod spec: cpu request")
                if (rc) return rc;
                if (__builtin_add_overflow(c, v, &c)) return fail(KS_ERR_RANGE, "cpu request sum overflows");
            }
            if (const char* q = kv_find(ct.requests, ct.n_requests, "memory")) {
                int64_t v;
                int rc = ksh_parse_memory_bytes(q, &v);
                if (rc) return rc;
                if (__builtin_add_overflow(m, v, &m)) return fail(KS_ERR_RANGE, "memory request sum overflows");
            }
        }
    }
    if (c > KS_MAX_CPU_MILLI || c < -KS_MAX_CPU_MILLI || m > KS_MAX_MEM_BYTES || m < -KS_MAX_MEM_BYTES)
        return fail(KS_ERR_RANGE, "pod requests outside +-2^36 millicores / +-2^55 bytes");
    *cpu = c;
    *mem = m;
    return KS_OK;
}

int ksh_is_pod_bound(const ks_pod_obj* pod) { return pod && pod->has_spec && pod->node_name != nullptr; } // util.rs:38-45

int ksh_context_create(int device, ksh_context** out) {
    if (!out) return fail(KS_ERR_INVALID, "out is NULL");
    *out = nullptr;
    ksh_context* c = new (std::nothrow) ksh_context();
    if (!c) return fail(KS_ERR_NOMEM, "out of host memory");
    if (device != KSH_DEVICE_NONE) { // KSH_DEVICE_NONE: packer only, every call that needs the device fails
        int rc = ks_snapshot_create(device, &c->snap);
        if (rc) {
            delete c;
            return rc;
        }
    }
    *out = c;
    return KS_OK;
}

void ksh_context_destroy(ksh_context* c) {
    if (!c) return;
    ks_snapshot_destroy(c->snap);
    delete c;
}

uint32_t ksh_context_num_nodes(const ksh_context* c) { return c ? c->N : 0; }
uint32_t ksh_context_label_words(const ksh_context* c) { return c ? c->W : 0; }
uint64_t ksh_context_num_bound(const ksh_context* c) { return c ? c->bnode.size() : 0; }

int ksh_context_export_packed(const ksh_context* c, int64_t* alloc_cpu, int64_t* alloc_mem, uint64_t* labels,
                              int32_t* bound_node, int64_t* bound_cpu, int64_t* bound_mem) {
    if (!c) return fail(KS_ERR_INVALID, "NULL argument");
    if (c->N && (!alloc_cpu || !alloc_mem || !labels)) return fail(KS_ERR_INVALID, "NULL node array");
    if (!c->bnode.empty() && (!bound_node || !bound_cpu || !bound_mem)) return fail(KS_ERR_INVALID, "NULL bound array");
    if (c->N) {
        std::memcpy(alloc_cpu, c->alloc_cpu.data(), (size_t)c->N * 8);
        std::memcpy(alloc_mem, c->alloc_mem.data(), (size_t)c->N * 8);
        fill_label_words(c, labels);
    }
    if (!c->bnode.empty()) {
        std::memcpy(bound_node, c->bnode.data(), c->bnode.size() * 4);
        std::memcpy(bound_cpu, c->bcpu.data(), c->bcpu.size() * 8);
        std::memcpy(bound_mem, c->bmem.data(), c->bmem.size() * 8);
    }
    return KS_OK;
}

ks_snapshot* ksh_context_snapshot(ksh_context* c) {
    if (!c || upload(c)) return nullptr;
    return c->snap;
}

int ksh_context_set_nodes(ksh_context* c, const ks_node_obj* nodes, uint32_t n) {
    if (!c || (n && !nodes)) return fail(KS_ERR_INVALID, "NULL argument");
    std::vector<int64_t> ac(n), am(n);
    std::vector<std::vector<std::string>> pairs(n);
    std::vector<std::string> names(n);
    std::unordered_set<std::string> all;
    for (uint32_t i = 0; i < n; i++) {
        const ks_node_obj& nd = nodes[i];
        names[i] = nd.name ? nd.name : "";
        ac[i] = 0;
        am[i] = 0; // status/allocatable None => PodResources::new() = (0,0)   (predicates.rs:27-28)
        if (nd.has_allocatable) {
            const char* q = kv_find(nd.allocatable, nd.n_allocatable, "cpu");
            if (!q) return fail(KS_ERR_MISSING, "node '" + names[i] + "': allocatable has no cpu (reference panics, predicates.rs:29)");
            int rc = ksh_parse_cpu_millicores(q, &ac[i]);
            if (rc) return rc;
            q = kv_find(nd.allocatable, nd.n_allocatable, "memory");
            if (!q) return fail(KS_ERR_MISSING, "node '" + names[i] + "': allocatable has no memory (reference panics, predicates.rs:30)");
            rc = ksh_parse_memory_bytes(q, &am[i]);
            if (rc) return rc;
            if (ac[i] > KS_MAX_CPU_MILLI || ac[i] < -KS_MAX_CPU_MILLI || am[i] > KS_MAX_MEM_BYTES || am[i] < -KS_MAX_MEM_BYTES)
                return fail(KS_ERR_RANGE, "node '" + names[i] + "': allocatable out of range");
        }
        if (nd.has_labels)
            for (uint32_t l = 0; l < nd.n_labels; l++) {
                pairs[i].push_back(pair_key(nd.labels[l].key, nd.labels[l].val));
                all.insert(pairs[i].back());
            }
    }
    c->N = n;
    c->alloc_cpu.swap(ac);
    c->alloc_mem.swap(am);
    c->node_pairs.swap(pairs);
    c->node_names.swap(names);
    c->all_pairs.swap(all);
    c->name2idx.clear();
    for (uint32_t i = 0; i < n; i++) c->name2idx.emplace(c->node_names[i], i); // first wins
    // keep dictionary bits only for pairs that still exist on some node
    std::unordered_map<std::string, uint32_t> kept;
    for (auto& kv : c->dict)
        if (c->all_pairs.count(kv.first)) kept.emplace(kv.first, (uint32_t)kept.size());
    c->dict.swap(kept);
    c->W = words_for_bits((uint32_t)c->dict.size());
    c->bnode.clear();
    c->bcpu.clear();
    c->bmem.clear();
    c->bkey.clear();
    c->bkey2pos.clear();
    c->dirty = true;
    return KS_OK;
}

int ksh_context_upsert_node(ksh_context* c, const ks_node_obj* node, uint32_t* out_idx) {
    if (!c || !node) return fail(KS_ERR_INVALID, "NULL argument");
    std::string name;
    int64_t ac, am;
    std::vector<std::string> pairs;
    int rc = parse_node(*node, &name, &ac, &am, &pairs);
    if (rc) return rc;
    auto it = c->name2idx.find(name);
    uint32_t idx;
    if (it == c->name2idx.end()) {
        idx = c->N++;
        c->node_names.push_back(name);
        c->alloc_cpu.push_back(ac);
        c->alloc_mem.push_back(am);
        c->node_pairs.push_back(pairs);
        c->name2idx.emplace(name, idx);
    } else {
        idx = it->second;
        c->alloc_cpu[idx] = ac;
        c->alloc_mem[idx] = am;
        c->node_pairs[idx] = pairs;
    }
    rebuild_pair_universe(c);
    c->dirty = true;
    if (out_idx) *out_idx = idx;
    return KS_OK;
}

int ksh_context_remove_node(ksh_context* c, const char* name) {
    if (!c || !name) return fail(KS_ERR_INVALID, "NULL argument");
    auto it = c->name2idx.find(name);
    if (it == c->name2idx.end()) return KS_OK;
    const uint32_t idx = it->second;
    c->node_names.erase(c->node_names.begin() + idx);
    c->alloc_cpu.erase(c->alloc_cpu.begin() + idx);
    c->alloc_mem.erase(c->alloc_mem.begin() + idx);
    c->node_pairs.erase(c->node_pairs.begin() + idx);
    c->N--;
    c->name2idx.clear();
    for (uint32_t i = 0; i < c->N; i++) c->name2idx.emplace(c->node_names[i], i);
    // pods bound to the removed node disappear from every later LIST; indices above it move down
    for (size_t p = 0; p < c->bnode.size();) {
        if ((uint32_t)c->bnode[p] == idx) {
            bound_erase(c, p);
        } else {
            if ((uint32_t)c->bnode[p] > idx) c->bnode[p]--;
            p++;
        }
    }
    rebuild_pair_universe(c);
    c->dirty = true;
    return KS_OK;
}

int ksh_context_pod_bound(ksh_context* c, const ks_pod_obj* pod) {
    if (!c || !pod) return fail(KS_ERR_INVALID, "NULL argument");
    if (!ksh_is_pod_bound(pod)) return KS_OK;
    auto it = c->name2idx.find(pod->node_name);
    if (it == c->name2idx.end()) return KS_OK;
    int64_t cpu, mem;
    int rc = ksh_total_pod_resources(pod, &cpu, &mem);
    if (rc) return rc;
    const std::string key = pod_key(pod);
    auto old = c->bkey2pos.find(key);
    if (old != c->bkey2pos.end()) bound_erase(c, old->second); // update of a pod already known
    bound_push(c, key, (int32_t)it->second, cpu, mem);
    c->dirty = true;
    return KS_OK;
}

int ksh_context_pod_deleted(ksh_context* c, const ks_pod_obj* pod) {
    if (!c || !pod) return fail(KS_ERR_INVALID, "NULL argument");
    auto it = c->bkey2pos.find(pod_key(pod));
    if (it == c->bkey2pos.end()) return KS_OK;
    bound_erase(c, it->second);
    c->dirty = true;
    return KS_OK;
}

const char* ksh_context_node_name(const ksh_context* c, uint32_t idx) {
    return (c && idx < c->N) ? c->node_names[idx].c_str() : nullptr;
}

int ksh_context_set_cluster_pods(ksh_context* c, const ks_pod_obj* pods, uint64_t n) {
    if (!c || (n && !pods)) return fail(KS_ERR_INVALID, "NULL argument");
    // validate first, then replace
    std::vector<int64_t> cpus(n), mems(n);
    for (uint64_t p = 0; p < n; p++) {
        if (!ksh_is_pod_bound(&pods[p]) || !c->name2idx.count(pods[p].node_name)) continue;
        int rc = ksh_total_pod_resources(&pods[p], &cpus[p], &mems[p]); // predicates.rs:37
        if (rc) return rc;
    }
    c->bnode.clear();
    c->bcpu.clear();
    c->bmem.clear();
    c->bkey.clear();
    c->bkey2pos.clear();
    for (uint64_t p = 0; p < n; p++) {
        if (!ksh_is_pod_bound(&pods[p])) continue;
        auto it = c->name2idx.find(pods[p].node_name); // field selector spec.nodeName=<node> (predicates.rs:22-25)
        if (it == c->name2idx.end()) continue;
        bound_push(c, pod_key(&pods[p]), (int32_t)it->second, cpus[p], mems[p]);
    }
    c->dirty = true;
    return KS_OK;
}

int ksh_pack_pods(ksh_context* c, const ks_pod_obj* pods, uint64_t n, int64_t* req_cpu, int64_t* req_mem,
                  uint64_t* sel, uint32_t stride) {
    if (!c || (n && (!pods || !req_cpu || !req_mem || !sel))) return fail(KS_ERR_INVALID, "NULL argument");
    int rc = grow_dictionary(c, pods, n);
    if (rc) return rc;
    if (stride < c->W) return fail(KS_ERR_INVALID, "sel_stride_words smaller than the dictionary's word count");
    const uint32_t absent = c->W * 64 - 1;
    for (uint64_t p = 0; p < n; p++) {
        rc = ksh_total_pod_resources(&pods[p], &req_cpu[p], &req_mem[p]);
        if (rc) return rc;
        uint64_t* row = sel + p * stride;
        for (uint32_t w = 0; w < stride; w++) row[w] = 0;
        if (pods[p].has_spec && pods[p].has_node_selector)
            for (uint32_t i = 0; i < pods[p].n_selector; i++) {
                auto it = c->dict.find(pair_key(pods[p].selector[i].key, pods[p].selector[i].val));
                const uint32_t bit = it == c->dict.end() ? absent : it->second;
                row[bit >> 6] |= 1ull << (bit & 63);
            }
    }
    return (int)c->W;
}

static int pack_and_upload(ksh_context* c, const ks_pod_obj* pods, uint64_t n, std::vector<int64_t>& rc_, std::vector<int64_t>& rm_,
                           std::vector<uint64_t>& sel) {
    int rc = grow_dictionary(c, pods, n);
    if (rc) return rc;
    rc_.resize(n);
    rm_.resize(n);
    sel.assign((size_t)n * c->W, 0);
    rc = ksh_pack_pods(c, pods, n, rc_.data(), rm_.data(), sel.data(), c->W);
    if (rc < 0) return rc;
    return upload(c);
}

int ksh_check_node_validity(ksh_context* c, const ks_pod_obj* pod, uint32_t node_idx) {
    if (!c || !pod) return fail(KS_ERR_INVALID, "NULL argument");
    if (node_idx >= c->N) return fail(KS_ERR_INVALID, "node index out of range");
    std::vector<int64_t> rc_, rm_;
    std::vector<uint64_t> sel;
    int rc = pack_and_upload(c, pod, 1, rc_, rm_, sel);
    if (rc) return rc;
    return ks_check_cell(c->snap, rc_[0], rm_[0], sel.data(), node_idx);
}

int ksh_select_nodes(ksh_context* c, const ks_pod_obj* pods, uint64_t n, int policy, int32_t* out_node_idx,
                     int64_t* out_score, uint32_t* out_cnt) {
    if (!c || (n && !pods)) return fail(KS_ERR_INVALID, "NULL argument");
    if (n == 0) return KS_OK;
    std::vector<int64_t> rc_, rm_;
    std::vector<uint64_t> sel;
    int rc = pack_and_upload(c, pods, n, rc_, rm_, sel);
    if (rc) return rc;
    ks_pods kp{n, rc_.data(), rm_.data(), sel.data(), KS_MEM_HOST};
    ks_bindings kb{out_node_idx, out_score, out_cnt, KS_MEM_HOST, nullptr, 0, KS_MEM_HOST, nullptr};
    return ks_select(c->snap, &kp, policy, KS_SELECT_AUTO, &kb, nullptr);
}

int ksh_select_node_for_pod(ksh_context* c, const ks_pod_obj* pods, uint64_t n, uint32_t attempts, uint64_t seed,
                            uint64_t first_pod_index, int32_t* out_node_idx, uint32_t* out_attempts,
                            int32_t* out_draw_node, uint8_t* out_draw_code) {
    if (!c || (n && !pods)) return fail(KS_ERR_INVALID, "NULL argument");
    if (n == 0) return KS_OK;
    std::vector<int64_t> rc_, rm_;
    std::vector<uint64_t> sel;
    int rc = pack_and_upload(c, pods, n, rc_, rm_, sel);
    if (rc) return rc;
    ks_pods kp{n, rc_.data(), rm_.data(), sel.data(), KS_MEM_HOST};
    return ks_select_sampling(c->snap, &kp, attempts, seed, first_pod_index, out_node_idx, out_attempts, out_draw_node,
                              out_draw_code);
}

int ksh_reconcile(ksh_context* c, const ks_pod_obj* pod, int policy, int32_t* node_idx, char* json, size_t cap) {
    if (!c || !pod || !node_idx) return fail(KS_ERR_INVALID, "NULL argument");
    *node_idx = -1;
    if (json && cap) json[0] = '\0';
    if (ksh_is_pod_bound(pod)) return KSH_RECONCILE_OK; // src/main.rs:74-76
    int32_t idx = -1;
    int rc = ksh_select_nodes(c, pod, 1, policy, &idx, nullptr, nullptr); // src/main.rs:78
    if (rc) return rc;
    if (idx < 0) return KSH_RECONCILE_NO_NODE_FOUND; // src/main.rs:116-118
    if (!pod->ns || !pod->name) {
        fail(KS_ERR_INVALID, "pod has no namespace/name (reference: unwrap panic, src/main.rs:80)");
        return KSH_RECONCILE_BINDING_OBJECT_FAILED;
    }
    int64_t cpu, mem;
    rc = ksh_total_pod_resources(pod, &cpu, &mem);
    if (rc) return rc;
    // what the next LIST would report once the binding is accepted (src/predicates.rs:34 after src/main.rs:103)
    rc = ks_snapshot_apply_bind(c->snap, idx, cpu, mem);
    if (rc) return rc;
    bound_push(c, pod_key(pod), idx, cpu, mem);
    *node_idx = idx;
    if (json && cap) { // corev1::Binding{metadata, target: ObjectReference{name}}  (src/main.rs:83-91)
        std::string s = "{\"apiVersion\":\"v1\",\"kind\":\"Binding\",\"metadata\":{\"name\":\"";
        json_escape(s, pod->name);
        s += "\",\"namespace\":\"";
        json_escape(s, pod->ns);
        s += "\"},\"target\":{\"name\":\"";
        json_escape(s, c->node_names[idx].c_str());
        s += "\"}}";
        std::snprintf(json, cap, "%s", s.c_str());
    }
    return KSH_RECONCILE_OK;
}

} // extern "C"
