// ksh_host.cpp — host layer over Pod/Node objects (include/ksched_host.h): quantity parsing, packing into
// SoA int64 + label bit columns, and the mirror of check_node_validity / select_node_for_pod / reconcile.
// It only packs and dispatches: every predicate is evaluated by the CUDA core (ks_* in ksched.h).
// Built for cluster scale: strings are interned once (no per-lookup allocation), label pairs and node names are
// integers afterwards, node/pod events cost O(labels of the object), and the bulk calls (LIST results, pod
// batches) are parsed by all host threads with results identical to the single-thread order.
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <vector>

#include "../../../include/ksched_host.h"

extern "C" void ks__set_error(const char* msg); // defined in ks_api.cu (not part of the public ABI)

namespace {

int fail(int code, const std::string& msg) {
    ks__set_error(msg.c_str());
    return code;
}

// Nothing may throw across the C ABI (ksched_host.h): every extern "C" entry point is a function-try-block ending here
#define KSH_CATCH                                                                                   \
    catch (const std::bad_alloc&) { return fail(KS_ERR_NOMEM, "out of host memory in the host layer"); } \
    catch (...) { return fail(KS_ERR_INVALID, "unexpected C++ exception in the host layer"); }

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
    { // the first 18 digits fit 64-bit arithmetic: every realistic quantity takes only this loop
        uint64_t m64 = 0;
        while (int_digits < 18 && s[i] >= '0' && s[i] <= '9') {
            m64 = m64 * 10 + (unsigned)(s[i++] - '0');
            int_digits++;
        }
        mant = m64;
    }
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

// ---- string interning: (a, sep, b) byte strings -> dense ids in insertion order; lookups allocate nothing ----
static inline uint64_t hash_bytes(uint64_t h, const char* s, size_t n) {
    for (size_t i = 0; i < n; i++) h = (h ^ (unsigned char)s[i]) * 0x100000001B3ull; // FNV-1a
    return h;
}
static inline uint64_t hash2(const char* a, size_t la, char sep, const char* b, size_t lb) {
    uint64_t h = hash_bytes(0xCBF29CE484222325ull, a, la);
    h = (h ^ (unsigned char)sep) * 0x100000001B3ull;
    h = hash_bytes(h, b, lb);
    h ^= h >> 32; // the table index uses the low bits
    return h;
}

struct Interner {
    struct Ent {
        uint64_t h;
        uint64_t off;
        uint32_t len;
    };
    std::vector<char> arena; // every string stored NUL-terminated
    std::vector<Ent> ents;
    std::vector<uint32_t> slots; // open addressing, linear probing; EMPTY = free
    static constexpr uint32_t EMPTY = 0xFFFFFFFFu;

    void clear() {
        arena.clear();
        ents.clear();
        slots.clear();
    }
    size_t size() const { return ents.size(); }
    const char* str(uint32_t id) const { return arena.data() + ents[id].off; }
    bool equal(const Ent& e, uint64_t h, const char* a, size_t la, char sep, const char* b, size_t lb) const {
        if (e.h != h || e.len != la + 1 + lb) return false;
        const char* s = arena.data() + e.off;
        return std::memcmp(s, a, la) == 0 && s[la] == sep && std::memcmp(s + la + 1, b, lb) == 0;
    }
    int64_t find_h(uint64_t h, const char* a, size_t la, char sep, const char* b, size_t lb) const {
        if (slots.empty()) return -1;
        const size_t mask = slots.size() - 1;
        for (size_t i = h & mask;; i = (i + 1) & mask) {
            const uint32_t id = slots[i];
            if (id == EMPTY) return -1;
            if (equal(ents[id], h, a, la, sep, b, lb)) return id;
        }
    }
    int64_t find(const char* a, char sep, const char* b) const {
        const size_t la = std::strlen(a), lb = std::strlen(b);
        return find_h(hash2(a, la, sep, b, lb), a, la, sep, b, lb);
    }
    void grow() {
        const size_t cap = slots.empty() ? 64 : slots.size() * 2;
        slots.assign(cap, EMPTY);
        for (uint32_t id = 0; id < ents.size(); id++) {
            size_t i = ents[id].h & (cap - 1);
            while (slots[i] != EMPTY) i = (i + 1) & (cap - 1);
            slots[i] = id;
        }
    }
    uint32_t intern_h(uint64_t h, const char* a, size_t la, char sep, const char* b, size_t lb) {
        const int64_t f = find_h(h, a, la, sep, b, lb);
        if (f >= 0) return (uint32_t)f;
        if ((ents.size() + 1) * 2 > slots.size()) grow();
        const uint32_t id = (uint32_t)ents.size();
        ents.push_back({h, (uint64_t)arena.size(), (uint32_t)(la + 1 + lb)});
        arena.insert(arena.end(), a, a + la);
        arena.push_back(sep);
        arena.insert(arena.end(), b, b + lb);
        arena.push_back('\0');
        const size_t mask = slots.size() - 1;
        size_t i = h & mask;
        while (slots[i] != EMPTY) i = (i + 1) & mask;
        slots[i] = id;
        return id;
    }
    uint32_t intern(const char* a, char sep, const char* b) {
        const size_t la = std::strlen(a), lb = std::strlen(b);
        return intern_h(hash2(a, la, sep, b, lb), a, la, sep, b, lb);
    }
};

static inline const char* nz(const char* s) { return s ? s : ""; }

// ---- bound pods keyed by "namespace/name": key hash -> position, with erase (positions move on swap-remove) ----
struct PosTable {
    // slot = (high 32 bits of the key hash) << 32 | (position + 2); low word 0 = empty, 1 = tombstone
    std::vector<uint64_t> slots;
    size_t used = 0; // live + tombstones
    void clear() {
        slots.clear();
        used = 0;
    }
};

struct ksh_context {
    ks_snapshot* snap = nullptr;
    uint32_t N = 0;
    // nodes: names are interned once; name id -> current index (-1 = gone), index -> name id
    Interner names;
    std::vector<int32_t> nameid2idx;
    std::vector<uint32_t> idx2nameid;
    std::vector<int64_t> alloc_cpu, alloc_mem;
    // label pairs "key\0value" interned to pair ids; per pair: how many nodes carry it, and its dictionary bit
    Interner pairs;
    std::vector<uint32_t> pair_refcnt;
    std::vector<int32_t> pair_bit;            // -1 = no selector has named this pair yet
    std::vector<uint32_t> bit2pair;           // dictionary, in order of assignment
    std::vector<std::vector<uint32_t>> node_pids; // per node: pair ids of its labels
    uint32_t W = 1;
    // bound pods (the LIST results of predicates.rs:21-34), swap-remove on delete
    std::vector<int32_t> bnode;
    std::vector<int64_t> bcpu, bmem;
    std::vector<uint64_t> bhash;  // hash of "ns/name"; 0 with blen 0 = anonymous pod, not tracked
    std::vector<uint64_t> boff;   // key bytes in bkeys
    std::vector<uint32_t> blen;
    std::vector<char> bkeys;
    PosTable btab;
    uint64_t live_pairs = 0;     // pairs carried by at least one node
    uint64_t live_key_bytes = 0; // bytes of bkeys that belong to current rows
    // scratch of the bulk calls, kept between calls (fresh pages are expensive to fault in)
    std::vector<int32_t> tmp_node;
    std::vector<int64_t> tmp_cpu, tmp_mem;
    std::vector<uint64_t> tmp_hash;
    std::vector<uint32_t> tmp_len;
    std::vector<int64_t> pk_cpu, pk_mem; // packed form of the last ksh_select_nodes batch
    std::vector<uint64_t> pk_sel;
    bool dirty = true; // device snapshot must be re-uploaded
    // one lock per context: every ksh_* entry that takes a context holds it for the whole call (entries call each other)
    mutable std::recursive_mutex mu;
};

static uint32_t words_for_bits(uint32_t real_bits) {
    // +1: the last bit of the last word is the shared "no node carries this pair" bit
    uint32_t w = 1;
    while ((uint64_t)w * 64 < (uint64_t)real_bits + 1) w <<= 1;
    return w;
}

// ---- threads ----
static unsigned host_threads() {
    static const unsigned t = [] {
        if (const char* e = std::getenv("KSH_THREADS")) {
            const int v = std::atoi(e);
            if (v > 0) return (unsigned)std::min(v, 256);
        }
        const unsigned hc = std::thread::hardware_concurrency();
        return std::max(1u, std::min(hc ? hc : 1u, 32u));
    }();
    return t;
}

// f(thread, begin, end) over [0,n) in contiguous ascending ranges: concatenating per-thread results in thread order
// reproduces the serial order
template <class F>
static unsigned parallel_ranges(uint64_t n, uint64_t min_per_thread, F f) {
    unsigned T = (unsigned)std::min<uint64_t>(host_threads(), std::max<uint64_t>(1, n / std::max<uint64_t>(1, min_per_thread)));
    if (T <= 1) {
        f(0u, (uint64_t)0, n);
        return 1;
    }
    std::vector<std::thread> th;
    th.reserve(T - 1);
    const uint64_t per = (n + T - 1) / T;
    std::atomic<bool> oom{false}; // an exception must not leave a worker thread: it is re-raised after the join
    for (unsigned t = 1; t < T; t++)
        th.emplace_back([=, &f, &oom] {
            try {
                f(t, std::min(n, t * per), std::min(n, (t + 1) * per));
            } catch (...) {
                oom = true;
            }
        });
    try {
        f(0u, (uint64_t)0, std::min(n, per));
    } catch (...) {
        oom = true;
    }
    for (auto& x : th) x.join();
    if (oom) throw std::bad_alloc();
    return T;
}

struct alignas(128) RangeError { // first failure of a range (ranges stop at their first failure); one cache line pair each
    int rc = KS_OK;
    std::string msg;
};
static int first_error(const std::vector<RangeError>& errs) {
    for (const RangeError& e : errs)
        if (e.rc) return fail(e.rc, e.msg);
    return KS_OK;
}

// ---- parsing without touching the caller-visible error string (usable from worker threads) ----
static int parse_cpu(const char* q, int64_t* out, std::string* err) {
    const int rc = parse_quantity(q, 3, out);
    if (rc) *err = std::string("cannot convert cpu quantity '") + (q ? q : "(null)") + "' to integer millicores";
    return rc;
}
static int parse_mem(const char* q, int64_t* out, std::string* err) {
    const int rc = parse_quantity(q, 0, out);
    if (rc) *err = std::string("cannot convert memory quantity '") + (q ? q : "(null)") + "' to integer bytes";
    return rc;
}

// src/util.rs:54-75: sum of resources.requests over spec.containers only (no initContainers, no limits)
static int total_pod_resources(const ks_pod_obj* pod, int64_t* cpu, int64_t* mem, std::string* err) {
    int64_t c = 0, m = 0;
    if (pod->has_spec) {
        for (uint32_t i = 0; i < pod->n_containers; i++) {
            const ks_container_obj& ct = pod->containers[i];
            if (!ct.has_requests) continue;
            if (const char* q = kv_find(ct.requests, ct.n_requests, "cpu")) {
                int64_t v;
                const int rc = parse_cpu(q, &v, err); // reference: .expect("invalid pod spec: cpu request")
                if (rc) return rc;
                if (__builtin_add_overflow(c, v, &c)) return *err = "cpu request sum overflows", KS_ERR_RANGE;
            }
            if (const char* q = kv_find(ct.requests, ct.n_requests, "memory")) {
                int64_t v;
                const int rc = parse_mem(q, &v, err);
                if (rc) return rc;
                if (__builtin_add_overflow(m, v, &m)) return *err = "memory request sum overflows", KS_ERR_RANGE;
            }
        }
    }
    if (c > KS_MAX_CPU_MILLI || c < -KS_MAX_CPU_MILLI || m > KS_MAX_MEM_BYTES || m < -KS_MAX_MEM_BYTES)
        return *err = "pod requests outside +-2^36 millicores / +-2^55 bytes", KS_ERR_RANGE;
    *cpu = c;
    *mem = m;
    return KS_OK;
}

// allocatable of one node; mirrors src/predicates.rs:27-32
static int parse_node_allocatable(const ks_node_obj& nd, int64_t* ac, int64_t* am, std::string* err) {
    *ac = 0;
    *am = 0; // status/allocatable None => PodResources::new() = (0,0)   (predicates.rs:27-28)
    if (!nd.has_allocatable) return KS_OK;
    const std::string name = nz(nd.name);
    const char* q = kv_find(nd.allocatable, nd.n_allocatable, "cpu");
    if (!q) return *err = "node '" + name + "': allocatable has no cpu (reference panics, predicates.rs:29)", KS_ERR_MISSING;
    int rc = parse_cpu(q, ac, err);
    if (rc) return rc;
    q = kv_find(nd.allocatable, nd.n_allocatable, "memory");
    if (!q) return *err = "node '" + name + "': allocatable has no memory (reference panics, predicates.rs:30)", KS_ERR_MISSING;
    rc = parse_mem(q, am, err);
    if (rc) return rc;
    if (*ac > KS_MAX_CPU_MILLI || *ac < -KS_MAX_CPU_MILLI || *am > KS_MAX_MEM_BYTES || *am < -KS_MAX_MEM_BYTES)
        return *err = "node '" + name + "': allocatable out of range", KS_ERR_RANGE;
    return KS_OK;
}

// ---- label pairs ----
static uint32_t intern_pair(ksh_context* c, uint64_t h, const char* k, size_t lk, const char* v, size_t lv) {
    const uint32_t pid = c->pairs.intern_h(h, k, lk, '\0', v, lv);
    if (pid == c->pair_refcnt.size()) {
        c->pair_refcnt.push_back(0);
        c->pair_bit.push_back(-1);
    }
    return pid;
}

static void node_pairs_release(ksh_context* c, uint32_t idx) {
    for (uint32_t pid : c->node_pids[idx])
        if (--c->pair_refcnt[pid] == 0) c->live_pairs--;
    c->node_pids[idx].clear();
}

static void node_pairs_set(ksh_context* c, uint32_t idx, const ks_node_obj& nd) {
    std::vector<uint32_t>& v = c->node_pids[idx];
    if (!nd.has_labels) return;
    for (uint32_t l = 0; l < nd.n_labels; l++) {
        const char *k = nz(nd.labels[l].key), *val = nz(nd.labels[l].val);
        const size_t lk = std::strlen(k), lv = std::strlen(val);
        const uint32_t pid = intern_pair(c, hash2(k, lk, '\0', val, lv), k, lk, val, lv);
        if (std::find(v.begin(), v.end(), pid) != v.end()) continue; // a map has each key once; be safe
        v.push_back(pid);
        if (c->pair_refcnt[pid]++ == 0) c->live_pairs++;
    }
}

// A long-running host sees label values, node names and pods come and go; what the interners and the key arena
// keep for objects that no longer exist is reclaimed once it outweighs the live part.
static constexpr size_t GC_SLACK = 1024;

static void gc_pairs(ksh_context* c) {
    if (c->pairs.size() <= 2 * c->live_pairs + GC_SLACK) return;
    Interner fresh;
    std::vector<uint32_t> refcnt;
    std::vector<int32_t> bit;
    std::vector<uint32_t> remap(c->pairs.size(), 0xFFFFFFFFu);
    auto move_pair = [&](uint32_t pid) {
        if (remap[pid] != 0xFFFFFFFFu) return remap[pid];
        const Interner::Ent& e = c->pairs.ents[pid];
        const char* k = c->pairs.str(pid);
        const size_t lk = std::strlen(k);
        const uint32_t id = fresh.intern_h(e.h, k, lk, '\0', k + lk + 1, e.len - lk - 1);
        refcnt.push_back(c->pair_refcnt[pid]);
        bit.push_back(c->pair_bit[pid]);
        return remap[pid] = id;
    };
    for (uint32_t& pid : c->bit2pair) pid = move_pair(pid); // dictionary entries survive, stale ones included
    for (auto& v : c->node_pids)
        for (uint32_t& pid : v) pid = move_pair(pid);
    c->pairs = std::move(fresh);
    c->pair_refcnt.swap(refcnt);
    c->pair_bit.swap(bit);
}

static void gc_names(ksh_context* c) {
    if (c->names.size() <= 2 * (size_t)c->N + GC_SLACK) return;
    Interner fresh;
    std::vector<int32_t> id2idx;
    for (uint32_t i = 0; i < c->N; i++) {
        const uint32_t old = c->idx2nameid[i];
        const char* name = c->names.str(old);
        const uint32_t id = fresh.intern_h(c->names.ents[old].h, name, std::strlen(name), '\0', "", 0);
        if (id == id2idx.size()) id2idx.push_back((int32_t)i); // rows are visited in index order: the first row of a name wins
        c->idx2nameid[i] = id;
    }
    c->names = std::move(fresh);
    c->nameid2idx.swap(id2idx);
}

static void gc_bound_keys(ksh_context* c) {
    if (c->bkeys.size() <= 2 * c->live_key_bytes + 64 * GC_SLACK) return;
    std::vector<char> fresh;
    fresh.reserve(c->live_key_bytes);
    for (size_t pos = 0; pos < c->bnode.size(); pos++) {
        if (c->blen[pos] == 0) continue;
        const uint64_t off = fresh.size();
        fresh.insert(fresh.end(), c->bkeys.begin() + (ptrdiff_t)c->boff[pos], c->bkeys.begin() + (ptrdiff_t)(c->boff[pos] + c->blen[pos]));
        c->boff[pos] = off;
    }
    c->bkeys.swap(fresh);
}

// Dictionary bits of pairs that no node carries any more are dead weight (a selector naming such a pair is
// infeasible everywhere either way); they are dropped, and the rest renumbered in order, when space is needed.
static void compact_dictionary(ksh_context* c) {
    std::vector<uint32_t> kept;
    for (uint32_t pid : c->bit2pair) {
        if (c->pair_refcnt[pid] > 0) {
            c->pair_bit[pid] = (int32_t)kept.size();
            kept.push_back(pid);
        } else {
            c->pair_bit[pid] = -1;
        }
    }
    if (kept.size() != c->bit2pair.size()) c->dirty = true;
    c->bit2pair.swap(kept);
}

static int assign_bit(ksh_context* c, uint32_t pid) {
    if (c->pair_bit[pid] >= 0) return KS_OK;
    if (words_for_bits((uint32_t)c->bit2pair.size() + 1) > KS_MAX_LABEL_WORDS) {
        compact_dictionary(c);
        if (words_for_bits((uint32_t)c->bit2pair.size() + 1) > KS_MAX_LABEL_WORDS)
            return fail(KS_ERR_RANGE, "more than 511 distinct (key,value) pairs referenced by selectors");
    }
    c->pair_bit[pid] = (int32_t)c->bit2pair.size();
    c->bit2pair.push_back(pid);
    c->dirty = true;
    return KS_OK;
}

static void update_words(ksh_context* c) {
    const uint32_t w = words_for_bits((uint32_t)c->bit2pair.size());
    if (w != c->W) {
        c->W = w;
        c->dirty = true;
    }
}

// node label words [N*W] under the current dictionary
static void fill_label_words(const ksh_context* c, uint64_t* lab) {
    std::memset(lab, 0, (size_t)c->N * c->W * sizeof(uint64_t));
    for (uint32_t n = 0; n < c->N; n++)
        for (uint32_t pid : c->node_pids[n]) {
            const int32_t bit = c->pair_bit[pid];
            if (bit >= 0) lab[(size_t)n * c->W + ((uint32_t)bit >> 6)] |= 1ull << (bit & 63);
        }
}

// ---- bound pods ----
static inline uint64_t pod_key_hash(const ks_pod_obj* pod, size_t* lns, size_t* lname) {
    const char *ns = nz(pod->ns), *name = nz(pod->name);
    *lns = std::strlen(ns);
    *lname = std::strlen(name);
    return hash2(ns, *lns, '/', name, *lname);
}

static bool bkey_equal(const ksh_context* c, size_t pos, uint64_t h, const ks_pod_obj* pod, size_t lns, size_t lname) {
    if (c->bhash[pos] != h || c->blen[pos] != lns + 1 + lname) return false;
    const char* s = c->bkeys.data() + c->boff[pos];
    return std::memcmp(s, nz(pod->ns), lns) == 0 && s[lns] == '/' && std::memcmp(s + lns + 1, nz(pod->name), lname) == 0;
}

static inline uint64_t btab_word(uint64_t h, size_t pos) { return (h & 0xFFFFFFFF00000000ull) | (uint32_t)(pos + 2); }
static inline uint32_t slot_low(uint64_t s) { return (uint32_t)s; }

static void btab_rebuild(ksh_context* c, size_t want_live) {
    size_t cap = 64;
    while (cap < want_live * 2 + 2) cap *= 2;
    c->btab.slots.assign(cap, 0);
    c->btab.used = 0;
    for (size_t pos = 0; pos < c->bnode.size(); pos++) {
        if (c->blen[pos] == 0) continue;
        size_t i = c->bhash[pos] & (cap - 1);
        while (c->btab.slots[i] != 0) i = (i + 1) & (cap - 1);
        c->btab.slots[i] = btab_word(c->bhash[pos], pos);
        c->btab.used++;
    }
}

static int64_t btab_find_slot(const ksh_context* c, uint64_t h, const ks_pod_obj* pod, size_t lns, size_t lname) {
    if (c->btab.slots.empty()) return -1;
    const size_t mask = c->btab.slots.size() - 1;
    for (size_t i = h & mask;; i = (i + 1) & mask) {
        const uint64_t s = c->btab.slots[i];
        if (s == 0) return -1;
        if (slot_low(s) >= 2 && (s >> 32) == (h >> 32) && bkey_equal(c, slot_low(s) - 2, h, pod, lns, lname)) return (int64_t)i;
    }
}

static int64_t btab_slot_of_pos(const ksh_context* c, size_t pos) {
    if (c->btab.slots.empty()) return -1;
    const size_t mask = c->btab.slots.size() - 1;
    for (size_t i = c->bhash[pos] & mask;; i = (i + 1) & mask) {
        const uint64_t s = c->btab.slots[i];
        if (s == 0) return -1;
        if (slot_low(s) == pos + 2) return (int64_t)i;
    }
}

static void bound_erase(ksh_context* c, size_t pos);

// append a bound pod; a tracked key that is already present REPLACES its row (last one wins, as a map would): the old
// row is removed so its requests are no longer charged, and the device copy is re-uploaded from the host truth
static void bound_push(ksh_context* c, const ks_pod_obj* pod, uint64_t h, size_t lns, size_t lname, int32_t node, int64_t cpu,
                       int64_t mem) {
    const bool tracked = lns + lname > 0; // a pod without namespace and name cannot be addressed by later events
    if (tracked) {
        const int64_t old = btab_find_slot(c, h, pod, lns, lname);
        if (old >= 0) {
            bound_erase(c, slot_low(c->btab.slots[(size_t)old]) - 2);
            c->dirty = true;
        }
    }
    const size_t pos = c->bnode.size();
    c->bnode.push_back(node);
    c->bcpu.push_back(cpu);
    c->bmem.push_back(mem);
    c->bhash.push_back(tracked ? h : 0);
    c->boff.push_back(c->bkeys.size());
    c->blen.push_back(tracked ? (uint32_t)(lns + 1 + lname) : 0);
    if (!tracked) return;
    c->live_key_bytes += lns + 1 + lname;
    c->bkeys.insert(c->bkeys.end(), nz(pod->ns), nz(pod->ns) + lns);
    c->bkeys.push_back('/');
    c->bkeys.insert(c->bkeys.end(), nz(pod->name), nz(pod->name) + lname);
    if ((c->btab.used + 1) * 2 > c->btab.slots.size()) { // also the first insertion
        btab_rebuild(c, c->bnode.size());                // places every row, the new one included
        return;
    }
    // one probe walk: an older row with the same key is re-pointed; else the first free slot on the way is taken
    const size_t mask = c->btab.slots.size() - 1;
    int64_t free_slot = -1;
    for (size_t i = h & mask;; i = (i + 1) & mask) {
        const uint64_t s = c->btab.slots[i];
        if (s == 0) {
            if (free_slot < 0) {
                free_slot = (int64_t)i;
                c->btab.used++;
            }
            break;
        }
        if (slot_low(s) == 1) {
            if (free_slot < 0) free_slot = (int64_t)i;
        } else if ((s >> 32) == (h >> 32) && bkey_equal(c, slot_low(s) - 2, h, pod, lns, lname)) {
            c->btab.slots[i] = btab_word(h, pos);
            return;
        }
    }
    c->btab.slots[(size_t)free_slot] = btab_word(h, pos);
}

static void bound_erase(ksh_context* c, size_t pos) { // swap-remove
    const size_t last = c->bnode.size() - 1;
    c->live_key_bytes -= c->blen[pos];
    if (c->blen[pos]) {
        const int64_t i = btab_slot_of_pos(c, pos);
        if (i >= 0) c->btab.slots[(size_t)i] = 1; // tombstone
    }
    if (pos != last) {
        if (c->blen[last]) {
            const int64_t i = btab_slot_of_pos(c, last);
            if (i >= 0) c->btab.slots[(size_t)i] = btab_word(c->bhash[last], pos);
        }
        c->bnode[pos] = c->bnode[last];
        c->bcpu[pos] = c->bcpu[last];
        c->bmem[pos] = c->bmem[last];
        c->bhash[pos] = c->bhash[last];
        c->boff[pos] = c->boff[last];
        c->blen[pos] = c->blen[last];
    }
    c->bnode.pop_back();
    c->bcpu.pop_back();
    c->bmem.pop_back();
    c->bhash.pop_back();
    c->boff.pop_back();
    c->blen.pop_back();
    if (c->bnode.empty()) c->bkeys.clear(); // otherwise erased keys wait for gc_bound_keys
}

static void bound_clear(ksh_context* c) {
    c->bnode.clear();
    c->bcpu.clear();
    c->bmem.clear();
    c->bhash.clear();
    c->boff.clear();
    c->blen.clear();
    c->bkeys.clear();
    c->btab.clear();
    c->live_key_bytes = 0;
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

static inline int32_t node_index_of(const ksh_context* c, const char* name) {
    const int64_t id = c->names.find(nz(name), '\0', "");
    return id < 0 ? -1 : c->nameid2idx[(size_t)id];
}

// ---- packing a batch of pending pods (src/util.rs:54-75 + the selector side of src/predicates.rs:45-61) ----
// One pass over the objects does all the string work: request totals (quantity parse) and ONE interner lookup per selector
// entry, whose pair ids are parked per thread in array order.  Then the dictionary grows (serial, tiny), and a second pass over
// the same ranges turns the parked pair ids into selector bits without touching a string again.  Every selector pair of the
// batch that some node carries gets a dictionary bit, first occurrence (in pod order) first, whatever the thread count.
constexpr uint64_t PACK_MIN_PER_THREAD = 2048; // both passes must cut [0,n) into the same ranges

struct PackScratch {
    std::vector<std::vector<int32_t>> pids;  // per thread: pair id (-1 = unknown pair) of every selector entry of its range, in order
    std::vector<std::vector<uint32_t>> need; // per thread: known pairs some node carries that have no dictionary bit yet
};

static int pack_scan(const ksh_context* c, const ks_pod_obj* pods, uint64_t n, int64_t* req_cpu, int64_t* req_mem, PackScratch& ps) {
    ps.pids.assign(host_threads(), {});
    ps.need.assign(host_threads(), {});
    std::vector<RangeError> errs(host_threads());
    parallel_ranges(n, PACK_MIN_PER_THREAD, [&](unsigned t, uint64_t b, uint64_t e) {
        std::vector<int32_t>& pids = ps.pids[t];
        std::vector<uint32_t>& need = ps.need[t];
        for (uint64_t p = b; p < e; p++) {
            const ks_pod_obj& pod = pods[p];
            const int rc = total_pod_resources(&pod, &req_cpu[p], &req_mem[p], &errs[t].msg);
            if (rc) {
                errs[t].rc = rc;
                return;
            }
            if (!(pod.has_spec && pod.has_node_selector)) continue;
            for (uint32_t i = 0; i < pod.n_selector; i++) {
                const int64_t pid = c->pairs.find(nz(pod.selector[i].key), '\0', nz(pod.selector[i].val));
                pids.push_back((int32_t)pid);
                if (pid < 0 || c->pair_bit[(size_t)pid] >= 0 || c->pair_refcnt[(size_t)pid] == 0) continue; // unknown pair, known bit, or absent everywhere
                if (need.empty() || need.back() != (uint32_t)pid) need.push_back((uint32_t)pid);
            }
        }
    });
    return first_error(errs);
}

static int grow_dictionary(ksh_context* c, const PackScratch& ps, const ks_pod_obj* pods, uint64_t n) {
    const std::vector<std::vector<uint32_t>>& need = ps.need;
    // Room check first.  The dictionary only has to cover the pairs THIS batch names (node columns are re-uploaded
    // whenever it changes), so when the pairs accumulated over the context's lifetime leave no room - e.g. hostname
    // selectors across more than 511 nodes over time - it is rebuilt from the current batch instead of failing.
    {
        std::vector<uint32_t> fresh;
        std::vector<uint8_t> seen(c->pairs.size(), 0);
        for (const auto& v : need)
            for (uint32_t pid : v)
                if (!seen[pid]) {
                    seen[pid] = 1;
                    fresh.push_back(pid);
                }
        if (words_for_bits((uint32_t)(c->bit2pair.size() + fresh.size())) > KS_MAX_LABEL_WORDS) {
            compact_dictionary(c);
            if (words_for_bits((uint32_t)(c->bit2pair.size() + fresh.size())) > KS_MAX_LABEL_WORDS) {
                for (uint32_t pid : c->bit2pair) c->pair_bit[pid] = -1;
                c->bit2pair.clear();
                c->dirty = true;
                std::fill(seen.begin(), seen.end(), 0);
                for (uint64_t p = 0; p < n; p++) { // rare path: serial re-scan, every live pair of the batch
                    const ks_pod_obj& pod = pods[p];
                    if (!(pod.has_spec && pod.has_node_selector)) continue;
                    for (uint32_t i = 0; i < pod.n_selector; i++) {
                        const int64_t pid = c->pairs.find(nz(pod.selector[i].key), '\0', nz(pod.selector[i].val));
                        if (pid < 0 || c->pair_refcnt[(size_t)pid] == 0 || seen[(size_t)pid]) continue;
                        seen[(size_t)pid] = 1;
                        if (words_for_bits((uint32_t)c->bit2pair.size() + 1) > KS_MAX_LABEL_WORDS)
                            return fail(KS_ERR_RANGE, "one batch names more than 511 distinct (key,value) pairs that nodes carry: split the batch");
                        c->pair_bit[(size_t)pid] = (int32_t)c->bit2pair.size();
                        c->bit2pair.push_back((uint32_t)pid);
                    }
                }
                update_words(c);
                return KS_OK;
            }
        }
    }
    for (const auto& v : need)
        for (uint32_t pid : v) {
            const int rc = assign_bit(c, pid);
            if (rc) return rc;
        }
    update_words(c);
    return KS_OK;
}

// selector words of the batch under the current dictionary, from the pair ids parked by pack_scan (same ranges, same order)
static void pack_selectors(const ksh_context* c, const ks_pod_obj* pods, uint64_t n, const PackScratch& ps, uint64_t* sel,
                           uint32_t stride) {
    const uint32_t absent = c->W * 64 - 1;
    parallel_ranges(n, PACK_MIN_PER_THREAD, [&](unsigned t, uint64_t b, uint64_t e) {
        const int32_t* pid = ps.pids[t].data();
        for (uint64_t p = b; p < e; p++) {
            uint64_t* row = sel + p * stride;
            for (uint32_t w = 0; w < stride; w++) row[w] = 0;
            if (!(pods[p].has_spec && pods[p].has_node_selector)) continue;
            for (uint32_t i = 0; i < pods[p].n_selector; i++) {
                const int32_t id = *pid++;
                const int32_t b_ = id < 0 ? -1 : c->pair_bit[(size_t)id];
                const uint32_t bit = b_ < 0 ? absent : (uint32_t)b_; // a pair no node carries: never-set bit
                row[bit >> 6] |= 1ull << (bit & 63);
            }
        }
    });
}

extern "C" {

int ksh_parse_cpu_millicores(const char* q, int64_t* out) try {
    if (!out) return fail(KS_ERR_INVALID, "out is NULL");
    std::string err;
    const int rc = parse_cpu(q, out, &err);
    return rc ? fail(rc, err) : KS_OK;
}
KSH_CATCH

int ksh_parse_memory_bytes(const char* q, int64_t* out) try {
    if (!out) return fail(KS_ERR_INVALID, "out is NULL");
    std::string err;
    const int rc = parse_mem(q, out, &err);
    return rc ? fail(rc, err) : KS_OK;
}
KSH_CATCH

int ksh_total_pod_resources(const ks_pod_obj* pod, int64_t* cpu, int64_t* mem) try {
    if (!pod || !cpu || !mem) return fail(KS_ERR_INVALID, "NULL argument");
    std::string err;
    const int rc = total_pod_resources(pod, cpu, mem, &err);
    return rc ? fail(rc, err) : KS_OK;
}
KSH_CATCH

int ksh_is_pod_bound(const ks_pod_obj* pod) { return pod && pod->has_spec && pod->node_name != nullptr; } // util.rs:38-45

int ksh_context_create(int device, ksh_context** out) try {
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
KSH_CATCH

void ksh_context_destroy(ksh_context* c) {
    if (!c) return;
    ks_snapshot_destroy(c->snap);
    delete c;
}

uint32_t ksh_context_num_nodes(const ksh_context* c) {
    if (!c) return 0;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    return c->N;
}
uint32_t ksh_context_label_words(const ksh_context* c) {
    if (!c) return 0;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    return c->W;
}
uint64_t ksh_context_num_bound(const ksh_context* c) {
    if (!c) return 0;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    return c->bnode.size();
}

int ksh_context_export_packed(const ksh_context* c, int64_t* alloc_cpu, int64_t* alloc_mem, uint64_t* labels,
                              int32_t* bound_node, int64_t* bound_cpu, int64_t* bound_mem) try {
    if (!c) return fail(KS_ERR_INVALID, "NULL argument");
    std::lock_guard<std::recursive_mutex> lk(c->mu);
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
KSH_CATCH

ks_snapshot* ksh_context_snapshot(ksh_context* c) {
    if (!c) return nullptr;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    if (upload(c)) return nullptr;
    return c->snap;
}

int ksh_context_set_nodes(ksh_context* c, const ks_node_obj* nodes, uint32_t n) try {
    if (!c || (n && !nodes)) return fail(KS_ERR_INVALID, "NULL argument");
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    // validate first (quantities parsed by all host threads), then replace
    std::vector<int64_t> ac(n), am(n);
    std::vector<RangeError> errs(host_threads());
    parallel_ranges(n, 2048, [&](unsigned t, uint64_t b, uint64_t e) {
        for (uint64_t i = b; i < e; i++) {
            const int rc = parse_node_allocatable(nodes[i], &ac[i], &am[i], &errs[t].msg);
            if (rc) {
                errs[t].rc = rc;
                return;
            }
        }
    });
    int rc = first_error(errs);
    if (rc) return rc;
    // selector pairs that already have a bit keep it if some new node still carries the pair
    std::vector<std::string> named; // "key\0value" of the dictionary, in bit order
    named.reserve(c->bit2pair.size());
    for (uint32_t pid : c->bit2pair) named.emplace_back(c->pairs.str(pid), c->pairs.ents[pid].len);
    c->names.clear();
    c->nameid2idx.clear();
    c->idx2nameid.assign(n, 0);
    c->pairs.clear();
    c->live_pairs = 0;
    c->pair_refcnt.clear();
    c->pair_bit.clear();
    c->bit2pair.clear();
    c->node_pids.assign(n, {});
    c->N = n;
    c->alloc_cpu.swap(ac);
    c->alloc_mem.swap(am);
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t id = c->names.intern(nz(nodes[i].name), '\0', "");
        if (id == c->nameid2idx.size()) c->nameid2idx.push_back((int32_t)i); // first node of a name wins
        c->idx2nameid[i] = id;
        node_pairs_set(c, i, nodes[i]);
    }
    for (const std::string& pk : named) {
        const char* k = pk.c_str();
        const char* v = k + std::strlen(k) + 1;
        const int64_t pid = c->pairs.find(k, '\0', v);
        if (pid >= 0 && c->pair_refcnt[(size_t)pid] > 0) {
            c->pair_bit[(size_t)pid] = (int32_t)c->bit2pair.size();
            c->bit2pair.push_back((uint32_t)pid);
        }
    }
    c->W = words_for_bits((uint32_t)c->bit2pair.size());
    bound_clear(c);
    c->dirty = true;
    return KS_OK;
}
KSH_CATCH

int ksh_context_upsert_node(ksh_context* c, const ks_node_obj* node, uint32_t* out_idx) try {
    if (!c || !node) return fail(KS_ERR_INVALID, "NULL argument");
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    int64_t ac, am;
    std::string err;
    const int rc = parse_node_allocatable(*node, &ac, &am, &err);
    if (rc) return fail(rc, err);
    const uint32_t id = c->names.intern(nz(node->name), '\0', "");
    if (id == c->nameid2idx.size()) c->nameid2idx.push_back(-1);
    uint32_t idx;
    if (c->nameid2idx[id] < 0) { // new node (or one that had been removed): appended
        idx = c->N++;
        c->nameid2idx[id] = (int32_t)idx;
        c->idx2nameid.push_back(id);
        c->alloc_cpu.push_back(ac);
        c->alloc_mem.push_back(am);
        c->node_pids.emplace_back();
    } else {
        idx = (uint32_t)c->nameid2idx[id];
        c->alloc_cpu[idx] = ac;
        c->alloc_mem[idx] = am;
        node_pairs_release(c, idx);
    }
    node_pairs_set(c, idx, *node);
    gc_pairs(c);
    c->dirty = true;
    if (out_idx) *out_idx = idx;
    return KS_OK;
}
KSH_CATCH

int ksh_context_remove_node(ksh_context* c, const char* name) try {
    if (!c || !name) return fail(KS_ERR_INVALID, "NULL argument");
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    const int32_t found = node_index_of(c, name);
    if (found < 0) return KS_OK;
    const uint32_t idx = (uint32_t)found;
    node_pairs_release(c, idx);
    c->nameid2idx[c->idx2nameid[idx]] = -1;
    c->idx2nameid.erase(c->idx2nameid.begin() + idx);
    c->alloc_cpu.erase(c->alloc_cpu.begin() + idx);
    c->alloc_mem.erase(c->alloc_mem.begin() + idx);
    c->node_pids.erase(c->node_pids.begin() + idx);
    c->N--;
    // later nodes move down by one index; a name carried by several rows points to its first remaining row
    for (uint32_t i = c->N; i-- > idx;) c->nameid2idx[c->idx2nameid[i]] = (int32_t)i;
    for (uint32_t i = 0; i < idx; i++)
        if (c->nameid2idx[c->idx2nameid[i]] > (int32_t)i) c->nameid2idx[c->idx2nameid[i]] = (int32_t)i;
    // pods bound to the removed node disappear from every later LIST; indices above it move down
    for (size_t p = 0; p < c->bnode.size();) {
        if ((uint32_t)c->bnode[p] == idx) {
            bound_erase(c, p);
        } else {
            if ((uint32_t)c->bnode[p] > idx) c->bnode[p]--;
            p++;
        }
    }
    gc_pairs(c);
    gc_names(c);
    gc_bound_keys(c);
    c->dirty = true;
    return KS_OK;
}
KSH_CATCH

int ksh_context_pod_bound(ksh_context* c, const ks_pod_obj* pod) try {
    if (!c || !pod) return fail(KS_ERR_INVALID, "NULL argument");
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    if (!ksh_is_pod_bound(pod)) return KS_OK;
    const int32_t node = node_index_of(c, pod->node_name);
    if (node < 0) return KS_OK;
    int64_t cpu, mem;
    std::string err;
    const int rc = total_pod_resources(pod, &cpu, &mem, &err);
    if (rc) return fail(rc, err);
    size_t lns, lname;
    const uint64_t h = pod_key_hash(pod, &lns, &lname);
    if (lns + lname > 0) {
        const int64_t slot = btab_find_slot(c, h, pod, lns, lname);
        if (slot >= 0) bound_erase(c, slot_low(c->btab.slots[(size_t)slot]) - 2); // update of a pod already known
    }
    bound_push(c, pod, h, lns, lname, node, cpu, mem);
    gc_bound_keys(c);
    c->dirty = true;
    return KS_OK;
}
KSH_CATCH

int ksh_context_pod_deleted(ksh_context* c, const ks_pod_obj* pod) try {
    if (!c || !pod) return fail(KS_ERR_INVALID, "NULL argument");
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    size_t lns, lname;
    const uint64_t h = pod_key_hash(pod, &lns, &lname);
    if (lns + lname == 0) return KS_OK;
    const int64_t slot = btab_find_slot(c, h, pod, lns, lname);
    if (slot < 0) return KS_OK;
    bound_erase(c, slot_low(c->btab.slots[(size_t)slot]) - 2);
    gc_bound_keys(c);
    c->dirty = true;
    return KS_OK;
}
KSH_CATCH

const char* ksh_context_node_name(const ksh_context* c, uint32_t idx) {
    if (!c) return nullptr;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    return idx < c->N ? c->names.str(c->idx2nameid[idx]) : nullptr;
}

int ksh_context_set_cluster_pods(ksh_context* c, const ks_pod_obj* pods, uint64_t n) try {
    if (!c || (n && !pods)) return fail(KS_ERR_INVALID, "NULL argument");
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    if (n > 0xFFFFFFF0ull) return fail(KS_ERR_RANGE, "too many pods");
    // validate first, then replace.  Per pod (all host threads): spec.nodeName -> node index (the field selector of
    // predicates.rs:22-25), request totals (predicates.rs:37), key hash.
    std::vector<int32_t>& node = c->tmp_node;
    std::vector<int64_t>&cpus = c->tmp_cpu, &mems = c->tmp_mem;
    std::vector<uint64_t>& hashes = c->tmp_hash;
    std::vector<uint32_t>& lens = c->tmp_len;
    node.resize(n);
    cpus.resize(n);
    mems.resize(n);
    hashes.resize(n);
    lens.resize(2 * n);
    std::vector<RangeError> errs(host_threads());
    parallel_ranges(n, 4096, [&](unsigned t, uint64_t b, uint64_t e) {
        for (uint64_t p = b; p < e; p++) {
            node[p] = ksh_is_pod_bound(&pods[p]) ? node_index_of(c, pods[p].node_name) : -1;
            if (node[p] < 0) continue;
            const int rc = total_pod_resources(&pods[p], &cpus[p], &mems[p], &errs[t].msg);
            if (rc) {
                errs[t].rc = rc;
                return;
            }
            size_t lns, lname;
            hashes[p] = pod_key_hash(&pods[p], &lns, &lname);
            lens[2 * p] = (uint32_t)lns;
            lens[2 * p + 1] = (uint32_t)lname;
        }
    });
    const int rc = first_error(errs);
    if (rc) return rc;
    bound_clear(c);
    uint64_t kept = 0, key_bytes = 0;
    for (uint64_t p = 0; p < n; p++)
        if (node[p] >= 0) {
            kept++;
            key_bytes += (uint64_t)lens[2 * p] + lens[2 * p + 1] + 1;
        }
    c->bnode.reserve(kept);
    c->bcpu.reserve(kept);
    c->bmem.reserve(kept);
    c->bhash.reserve(kept);
    c->boff.reserve(kept);
    c->blen.reserve(kept);
    c->bkeys.reserve(key_bytes);
    btab_rebuild(c, kept); // sized once for the whole list
    for (uint64_t p = 0; p < n; p++)
        if (node[p] >= 0) bound_push(c, &pods[p], hashes[p], lens[2 * p], lens[2 * p + 1], node[p], cpus[p], mems[p]);
    c->dirty = true;
    return KS_OK;
}
KSH_CATCH

int ksh_pack_pods(ksh_context* c, const ks_pod_obj* pods, uint64_t n, int64_t* req_cpu, int64_t* req_mem,
                  uint64_t* sel, uint32_t stride) try {
    if (!c || (n && (!pods || !req_cpu || !req_mem || !sel))) return fail(KS_ERR_INVALID, "NULL argument");
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    PackScratch ps;
    int rc = pack_scan(c, pods, n, req_cpu, req_mem, ps);
    if (rc) return rc;
    rc = grow_dictionary(c, ps, pods, n);
    if (rc) return rc;
    if (stride < c->W) return fail(KS_ERR_INVALID, "sel_stride_words smaller than the dictionary's word count");
    pack_selectors(c, pods, n, ps, sel, stride);
    return (int)c->W;
}
KSH_CATCH

static int pack_and_upload(ksh_context* c, const ks_pod_obj* pods, uint64_t n, std::vector<int64_t>& rc_, std::vector<int64_t>& rm_,
                           std::vector<uint64_t>& sel) {
    rc_.resize(n);
    rm_.resize(n);
    PackScratch ps;
    int rc = pack_scan(c, pods, n, rc_.data(), rm_.data(), ps);
    if (rc) return rc;
    rc = grow_dictionary(c, ps, pods, n);
    if (rc) return rc;
    sel.resize((size_t)n * c->W);
    pack_selectors(c, pods, n, ps, sel.data(), c->W);
    return upload(c);
}

int ksh_check_node_validity(ksh_context* c, const ks_pod_obj* pod, uint32_t node_idx) try {
    if (!c || !pod) return fail(KS_ERR_INVALID, "NULL argument");
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    if (node_idx >= c->N) return fail(KS_ERR_INVALID, "node index out of range");
    std::vector<int64_t> rc_, rm_;
    std::vector<uint64_t> sel;
    int rc = pack_and_upload(c, pod, 1, rc_, rm_, sel);
    if (rc) return rc;
    return ks_check_cell(c->snap, rc_[0], rm_[0], sel.data(), node_idx);
}
KSH_CATCH

int ksh_select_nodes(ksh_context* c, const ks_pod_obj* pods, uint64_t n, int policy, int32_t* out_node_idx,
                     int64_t* out_score, uint32_t* out_cnt) try {
    if (!c || (n && !pods)) return fail(KS_ERR_INVALID, "NULL argument");
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    if (n == 0) return KS_OK;
    // packed form of the batch: kept in the context between calls (24 bytes per pod; fresh pages cost more than the packing)
    std::vector<int64_t>&rc_ = c->pk_cpu, &rm_ = c->pk_mem;
    std::vector<uint64_t>& sel = c->pk_sel;
    int rc = pack_and_upload(c, pods, n, rc_, rm_, sel);
    if (rc) return rc;
    ks_pods kp{n, rc_.data(), rm_.data(), sel.data(), KS_MEM_HOST};
    ks_bindings kb{out_node_idx, out_score, out_cnt, KS_MEM_HOST, nullptr, 0, KS_MEM_HOST, nullptr, nullptr};
    return ks_select(c->snap, &kp, policy, KS_SELECT_AUTO, &kb, nullptr);
}
KSH_CATCH

int ksh_select_node_for_pod(ksh_context* c, const ks_pod_obj* pods, uint64_t n, uint32_t attempts, uint64_t seed,
                            uint64_t first_pod_index, int32_t* out_node_idx, uint32_t* out_attempts,
                            int32_t* out_draw_node, uint8_t* out_draw_code) try {
    if (!c || (n && !pods)) return fail(KS_ERR_INVALID, "NULL argument");
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    if (n == 0) return KS_OK;
    std::vector<int64_t> rc_, rm_;
    std::vector<uint64_t> sel;
    int rc = pack_and_upload(c, pods, n, rc_, rm_, sel);
    if (rc) return rc;
    ks_pods kp{n, rc_.data(), rm_.data(), sel.data(), KS_MEM_HOST};
    return ks_select_sampling(c->snap, &kp, attempts, seed, first_pod_index, out_node_idx, out_attempts, out_draw_node,
                              out_draw_code);
}
KSH_CATCH

// A pod that comes back to reconcile() unbound although this context bound it earlier (the caller's POST failed or
// raced, error_policy requeued it - src/main.rs:105-108,122-125): the reference's next LIST would not show it, so
// the earlier charge is dropped before the new selection.  Returns true if a row was released.
static bool release_earlier_binding(ksh_context* c, const ks_pod_obj* pod) {
    size_t lns, lname;
    const uint64_t h = pod_key_hash(pod, &lns, &lname);
    if (lns + lname == 0) return false;
    const int64_t slot = btab_find_slot(c, h, pod, lns, lname);
    if (slot < 0) return false;
    bound_erase(c, slot_low(c->btab.slots[(size_t)slot]) - 2);
    c->dirty = true; // the device's free[] still carries the old charge: re-upload before the next evaluation
    return true;
}

// corev1::Binding{metadata, target: ObjectReference{name}}  (src/main.rs:83-91), the body of
// POST /api/v1/namespaces/{ns}/pods/{name}/binding
static std::string binding_body(const ksh_context* c, const ks_pod_obj* pod, uint32_t node_idx) {
    std::string s = "{\"apiVersion\":\"v1\",\"kind\":\"Binding\",\"metadata\":";
    if (pod->metadata_json && pod->metadata_json[0]) { // src/main.rs:88: metadata: pod.metadata.clone()
        s += pod->metadata_json;
    } else {
        s += "{\"name\":\"";
        json_escape(s, pod->name);
        s += "\",\"namespace\":\"";
        json_escape(s, pod->ns);
        s += "\"}";
    }
    s += ",\"target\":{\"name\":\"";
    json_escape(s, ksh_context_node_name(c, node_idx));
    s += "\"}}";
    return s;
}

int ksh_reconcile(ksh_context* c, const ks_pod_obj* pod, int policy, int32_t* node_idx, char* json, size_t cap) try {
    if (!c || !pod || !node_idx) return fail(KS_ERR_INVALID, "NULL argument");
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    *node_idx = -1;
    if (json && cap) json[0] = '\0';
    if (ksh_is_pod_bound(pod)) return KSH_RECONCILE_OK; // src/main.rs:74-76
    release_earlier_binding(c, pod);
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
    if (rc) {
        c->dirty = true; // the device may or may not carry the bind: re-upload from the host truth next time
        return rc;
    }
    {
        size_t lns, lname;
        const uint64_t h = pod_key_hash(pod, &lns, &lname);
        bound_push(c, pod, h, lns, lname, idx, cpu, mem);
    }
    *node_idx = idx;
    if (json && cap) std::snprintf(json, cap, "%s", binding_body(c, pod, (uint32_t)idx).c_str());
    return KSH_RECONCILE_OK;
}
KSH_CATCH

// reconcile() for a drained queue: one pack, one micro-batch loop on the device (select -> claims resolved in array order ->
// losers re-selected against what is left), then the host-side commit of every bind (what the next LIST would show)
int ksh_reconcile_batch(ksh_context* c, const ks_pod_obj* pods, uint64_t n, int policy, int32_t* out_status, int32_t* out_node_idx,
                        char* json, size_t cap, int64_t* out_json_off, uint32_t* out_rounds) try {
    if (!c || (n && (!pods || !out_status || !out_node_idx))) return fail(KS_ERR_INVALID, "NULL argument");
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    if (out_rounds) *out_rounds = 0;
    std::vector<uint64_t> todo;
    std::vector<ks_pod_obj> sub; // shallow copies: the pods that actually go to the device, in arrival order
    for (uint64_t i = 0; i < n; i++) {
        out_node_idx[i] = -1;
        if (out_json_off) out_json_off[i] = -1;
        if (ksh_is_pod_bound(&pods[i])) {
            out_status[i] = KSH_RECONCILE_OK; // src/main.rs:74-76
        } else if (!pods[i].ns || !pods[i].name) {
            out_status[i] = KSH_RECONCILE_BINDING_OBJECT_FAILED; // reference: unwrap panic, src/main.rs:80
        } else {
            out_status[i] = KSH_RECONCILE_NO_NODE_FOUND; // until bound below (src/main.rs:116-118)
            release_earlier_binding(c, &pods[i]);
            todo.push_back(i);
            sub.push_back(pods[i]);
        }
    }
    if (todo.empty()) return KS_OK;
    std::vector<int64_t> rc_, rm_;
    std::vector<uint64_t> sel;
    int rc = pack_and_upload(c, sub.data(), sub.size(), rc_, rm_, sel);
    if (rc) return rc;
    std::vector<int32_t> idx(sub.size(), -1);
    ks_pods kp{sub.size(), rc_.data(), rm_.data(), sel.data(), KS_MEM_HOST};
    rc = ks_stream_bind(c->snap, &kp, policy, idx.data(), nullptr, out_rounds); // commits capacity on the device
    if (rc) {
        c->dirty = true; // some rounds may have committed claims on the device: the host list is the truth
        return rc;
    }
    size_t used = 0;
    for (size_t k = 0; k < todo.size(); k++) {
        if (idx[k] < 0) continue;
        const uint64_t i = todo[k];
        size_t lns, lname;
        const uint64_t h = pod_key_hash(&pods[i], &lns, &lname);
        bound_push(c, &pods[i], h, lns, lname, idx[k], rc_[k], rm_[k]);
        out_status[i] = KSH_RECONCILE_OK;
        out_node_idx[i] = idx[k];
        if (json && out_json_off) {
            const std::string body = binding_body(c, &pods[i], (uint32_t)idx[k]);
            if (used + body.size() + 1 <= cap) { // bodies back to back, each NUL-terminated; -1 = did not fit
                std::memcpy(json + used, body.c_str(), body.size() + 1);
                out_json_off[i] = (int64_t)used;
                used += body.size() + 1;
            }
        }
    }
    return KS_OK;
}
KSH_CATCH

} // extern "C"
