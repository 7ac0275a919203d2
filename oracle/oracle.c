/* oracle.c — CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 * Every function cites the reference lines it follows (relative to /root/reference).
 * PARITY: nodeSelector pinned by src/predicates/test.rs:42-58; resource fit "parity unpinned". */
#define _GNU_SOURCE
#include "oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

typedef __int128 i128;

/* ------------------------------------------------------------------------------------------------
 * Quantity parsing.  Stands in for kube_quantity 0.6.1 `ParsedQuantity::try_from` (crate source absent;
 * Cargo.lock:788-797).  Grammar restated from the Kubernetes API conventions:
 *   quantity ::= sign? digits ('.' digits?)? suffix?        (or '.' digits)
 *   suffix   ::= Ki|Mi|Gi|Ti|Pi|Ei | n|u|m|k|M|G|T|P|E | (e|E) sign? digits
 * The value is returned as an exact integer count of 1/1000 units; anything finer is ORC_ERR_INEXACT.
 * On the exact domain (integer cores, integer millicores, integer bytes) this is the value any exact
 * decimal implementation (rust_decimal 1.30.0) holds, so +=, -= and <= on it match int64 arithmetic.
 * ---------------------------------------------------------------------------------------------- */
int orc_parse_quantity(const char* s, int64_t* out_milli) {
    if (!s || !out_milli) return ORC_ERR_PARSE;
    const char* p = s;
    int neg = 0;
    if (*p == '+' || *p == '-') {
        neg = (*p == '-');
        p++;
    }
    const i128 LIM = ((i128)1) << 100;
    i128 mant = 0;
    int ndig = 0, nfrac = 0, seen_dot = 0;
    for (;; p++) {
        if (*p >= '0' && *p <= '9') {
            if (mant >= LIM) return ORC_ERR_RANGE;
            mant = mant * 10 + (*p - '0');
            ndig++;
            if (seen_dot) nfrac++;
        } else if (*p == '.' && !seen_dot) {
            seen_dot = 1;
        } else {
            break;
        }
    }
    if (ndig == 0) return ORC_ERR_PARSE;
    int bin_shift = 0, dec_exp = 0;
    if (*p == 0) {
        /* no suffix */
    } else if (p[1] == 'i' && (p[0] == 'K' || p[0] == 'M' || p[0] == 'G' || p[0] == 'T' || p[0] == 'P' ||
                               p[0] == 'E')) {
        switch (p[0]) {
            case 'K': bin_shift = 10; break;
            case 'M': bin_shift = 20; break;
            case 'G': bin_shift = 30; break;
            case 'T': bin_shift = 40; break;
            case 'P': bin_shift = 50; break;
            default: bin_shift = 60; break;
        }
        p += 2;
    } else if ((*p == 'e' || *p == 'E') &&
               ((p[1] >= '0' && p[1] <= '9') || ((p[1] == '+' || p[1] == '-') && p[2] >= '0' && p[2] <= '9'))) {
        p++;
        int eneg = 0;
        if (*p == '+' || *p == '-') {
            eneg = (*p == '-');
            p++;
        }
        int e = 0;
        while (*p >= '0' && *p <= '9') {
            if (e > 1000) return ORC_ERR_RANGE;
            e = e * 10 + (*p - '0');
            p++;
        }
        dec_exp = eneg ? -e : e;
    } else {
        switch (*p) {
            case 'n': dec_exp = -9; break;
            case 'u': dec_exp = -6; break;
            case 'm': dec_exp = -3; break;
            case 'k': dec_exp = 3; break;
            case 'M': dec_exp = 6; break;
            case 'G': dec_exp = 9; break;
            case 'T': dec_exp = 12; break;
            case 'P': dec_exp = 15; break;
            case 'E': dec_exp = 18; break;
            default: return ORC_ERR_PARSE;
        }
        p++;
    }
    if (*p != 0) return ORC_ERR_PARSE;
    i128 v = mant;
    if (v != 0 && bin_shift) {
        if (v >= (LIM >> bin_shift)) return ORC_ERR_RANGE;
        v <<= bin_shift;
    }
    int e10 = dec_exp + 3 - nfrac; /* +3: result is in 1/1000 units */
    if (v != 0) {
        for (; e10 > 0; e10--) {
            if (v >= LIM) return ORC_ERR_RANGE;
            v *= 10;
        }
        for (; e10 < 0; e10++) {
            if (v % 10 != 0) return ORC_ERR_INEXACT;
            v /= 10;
        }
    }
    if (v > (i128)INT64_MAX) return ORC_ERR_RANGE;
    *out_milli = neg ? -(int64_t)v : (int64_t)v;
    return ORC_OK;
}

static const char* kv_get(const ks_kv* kv, uint32_t n, const char* key) {
    for (uint32_t i = 0; i < n; i++)
        if (kv[i].key && strcmp(kv[i].key, key) == 0) return kv[i].val;
    return NULL;
}

/* src/util.rs:54-75 — sum over spec.containers of resources.requests["cpu"|"memory"]; starts from
 * PodResources::new() = ("0","0") (src/util.rs:22-29). */
int orc_total_pod_resources(const ks_pod_obj* pod, int64_t out[2]) {
    int64_t cpu = 0, mem = 0; /* util.rs:55 */
    if (pod->has_spec) {      /* util.rs:57 */
        for (uint32_t c = 0; c < pod->n_containers; c++) { /* util.rs:58 */
            const ks_container_obj* ct = &pod->containers[c];
            if (!ct->has_requests) continue; /* util.rs:59-63 pattern does not match */
            const char* q = kv_get(ct->requests, ct->n_requests, "cpu"); /* util.rs:64 */
            if (q) {
                int64_t v;
                int rc = orc_parse_quantity(q, &v); /* util.rs:65 .expect */
                if (rc) return rc;
                if (__builtin_add_overflow(cpu, v, &cpu)) return ORC_ERR_RANGE;
            }
            q = kv_get(ct->requests, ct->n_requests, "memory"); /* util.rs:67 */
            if (q) {
                int64_t v;
                int rc = orc_parse_quantity(q, &v); /* util.rs:68 .expect */
                if (rc) return rc;
                if (__builtin_add_overflow(mem, v, &mem)) return ORC_ERR_RANGE;
            }
        }
    }
    out[0] = cpu;
    out[1] = mem;
    return ORC_OK;
}

/* src/util.rs:38-45 */
int orc_is_pod_bound(const ks_pod_obj* pod) { return pod->has_spec && pod->node_name != NULL; }

/* src/predicates.rs:45-61 */
int orc_does_node_selector_match(const ks_pod_obj* pod, const ks_node_obj* node) {
    int matches = 1;                                   /* :46 */
    if (pod->has_spec && pod->has_node_selector) {     /* :47 */
        for (uint32_t i = 0; i < pod->n_selector; i++) { /* :48 */
            if (node->has_labels) {                    /* :49 */
                const char* v = kv_get(node->labels, node->n_labels, pod->selector[i].key);
                if (v == NULL || strcmp(v, pod->selector[i].val) != 0) { /* :50 labels.get(pk) != Some(pv) */
                    matches = 0;
                    break;
                }
            } else { /* :54-57 */
                matches = 0;
                break;
            }
        }
    }
    return matches;
}

/* ------------------------------------------------------------------------------------------------
 * Cluster: node store + the pods a LIST with field selector spec.nodeName=<node> would return
 * (src/predicates.rs:21-25,34).  The index maps node name -> bound pods; evaluation still re-parses and
 * re-sums every bound pod per cell, as the reference does (src/predicates.rs:36-38).
 * ---------------------------------------------------------------------------------------------- */
struct orc_cluster {
    const ks_node_obj* nodes;
    uint32_t n_nodes;
    const ks_pod_obj* pods;
    uint64_t n_pods;
    uint64_t* bound_off; /* n_nodes+1 */
    uint64_t* bound_idx; /* indices into pods */
    uint32_t* htab;      /* open addressing: node idx+1 */
    uint32_t hcap;
};

static uint64_t str_hash(const char* s) {
    uint64_t h = 1469598103934665603ull;
    for (; *s; s++) h = (h ^ (unsigned char)*s) * 1099511628211ull;
    return h;
}

static int64_t cluster_find_node(const orc_cluster* c, const char* name) {
    if (!name || c->hcap == 0) return -1;
    uint64_t h = str_hash(name) & (c->hcap - 1);
    while (c->htab[h]) {
        uint32_t i = c->htab[h] - 1;
        if (c->nodes[i].name && strcmp(c->nodes[i].name, name) == 0) return i;
        h = (h + 1) & (c->hcap - 1);
    }
    return -1;
}

orc_cluster* orc_cluster_create(const ks_node_obj* nodes, uint32_t n_nodes, const ks_pod_obj* all_pods,
                                uint64_t n_all_pods) {
    orc_cluster* c = (orc_cluster*)calloc(1, sizeof(*c));
    if (!c) return NULL;
    c->nodes = nodes;
    c->n_nodes = n_nodes;
    c->pods = all_pods;
    c->n_pods = n_all_pods;
    uint32_t cap = 16;
    while (cap < 2 * (uint64_t)n_nodes + 1) cap <<= 1;
    c->hcap = cap;
    c->htab = (uint32_t*)calloc(cap, sizeof(uint32_t));
    c->bound_off = (uint64_t*)calloc((size_t)n_nodes + 2, sizeof(uint64_t));
    if (!c->htab || !c->bound_off) {
        orc_cluster_destroy(c);
        return NULL;
    }
    for (uint32_t i = 0; i < n_nodes; i++) {
        if (!nodes[i].name) continue;
        /* Node names are unique in the reference: reflector::Store<Node> is keyed by the object reference
         * (src/main.rs:133-139), so state() never holds two nodes of one name.  Input that repeats a name is outside
         * its domain; here (and in the product packer) pods are charged to the first node of the name. */
        if (cluster_find_node(c, nodes[i].name) >= 0) continue;
        uint64_t h = str_hash(nodes[i].name) & (cap - 1);
        while (c->htab[h]) h = (h + 1) & (cap - 1);
        c->htab[h] = i + 1;
    }
    int64_t* owner = (int64_t*)malloc(sizeof(int64_t) * (n_all_pods ? n_all_pods : 1));
    if (!owner) {
        orc_cluster_destroy(c);
        return NULL;
    }
    for (uint64_t p = 0; p < n_all_pods; p++) {
        owner[p] = orc_is_pod_bound(&all_pods[p]) ? cluster_find_node(c, all_pods[p].node_name) : -1;
        if (owner[p] >= 0) c->bound_off[owner[p] + 1]++;
    }
    for (uint32_t i = 0; i < n_nodes; i++) c->bound_off[i + 1] += c->bound_off[i];
    uint64_t total = c->bound_off[n_nodes];
    c->bound_idx = (uint64_t*)malloc(sizeof(uint64_t) * (total ? total : 1));
    uint64_t* cur = (uint64_t*)malloc(sizeof(uint64_t) * ((size_t)n_nodes + 1));
    if (!c->bound_idx || !cur) {
        free(owner);
        free(cur);
        orc_cluster_destroy(c);
        return NULL;
    }
    memcpy(cur, c->bound_off, sizeof(uint64_t) * n_nodes);
    for (uint64_t p = 0; p < n_all_pods; p++)
        if (owner[p] >= 0) c->bound_idx[cur[owner[p]]++] = p;
    free(owner);
    free(cur);
    return c;
}

void orc_cluster_destroy(orc_cluster* c) {
    if (!c) return;
    free(c->bound_off);
    free(c->bound_idx);
    free(c->htab);
    free(c);
}

/* allocatable as can_pod_fit reads it: (0,0) unless status.allocatable is Some (predicates.rs:27-32) */
static int node_allocatable(const ks_node_obj* node, int64_t out[2]) {
    out[0] = 0;
    out[1] = 0; /* :27 PodResources::new() */
    if (node->has_allocatable) { /* :28 */
        const char* q = kv_get(node->allocatable, node->n_allocatable, "cpu");
        if (!q) return ORC_ERR_MISSING; /* :29 index panic */
        int rc = orc_parse_quantity(q, &out[0]);
        if (rc) return rc;
        q = kv_get(node->allocatable, node->n_allocatable, "memory");
        if (!q) return ORC_ERR_MISSING; /* :30 */
        rc = orc_parse_quantity(q, &out[1]);
        if (rc) return rc;
    }
    return ORC_OK;
}

/* src/predicates.rs:27-38 */
int orc_node_available(const orc_cluster* c, uint32_t node_idx, int64_t out[2]) {
    if (node_idx >= c->n_nodes) return ORC_ERR_INVALID;
    int rc = node_allocatable(&c->nodes[node_idx], out);
    if (rc) return rc;
    for (uint64_t k = c->bound_off[node_idx]; k < c->bound_off[node_idx + 1]; k++) { /* :36 */
        int64_t r[2];
        rc = orc_total_pod_resources(&c->pods[c->bound_idx[k]], r); /* :37 */
        if (rc) return rc;
        if (__builtin_sub_overflow(out[0], r[0], &out[0])) return ORC_ERR_RANGE; /* util.rs:33 */
        if (__builtin_sub_overflow(out[1], r[1], &out[1])) return ORC_ERR_RANGE; /* util.rs:34 */
    }
    return ORC_OK;
}

/* src/predicates.rs:20-43 */
int orc_can_pod_fit(const orc_cluster* c, const ks_pod_obj* pod, uint32_t node_idx) {
    int64_t avail[2], req[2];
    int rc = orc_node_available(c, node_idx, avail);
    if (rc) return rc;
    rc = orc_total_pod_resources(pod, req); /* :40 */
    if (rc) return rc;
    return req[0] <= avail[0] && req[1] <= avail[1]; /* :42 */
}

/* src/predicates.rs:63-77 — fit first, then selector */
int orc_check_node_validity(const orc_cluster* c, const ks_pod_obj* pod, uint32_t node_idx) {
    int fit = orc_can_pod_fit(c, pod, node_idx); /* :68 */
    if (fit < 0) return fit;
    if (!fit) return ORC_CELL_NOT_ENOUGH_RESOURCES;                                              /* :69 */
    if (!orc_does_node_selector_match(pod, &c->nodes[node_idx])) return ORC_CELL_NODE_SELECTOR_MISMATCH; /* :72-73 */
    return ORC_CELL_OK;                                                                          /* :76 */
}

/* Score (spec extension; no reference lines).  Inputs in milli units; memory converted to bytes. */
static int score_from_milli(int policy, const int64_t avail[2], const int64_t alloc[2], const int64_t req[2],
                            int64_t* out) {
    if (policy == ORC_SCORE_LEFTOVER) {
        i128 dm = (i128)avail[1] - req[1];
        if (dm % 1000 != 0) return ORC_ERR_INEXACT;
        i128 s = ((i128)avail[0] - req[0]) * ((i128)1 << 22) + dm / 1000;
        if (s > INT64_MAX || s < INT64_MIN) return ORC_ERR_RANGE;
        *out = (int64_t)s;
        return ORC_OK;
    }
    if (policy == ORC_SCORE_LEAST_ALLOCATED) {
        i128 pc = 0, pm = 0;
        if (alloc[0] > 0) pc = (((i128)avail[0] - req[0]) * 100) / alloc[0];
        if (alloc[1] > 0) pm = (((i128)avail[1] - req[1]) * 100) / alloc[1];
        *out = (int64_t)((pc + pm) / 2);
        return ORC_OK;
    }
    return ORC_ERR_INVALID;
}

int orc_score_cell(const orc_cluster* c, int policy, const ks_pod_obj* pod, uint32_t node_idx, int64_t* out) {
    int64_t avail[2], alloc[2], req[2];
    int rc = orc_node_available(c, node_idx, avail);
    if (rc) return rc;
    rc = node_allocatable(&c->nodes[node_idx], alloc);
    if (rc) return rc;
    rc = orc_total_pod_resources(pod, req);
    if (rc) return rc;
    return score_from_milli(policy, avail, alloc, req, out);
}

int orc_online_cores(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

/* ---- threaded drivers ---- */
typedef struct {
    int tid, nthreads;
    int err;
    /* faithful */
    const orc_cluster* c;
    const ks_pod_obj* pods;
    /* packed */
    uint32_t n_nodes, W;
    const int64_t *free_cpu, *free_mem, *alloc_cpu, *alloc_mem, *req_cpu, *req_mem;
    const uint64_t *node_labels, *pod_sel;
    /* common */
    uint64_t n_pods;
    int policy;
    int32_t* out_node_idx;
    int64_t* out_score;
    uint32_t* out_cnt;
    uint8_t* out_mask;
    uint64_t mask_row_bytes;
    uint8_t* out_codes;
} job_t;

static void* faithful_worker(void* arg) {
    job_t* j = (job_t*)arg;
    const orc_cluster* c = j->c;
    uint32_t N = c->n_nodes;
    for (uint64_t p = j->tid; p < j->n_pods; p += j->nthreads) {
        const ks_pod_obj* pod = &j->pods[p];
        int32_t best = -1;
        int64_t best_score = 0;
        uint32_t cnt = 0;
        uint8_t* row = j->out_mask ? j->out_mask + p * j->mask_row_bytes : NULL;
        if (row) memset(row, 0, j->mask_row_bytes);
        for (uint32_t n = 0; n < N; n++) {
            int code = orc_check_node_validity(c, pod, n);
            if (code < 0) {
                j->err = code;
                return NULL;
            }
            if (j->out_codes) j->out_codes[p * N + n] = (uint8_t)code;
            if (code != ORC_CELL_OK) continue;
            cnt++;
            if (row) row[n >> 3] |= (uint8_t)(1u << (n & 7));
            int64_t s;
            int rc = orc_score_cell(c, j->policy, pod, n, &s);
            if (rc) {
                j->err = rc;
                return NULL;
            }
            if (best < 0 || s > best_score) { /* ties -> lowest node index */
                best = (int32_t)n;
                best_score = s;
            }
        }
        if (j->out_node_idx) j->out_node_idx[p] = best;
        if (j->out_score) j->out_score[p] = best < 0 ? 0 : best_score;
        if (j->out_cnt) j->out_cnt[p] = cnt;
    }
    return NULL;
}

static int run_jobs(job_t* proto, int nthreads, void* (*fn)(void*)) {
    if (nthreads <= 0) nthreads = orc_online_cores();
    if (nthreads > 256) nthreads = 256;
    if ((uint64_t)nthreads > proto->n_pods) nthreads = proto->n_pods ? (int)proto->n_pods : 1;
    job_t* jobs = (job_t*)malloc(sizeof(job_t) * nthreads);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
    if (!jobs || !th) {
        free(jobs);
        free(th);
        return ORC_ERR_INVALID;
    }
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = *proto;
        jobs[t].tid = t;
        jobs[t].nthreads = nthreads;
        jobs[t].err = 0;
    }
    for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, fn, &jobs[t]);
    fn(&jobs[0]);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
    int err = 0;
    for (int t = 0; t < nthreads; t++)
        if (jobs[t].err && !err) err = jobs[t].err;
    free(jobs);
    free(th);
    return err;
}

int orc_run_faithful(const orc_cluster* c, const ks_pod_obj* pods, uint64_t n_pods, int policy,
                     int32_t* out_node_idx, int64_t* out_score, uint32_t* out_cnt, uint8_t* out_mask,
                     uint64_t mask_row_bytes, uint8_t* out_codes, int nthreads) {
    if (!c || (!pods && n_pods)) return ORC_ERR_INVALID;
    if (out_mask && mask_row_bytes * 8 < c->n_nodes) return ORC_ERR_INVALID;
    job_t j;
    memset(&j, 0, sizeof(j));
    j.c = c;
    j.pods = pods;
    j.n_pods = n_pods;
    j.policy = policy;
    j.out_node_idx = out_node_idx;
    j.out_score = out_score;
    j.out_cnt = out_cnt;
    j.out_mask = out_mask;
    j.mask_row_bytes = mask_row_bytes;
    j.out_codes = out_codes;
    return run_jobs(&j, nthreads, faithful_worker);
}

/* K0 restated: free = alloc - sum of bound requests (src/predicates.rs:27-38, src/util.rs:31-36) */
int orc_free_reduce(uint32_t n_nodes, const int64_t* alloc_cpu, const int64_t* alloc_mem, uint64_t n_bound,
                    const int32_t* bound_node, const int64_t* bound_cpu, const int64_t* bound_mem,
                    int64_t* free_cpu, int64_t* free_mem) {
    for (uint32_t n = 0; n < n_nodes; n++) {
        free_cpu[n] = alloc_cpu[n];
        free_mem[n] = alloc_mem[n];
    }
    for (uint64_t b = 0; b < n_bound; b++) {
        int32_t n = bound_node[b];
        if (n < 0 || (uint32_t)n >= n_nodes) return ORC_ERR_INVALID;
        free_cpu[n] -= bound_cpu[b];
        free_mem[n] -= bound_mem[b];
    }
    return ORC_OK;
}

static void* packed_worker(void* arg) {
    job_t* j = (job_t*)arg;
    uint32_t N = j->n_nodes, W = j->W;
    for (uint64_t p = j->tid; p < j->n_pods; p += j->nthreads) {
        int64_t rc = j->req_cpu[p], rm = j->req_mem[p];
        const uint64_t* sel = j->pod_sel + p * W;
        int32_t best = -1;
        int64_t best_score = 0;
        uint32_t cnt = 0;
        uint8_t* row = j->out_mask ? j->out_mask + p * j->mask_row_bytes : NULL;
        if (row) memset(row, 0, j->mask_row_bytes);
        for (uint32_t n = 0; n < N; n++) {
            int fit = rc <= j->free_cpu[n] && rm <= j->free_mem[n]; /* predicates.rs:42 */
            uint64_t miss = 0;
            for (uint32_t w = 0; w < W; w++) miss |= sel[w] & ~j->node_labels[(uint64_t)n * W + w];
            int code = !fit ? ORC_CELL_NOT_ENOUGH_RESOURCES : (miss ? ORC_CELL_NODE_SELECTOR_MISMATCH : ORC_CELL_OK);
            if (j->out_codes) j->out_codes[p * N + n] = (uint8_t)code;
            if (code) continue;
            cnt++;
            if (row) row[n >> 3] |= (uint8_t)(1u << (n & 7));
            int64_t s;
            if (j->policy == ORC_SCORE_LEFTOVER) {
                s = (j->free_cpu[n] - rc) * ((int64_t)1 << 22) + (j->free_mem[n] - rm);
            } else {
                int64_t pc = j->alloc_cpu[n] > 0 ? (int64_t)((((i128)j->free_cpu[n] - rc) * 100) / j->alloc_cpu[n]) : 0;
                int64_t pm = j->alloc_mem[n] > 0 ? (int64_t)((((i128)j->free_mem[n] - rm) * 100) / j->alloc_mem[n]) : 0;
                s = (pc + pm) / 2;
            }
            if (best < 0 || s > best_score) {
                best = (int32_t)n;
                best_score = s;
            }
        }
        if (j->out_node_idx) j->out_node_idx[p] = best;
        if (j->out_score) j->out_score[p] = best < 0 ? 0 : best_score;
        if (j->out_cnt) j->out_cnt[p] = cnt;
    }
    return NULL;
}

int orc_run_packed(uint32_t n_nodes, uint32_t label_words, const int64_t* free_cpu, const int64_t* free_mem,
                   const int64_t* alloc_cpu, const int64_t* alloc_mem, const uint64_t* node_labels,
                   uint64_t n_pods, const int64_t* req_cpu, const int64_t* req_mem, const uint64_t* pod_sel,
                   int policy, int32_t* out_node_idx, int64_t* out_score, uint32_t* out_cnt,
                   uint8_t* out_mask, uint64_t mask_row_bytes, uint8_t* out_codes, int nthreads) {
    if (policy != ORC_SCORE_LEFTOVER && policy != ORC_SCORE_LEAST_ALLOCATED) return ORC_ERR_INVALID;
    if (out_mask && mask_row_bytes * 8 < n_nodes) return ORC_ERR_INVALID;
    job_t j;
    memset(&j, 0, sizeof(j));
    j.n_nodes = n_nodes;
    j.W = label_words;
    j.free_cpu = free_cpu;
    j.free_mem = free_mem;
    j.alloc_cpu = alloc_cpu;
    j.alloc_mem = alloc_mem;
    j.node_labels = node_labels;
    j.req_cpu = req_cpu;
    j.req_mem = req_mem;
    j.pod_sel = pod_sel;
    j.n_pods = n_pods;
    j.policy = policy;
    j.out_node_idx = out_node_idx;
    j.out_score = out_score;
    j.out_cnt = out_cnt;
    j.out_mask = out_mask;
    j.mask_row_bytes = mask_row_bytes;
    j.out_codes = out_codes;
    return run_jobs(&j, nthreads, packed_worker);
}

static uint64_t splitmix64(uint64_t* s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* src/main.rs:49-71 with a seeded generator: up to `attempts` draws with replacement; an empty node store
 * wastes the attempt (:56,60); the first draw passing check_node_validity wins (:61-66). */
int32_t orc_select_sampling(const orc_cluster* c, const ks_pod_obj* pod, uint32_t attempts, uint64_t* rng_state,
                            uint32_t* cells_evaluated) {
    uint32_t cells = 0;
    int32_t node = -1;
    for (uint32_t a = 0; a < attempts; a++) { /* :53 */
        if (c->n_nodes == 0) continue;        /* :56 choose() on empty -> None */
        uint32_t cand = (uint32_t)(splitmix64(rng_state) % c->n_nodes);
        cells++;
        if (orc_check_node_validity(c, pod, cand) == ORC_CELL_OK) { /* :61 */
            node = (int32_t)cand;                                  /* :64 */
            break;
        }
    }
    if (cells_evaluated) *cells_evaluated = cells;
    return node;
}

/* Same policy on the packed form, one generator state per pod (rng_state[p] is advanced in place). */
int orc_select_sampling_packed(uint32_t n_nodes, uint32_t W, const int64_t* free_cpu, const int64_t* free_mem,
                               const uint64_t* node_labels, uint64_t n_pods, const int64_t* req_cpu,
                               const int64_t* req_mem, const uint64_t* pod_sel, uint32_t attempts, uint64_t* rng_state,
                               int32_t* out_node_idx, uint32_t* out_cells, int32_t* out_draw_node,
                               uint8_t* out_draw_code) {
    for (uint64_t p = 0; p < n_pods; p++) {
        const uint64_t* sel = pod_sel + p * W;
        int32_t node = -1;
        uint32_t cells = 0;
        for (uint32_t a = 0; a < attempts; a++) {
            if (out_draw_node) out_draw_node[p * attempts + a] = -1;
            if (out_draw_code) out_draw_code[p * attempts + a] = 0xff;
        }
        for (uint32_t a = 0; a < attempts && node < 0; a++) { /* main.rs:53 */
            if (n_nodes == 0) continue;                         /* :56 */
            uint32_t n = (uint32_t)(splitmix64(&rng_state[p]) % n_nodes);
            int fit = req_cpu[p] <= free_cpu[n] && req_mem[p] <= free_mem[n]; /* predicates.rs:42 */
            uint64_t miss = 0;
            for (uint32_t w = 0; w < W; w++) miss |= sel[w] & ~node_labels[(uint64_t)n * W + w];
            int code = !fit ? ORC_CELL_NOT_ENOUGH_RESOURCES : (miss ? ORC_CELL_NODE_SELECTOR_MISMATCH : ORC_CELL_OK);
            if (out_draw_node) out_draw_node[p * attempts + cells] = (int32_t)n;
            if (out_draw_code) out_draw_code[p * attempts + cells] = (uint8_t)code;
            cells++;
            if (code == ORC_CELL_OK) node = (int32_t)n; /* main.rs:63-65 */
        }
        out_node_idx[p] = node;
        if (out_cells) out_cells[p] = cells;
    }
    return ORC_OK;
}

/* ---- streaming restated (see oracle.h) ---- */
int orc_commit_claims(uint32_t n_nodes, int64_t* free_cpu, int64_t* free_mem, uint64_t n_claims,
                      const int32_t* claim_node, const int64_t* req_cpu, const int64_t* req_mem, uint8_t* accepted) {
    for (uint64_t i = 0; i < n_claims; i++) {
        int32_t n = claim_node[i];
        accepted[i] = 0;
        if (n < 0 || (uint32_t)n >= n_nodes) continue;
        if (req_cpu[i] <= free_cpu[n] && req_mem[i] <= free_mem[n]) { /* predicates.rs:42 on what is left */
            accepted[i] = 1;
            free_cpu[n] -= req_cpu[i]; /* util.rs:33 */
            free_mem[n] -= req_mem[i]; /* util.rs:34 */
        }
    }
    return ORC_OK;
}

int orc_stream_bind_packed(uint32_t n_nodes, uint32_t W, int64_t* free_cpu, int64_t* free_mem,
                           const int64_t* alloc_cpu, const int64_t* alloc_mem, const uint64_t* node_labels,
                           uint64_t n_pods, const int64_t* req_cpu, const int64_t* req_mem, const uint64_t* pod_sel,
                           int policy, int32_t* out_node_idx, int64_t* out_score, uint32_t* out_rounds) {
    uint64_t* pending = (uint64_t*)malloc(sizeof(uint64_t) * (n_pods ? n_pods : 1));
    int64_t* rc = (int64_t*)malloc(sizeof(int64_t) * (n_pods ? n_pods : 1));
    int64_t* rm = (int64_t*)malloc(sizeof(int64_t) * (n_pods ? n_pods : 1));
    int64_t* sc = (int64_t*)malloc(sizeof(int64_t) * (n_pods ? n_pods : 1));
    int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * (n_pods ? n_pods : 1));
    uint8_t* acc = (uint8_t*)malloc(n_pods ? n_pods : 1);
    uint64_t* sel = (uint64_t*)malloc(sizeof(uint64_t) * (n_pods ? n_pods * W : 1));
    if (!pending || !rc || !rm || !sc || !idx || !acc || !sel) return ORC_ERR_INVALID;
    uint64_t m = n_pods;
    for (uint64_t i = 0; i < n_pods; i++) {
        pending[i] = i;
        out_node_idx[i] = -1;
        if (out_score) out_score[i] = 0;
    }
    uint32_t rounds = 0;
    int err = ORC_OK;
    while (m > 0 && rounds <= n_pods + 1) {
        for (uint64_t k = 0; k < m; k++) {
            rc[k] = req_cpu[pending[k]];
            rm[k] = req_mem[pending[k]];
            for (uint32_t w = 0; w < W; w++) sel[k * W + w] = pod_sel[pending[k] * W + w];
        }
        err = orc_run_packed(n_nodes, W, free_cpu, free_mem, alloc_cpu, alloc_mem, node_labels, m, rc, rm, sel, policy,
                             idx, sc, NULL, NULL, 0, NULL, 1);
        if (err) break;
        orc_commit_claims(n_nodes, free_cpu, free_mem, m, idx, rc, rm, acc);
        uint64_t next = 0;
        for (uint64_t k = 0; k < m; k++) {
            if (idx[k] < 0) continue;
            if (acc[k]) {
                out_node_idx[pending[k]] = idx[k];
                if (out_score) out_score[pending[k]] = sc[k];
            } else {
                pending[next++] = pending[k];
            }
        }
        m = next;
        rounds++;
    }
    if (out_rounds) *out_rounds = rounds;
    free(pending); free(rc); free(rm); free(sc); free(idx); free(acc); free(sel);
    return err;
}
