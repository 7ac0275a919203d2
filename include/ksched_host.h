/* ksched_host.h — host layer of libksched.so: the reference's scheduling surface over Pod/Node OBJECTS.
 *
 * The reference's host is Rust (async fns inside a binary crate).  No Rust toolchain exists in this image,
 * so this layer is C++ behind a C ABI that keeps the reference's names, argument meaning and error
 * behaviour; a Rust controller would bind it with the `extern "C"` block shown in INTEGRATION.md and keep
 * its reconcile() loop.  Mapping (paths relative to /root/reference):
 *   ksh_total_pod_resources   <- total_pod_resources            src/util.rs:54-75
 *   ksh_is_pod_bound          <- is_pod_bound                   src/util.rs:38-45
 *   ksh_context               <- Context{client,node_store}     src/util.rs:12-15 (+ the LIST of src/predicates.rs:21-38)
 *   ksh_check_node_validity   <- check_node_validity            src/predicates.rs:63-77
 *   ksh_select_nodes          <- select_node_for_pod, batched   src/main.rs:51-71 (argmax score instead of 5 random draws)
 *   ksh_select_node_for_pod   <- select_node_for_pod, seeded    src/main.rs:49-71 (the reference's own 5-draw policy)
 *   ksh_reconcile_batch       <- reconcile over a drained queue  src/main.rs:73-120 (micro-batch with capacity commit)
 *   ksh_reconcile             <- reconcile                      src/main.rs:73-120 (builds the Binding, does not POST it)
 *   KSH_RECONCILE_*           <- ReconcileError                 src/error.rs:5-15
 * The layer packs objects into the SoA/bitmask form of ksched.h and calls the CUDA core; it never evaluates
 * a predicate on the CPU.  Malformed quantities / missing allocatable keys (reference: panic) return
 * KS_ERR_PARSE / KS_ERR_MISSING at pack time.
 */
#ifndef KSCHED_HOST_H
#define KSCHED_HOST_H

#include <stddef.h>
#include <stdint.h>

#include "ks_objects.h"
#include "ksched.h"

#ifdef __cplusplus
extern "C" {
#endif

/* reconcile outcomes: Ok(Action::await_change()) or ReconcileError (src/error.rs:5-15) */
#define KSH_RECONCILE_OK 0                   /* pod already bound (src/main.rs:74-76) or binding produced */
#define KSH_RECONCILE_NO_NODE_FOUND 1        /* ReconcileError::NoNodeFound (src/main.rs:116-118) */
#define KSH_RECONCILE_BINDING_OBJECT_FAILED 2 /* ReconcileError::CreateBindingObjectFailed (src/main.rs:111-114);
                                                 also a pod without namespace (reference: unwrap panic, :80) */
#define KSH_RECONCILE_BINDING_FAILED 3       /* ReconcileError::CreateBindingFailed: transport, caller side (:105-108) */

/* Kubernetes quantity strings, exact: cpu -> millicores, memory -> bytes.
 * KS_ERR_PARSE (malformed), KS_ERR_INEXACT (finer than 1m / 1 byte), KS_ERR_RANGE. */
int ksh_parse_cpu_millicores(const char* quantity, int64_t* out);
int ksh_parse_memory_bytes(const char* quantity, int64_t* out);

int ksh_total_pod_resources(const ks_pod_obj* pod, int64_t* cpu_millicores, int64_t* mem_bytes);
int ksh_is_pod_bound(const ks_pod_obj* pod);

typedef struct ksh_context ksh_context;
/* device = CUDA ordinal, or KSH_DEVICE_NONE for a packing-only context: objects -> SoA/bitmask arrays
 * (ksh_context_set_* / upsert / pod events, ksh_pack_pods, ksh_context_export_packed) work on the host; every call
 * that evaluates a predicate or selects a node returns KS_ERR_NO_DEVICE — nothing is ever computed on the CPU. */
#define KSH_DEVICE_NONE (-1)
int ksh_context_create(int device, ksh_context** out);
void ksh_context_destroy(ksh_context* ctx);
/* node store contents (what reflector::Store<Node>::state() returns, src/main.rs:56).  The store is keyed by name, so
 * names are unique; if a name repeats anyway, pods and events address the first node of that name. */
int ksh_context_set_nodes(ksh_context* ctx, const ks_node_obj* nodes, uint32_t n_nodes);
/* every pod object the API server holds; those with spec.nodeName naming a known node are the LIST results
 * of src/predicates.rs:22-34 (any phase) and are charged to that node; the rest are ignored here. */
int ksh_context_set_cluster_pods(ksh_context* ctx, const ks_pod_obj* pods, uint64_t n_pods);
/* Incremental node-store / informer events (what the reflector applies between reconciles, src/main.rs:133-139):
 * - upsert: add a node or replace the one with the same metadata.name; *out_idx = its index
 * - remove: delete by name; later nodes move down by one index (indices are positions in the current store;
 *   ksh_context_node_name maps an index back to the name a Binding needs)
 * - pod_bound / pod_deleted: a pod with spec.nodeName appeared / went away; capacity is charged / returned.
 *   Pods are identified by namespace/name; unknown pods or nodes are ignored (KS_OK), like a LIST that no
 *   longer shows them.
 * The device snapshot is refreshed lazily on the next select/check call. */
int ksh_context_upsert_node(ksh_context* ctx, const ks_node_obj* node, uint32_t* out_idx);
int ksh_context_remove_node(ksh_context* ctx, const char* name);
int ksh_context_pod_bound(ksh_context* ctx, const ks_pod_obj* pod);
int ksh_context_pod_deleted(ksh_context* ctx, const ks_pod_obj* pod);
const char* ksh_context_node_name(const ksh_context* ctx, uint32_t node_idx); /* NULL if out of range; borrowed,
                                                                                 valid until the next ksh_context_* mutation */
uint32_t ksh_context_num_nodes(const ksh_context* ctx);
uint32_t ksh_context_label_words(const ksh_context* ctx);
uint64_t ksh_context_num_bound(const ksh_context* ctx);
/* The packed node side exactly as the next device upload sends it (arguments of ks_snapshot_set_nodes /
 * ks_snapshot_set_bound in ksched.h): allocatable [N] in millicores / bytes, label words [N * label_words] under the
 * current dictionary, bound-pod triples [num_bound].  Lets a host feed the packed core itself, and the CPU tests
 * check the packer against the oracle without a GPU. */
int ksh_context_export_packed(const ksh_context* ctx, int64_t* alloc_cpu, int64_t* alloc_mem, uint64_t* labels,
                              int32_t* bound_node, int64_t* bound_cpu, int64_t* bound_mem);
ks_snapshot* ksh_context_snapshot(ksh_context* ctx); /* borrowed; valid until the next ksh_context_* mutation */

/* Pack P pod objects into caller buffers (req_cpu[P], req_mem[P], sel[P*W]) with the context's current label
 * dictionary, growing the dictionary (and re-uploading node columns) when a selector names a new pair.
 * Returns W (>0) on success so the caller can size `sel` = P * 8 words up front (KS_MAX_LABEL_WORDS). */
int ksh_pack_pods(ksh_context* ctx, const ks_pod_obj* pods, uint64_t n_pods, int64_t* req_cpu, int64_t* req_mem,
                  uint64_t* sel, uint32_t sel_stride_words);

int ksh_check_node_validity(ksh_context* ctx, const ks_pod_obj* pod, uint32_t node_idx);
/* select_node_for_pod for a whole batch (src/main.rs:51-71 with the argmax over ALL feasible nodes instead of <= 5 draws):
 * out_node_idx[p] < 0 = None.  A pod whose quantities do not parse fails the call before anything is evaluated (status +
 * ks_last_error(), the reference panics there: src/util.rs:65,68).  The context keeps the packed form of the last batch
 * (24 + 8W bytes per pod) until it is destroyed, so that repeated calls do not fault in fresh pages. */
int ksh_select_nodes(ksh_context* ctx, const ks_pod_obj* pods, uint64_t n_pods, int policy, int32_t* out_node_idx,
                     int64_t* out_score, uint32_t* out_feasible_cnt);

/* select_node_for_pod with the reference's OWN policy (src/main.rs:49-71): <= attempts seeded random draws per pod,
 * first valid draw wins, -1 = None.  Arguments as ks_select_sampling (ksched.h); pass KS_REFERENCE_ATTEMPTS for the
 * reference's ATTEMPTS.  out_draw_code holds the InvalidNodeReason of each failed draw (what src/main.rs:62 logs). */
int ksh_select_node_for_pod(ksh_context* ctx, const ks_pod_obj* pods, uint64_t n_pods, uint32_t attempts, uint64_t seed,
                            uint64_t first_pod_index, int32_t* out_node_idx, uint32_t* out_attempts,
                            int32_t* out_draw_node, uint8_t* out_draw_code);

/* One reconcile: skip bound pods, select, emit `POST /api/v1/namespaces/{ns}/pods/{name}/binding` body, and
 * charge the pod to the chosen node in the snapshot (what the next LIST would show).  *node_idx = -1 if none.
 * binding_json may be NULL; otherwise receives a NUL-terminated JSON document (truncated to cap). */
int ksh_reconcile(ksh_context* ctx, const ks_pod_obj* pod, int policy, int32_t* node_idx, char* binding_json,
                  size_t cap);

/* reconcile() for a drained controller queue, pods in arrival order.  Pods already bound are skipped (KSH_RECONCILE_OK,
 * node -1; src/main.rs:74-76); pods without namespace/name get KSH_RECONCILE_BINDING_OBJECT_FAILED (reference: unwrap
 * panic, :80) and are not scheduled; the rest are packed once and bound by the micro-batch loop of ks_stream_bind (ksched.h):
 * every round selects on the current capacity, claims are accepted per node in arrival order while they still fit, losers
 * are re-selected against what is left — capacity is never oversubscribed, a pod ends with KSH_RECONCILE_NO_NODE_FOUND
 * only if no node fits it at its turn.  Accepted binds are committed to the snapshot and to the context's bound-pod list.
 * out_status[n], out_node_idx[n] (-1 = none).  Binding bodies (see ksh_reconcile) are written back to back, NUL-terminated,
 * into json (cap bytes); out_json_off[i] = offset of pod i's body or -1 (not bound, or the buffer was full).  json /
 * out_json_off / out_rounds may be NULL. */
int ksh_reconcile_batch(ksh_context* ctx, const ks_pod_obj* pods, uint64_t n_pods, int policy, int32_t* out_status,
                        int32_t* out_node_idx, char* json, size_t cap, int64_t* out_json_off, uint32_t* out_rounds);

#ifdef __cplusplus
}
#endif
#endif /* KSCHED_HOST_H */
