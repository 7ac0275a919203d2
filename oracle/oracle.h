/* oracle.h — CPU restatement of the reference's scheduling hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing in the product (kube-scheduler-rs-reference_b200/, include/ksched*.h) may include, link,
 * load or execute this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, as the checker / the timed CPU baseline.
 *
 * What it restates (file:line are relative to /root/reference):
 *   orc_parse_quantity        kube_quantity 0.6.1 `TryFrom<&Quantity> for ParsedQuantity`
 *                             (third-party crate, Cargo.lock:788-797, source NOT in /root/reference;
 *                             call sites src/util.rs:25-26,65,68 and src/predicates.rs:29-31).
 *                             Restated from the published Kubernetes quantity grammar.
 *   orc_total_pod_resources   src/util.rs:54-75
 *   orc_can_pod_fit           src/predicates.rs:20-43   (the LIST at :34 is an in-memory index)
 *   orc_does_node_selector_match  src/predicates.rs:45-61
 *   orc_check_node_validity   src/predicates.rs:63-77
 *   orc_is_pod_bound          src/util.rs:38-45
 *   orc_select_*              replaces src/main.rs:51-71 by "argmax score over the feasible set"
 *                             (spec extension: the reference samples <=5 random nodes, has no score).
 *
 * PARITY STATUS
 *   does_node_selector_match : PINNED by the reference's three tests (src/predicates/test.rs:42-58),
 *                              reproduced in tests/test_oracle_golden.py.
 *   resource fit / quantity  : PARITY UNPINNED.  The reference has no test for can_pod_fit,
 *                              total_pod_resources or PodResources, cannot be compiled here (no rustc),
 *                              and the arithmetic lives in the absent kube_quantity crate.  Bit-exactness
 *                              is claimed only on the exact domain (integer cores / integer millicores
 *                              for cpu, plain integer bytes for memory) where any exact decimal
 *                              implementation must agree with int64 arithmetic.  Golden vector GV-1
 *                              (tests/golden/gv1.json) is hand-derived from the reference *code*.
 *   score / argmax           : no reference definition exists; the oracle is the definition.
 */
#ifndef ORACLE_H
#define ORACLE_H

#include <stdint.h>
#include "../include/ks_objects.h"

#ifdef __cplusplus
extern "C" {
#endif

/* status codes */
#define ORC_OK 0
#define ORC_ERR_PARSE (-3)   /* reference: .expect("invalid ... spec") panic, src/util.rs:65,68; predicates.rs:29,31 */
#define ORC_ERR_INEXACT (-6) /* finer than 1/1000 unit: outside the representable exact domain */
#define ORC_ERR_RANGE (-5)
#define ORC_ERR_MISSING (-7) /* allocatable lacks "cpu"/"memory": reference BTreeMap index panic, predicates.rs:29-30 */
#define ORC_ERR_INVALID (-1)

/* cell codes: Ok(()) / Err(InvalidNodeReason::*)   src/predicates.rs:14-18,63-77 */
#define ORC_CELL_OK 0
#define ORC_CELL_NOT_ENOUGH_RESOURCES 1
#define ORC_CELL_NODE_SELECTOR_MISMATCH 2

/* score policies (spec extension, see DESIGN.md "Score") */
#define ORC_SCORE_LEFTOVER 0        /* (free_cpu-req_cpu)[millicores]*2^22 + (free_mem-req_mem)[bytes] */
#define ORC_SCORE_LEAST_ALLOCATED 1 /* upstream NodeResourcesFit/LeastAllocated, integer floor, 0..100 */

/* Quantity string -> exact count of 1/1000 units (cpu: millicores, memory: milli-bytes). */
int orc_parse_quantity(const char* s, int64_t* out_milli);

/* src/util.rs:54-75.  out[0] = cpu millicores, out[1] = memory in milli-bytes. */
int orc_total_pod_resources(const ks_pod_obj* pod, int64_t out[2]);
int orc_is_pod_bound(const ks_pod_obj* pod);
int orc_does_node_selector_match(const ks_pod_obj* pod, const ks_node_obj* node);

/* The cluster = node store (src/main.rs:133-139) + every pod object the API server would LIST. */
typedef struct orc_cluster orc_cluster;
orc_cluster* orc_cluster_create(const ks_node_obj* nodes, uint32_t n_nodes, const ks_pod_obj* all_pods,
                                uint64_t n_all_pods);
void orc_cluster_destroy(orc_cluster*);

/* src/predicates.rs:20-43: returns 1 fits / 0 does not / <0 error (reference would panic). */
int orc_can_pod_fit(const orc_cluster*, const ks_pod_obj* pod, uint32_t node_idx);
/* src/predicates.rs:63-77: ORC_CELL_* or <0. */
int orc_check_node_validity(const orc_cluster*, const ks_pod_obj* pod, uint32_t node_idx);
/* free resources of one node exactly as can_pod_fit derives them (:27-38); milli units. */
int orc_node_available(const orc_cluster*, uint32_t node_idx, int64_t out[2]);
/* score of one (pod,node) cell; defined for feasible cells. */
int orc_score_cell(const orc_cluster*, int policy, const ks_pod_obj* pod, uint32_t node_idx, int64_t* out);

/* Faithful batched evaluation: every cell goes through orc_check_node_validity (string parse and
 * bound-pod re-summation per cell, like the reference).  Outputs may be NULL.
 *   out_codes : P*N bytes of ORC_CELL_*           out_mask : P rows of mask_row_bytes, bit n%8 of byte n/8
 *   out_node_idx : argmax-score node, ties -> lowest index, -1 if none      out_cnt : feasible count
 * nthreads <= 0 -> all online cores.  Returns ORC_OK or the first error. */
int orc_run_faithful(const orc_cluster*, const ks_pod_obj* pods, uint64_t n_pods, int policy,
                     int32_t* out_node_idx, int64_t* out_score, uint32_t* out_cnt, uint8_t* out_mask,
                     uint64_t mask_row_bytes, uint8_t* out_codes, int nthreads);

/* Packed flavour: the same answer from the SoA int64 + label-bitmask form the GPU consumes
 * (cpu in millicores, memory in BYTES).  free_* already include the bound-pod subtraction. */
int orc_free_reduce(uint32_t n_nodes, const int64_t* alloc_cpu, const int64_t* alloc_mem, uint64_t n_bound,
                    const int32_t* bound_node, const int64_t* bound_cpu, const int64_t* bound_mem,
                    int64_t* free_cpu, int64_t* free_mem);
int orc_run_packed(uint32_t n_nodes, uint32_t label_words, const int64_t* free_cpu, const int64_t* free_mem,
                   const int64_t* alloc_cpu, const int64_t* alloc_mem, const uint64_t* node_labels,
                   uint64_t n_pods, const int64_t* req_cpu, const int64_t* req_mem, const uint64_t* pod_sel,
                   int policy, int32_t* out_node_idx, int64_t* out_score, uint32_t* out_cnt,
                   uint8_t* out_mask, uint64_t mask_row_bytes, uint8_t* out_codes, int nthreads);

/* Reference-policy model: <=attempts uniform draws with replacement, first valid wins
 * (src/main.rs:49-71) driven by a splitmix64 stream instead of thread_rng.  Returns node idx or -1. */
int32_t orc_select_sampling(const orc_cluster*, const ks_pod_obj* pod, uint32_t attempts, uint64_t* rng_state,
                            uint32_t* cells_evaluated);

int orc_select_sampling_packed(uint32_t n_nodes, uint32_t label_words, const int64_t* free_cpu, const int64_t* free_mem,
                               const uint64_t* node_labels, uint64_t n_pods, const int64_t* req_cpu,
                               const int64_t* req_mem, const uint64_t* pod_sel, uint32_t attempts,
                               uint64_t* rng_state /* [n_pods], advanced in place */, int32_t* out_node_idx,
                               uint32_t* out_cells, int32_t* out_draw_node, uint8_t* out_draw_code);

/* Streaming (config C5) restated: what the reference gets from re-LISTing per cell (src/predicates.rs:34-38).
 * orc_commit_claims: claims in arrival (array) order; claim i is accepted iff its request still fits the
 * node's remaining free (predicates.rs:42), then free -= request (util.rs:31-36).  claim_node < 0 = no claim.
 * orc_stream_bind_packed: repeat {select for pending pods; commit; losers stay pending} until none pending. */
int orc_commit_claims(uint32_t n_nodes, int64_t* free_cpu, int64_t* free_mem, uint64_t n_claims,
                      const int32_t* claim_node, const int64_t* req_cpu, const int64_t* req_mem, uint8_t* accepted);
int orc_stream_bind_packed(uint32_t n_nodes, uint32_t label_words, int64_t* free_cpu, int64_t* free_mem,
                           const int64_t* alloc_cpu, const int64_t* alloc_mem, const uint64_t* node_labels,
                           uint64_t n_pods, const int64_t* req_cpu, const int64_t* req_mem, const uint64_t* pod_sel,
                           int policy, int32_t* out_node_idx, int64_t* out_score, uint32_t* out_rounds);

int orc_online_cores(void);

#ifdef __cplusplus
}
#endif
#endif
