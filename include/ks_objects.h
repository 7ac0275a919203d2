/* ks_objects.h — borrowed, read-only views of the Kubernetes objects the hot path reads.
 *
 * These plain-C structs stand in for the k8s-openapi types the reference borrows in
 *   check_node_validity(pod: &corev1::Pod, node: &corev1::Node, ctx: &Context)
 *   (/root/reference/src/predicates.rs:63-67).
 * Only the fields the hot path touches exist; "Option<T> is None" is modelled by the has_* flags so
 * that every None/Some/empty distinction the reference branches on survives:
 *   pod.spec                         -> has_spec                (src/util.rs:39,57; src/predicates.rs:47)
 *   pod.spec.node_name               -> node_name (NULL = None) (src/util.rs:40)
 *   pod.spec.containers[]            -> containers[]            (src/util.rs:58)
 *   container.resources.requests     -> has_requests + requests (src/util.rs:59-62; limits are ignored there)
 *   pod.spec.node_selector           -> has_node_selector       (src/predicates.rs:47)
 *   node.metadata.labels             -> has_labels              (src/predicates.rs:49,54)
 *   node.status.allocatable          -> has_allocatable         (src/predicates.rs:28)
 * Quantities stay strings (k8s_openapi Quantity(pub String)); parsing is the consumer's job.
 * The same structs are consumed by the product host layer (ksh_* in ksched_host.h) and by the CPU
 * oracle (oracle/oracle.h), so both sides see "the same synthetic Pod/Node objects".
 */
#ifndef KS_OBJECTS_H
#define KS_OBJECTS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ks_kv {
    const char* key;
    const char* val;
} ks_kv;

typedef struct ks_container_obj {
    int32_t has_requests; /* resources: Some(ResourceRequirements{requests: Some(..)}) */
    uint32_t n_requests;
    const ks_kv* requests; /* e.g. {"cpu","250m"}, {"memory","268435456"} */
} ks_container_obj;

typedef struct ks_pod_obj {
    const char* ns;   /* metadata.namespace (NULL = None) */
    const char* name; /* metadata.name */
    int32_t has_spec;
    const char* node_name; /* spec.nodeName; NULL = None (pod is unbound) */
    uint32_t n_containers;
    const ks_container_obj* containers;
    int32_t has_node_selector;
    uint32_t n_selector;
    const ks_kv* selector;
    const char* metadata_json; /* optional: the pod's whole ObjectMeta, pre-serialised by the caller as one JSON object
                                  ("{...}").  The reference sends `metadata: pod.metadata.clone()` in the Binding
                                  (src/main.rs:88-91); when set, ksh_reconcile emits it verbatim, else {name, namespace}. */
} ks_pod_obj;

typedef struct ks_node_obj {
    const char* name; /* metadata.name */
    int32_t has_labels;
    uint32_t n_labels;
    const ks_kv* labels;
    int32_t has_allocatable; /* status: Some(NodeStatus{allocatable: Some(..)}) */
    uint32_t n_allocatable;
    const ks_kv* allocatable; /* must hold "cpu" and "memory" when present (reference panics otherwise) */
} ks_node_obj;

#ifdef __cplusplus
}
#endif
#endif /* KS_OBJECTS_H */
