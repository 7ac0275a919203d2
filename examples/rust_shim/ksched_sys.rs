// ksched_sys.rs — `extern "C"` binding of libksched.so for the reference's Rust host.
//
// NOT COMPILED IN THIS REPOSITORY: the build image has no rustc/cargo and the reference's 229 crates are not
// vendored.  Hand-written from include/ksched.h, include/ksched_host.h and include/ks_objects.h (bindgen over those
// headers produces the same items); tests/test_abi.py keeps this file honest by checking that every function declared
// here is exported by libksched.so, and the same ABI is driven from C by examples/reconcile_loop.c and from Python by
// kube-scheduler-rs-reference_b200/_capi.py + host.py.
//
// Drop into the reference as src/ksched_sys.rs (`mod ksched_sys;` in src/main.rs) together with src/ksched.rs
// (the safe wrapper next to this file).  build.rs:
//     println!("cargo:rustc-link-search=native=/opt/ksched/lib");
//     println!("cargo:rustc-link-lib=dylib=ksched");
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_int, c_void};

pub const KS_OK: c_int = 0;
pub const KS_SCORE_LEFTOVER: c_int = 0;
pub const KS_SCORE_LEAST_ALLOCATED: c_int = 1;
pub const KS_MEM_HOST: i32 = 0;
pub const KS_MEM_DEVICE: i32 = 1;
pub const KS_CELL_OK: c_int = 0;
pub const KS_CELL_NOT_ENOUGH_RESOURCES: c_int = 1; // InvalidNodeReason::NotEnoughResources  (src/predicates.rs:16)
pub const KS_CELL_NODE_SELECTOR_MISMATCH: c_int = 2; // InvalidNodeReason::NodeSelectorMismatch (src/predicates.rs:17)
pub const KSH_RECONCILE_OK: c_int = 0;
pub const KSH_RECONCILE_NO_NODE_FOUND: c_int = 1; // ReconcileError::NoNodeFound (src/error.rs)
pub const KSH_RECONCILE_BINDING_OBJECT_FAILED: c_int = 2; // ReconcileError::CreateBindingObjectFailed
pub const KS_REFERENCE_ATTEMPTS: u32 = 5; // ATTEMPTS, src/main.rs:49
pub const KS_MAX_PEERS: usize = 15;

#[repr(C)]
pub struct ks_kv {
    pub key: *const c_char,
    pub val: *const c_char,
}
#[repr(C)]
pub struct ks_container_obj {
    pub has_requests: i32,
    pub n_requests: u32,
    pub requests: *const ks_kv,
}
#[repr(C)]
pub struct ks_pod_obj {
    pub ns: *const c_char,
    pub name: *const c_char,
    pub has_spec: i32,
    pub node_name: *const c_char,
    pub n_containers: u32,
    pub containers: *const ks_container_obj,
    pub has_node_selector: i32,
    pub n_selector: u32,
    pub selector: *const ks_kv,
    pub metadata_json: *const c_char, // serde_json::to_string(&pod.metadata): emitted verbatim in the Binding (src/main.rs:88)
}
#[repr(C)]
pub struct ks_node_obj {
    pub name: *const c_char,
    pub has_labels: i32,
    pub n_labels: u32,
    pub labels: *const ks_kv,
    pub has_allocatable: i32,
    pub n_allocatable: u32,
    pub allocatable: *const ks_kv,
}
#[repr(C)]
pub struct ks_pods {
    pub n: u64,
    pub req_cpu: *const i64,
    pub req_mem: *const i64,
    pub sel: *const u64,
    pub mem_space: i32,
}
#[repr(C)]
pub struct ks_exchange {
    pub world: u32,
    pub rank: u32,
    pub n_peers: u32,
    pub peer_node_idx: [*mut i32; KS_MAX_PEERS],
    pub peer_score: [*mut i64; KS_MAX_PEERS],
    pub peer_flag: [*mut u32; KS_MAX_PEERS],
    pub local_flags: *mut u32,
    pub local_state: *mut u32,
}
#[repr(C)]
pub struct ks_bindings {
    pub node_idx: *mut i32,
    pub score: *mut i64,
    pub feasible_cnt: *mut u32,
    pub mem_space: i32,
    pub mask: *mut u8,
    pub mask_row_bytes: u64,
    pub mask_space: i32,
    pub bindings_ready_event: *mut c_void, // optional cudaEvent_t
    pub exchange: *const ks_exchange,      // optional fused all-gather over NVLink
}
pub enum ks_snapshot {}
pub enum ks_stream {}
pub enum ksh_context {}

extern "C" {
    pub fn ks_last_error() -> *const c_char;
    pub fn ks_version() -> c_int;
    pub fn ks_device_count() -> c_int;
    pub fn ks_mask_row_bytes(n_nodes: u32) -> u64;
    // ---- packed core (include/ksched.h) ----
    pub fn ks_snapshot_create(device: c_int, out: *mut *mut ks_snapshot) -> c_int;
    pub fn ks_snapshot_destroy(s: *mut ks_snapshot);
    pub fn ks_snapshot_set_nodes(s: *mut ks_snapshot, n: u32, w: u32, cpu: *const i64, mem: *const i64, labels: *const u64) -> c_int;
    pub fn ks_snapshot_set_bound(s: *mut ks_snapshot, b: u64, node: *const i32, cpu: *const i64, mem: *const i64) -> c_int;
    pub fn ks_snapshot_apply_bind(s: *mut ks_snapshot, node: i32, cpu: i64, mem: i64) -> c_int;
    pub fn ks_snapshot_get_free(s: *mut ks_snapshot, cpu: *mut i64, mem: *mut i64) -> c_int;
    pub fn ks_check_cell(s: *mut ks_snapshot, cpu: i64, mem: i64, sel: *const u64, node: u32) -> c_int;
    pub fn ks_check_cells(s: *mut ks_snapshot, pods: *const ks_pods, codes: *mut u8) -> c_int;
    pub fn ks_select(s: *mut ks_snapshot, pods: *const ks_pods, policy: c_int, flags: u32, out: *mut ks_bindings, stream: *mut c_void) -> c_int;
    pub fn ks_select_sampling(s: *mut ks_snapshot, pods: *const ks_pods, attempts: u32, seed: u64, first_pod_index: u64,
                              node_idx: *mut i32, attempts_used: *mut u32, draw_node: *mut i32, draw_code: *mut u8) -> c_int;
    pub fn ks_snapshot_commit_claims(s: *mut ks_snapshot, n: u64, node: *const i32, cpu: *const i64, mem: *const i64, accepted: *mut u8) -> c_int;
    pub fn ks_stream_bind(s: *mut ks_snapshot, pods: *const ks_pods, policy: c_int, node_idx: *mut i32, score: *mut i64, rounds: *mut u32) -> c_int;
    // asynchronous streaming surface = the Controller's work queue (src/main.rs:73,141-148)
    pub fn ks_stream_open(s: *mut ks_snapshot, policy: c_int, max_batch: u32, out: *mut *mut ks_stream) -> c_int;
    pub fn ks_stream_submit(q: *mut ks_stream, n: u64, cpu: *const i64, mem: *const i64, sel: *const u64, tickets: *const u64) -> c_int;
    pub fn ks_stream_poll(q: *mut ks_stream, max: u64, ticket: *mut u64, node_idx: *mut i32, score: *mut i64, out_n: *mut u64) -> c_int;
    pub fn ks_stream_flush(q: *mut ks_stream) -> c_int;
    pub fn ks_stream_close(q: *mut ks_stream);
    // multi-GPU: CUDA-IPC gather buffers for ks_bindings.exchange
    pub fn ks_ipc_alloc(device: c_int, bytes: u64, out_ptr: *mut *mut c_void, out_handle: *mut u8) -> c_int;
    pub fn ks_ipc_open(device: c_int, handle: *const u8, out_ptr: *mut *mut c_void) -> c_int;
    pub fn ks_ipc_close(device: c_int, ptr: *mut c_void) -> c_int;
    pub fn ks_ipc_free(device: c_int, ptr: *mut c_void) -> c_int;
    pub fn ks_exchange_check(s: *mut ks_snapshot) -> c_int;
    // ---- object layer (include/ksched_host.h) ----
    pub fn ksh_context_create(device: c_int, out: *mut *mut ksh_context) -> c_int;
    pub fn ksh_context_destroy(ctx: *mut ksh_context);
    pub fn ksh_context_set_nodes(ctx: *mut ksh_context, nodes: *const ks_node_obj, n: u32) -> c_int;
    pub fn ksh_context_set_cluster_pods(ctx: *mut ksh_context, pods: *const ks_pod_obj, n: u64) -> c_int;
    pub fn ksh_context_upsert_node(ctx: *mut ksh_context, node: *const ks_node_obj, out_idx: *mut u32) -> c_int;
    pub fn ksh_context_remove_node(ctx: *mut ksh_context, name: *const c_char) -> c_int;
    pub fn ksh_context_pod_bound(ctx: *mut ksh_context, pod: *const ks_pod_obj) -> c_int;
    pub fn ksh_context_pod_deleted(ctx: *mut ksh_context, pod: *const ks_pod_obj) -> c_int;
    pub fn ksh_context_node_name(ctx: *const ksh_context, node_idx: u32) -> *const c_char;
    pub fn ksh_total_pod_resources(pod: *const ks_pod_obj, cpu: *mut i64, mem: *mut i64) -> c_int;
    pub fn ksh_is_pod_bound(pod: *const ks_pod_obj) -> c_int;
    pub fn ksh_check_node_validity(ctx: *mut ksh_context, pod: *const ks_pod_obj, node_idx: u32) -> c_int;
    pub fn ksh_select_nodes(ctx: *mut ksh_context, pods: *const ks_pod_obj, n: u64, policy: c_int,
                            node_idx: *mut i32, score: *mut i64, cnt: *mut u32) -> c_int;
    pub fn ksh_select_node_for_pod(ctx: *mut ksh_context, pods: *const ks_pod_obj, n: u64, attempts: u32, seed: u64,
                                   first_pod_index: u64, node_idx: *mut i32, attempts_used: *mut u32,
                                   draw_node: *mut i32, draw_code: *mut u8) -> c_int;
    pub fn ksh_reconcile(ctx: *mut ksh_context, pod: *const ks_pod_obj, policy: c_int, node_idx: *mut i32,
                         binding_json: *mut c_char, cap: usize) -> c_int;
    pub fn ksh_reconcile_batch(ctx: *mut ksh_context, pods: *const ks_pod_obj, n: u64, policy: c_int, status: *mut i32,
                               node_idx: *mut i32, json: *mut c_char, cap: usize, json_off: *mut i64, rounds: *mut u32) -> c_int;
}
