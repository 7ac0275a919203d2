// ksched.rs — safe wrapper that keeps the reference's own signatures:
//     check_node_validity(&Pod, &Node-index, &Gpu) -> Result<(), InvalidNodeReason>      src/predicates.rs:63-77
//     select_node_for_pod(&Pod, &Gpu) -> Option<node name>                               src/main.rs:51-71
//     reconcile body: Binding JSON for the POST                                          src/main.rs:79-103
// NOT COMPILED IN THIS REPOSITORY (no Rust toolchain in the build image); see ksched_sys.rs.
use std::ffi::{CStr, CString};
use std::os::raw::c_char;
use std::ptr;

use k8s_openapi::api::core::v1 as corev1;

use crate::ksched_sys::*;
use crate::predicates::InvalidNodeReason;

/// Owns the CStrings a `ks_pod_obj` view points into (the library copies what it keeps; views are per call).
pub struct PodView {
    _strings: Vec<CString>,
    _kvs: Vec<Vec<ks_kv>>,
    _containers: Vec<ks_container_obj>,
    pub obj: ks_pod_obj,
}

fn cs(v: &mut Vec<CString>, s: &str) -> *const c_char {
    v.push(CString::new(s).expect("NUL in a Kubernetes string"));
    v.last().unwrap().as_ptr()
}

impl PodView {
    pub fn new(pod: &corev1::Pod) -> PodView {
        let mut strings = Vec::new();
        let mut kvs: Vec<Vec<ks_kv>> = Vec::new();
        let mut containers = Vec::new();
        let ns = pod.metadata.namespace.as_deref().map(|s| cs(&mut strings, s)).unwrap_or(ptr::null());
        let name = pod.metadata.name.as_deref().map(|s| cs(&mut strings, s)).unwrap_or(ptr::null());
        let meta = cs(&mut strings, &serde_json::to_string(&pod.metadata).expect("ObjectMeta serialises"));
        let (mut node_name, mut has_sel, mut sel_ptr, mut n_sel) = (ptr::null(), 0, ptr::null(), 0u32);
        if let Some(spec) = &pod.spec {
            if let Some(nn) = &spec.node_name {
                node_name = cs(&mut strings, nn);
            }
            for c in &spec.containers {
                // only resources.requests is read (src/util.rs:58-62)
                match c.resources.as_ref().and_then(|r| r.requests.as_ref()) {
                    Some(req) => {
                        let v: Vec<ks_kv> = req.iter().map(|(k, q)| ks_kv { key: cs(&mut strings, k), val: cs(&mut strings, &q.0) }).collect();
                        kvs.push(v);
                        let v = kvs.last().unwrap();
                        containers.push(ks_container_obj { has_requests: 1, n_requests: v.len() as u32, requests: v.as_ptr() });
                    },
                    None => containers.push(ks_container_obj { has_requests: 0, n_requests: 0, requests: ptr::null() }),
                }
            }
            if let Some(sel) = &spec.node_selector {
                let v: Vec<ks_kv> = sel.iter().map(|(k, val)| ks_kv { key: cs(&mut strings, k), val: cs(&mut strings, val) }).collect();
                kvs.push(v);
                let v = kvs.last().unwrap();
                has_sel = 1;
                n_sel = v.len() as u32;
                sel_ptr = v.as_ptr();
            }
        }
        let obj = ks_pod_obj {
            ns,
            name,
            has_spec: pod.spec.is_some() as i32,
            node_name,
            n_containers: containers.len() as u32,
            containers: containers.as_ptr(),
            has_node_selector: has_sel,
            n_selector: n_sel,
            selector: sel_ptr,
            metadata_json: meta,
        };
        PodView { _strings: strings, _kvs: kvs, _containers: containers, obj }
    }
}

/// Replaces Context.node_store + the per-cell LIST (src/util.rs:12-15, src/predicates.rs:21-38).
pub struct Gpu {
    ctx: *mut ksh_context,
}
unsafe impl Send for Gpu {} // the library serialises per handle (one mutex per context)
unsafe impl Sync for Gpu {}

fn last_error() -> String {
    unsafe { CStr::from_ptr(ks_last_error()).to_string_lossy().into_owned() }
}

impl Gpu {
    pub fn new(device: i32) -> Result<Gpu, String> {
        let mut ctx = ptr::null_mut();
        if unsafe { ksh_context_create(device, &mut ctx) } != KS_OK {
            return Err(last_error());
        }
        Ok(Gpu { ctx })
    }

    /// src/predicates.rs:63-77 — same Result, same reason precedence (fit first, then selector)
    pub fn check_node_validity(&self, pod: &corev1::Pod, node_idx: u32) -> Result<(), InvalidNodeReason> {
        let v = PodView::new(pod);
        match unsafe { ksh_check_node_validity(self.ctx, &v.obj, node_idx) } {
            KS_CELL_OK => Ok(()),
            KS_CELL_NOT_ENOUGH_RESOURCES => Err(InvalidNodeReason::NotEnoughResources),
            KS_CELL_NODE_SELECTOR_MISMATCH => Err(InvalidNodeReason::NodeSelectorMismatch),
            e => panic!("ksh_check_node_validity failed ({}): {}", e, last_error()), // the reference panics on malformed objects too
        }
    }

    /// src/main.rs:51-71 — `None` when no node is feasible.  `reference_policy` keeps the <=ATTEMPTS random draws
    /// (seeded), otherwise the argmax of the leftover score over every feasible node.
    pub fn select_node_for_pod(&self, pod: &corev1::Pod, reference_policy: Option<(u64, u64)>) -> Option<String> {
        let v = PodView::new(pod);
        let mut idx: i32 = -1;
        let rc = match reference_policy {
            Some((seed, pod_counter)) => unsafe {
                ksh_select_node_for_pod(self.ctx, &v.obj, 1, KS_REFERENCE_ATTEMPTS, seed, pod_counter, &mut idx, ptr::null_mut(),
                                        ptr::null_mut(), ptr::null_mut())
            },
            None => unsafe { ksh_select_nodes(self.ctx, &v.obj, 1, KS_SCORE_LEFTOVER, &mut idx, ptr::null_mut(), ptr::null_mut()) },
        };
        if rc != KS_OK {
            panic!("select failed ({}): {}", rc, last_error());
        }
        if idx < 0 {
            return None;
        }
        Some(unsafe { CStr::from_ptr(ksh_context_node_name(self.ctx, idx as u32)) }.to_string_lossy().into_owned())
    }

    /// src/main.rs:79-103 up to (not including) the POST: Ok(Some(body)) = bind, Ok(None) = already bound,
    /// Err(1) = NoNodeFound, Err(2) = CreateBindingObjectFailed.  A failed POST needs no rollback call: the requeued
    /// pod comes back through reconcile(), which releases the earlier charge before it selects again.
    pub fn reconcile(&self, pod: &corev1::Pod) -> Result<Option<String>, i32> {
        let v = PodView::new(pod);
        let mut idx: i32 = -1;
        let mut buf = vec![0u8; 16384];
        match unsafe { ksh_reconcile(self.ctx, &v.obj, KS_SCORE_LEFTOVER, &mut idx, buf.as_mut_ptr() as *mut c_char, buf.len()) } {
            KSH_RECONCILE_OK if idx < 0 => Ok(None),
            KSH_RECONCILE_OK => Ok(Some(unsafe { CStr::from_ptr(buf.as_ptr() as *const c_char) }.to_string_lossy().into_owned())),
            e => Err(e),
        }
    }
}

impl Drop for Gpu {
    fn drop(&mut self) {
        unsafe { ksh_context_destroy(self.ctx) }
    }
}
