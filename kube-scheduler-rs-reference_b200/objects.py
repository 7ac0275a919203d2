"""ctypes mirrors of include/ks_objects.h and builders that render Python specs / synth.Cluster rows as
Kubernetes-like Pod/Node objects with quantity strings.  The same arrays are handed to the C++ host layer
(ksh_*) and to the CPU oracle, so both evaluate "the same synthetic Pod/Node objects"."""
import ctypes as C


class ks_kv(C.Structure):
    _fields_ = [("key", C.c_char_p), ("val", C.c_char_p)]


class ks_container_obj(C.Structure):
    _fields_ = [("has_requests", C.c_int32), ("n_requests", C.c_uint32), ("requests", C.POINTER(ks_kv))]


class ks_pod_obj(C.Structure):
    _fields_ = [("ns", C.c_char_p), ("name", C.c_char_p), ("has_spec", C.c_int32), ("node_name", C.c_char_p),
                ("n_containers", C.c_uint32), ("containers", C.POINTER(ks_container_obj)),
                ("has_node_selector", C.c_int32), ("n_selector", C.c_uint32), ("selector", C.POINTER(ks_kv)),
                ("metadata_json", C.c_char_p)]


class ks_node_obj(C.Structure):
    _fields_ = [("name", C.c_char_p), ("has_labels", C.c_int32), ("n_labels", C.c_uint32),
                ("labels", C.POINTER(ks_kv)), ("has_allocatable", C.c_int32), ("n_allocatable", C.c_uint32),
                ("allocatable", C.POINTER(ks_kv))]


def _b(s):
    return s if isinstance(s, bytes) else str(s).encode()


class ObjectArena:
    """Owns every ctypes buffer referenced by the object arrays it builds (keeps them alive)."""

    def __init__(self):
        self._keep = []

    def kvs(self, mapping):
        items = list(mapping.items()) if isinstance(mapping, dict) else list(mapping)
        arr = (ks_kv * max(len(items), 1))()
        for i, (k, v) in enumerate(items):
            arr[i].key = _b(k)
            arr[i].val = _b(v)
        self._keep.append(arr)
        return arr, len(items)

    def pods(self, specs):
        """specs: list of dicts with keys
             name, ns (default 'default'), spec (default True), node_name (None = unbound),
             containers: list of (None | dict of requests)   [None = container without resources.requests]
             selector: None | dict
             metadata_json: None | str  (the pod's ObjectMeta as one JSON object, emitted verbatim in the Binding)"""
        arr = (ks_pod_obj * max(len(specs), 1))()
        for i, s in enumerate(specs):
            o = arr[i]
            ns, name = s.get("ns", "default"), s.get("name", f"pod{i}")
            o.ns = _b(ns) if ns is not None else None      # None = metadata.namespace / name absent
            o.name = _b(name) if name is not None else None
            o.has_spec = 1 if s.get("spec", True) else 0
            nn = s.get("node_name")
            o.node_name = _b(nn) if nn is not None else None
            conts = s.get("containers", [])
            carr = (ks_container_obj * max(len(conts), 1))()
            for j, c in enumerate(conts):
                if c is None:
                    carr[j].has_requests = 0
                    carr[j].n_requests = 0
                    carr[j].requests = None
                else:
                    kv, n = self.kvs(c)
                    carr[j].has_requests = 1
                    carr[j].n_requests = n
                    carr[j].requests = kv
            self._keep.append(carr)
            o.n_containers = len(conts)
            o.containers = carr
            mj = s.get("metadata_json")
            o.metadata_json = _b(mj) if mj is not None else None
            sel = s.get("selector")
            if sel is None:
                o.has_node_selector = 0
                o.n_selector = 0
                o.selector = None
            else:
                kv, n = self.kvs(sel)
                o.has_node_selector = 1
                o.n_selector = n
                o.selector = kv
        self._keep.append(arr)
        return arr

    def nodes(self, specs):
        """specs: list of dicts: name, labels (None | dict), allocatable (None | dict with cpu, memory)"""
        arr = (ks_node_obj * max(len(specs), 1))()
        for i, s in enumerate(specs):
            o = arr[i]
            o.name = _b(s.get("name", f"node{i}"))
            lab = s.get("labels")
            if lab is None:
                o.has_labels = 0
                o.n_labels = 0
                o.labels = None
            else:
                kv, n = self.kvs(lab)
                o.has_labels = 1
                o.n_labels = n
                o.labels = kv
            al = s.get("allocatable")
            if al is None:
                o.has_allocatable = 0
                o.n_allocatable = 0
                o.allocatable = None
            else:
                kv, n = self.kvs(al)
                o.has_allocatable = 1
                o.n_allocatable = n
                o.allocatable = kv
        self._keep.append(arr)
        return arr


def _cpu_str(milli, variant):
    milli = int(milli)
    if milli % 1000 == 0 and variant % 2 == 0:
        return str(milli // 1000)  # integer cores
    return f"{milli}m"             # integer millicores


def cluster_specs(cl, pod_slice=None):
    """Render a synth.Cluster as (node_specs, bound_pod_specs, pending_pod_specs) in the exact domain."""
    nodes = []
    for n in range(cl.N):
        labels = {f"k{k}": f"v{int(cl.node_vals[n, k])}" for k in range(cl.n_keys)}
        nodes.append({"name": f"node-{n}", "labels": labels,
                      "allocatable": {"cpu": _cpu_str(cl.alloc_cpu[n], n), "memory": str(int(cl.alloc_mem[n]))}})
    bound = []
    for b in range(cl.B):
        bound.append({"name": f"bound-{b}", "ns": "load", "node_name": f"node-{int(cl.bound_node[b])}",
                      "containers": [{"cpu": _cpu_str(cl.bound_cpu[b], b), "memory": str(int(cl.bound_mem[b]))}]})
    pods = []
    rng = range(cl.P) if pod_slice is None else range(*pod_slice.indices(cl.P))
    for p in rng:
        c = int(cl.n_containers[p])
        tc, tm = int(cl.req_cpu[p]), int(cl.req_mem[p])
        conts = []
        for j in range(c):
            pc = tc // c if j < c - 1 else tc - (tc // c) * (c - 1)
            pm = tm // c if j < c - 1 else tm - (tm // c) * (c - 1)
            conts.append({"cpu": _cpu_str(pc, p + j), "memory": str(pm)})
        if p % 7 == 0:
            conts.append(None)  # a container without resources.requests contributes nothing (util.rs:59-63)
        sel = None
        if int(cl.sel_n[p]) > 0:
            sel = {}
            for j in range(int(cl.sel_n[p])):
                v = int(cl.sel_vals[p, j])
                sel[f"k{int(cl.sel_keys[p, j])}"] = "v-absent" if v == 99 else f"v{v}"
        pods.append({"name": f"pod-{p}", "ns": "work", "containers": conts, "selector": sel})
    return nodes, bound, pods
