"""Python face of the C++ host layer (include/ksched_host.h): the reference's surface over Pod/Node OBJECTS —
check_node_validity / select_node_for_pod (batched) / reconcile, plus the quantity parser and packer."""
import ctypes as C

import numpy as np

from . import _capi as capi
from ._capi import KsError, lib
from .objects import ks_node_obj, ks_pod_obj

_vp = C.c_void_p
for _name, _res, _args in [
    ("ksh_parse_cpu_millicores", C.c_int, [C.c_char_p, C.POINTER(C.c_int64)]),
    ("ksh_parse_memory_bytes", C.c_int, [C.c_char_p, C.POINTER(C.c_int64)]),
    ("ksh_total_pod_resources", C.c_int, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    ("ksh_is_pod_bound", C.c_int, [_vp]),
    ("ksh_context_create", C.c_int, [C.c_int, C.POINTER(_vp)]),
    ("ksh_context_destroy", None, [_vp]),
    ("ksh_context_set_nodes", C.c_int, [_vp, _vp, C.c_uint32]),
    ("ksh_context_set_cluster_pods", C.c_int, [_vp, _vp, C.c_uint64]),
    ("ksh_context_upsert_node", C.c_int, [_vp, _vp, C.POINTER(C.c_uint32)]),
    ("ksh_context_remove_node", C.c_int, [_vp, C.c_char_p]),
    ("ksh_context_pod_bound", C.c_int, [_vp, _vp]),
    ("ksh_context_pod_deleted", C.c_int, [_vp, _vp]),
    ("ksh_context_node_name", C.c_char_p, [_vp, C.c_uint32]),
    ("ksh_context_num_nodes", C.c_uint32, [_vp]),
    ("ksh_context_label_words", C.c_uint32, [_vp]),
    ("ksh_context_num_bound", C.c_uint64, [_vp]),
    ("ksh_context_export_packed", C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    ("ksh_context_snapshot", _vp, [_vp]),
    ("ksh_pack_pods", C.c_int, [_vp, _vp, C.c_uint64, _vp, _vp, _vp, C.c_uint32]),
    ("ksh_check_node_validity", C.c_int, [_vp, _vp, C.c_uint32]),
    ("ksh_select_nodes", C.c_int, [_vp, _vp, C.c_uint64, C.c_int, _vp, _vp, _vp]),
    ("ksh_select_node_for_pod", C.c_int, [_vp, _vp, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, _vp, _vp, _vp, _vp]),
    ("ksh_reconcile_batch", C.c_int, [_vp, _vp, C.c_uint64, C.c_int, _vp, _vp, _vp, C.c_size_t, _vp, C.POINTER(C.c_uint32)]),
    ("ksh_reconcile", C.c_int, [_vp, _vp, C.c_int, C.POINTER(C.c_int32), C.c_char_p, C.c_size_t]),
]:
    _f = getattr(lib, _name)
    _f.restype = _res
    _f.argtypes = _args

KSH_DEVICE_NONE = -1  # packing-only context: objects -> SoA arrays on the host, no predicate is ever evaluated
KSH_RECONCILE_OK, KSH_RECONCILE_NO_NODE_FOUND, KSH_RECONCILE_BINDING_OBJECT_FAILED = 0, 1, 2


def _addr(arr, i=0):
    return C.addressof(arr) + i * C.sizeof(arr._type_)


def parse_cpu_millicores(s):
    out = C.c_int64()
    rc = lib.ksh_parse_cpu_millicores(s.encode() if isinstance(s, str) else s, C.byref(out))
    return rc, out.value


def parse_memory_bytes(s):
    out = C.c_int64()
    rc = lib.ksh_parse_memory_bytes(s.encode() if isinstance(s, str) else s, C.byref(out))
    return rc, out.value


def total_pod_resources(pods, i=0):
    c, m = C.c_int64(), C.c_int64()
    rc = lib.ksh_total_pod_resources(_addr(pods, i), C.byref(c), C.byref(m))
    return rc, c.value, m.value


def is_pod_bound(pods, i=0):
    return bool(lib.ksh_is_pod_bound(_addr(pods, i)))


class Context:
    """ksh_context = the reference's Context{client, node_store} for this path (src/util.rs:12-15)."""

    def __init__(self, device=0):
        h = _vp()
        rc = lib.ksh_context_create(int(device), C.byref(h))
        if rc != capi.KS_OK:
            raise KsError(rc, "ksh_context_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib.ksh_context_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_nodes(self, nodes, n):
        rc = lib.ksh_context_set_nodes(self._h, C.addressof(nodes), n)
        if rc != capi.KS_OK:
            raise KsError(rc, "ksh_context_set_nodes")

    def set_cluster_pods(self, pods, n):
        rc = lib.ksh_context_set_cluster_pods(self._h, C.addressof(pods), n)
        if rc != capi.KS_OK:
            raise KsError(rc, "ksh_context_set_cluster_pods")

    @property
    def label_words(self):
        return int(lib.ksh_context_label_words(self._h))

    @property
    def n_nodes(self):
        return int(lib.ksh_context_num_nodes(self._h))

    @property
    def num_bound(self):
        return int(lib.ksh_context_num_bound(self._h))

    def upsert_node(self, nodes, i=0):
        idx = C.c_uint32()
        rc = lib.ksh_context_upsert_node(self._h, _addr(nodes, i), C.byref(idx))
        if rc != capi.KS_OK:
            raise KsError(rc, "ksh_context_upsert_node")
        return idx.value

    def remove_node(self, name):
        rc = lib.ksh_context_remove_node(self._h, name.encode())
        if rc != capi.KS_OK:
            raise KsError(rc, "ksh_context_remove_node")

    def pod_bound(self, pods, i=0):
        rc = lib.ksh_context_pod_bound(self._h, _addr(pods, i))
        if rc != capi.KS_OK:
            raise KsError(rc, "ksh_context_pod_bound")

    def pod_deleted(self, pods, i=0):
        rc = lib.ksh_context_pod_deleted(self._h, _addr(pods, i))
        if rc != capi.KS_OK:
            raise KsError(rc, "ksh_context_pod_deleted")

    def node_name(self, idx):
        n = lib.ksh_context_node_name(self._h, int(idx))
        return None if n is None else n.decode()

    def pack_pods(self, pods, n):
        rc_ = np.empty(n, np.int64)
        rm_ = np.empty(n, np.int64)
        sel = np.zeros((n, 8), np.uint64)
        w = lib.ksh_pack_pods(self._h, C.addressof(pods), n, rc_.ctypes.data, rm_.ctypes.data, sel.ctypes.data, 8)
        if w < 0:
            raise KsError(w, "ksh_pack_pods")
        return rc_, rm_, np.ascontiguousarray(sel[:, :w])

    def export_packed(self):
        """(alloc_cpu, alloc_mem, labels[N,W], bound_node, bound_cpu, bound_mem): what the next device upload sends."""
        n, w, b = self.n_nodes, self.label_words, int(lib.ksh_context_num_bound(self._h))
        ac, am = np.empty(n, np.int64), np.empty(n, np.int64)
        lab = np.zeros((n, w), np.uint64)
        bn, bc, bm = np.empty(b, np.int32), np.empty(b, np.int64), np.empty(b, np.int64)
        rc = lib.ksh_context_export_packed(self._h, ac.ctypes.data, am.ctypes.data, lab.ctypes.data, bn.ctypes.data,
                                           bc.ctypes.data, bm.ctypes.data)
        if rc != capi.KS_OK:
            raise KsError(rc, "ksh_context_export_packed")
        return ac, am, lab, bn, bc, bm

    def check_node_validity(self, pods, i, node_idx):
        rc = lib.ksh_check_node_validity(self._h, _addr(pods, i), int(node_idx))
        if rc < 0:
            raise KsError(rc, "ksh_check_node_validity")
        return rc

    def select_nodes(self, pods, n, policy=capi.KS_SCORE_LEFTOVER, first=0):
        idx = np.empty(n, np.int32)
        score = np.empty(n, np.int64)
        cnt = np.empty(n, np.uint32)
        rc = lib.ksh_select_nodes(self._h, _addr(pods, first), n, int(policy), idx.ctypes.data, score.ctypes.data,
                                  cnt.ctypes.data)
        if rc != capi.KS_OK:
            raise KsError(rc, "ksh_select_nodes")
        return idx, score, cnt

    def select_node_for_pod(self, pods, n, attempts=capi.KS_REFERENCE_ATTEMPTS, seed=0, first=0):
        """select_node_for_pod with the reference's own seeded <=attempts-draw policy (src/main.rs:49-71)."""
        idx = np.empty(n, np.int32)
        used = np.empty(n, np.uint32)
        dn = np.empty((n, attempts), np.int32)
        dc = np.empty((n, attempts), np.uint8)
        rc = lib.ksh_select_node_for_pod(self._h, _addr(pods, first), n, int(attempts), int(seed), int(first),
                                         idx.ctypes.data, used.ctypes.data, dn.ctypes.data, dc.ctypes.data)
        if rc != capi.KS_OK:
            raise KsError(rc, "ksh_select_node_for_pod")
        return idx, used, dn, dc

    def reconcile(self, pods, i, policy=capi.KS_SCORE_LEFTOVER):
        node = C.c_int32(-1)
        buf = C.create_string_buffer(1024)
        rc = lib.ksh_reconcile(self._h, _addr(pods, i), int(policy), C.byref(node), buf, len(buf))
        if rc < 0:
            raise KsError(rc, "ksh_reconcile")
        return rc, node.value, buf.value.decode()

    def reconcile_batch(self, pods, n, policy=capi.KS_SCORE_LEFTOVER, json_cap=None):
        """reconcile() over a drained queue: (status[n], node_idx[n], bodies[n] (str | None), rounds)."""
        status = np.empty(n, np.int32)
        node = np.empty(n, np.int32)
        off = np.empty(n, np.int64)
        cap = int(json_cap if json_cap is not None else 512 * max(n, 1))
        buf = C.create_string_buffer(cap)
        rounds = C.c_uint32()
        rc = lib.ksh_reconcile_batch(self._h, C.addressof(pods), n, int(policy), status.ctypes.data, node.ctypes.data,
                                     C.addressof(buf), cap, off.ctypes.data, C.byref(rounds))
        if rc != capi.KS_OK:
            raise KsError(rc, "ksh_reconcile_batch")
        bodies = [None if o < 0 else C.string_at(C.addressof(buf) + int(o)).decode() for o in off]
        return status, node, bodies, rounds.value
