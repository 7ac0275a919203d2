"""Snapshot — Python handle on ks_snapshot (the device-resident node table that replaces
Context.node_store + the per-cell LIST of the reference, /root/reference/src/util.rs:12-15,
src/predicates.rs:21-38).  numpy arrays = host buffers; objects with .data_ptr() (torch tensors) may be
passed for host-pinned or device buffers through select_raw()."""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _capi as capi
from ._capi import KsError, ks_bindings, ks_pods, lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return int(a)


def _c(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


@dataclass
class SelectResult:
    node_idx: np.ndarray
    score: np.ndarray
    feasible_cnt: np.ndarray
    mask: np.ndarray  # uint8 [P, row_bytes] or None
    path: str


class Snapshot:
    def __init__(self, device=0):
        h = C.c_void_p()
        rc = lib.ks_snapshot_create(int(device), C.byref(h))
        if rc != capi.KS_OK:
            raise KsError(rc, "ks_snapshot_create")
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            lib.ks_snapshot_destroy(self._h)
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

    @property
    def n_nodes(self):
        return int(lib.ks_snapshot_num_nodes(self._h))

    @property
    def label_words(self):
        return int(lib.ks_snapshot_label_words(self._h))

    def set_nodes(self, alloc_cpu, alloc_mem, labels):
        alloc_cpu = _c(alloc_cpu, np.int64)
        alloc_mem = _c(alloc_mem, np.int64)
        labels = _c(labels, np.uint64)
        n = alloc_cpu.shape[0]
        labels = labels.reshape(n, -1) if n else labels.reshape(0, max(1, labels.shape[-1] if labels.ndim > 1 else 1))
        w = labels.shape[1]
        rc = lib.ks_snapshot_set_nodes(self._h, n, w, _ptr(alloc_cpu), _ptr(alloc_mem), _ptr(labels))
        if rc != capi.KS_OK:
            raise KsError(rc, "ks_snapshot_set_nodes")

    def set_bound(self, node_idx, req_cpu, req_mem):
        node_idx = _c(node_idx, np.int32)
        req_cpu = _c(req_cpu, np.int64)
        req_mem = _c(req_mem, np.int64)
        rc = lib.ks_snapshot_set_bound(self._h, node_idx.shape[0], _ptr(node_idx), _ptr(req_cpu), _ptr(req_mem))
        if rc != capi.KS_OK:
            raise KsError(rc, "ks_snapshot_set_bound")

    def apply_bind(self, node_idx, req_cpu, req_mem):
        rc = lib.ks_snapshot_apply_bind(self._h, int(node_idx), int(req_cpu), int(req_mem))
        if rc != capi.KS_OK:
            raise KsError(rc, "ks_snapshot_apply_bind")

    def free(self):
        n = self.n_nodes
        fc = np.empty(n, np.int64)
        fm = np.empty(n, np.int64)
        rc = lib.ks_snapshot_get_free(self._h, _ptr(fc), _ptr(fm))
        if rc != capi.KS_OK:
            raise KsError(rc, "ks_snapshot_get_free")
        return fc, fm

    def check_cell(self, req_cpu, req_mem, sel, node_idx):
        sel = _c(sel, np.uint64).reshape(-1)
        rc = lib.ks_check_cell(self._h, int(req_cpu), int(req_mem), _ptr(sel), int(node_idx))
        if rc < 0:
            raise KsError(rc, "ks_check_cell")
        return rc

    def _pods(self, req_cpu, req_mem, sel):
        req_cpu = _c(req_cpu, np.int64)
        req_mem = _c(req_mem, np.int64)
        sel = _c(sel, np.uint64).reshape(req_cpu.shape[0], -1 if req_cpu.shape[0] else self.label_words)
        if sel.shape[1] != self.label_words:
            raise ValueError(f"sel has {sel.shape[1]} words per pod, snapshot has {self.label_words}")
        pods = ks_pods(req_cpu.shape[0], _ptr(req_cpu), _ptr(req_mem), _ptr(sel), capi.KS_MEM_HOST)
        return pods, (req_cpu, req_mem, sel)

    def check_cells(self, req_cpu, req_mem, sel):
        pods, keep = self._pods(req_cpu, req_mem, sel)
        codes = np.empty((pods.n, self.n_nodes), np.uint8)
        rc = lib.ks_check_cells(self._h, C.byref(pods), _ptr(codes))
        if rc != capi.KS_OK:
            raise KsError(rc, "ks_check_cells")
        return codes

    def select(self, req_cpu, req_mem, sel, policy=capi.KS_SCORE_LEFTOVER, flags=capi.KS_SELECT_AUTO, want_mask=False):
        """Batched select_node_for_pod over host (numpy) pods; returns host results."""
        pods, keep = self._pods(req_cpu, req_mem, sel)
        p = int(pods.n)
        idx = np.empty(p, np.int32)
        score = np.empty(p, np.int64)
        cnt = np.empty(p, np.uint32)
        mask = None
        row = capi.mask_row_bytes(self.n_nodes)
        if want_mask:
            mask = np.zeros((p, row), np.uint8)
        out = ks_bindings(_ptr(idx), _ptr(score), _ptr(cnt), capi.KS_MEM_HOST, _ptr(mask), row, capi.KS_MEM_HOST, None, None)
        rc = lib.ks_select(self._h, C.byref(pods), int(policy), int(flags), C.byref(out), None)
        if rc != capi.KS_OK:
            raise KsError(rc, "ks_select")
        return SelectResult(idx, score, cnt, mask, self.last_path())

    def select_raw(self, n_pods, req_cpu, req_mem, sel, pods_space, node_idx, score, cnt, out_space, mask=None,
                   mask_row_bytes=0, mask_space=capi.KS_MEM_DEVICE, policy=capi.KS_SCORE_LEFTOVER,
                   flags=capi.KS_SELECT_AUTO, stream=None, ready_event=None, exchange=None):
        """Pointer-level call (torch tensors / raw addresses); the caller keeps the buffers alive.
        exchange: a multigpu.PeerExchange (fused all-gather of node_idx / score into every rank's gather buffer)."""
        pods = ks_pods(int(n_pods), _ptr(req_cpu), _ptr(req_mem), _ptr(sel), int(pods_space))
        out = ks_bindings(_ptr(node_idx), _ptr(score), _ptr(cnt), int(out_space), _ptr(mask), int(mask_row_bytes),
                          int(mask_space), C.c_void_p(ready_event) if ready_event else None,
                          C.pointer(exchange.desc) if exchange is not None else None)
        rc = lib.ks_select(self._h, C.byref(pods), int(policy), int(flags), C.byref(out),
                           C.c_void_p(stream) if stream else None)
        if rc != capi.KS_OK:
            raise KsError(rc, "ks_select")

    def commit_claims(self, claim_node, req_cpu, req_mem):
        """K3: accept claims per node in arrival order while they fit; decrements free[] on the device."""
        claim_node = _c(claim_node, np.int32)
        req_cpu = _c(req_cpu, np.int64)
        req_mem = _c(req_mem, np.int64)
        acc = np.zeros(claim_node.shape[0], np.uint8)
        rc = lib.ks_snapshot_commit_claims(self._h, claim_node.shape[0], _ptr(claim_node), _ptr(req_cpu),
                                           _ptr(req_mem), _ptr(acc))
        if rc != capi.KS_OK:
            raise KsError(rc, "ks_snapshot_commit_claims")
        return acc

    def stream_bind(self, req_cpu, req_mem, sel, policy=capi.KS_SCORE_LEFTOVER):
        """One streaming micro-batch: select + commit rounds until every pod is bound or infeasible."""
        pods, keep = self._pods(req_cpu, req_mem, sel)
        p = int(pods.n)
        idx = np.empty(p, np.int32)
        score = np.empty(p, np.int64)
        rounds = C.c_uint32()
        rc = lib.ks_stream_bind(self._h, C.byref(pods), int(policy), _ptr(idx), _ptr(score), C.byref(rounds))
        if rc != capi.KS_OK:
            raise KsError(rc, "ks_stream_bind")
        return idx, score, rounds.value

    def select_sampling(self, req_cpu, req_mem, sel, attempts=capi.KS_REFERENCE_ATTEMPTS, seed=0, first_pod_index=0):
        """The reference's own policy (src/main.rs:49-71), seeded: (node_idx, attempts_used, draw_node, draw_code)."""
        pods, keep = self._pods(req_cpu, req_mem, sel)
        p = int(pods.n)
        idx = np.empty(p, np.int32)
        used = np.empty(p, np.uint32)
        dn = np.empty((p, attempts), np.int32)
        dc = np.empty((p, attempts), np.uint8)
        rc = lib.ks_select_sampling(self._h, C.byref(pods), int(attempts), int(seed), int(first_pod_index), _ptr(idx),
                                    _ptr(used), _ptr(dn), _ptr(dc))
        if rc != capi.KS_OK:
            raise KsError(rc, "ks_select_sampling")
        return idx, used, dn, dc

    def last_timings(self):
        ms = (C.c_float * 3)()
        rc = lib.ks_last_timings(self._h, ms)
        if rc != capi.KS_OK:
            raise KsError(rc, "ks_last_timings")
        return float(ms[0]), float(ms[1]), float(ms[2])

    def last_path(self):
        return lib.ks_last_path(self._h).decode()

    def last_trace(self):
        """Timeline of the last bit-parallel select (ks_last_trace; needs KS_TRACE=1 in the environment): microseconds
        from the start of the pod-rank kernel, {name: (start, end)}; kernels that did not run are left out."""
        ns = (C.c_uint64 * 16)()
        rc = lib.ks_last_trace(self._h, ns)
        if rc != capi.KS_OK:
            raise KsError(rc, "ks_last_trace")
        t0 = ns[0]
        out = {}
        for name, a, b in (("pod_ranks", 0, 1), ("argmax1", 2, 3), ("argmax2", 4, 5), ("mask", 6, 7)):
            if ns[a] and ns[b]:
                out[name] = ((ns[a] - t0) / 1e3, (ns[b] - t0) / 1e3)
        if ns[8]:
            out["mask_first_cta_end"] = (ns[8] - t0) / 1e3
        out["t0_ns"] = int(t0)
        return out

    def exchange_check(self):
        """Synchronise and raise if a fused all-gather (ks_exchange) timed out waiting for a peer."""
        rc = lib.ks_exchange_check(self._h)
        if rc != capi.KS_OK:
            raise KsError(rc, "ks_exchange_check")


class Stream:
    """ks_stream: the reference's Controller queue in front of one snapshot (asynchronous submit / poll)."""

    def __init__(self, snap, policy=capi.KS_SCORE_LEFTOVER, max_batch=0):
        h = C.c_void_p()
        rc = lib.ks_stream_open(snap._h, int(policy), int(max_batch), C.byref(h))
        if rc != capi.KS_OK:
            raise KsError(rc, "ks_stream_open")
        self._h, self._snap = h, snap
        self._cap = 4096
        self._t = np.empty(self._cap, np.uint64)
        self._i = np.empty(self._cap, np.int32)
        self._s = np.empty(self._cap, np.int64)

    def submit(self, req_cpu, req_mem, sel, tickets):
        req_cpu, req_mem = _c(req_cpu, np.int64), _c(req_mem, np.int64)
        sel, tickets = _c(sel, np.uint64), _c(tickets, np.uint64)
        rc = lib.ks_stream_submit(self._h, req_cpu.shape[0], _ptr(req_cpu), _ptr(req_mem), _ptr(sel), _ptr(tickets))
        if rc != capi.KS_OK:
            raise KsError(rc, "ks_stream_submit")

    def poll(self):
        """(tickets, node_idx, score) of the pods finished since the last poll (possibly empty); never blocks."""
        n = C.c_uint64()
        rc = lib.ks_stream_poll(self._h, self._cap, _ptr(self._t), _ptr(self._i), _ptr(self._s), C.byref(n))
        if rc != capi.KS_OK:
            raise KsError(rc, "ks_stream_poll")
        k = int(n.value)
        return self._t[:k].copy(), self._i[:k].copy(), self._s[:k].copy()

    def flush(self):
        rc = lib.ks_stream_flush(self._h)
        if rc != capi.KS_OK:
            raise KsError(rc, "ks_stream_flush")

    def stats(self):
        b, r, m = C.c_uint64(), C.c_uint64(), C.c_uint64()
        lib.ks_stream_stats(self._h, C.byref(b), C.byref(r), C.byref(m))
        return int(b.value), int(r.value), int(m.value)

    def close(self):
        if getattr(self, "_h", None):
            lib.ks_stream_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
