"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package never imports this module."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("oracle.c", "oracle.h")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return LIB_PATH


if not os.path.exists(LIB_PATH):
    build()
lib = C.CDLL(LIB_PATH)

ORC_SCORE_LEFTOVER, ORC_SCORE_LEAST_ALLOCATED = 0, 1
ERR_PARSE, ERR_RANGE, ERR_INEXACT, ERR_MISSING = -3, -5, -6, -7

_vp = C.c_void_p
lib.orc_parse_quantity.restype = C.c_int
lib.orc_parse_quantity.argtypes = [C.c_char_p, C.POINTER(C.c_int64)]
lib.orc_total_pod_resources.restype = C.c_int
lib.orc_total_pod_resources.argtypes = [_vp, C.POINTER(C.c_int64)]
lib.orc_is_pod_bound.restype = C.c_int
lib.orc_is_pod_bound.argtypes = [_vp]
lib.orc_does_node_selector_match.restype = C.c_int
lib.orc_does_node_selector_match.argtypes = [_vp, _vp]
lib.orc_cluster_create.restype = _vp
lib.orc_cluster_create.argtypes = [_vp, C.c_uint32, _vp, C.c_uint64]
lib.orc_cluster_destroy.restype = None
lib.orc_cluster_destroy.argtypes = [_vp]
lib.orc_can_pod_fit.restype = C.c_int
lib.orc_can_pod_fit.argtypes = [_vp, _vp, C.c_uint32]
lib.orc_check_node_validity.restype = C.c_int
lib.orc_check_node_validity.argtypes = [_vp, _vp, C.c_uint32]
lib.orc_node_available.restype = C.c_int
lib.orc_node_available.argtypes = [_vp, C.c_uint32, C.POINTER(C.c_int64)]
lib.orc_score_cell.restype = C.c_int
lib.orc_score_cell.argtypes = [_vp, C.c_int, _vp, C.c_uint32, C.POINTER(C.c_int64)]
lib.orc_run_faithful.restype = C.c_int
lib.orc_run_faithful.argtypes = [_vp, _vp, C.c_uint64, C.c_int, _vp, _vp, _vp, _vp, C.c_uint64, _vp, C.c_int]
lib.orc_free_reduce.restype = C.c_int
lib.orc_free_reduce.argtypes = [C.c_uint32, _vp, _vp, C.c_uint64, _vp, _vp, _vp, _vp, _vp]
lib.orc_run_packed.restype = C.c_int
lib.orc_run_packed.argtypes = [C.c_uint32, C.c_uint32, _vp, _vp, _vp, _vp, _vp, C.c_uint64, _vp, _vp, _vp, C.c_int,
                               _vp, _vp, _vp, _vp, C.c_uint64, _vp, C.c_int]
lib.orc_select_sampling.restype = C.c_int32
lib.orc_select_sampling.argtypes = [_vp, _vp, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
lib.orc_select_sampling_packed.restype = C.c_int
lib.orc_select_sampling_packed.argtypes = [C.c_uint32, C.c_uint32, _vp, _vp, _vp, C.c_uint64, _vp, _vp, _vp, C.c_uint32, _vp, _vp,
                                           _vp, _vp, _vp]
lib.orc_online_cores.restype = C.c_int
lib.orc_commit_claims.restype = C.c_int
lib.orc_commit_claims.argtypes = [C.c_uint32, _vp, _vp, C.c_uint64, _vp, _vp, _vp, _vp]
lib.orc_stream_bind_packed.restype = C.c_int
lib.orc_stream_bind_packed.argtypes = [C.c_uint32, C.c_uint32, _vp, _vp, _vp, _vp, _vp, C.c_uint64, _vp, _vp, _vp, C.c_int,
                                       _vp, _vp, C.POINTER(C.c_uint32)]


def parse_quantity(s):
    out = C.c_int64()
    rc = lib.orc_parse_quantity(s.encode() if isinstance(s, str) else s, C.byref(out))
    return rc, out.value


def _p(a):
    return None if a is None else a.ctypes.data


def _addr(obj_array, i=0):
    return C.addressof(obj_array) + i * C.sizeof(obj_array._type_)


class Cluster:
    """orc_cluster over ctypes object arrays (nodes, all pods incl. bound ones)."""

    def __init__(self, nodes, n_nodes, all_pods, n_all_pods):
        self._nodes, self._pods = nodes, all_pods
        self.n_nodes = n_nodes
        self._h = lib.orc_cluster_create(C.addressof(nodes), n_nodes, C.addressof(all_pods), n_all_pods)
        if not self._h:
            raise MemoryError("orc_cluster_create")

    def close(self):
        if self._h:
            lib.orc_cluster_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def check(self, pods, i, node_idx):
        return lib.orc_check_node_validity(self._h, _addr(pods, i), node_idx)

    def fit(self, pods, i, node_idx):
        return lib.orc_can_pod_fit(self._h, _addr(pods, i), node_idx)

    def available(self, node_idx):
        out = (C.c_int64 * 2)()
        rc = lib.orc_node_available(self._h, node_idx, out)
        return rc, (out[0], out[1])

    def score(self, policy, pods, i, node_idx):
        out = C.c_int64()
        rc = lib.orc_score_cell(self._h, policy, _addr(pods, i), node_idx, C.byref(out))
        return rc, out.value

    def run(self, pods, n_pods, policy=0, want_mask=True, want_codes=False, nthreads=0, first=0):
        N = self.n_nodes
        idx = np.empty(n_pods, np.int32)
        score = np.empty(n_pods, np.int64)
        cnt = np.empty(n_pods, np.uint32)
        row = 32 * ((N + 255) // 256)
        mask = np.zeros((n_pods, row), np.uint8) if want_mask else None
        codes = np.empty((n_pods, N), np.uint8) if want_codes else None
        rc = lib.orc_run_faithful(self._h, _addr(pods, first), n_pods, policy, _p(idx), _p(score), _p(cnt), _p(mask),
                                  row, _p(codes), nthreads)
        if rc:
            raise RuntimeError(f"orc_run_faithful -> {rc}")
        return idx, score, cnt, mask, codes

    def sampling(self, pods, i, attempts, seed):
        st = C.c_uint64(seed)
        cells = C.c_uint32()
        n = lib.orc_select_sampling(self._h, _addr(pods, i), attempts, C.byref(st), C.byref(cells))
        return n, cells.value


def free_reduce(alloc_cpu, alloc_mem, bnode, bcpu, bmem):
    N = alloc_cpu.shape[0]
    fc = np.empty(N, np.int64)
    fm = np.empty(N, np.int64)
    a = [np.ascontiguousarray(x) for x in (alloc_cpu, alloc_mem, bnode.astype(np.int32), bcpu, bmem)]
    rc = lib.orc_free_reduce(N, _p(a[0]), _p(a[1]), a[2].shape[0], _p(a[2]), _p(a[3]), _p(a[4]), _p(fc), _p(fm))
    if rc:
        raise RuntimeError(f"orc_free_reduce -> {rc}")
    return fc, fm


def run_packed(free_cpu, free_mem, alloc_cpu, alloc_mem, labels, req_cpu, req_mem, sel, policy=0, want_mask=True,
               want_codes=False, nthreads=0):
    N = free_cpu.shape[0]
    P = req_cpu.shape[0]
    labels = np.ascontiguousarray(labels, np.uint64).reshape(N, -1) if N else np.zeros((0, 1), np.uint64)
    W = labels.shape[1]
    sel = np.ascontiguousarray(sel, np.uint64).reshape(P, W)
    arrs = [np.ascontiguousarray(x, np.int64) for x in (free_cpu, free_mem, alloc_cpu, alloc_mem, req_cpu, req_mem)]
    idx = np.empty(P, np.int32)
    score = np.empty(P, np.int64)
    cnt = np.empty(P, np.uint32)
    row = 32 * ((N + 255) // 256)
    mask = np.zeros((P, row), np.uint8) if want_mask else None
    codes = np.empty((P, N), np.uint8) if want_codes else None
    rc = lib.orc_run_packed(N, W, _p(arrs[0]), _p(arrs[1]), _p(arrs[2]), _p(arrs[3]), _p(labels), P, _p(arrs[4]),
                            _p(arrs[5]), _p(sel), policy, _p(idx), _p(score), _p(cnt), _p(mask), row, _p(codes),
                            nthreads)
    if rc:
        raise RuntimeError(f"orc_run_packed -> {rc}")
    return idx, score, cnt, mask, codes


def sampling_packed(free_cpu, free_mem, labels, req_cpu, req_mem, sel, attempts, states):
    """Reference policy on the packed form; states = uint64[P] generator start states (not modified)."""
    N = free_cpu.shape[0]
    P = req_cpu.shape[0]
    labels = np.ascontiguousarray(labels, np.uint64).reshape(N, -1) if N else np.zeros((0, 1), np.uint64)
    W = labels.shape[1]
    sel = np.ascontiguousarray(sel, np.uint64).reshape(P, W)
    a = [np.ascontiguousarray(x, np.int64) for x in (free_cpu, free_mem, req_cpu, req_mem)]
    st = np.array(states, np.uint64, copy=True)
    idx = np.empty(P, np.int32)
    cells = np.empty(P, np.uint32)
    dn = np.empty((P, attempts), np.int32)
    dc = np.empty((P, attempts), np.uint8)
    rc = lib.orc_select_sampling_packed(N, W, _p(a[0]), _p(a[1]), _p(labels), P, _p(a[2]), _p(a[3]), _p(sel), attempts,
                                        _p(st), _p(idx), _p(cells), _p(dn), _p(dc))
    if rc:
        raise RuntimeError(f"orc_select_sampling_packed -> {rc}")
    return idx, cells, dn, dc


def commit_claims(free_cpu, free_mem, claim_node, req_cpu, req_mem):
    """In-place on free_cpu/free_mem (int64 numpy); returns accepted uint8[n]."""
    n = len(claim_node)
    a = [np.ascontiguousarray(claim_node, np.int32), np.ascontiguousarray(req_cpu, np.int64),
         np.ascontiguousarray(req_mem, np.int64)]
    acc = np.zeros(n, np.uint8)
    rc = lib.orc_commit_claims(free_cpu.shape[0], _p(free_cpu), _p(free_mem), n, _p(a[0]), _p(a[1]), _p(a[2]), _p(acc))
    if rc:
        raise RuntimeError(f"orc_commit_claims -> {rc}")
    return acc


def stream_bind_packed(free_cpu, free_mem, alloc_cpu, alloc_mem, labels, req_cpu, req_mem, sel, policy=0):
    """In-place on free_cpu/free_mem; returns (node_idx, score, rounds)."""
    N = free_cpu.shape[0]
    P = len(req_cpu)
    labels = np.ascontiguousarray(labels, np.uint64).reshape(N, -1)
    W = labels.shape[1]
    sel = np.ascontiguousarray(sel, np.uint64).reshape(P, W)
    a = [np.ascontiguousarray(x, np.int64) for x in (alloc_cpu, alloc_mem, req_cpu, req_mem)]
    idx = np.empty(P, np.int32)
    score = np.empty(P, np.int64)
    rounds = C.c_uint32()
    rc = lib.orc_stream_bind_packed(N, W, _p(free_cpu), _p(free_mem), _p(a[0]), _p(a[1]), _p(labels), P, _p(a[2]),
                                    _p(a[3]), _p(sel), policy, _p(idx), _p(score), C.byref(rounds))
    if rc:
        raise RuntimeError(f"orc_stream_bind_packed -> {rc}")
    return idx, score, rounds.value
