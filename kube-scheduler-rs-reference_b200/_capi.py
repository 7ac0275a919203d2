"""ctypes binding of include/ksched.h (+ ksched_host.h once built).  Fails loudly if the library is missing."""
import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libksched.so")
INCLUDE_DIR = os.path.join(os.path.dirname(_HERE), "include")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(nvcc, sm_100a).  There is no Python/CPU fallback for the scheduling kernels."
    )
lib = C.CDLL(LIB_PATH)

KS_OK = 0
KS_MEM_HOST, KS_MEM_DEVICE = 0, 1
KS_SCORE_LEFTOVER, KS_SCORE_LEAST_ALLOCATED = 0, 1
KS_SELECT_AUTO, KS_SELECT_FORCE_DIRECT, KS_SELECT_FORCE_BITPAR, KS_SELECT_TIMING, KS_SELECT_NO_GRAPH = 0, 1, 2, 4, 8
KS_CELL_OK, KS_CELL_NOT_ENOUGH_RESOURCES, KS_CELL_NODE_SELECTOR_MISMATCH = 0, 1, 2
KS_ERR_NO_DEVICE = -8


class KsError(RuntimeError):
    def __init__(self, code, where):
        self.code = code
        msg = lib.ks_last_error().decode("utf-8", "replace")
        super().__init__(f"{where} failed with {code}: {msg}")


class ks_pods(C.Structure):
    _fields_ = [("n", C.c_uint64), ("req_cpu", C.c_void_p), ("req_mem", C.c_void_p), ("sel", C.c_void_p),
                ("mem_space", C.c_int32)]


KS_MAX_PEERS = 15


class ks_exchange(C.Structure):
    _fields_ = [("world", C.c_uint32), ("rank", C.c_uint32), ("n_peers", C.c_uint32),
                ("peer_node_idx", C.c_void_p * KS_MAX_PEERS), ("peer_score", C.c_void_p * KS_MAX_PEERS),
                ("peer_flag", C.c_void_p * KS_MAX_PEERS), ("local_flags", C.c_void_p), ("local_state", C.c_void_p)]


class ks_bindings(C.Structure):
    _fields_ = [("node_idx", C.c_void_p), ("score", C.c_void_p), ("feasible_cnt", C.c_void_p),
                ("mem_space", C.c_int32), ("mask", C.c_void_p), ("mask_row_bytes", C.c_uint64),
                ("mask_space", C.c_int32), ("bindings_ready_event", C.c_void_p), ("exchange", C.POINTER(ks_exchange))]


def _proto(name, restype, *argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = list(argtypes)
    return fn


_proto("ks_last_error", C.c_char_p)
_proto("ks_version", C.c_int)
_proto("ks_device_count", C.c_int)
_proto("ks_launch_count", C.c_uint64)
_proto("ks_mask_row_bytes", C.c_uint64, C.c_uint32)
_proto("ks_mask_row_bytes_aligned", C.c_uint64, C.c_uint32)
_proto("ks_snapshot_create", C.c_int, C.c_int, C.POINTER(C.c_void_p))
_proto("ks_snapshot_destroy", None, C.c_void_p)
_proto("ks_snapshot_set_nodes", C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p)
_proto("ks_snapshot_set_bound", C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p)
_proto("ks_snapshot_apply_bind", C.c_int, C.c_void_p, C.c_int32, C.c_int64, C.c_int64)
_proto("ks_snapshot_get_free", C.c_int, C.c_void_p, C.c_void_p, C.c_void_p)
_proto("ks_snapshot_num_nodes", C.c_uint32, C.c_void_p)
_proto("ks_snapshot_label_words", C.c_uint32, C.c_void_p)
_proto("ks_check_cell", C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_uint32)
_proto("ks_check_cells", C.c_int, C.c_void_p, C.POINTER(ks_pods), C.c_void_p)
_proto("ks_select", C.c_int, C.c_void_p, C.POINTER(ks_pods), C.c_int, C.c_uint32, C.POINTER(ks_bindings), C.c_void_p)
_proto("ks_last_timings", C.c_int, C.c_void_p, C.POINTER(C.c_float))
_proto("ks_last_path", C.c_char_p, C.c_void_p)
_proto("ks_last_trace", C.c_int, C.c_void_p, C.POINTER(C.c_uint64))
_proto("ks_snapshot_commit_claims", C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p)
_proto("ks_stream_bind", C.c_int, C.c_void_p, C.POINTER(ks_pods), C.c_int, C.c_void_p, C.c_void_p,
       C.POINTER(C.c_uint32))

_proto("ks_stream_open", C.c_int, C.c_void_p, C.c_int, C.c_uint32, C.POINTER(C.c_void_p))
_proto("ks_stream_submit", C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p)
_proto("ks_stream_poll", C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64))
_proto("ks_stream_flush", C.c_int, C.c_void_p)
_proto("ks_stream_stats", C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64))
_proto("ks_stream_close", None, C.c_void_p)
_proto("ks_exchange_check", C.c_int, C.c_void_p)
_proto("ks_ipc_alloc", C.c_int, C.c_int, C.c_uint64, C.POINTER(C.c_void_p), C.c_char_p)
_proto("ks_ipc_open", C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_void_p))
_proto("ks_ipc_close", C.c_int, C.c_int, C.c_void_p)
_proto("ks_ipc_free", C.c_int, C.c_int, C.c_void_p)
_proto("ks_measure_write_bandwidth", C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_double))
_proto("ks_device_read", C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64)
_proto("ks_select_sampling", C.c_int, C.c_void_p, C.POINTER(ks_pods), C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p,
       C.c_void_p, C.c_void_p, C.c_void_p)
KS_REFERENCE_ATTEMPTS = 5
_SAMPLING_MULT = 0xD1B54A32D192ED03


def sampling_stream(seed, p):
    """KS_SAMPLING_STREAM(seed, p) of include/ksched.h: start state of pod p's splitmix64 draw stream."""
    return (seed ^ (((p + 1) * _SAMPLING_MULT) & 0xFFFFFFFFFFFFFFFF)) & 0xFFFFFFFFFFFFFFFF


def declared_symbols():
    """Every function name declared in include/*.h (used by the CPU test that checks the export table)."""
    names = []
    for fn in sorted(os.listdir(INCLUDE_DIR)):
        if not fn.endswith(".h"):
            continue
        text = open(os.path.join(INCLUDE_DIR, fn)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names += re.findall(r"\b(ksh?_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def measure_write_bandwidth(device, dev_ptr, nbytes, iters=4):
    """GB/s of a store-only fill over a device buffer (its contents are overwritten)."""
    out = C.c_double()
    rc = lib.ks_measure_write_bandwidth(int(device), C.c_void_p(int(dev_ptr)), int(nbytes), int(iters), C.byref(out))
    if rc != KS_OK:
        raise KsError(rc, "ks_measure_write_bandwidth")
    return float(out.value)


def mask_row_bytes(n_nodes):
    return int(lib.ks_mask_row_bytes(int(n_nodes)))


def mask_row_bytes_aligned(n_nodes):
    return int(lib.ks_mask_row_bytes_aligned(int(n_nodes)))


def device_count():
    return int(lib.ks_device_count())


def launch_count():
    return int(lib.ks_launch_count())
