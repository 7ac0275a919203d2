"""Deterministic synthetic clusters, all inside the exact quantity domain of SURVEY §8c:
cpu = integer cores / integer millicores, memory = plain integer bytes.

Generation is counter-based splitmix64 (stateless, vectorised numpy), so value i of stream s of a seed is the
same everywhere.  One Cluster can be rendered two ways that must agree:
  * packed():  the SoA int64 + label-bitmask form the GPU kernels consume (built directly from the integers);
  * objects.build_*(): Kubernetes-like Pod/Node objects with quantity STRINGS (what the reference and the
    C++ host packer / the CPU oracle consume).
"""
from dataclasses import dataclass

import numpy as np

U64 = np.uint64
ABSENT = 99  # selector value index meaning "a (key,value) pair no node carries"

NODE_CPU_CORES = np.array([4, 8, 16, 32, 64, 96], np.int64)
NODE_MEM_GIB = np.array([16, 32, 64, 128, 256, 384], np.int64)

SEEDS = {"c2": 0xB2000002, "c3": 0xB2000003, "c5": 0xB2000005}
SHAPES = {"c2": (100_000, 10_000), "c3": (1_000_000, 50_000)}


def _mix(z):
    z = z.astype(U64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> U64(30))) * U64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> U64(27))) * U64(0x94D049BB133111EB)
        return z ^ (z >> U64(31))


def stream(seed, sid, n):
    """n pseudo-random uint64: splitmix64 finaliser over (seed, stream id, counter)."""
    with np.errstate(over="ignore"):
        base = U64(seed & 0xFFFFFFFFFFFFFFFF) + U64(sid) * U64(0xD1B54A32D192ED03)
        idx = (np.arange(n, dtype=U64) + U64(1)) * U64(0x9E3779B97F4A7C15)
        return _mix(base + idx)


@dataclass
class Cluster:
    n_keys: int
    alloc_cpu: np.ndarray      # [N] millicores
    alloc_mem: np.ndarray      # [N] bytes
    node_vals: np.ndarray      # [N, n_keys] value index of each label key on the node
    bound_node: np.ndarray     # [B] int32
    bound_cpu: np.ndarray      # [B]
    bound_mem: np.ndarray      # [B]
    req_cpu: np.ndarray        # [P]
    req_mem: np.ndarray        # [P]
    n_containers: np.ndarray   # [P] 1..3 (object rendering only)
    sel_n: np.ndarray          # [P] selector size 0..3
    sel_keys: np.ndarray       # [P,3]
    sel_vals: np.ndarray       # [P,3] value index or ABSENT
    seed: int = 0

    @property
    def P(self):
        return int(self.req_cpu.shape[0])

    @property
    def N(self):
        return int(self.alloc_cpu.shape[0])

    @property
    def B(self):
        return int(self.bound_node.shape[0])

    @property
    def label_words(self):
        return max(1, (self.n_keys * 8) // 64)

    def n_vals(self, key):
        return 7 if key == 0 else 8  # key 0 has 7 values so that 63 pairs + the absent bit fill one u64 word

    def pair_bit(self, keys, vals):
        keys = np.asarray(keys, np.int64)
        vals = np.asarray(vals, np.int64)
        bit = np.where(keys == 0, vals, 7 + (keys - 1) * 8 + vals)
        return np.where(vals == ABSENT, self.label_words * 64 - 1, bit)

    def packed(self):
        """(alloc_cpu, alloc_mem, labels[N,W], bound_node, bound_cpu, bound_mem, req_cpu, req_mem, sel[P,W])"""
        W = self.label_words
        labels = np.zeros((self.N, W), U64)
        for k in range(self.n_keys):
            bit = self.pair_bit(np.full(self.N, k), self.node_vals[:, k])
            np.bitwise_or.at(labels, (np.arange(self.N), bit // 64), U64(1) << (bit % 64).astype(U64))
        sel = np.zeros((self.P, W), U64)
        for j in range(3):
            on = self.sel_n > j
            bit = self.pair_bit(self.sel_keys[:, j], self.sel_vals[:, j])
            rows = np.nonzero(on)[0]
            np.bitwise_or.at(sel, (rows, bit[on] // 64), U64(1) << (bit[on] % 64).astype(U64))
        return (self.alloc_cpu, self.alloc_mem, labels, self.bound_node, self.bound_cpu, self.bound_mem,
                self.req_cpu, self.req_mem, sel)

    def free(self):
        fc = self.alloc_cpu.copy()
        fm = self.alloc_mem.copy()
        np.subtract.at(fc, self.bound_node, self.bound_cpu)
        np.subtract.at(fm, self.bound_node, self.bound_mem)
        return fc, fm

    def take_pods(self, start, count):
        """Same cluster, pods [start, start+count) only (multi-GPU shards, CPU-baseline samples)."""
        sl = slice(start, start + count)
        return Cluster(self.n_keys, self.alloc_cpu, self.alloc_mem, self.node_vals, self.bound_node, self.bound_cpu,
                       self.bound_mem, self.req_cpu[sl], self.req_mem[sl], self.n_containers[sl], self.sel_n[sl],
                       self.sel_keys[sl], self.sel_vals[sl], self.seed)


def make(n_pods, n_nodes, seed, n_keys=8, bound_per_node=10, selectors=True):
    """SURVEY §8d generator.  n_keys=8 -> W=1 (63 pairs + absent bit); n_keys=32 -> W=4."""
    N, P = int(n_nodes), int(n_pods)
    alloc_cpu = NODE_CPU_CORES[(stream(seed, 1, N) % U64(6)).astype(np.int64)] * 1000
    alloc_mem = NODE_MEM_GIB[(stream(seed, 2, N) % U64(6)).astype(np.int64)] << 30
    node_vals = np.zeros((N, n_keys), np.int64)
    for k in range(n_keys):
        nv = 7 if k == 0 else 8
        node_vals[:, k] = (stream(seed, 100 + k, N) % U64(nv)).astype(np.int64)
    # pre-bound load: bound_per_node pods per node sized for utilisation ~U(0,0.9); 1 node in 200 over-committed
    r = stream(seed, 3, N)
    util_pm = (r % U64(900)).astype(np.int64)
    over = (stream(seed, 4, N) % U64(200)) == 0
    util_pm = np.where(over, 1000 + (r % U64(100)).astype(np.int64), util_pm)
    b = np.arange(N * bound_per_node, dtype=np.int64)
    bound_node = (b % max(N, 1)).astype(np.int32)
    per_cpu = alloc_cpu * util_pm // (1000 * bound_per_node) if bound_per_node else alloc_cpu * 0
    per_mem = (((alloc_mem >> 20) * util_pm) // (1000 * bound_per_node)) << 20 if bound_per_node else alloc_mem * 0
    bound_cpu = per_cpu[bound_node] if N else np.zeros(0, np.int64)
    bound_mem = per_mem[bound_node] if N else np.zeros(0, np.int64)
    # pending pods
    req_cpu = 50 * (1 + (stream(seed, 10, P) % U64(80)).astype(np.int64))
    req_mem = (1 + (stream(seed, 11, P) % U64(256)).astype(np.int64)) << 26
    n_containers = (1 + stream(seed, 12, P) % U64(3)).astype(np.int64)
    rs = (stream(seed, 13, P) % U64(100)).astype(np.int64)
    sel_n = np.where(rs < 50, 0, np.where(rs < 80, 1, np.where(rs < 95, 2, 3)))
    a = (stream(seed, 14, P) % U64(n_keys)).astype(np.int64)
    step = 2 * (stream(seed, 15, P) % U64(max(n_keys // 2, 1))).astype(np.int64) + 1
    sel_keys = np.stack([(a + j * step) % n_keys for j in range(3)], axis=1)
    sel_vals = np.zeros((P, 3), np.int64)
    for j in range(3):
        nv = np.where(sel_keys[:, j] == 0, 7, 8)
        sel_vals[:, j] = (stream(seed, 16 + j, P) % U64(8)).astype(np.int64) % nv
    absent = (stream(seed, 19, P) % U64(100)) == 0
    sel_n = np.where(absent, np.maximum(sel_n, 1), sel_n)
    sel_vals[:, 0] = np.where(absent, ABSENT, sel_vals[:, 0])
    if not selectors:
        sel_n = np.zeros(P, np.int64)
    return Cluster(n_keys, alloc_cpu, alloc_mem, node_vals, bound_node, bound_cpu, bound_mem, req_cpu, req_mem,
                   n_containers, sel_n, sel_keys, sel_vals, seed)


def config(name):
    """Named configurations of BASELINE.json: 'c2' = 100k x 10k, 'c3' = 1M x 50k."""
    p, n = SHAPES[name]
    return make(p, n, SEEDS[name])
