"""Host packer throughput: Pod/Node OBJECTS (strings) -> the SoA int64 / label-bitmask arrays the
device consumes, through a packing-only context (KSH_DEVICE_NONE).  Pure host work, single thread; runs without a GPU.
Prints one JSON line.  The objects are the synthetic cluster of bench.py (seed and shapes of BASELINE.json configs[1])."""
import argparse
import json
import sys
import time

import ksched_pkg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2", choices=["c2", "c3"])
    ap.add_argument("--pods", type=int, default=0, help="pod sample (0 = the workload's P)")
    args = ap.parse_args()
    ks = ksched_pkg.load()
    cl = ks.synth.config(args.workload)
    P = args.pods or cl.P
    t0 = time.perf_counter()
    nodes_s, bound_s, pods_s = ks.objects.cluster_specs(cl)
    arena = ks.objects.ObjectArena()
    nodes, bound, pods = arena.nodes(nodes_s), arena.pods(bound_s), arena.pods(pods_s[:P])
    t_objs = time.perf_counter() - t0
    with ks.host.Context(ks.host.KSH_DEVICE_NONE) as ctx:
        t0 = time.perf_counter()
        ctx.set_nodes(nodes, cl.N)
        t_nodes = time.perf_counter() - t0
        t0 = time.perf_counter()
        ctx.set_cluster_pods(bound, cl.B)
        t_bound = time.perf_counter() - t0
        t0 = time.perf_counter()
        rc, rm, sel = ctx.pack_pods(pods, P)
        t_pods = time.perf_counter() - t0
        t0 = time.perf_counter()
        ctx.export_packed()
        t_exp = time.perf_counter() - t0
        W = ctx.label_words
    print(json.dumps({
        "metric": "host_packer_objects_per_sec", "workload": args.workload, "threads": 1,
        "nodes": cl.N, "nodes_per_s": cl.N / t_nodes, "bound_pods": cl.B, "bound_pods_per_s": cl.B / t_bound,
        "pods": P, "pods_per_s": P / t_pods, "export_ms": 1e3 * t_exp, "label_words": W,
        "ms": {"set_nodes": 1e3 * t_nodes, "set_cluster_pods": 1e3 * t_bound, "pack_pods": 1e3 * t_pods},
        "python_object_build_s": t_objs,
        "note": "quantity strings parsed exactly (cpu -> millicores, memory -> bytes), selector-driven label dictionary",
    }))
    return 0


if __name__ == "__main__":
    sys.exit(main())
