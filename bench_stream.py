#!/usr/bin/env python
"""bench_stream.py — BASELINE.json config C5: streaming reconcile.  Poisson arrivals (default 10k pods/s) of
unschedulable pods against a resident 50k-node snapshot; each loop iteration takes every pod that has arrived,
runs one micro-batch (select with the per-cell kernel + K3 capacity commit, losers retried) and records
bind latency = completion time - arrival time.  Prints one JSON line with p50/p99 latency.
Single process = 1 GPU; under torchrun every rank holds a replica, takes arrivals i % world == rank and the
ranks exchange claims with one all-gather per round (multigpu.stream_bind_distributed)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rate", type=float, default=10000.0)
    ap.add_argument("--seconds", type=float, default=10.0)  # >= 10 s of arrivals
    ap.add_argument("--nodes", type=int, default=50000)
    ap.add_argument("--policy", default="leftover")
    args = ap.parse_args()
    import torch
    import ksched_pkg
    ks = ksched_pkg.load()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    seed = ks.synth.SEEDS["c5"]
    n = int(args.rate * args.seconds)
    cl = ks.synth.make(n, args.nodes, seed)
    ac, am, lab, bn, bc, bm, rc, rm, sel = cl.packed()
    gaps = -np.log(1.0 - (ks.synth.stream(seed, 77, n) >> np.uint64(11)).astype(np.float64) / float(1 << 53)) / args.rate
    t_arr = np.cumsum(gaps)
    policy = 0 if args.policy == "leftover" else 1
    snap = ks.Snapshot(local)
    snap.set_nodes(ac, am, lab)
    snap.set_bound(bn, bc, bm)
    mine = np.arange(n) % world == rank
    ids = np.nonzero(mine)[0]
    snap.select(rc[:64], rm[:64], sel[:64], flags=1)  # warm-up (allocations, module load); no commit
    if world > 1:
        dist.barrier()
    lat, batches, rounds_l = [], [], []
    bound = 0
    nxt = 0
    t0 = time.perf_counter()
    while True:
        now = time.perf_counter() - t0
        hi = int(np.searchsorted(t_arr[ids], now, side="right"))
        if world == 1:
            if nxt >= len(ids):
                break
            if hi == nxt:
                continue  # nothing has arrived yet
        b = ids[nxt:hi]
        all_done = False
        if world == 1:
            idx, _, rounds = snap.stream_bind(rc[b], rm[b], sel[b], policy=policy)
        else:  # lockstep over ranks: a rank without arrivals takes part with an empty batch
            idx, rounds, all_done = ks.multigpu.stream_bind_distributed(snap, rc[b], rm[b], sel[b], b, policy=policy,
                                                                        done=hi >= len(ids))
        done = time.perf_counter() - t0
        if len(b):
            lat.extend((done - t_arr[b]).tolist())
            batches.append(len(b))
            rounds_l.append(rounds)
            bound += int((idx >= 0).sum())
        nxt = hi
        if world > 1 and all_done:
            break
    total = time.perf_counter() - t0
    lat_ms = np.asarray(lat) * 1e3
    fc, fm = snap.free()
    fc0, fm0 = cl.free()  # the synthetic cluster starts with a few oversubscribed nodes by construction
    newly_over = int((((fc < 0) & (fc0 >= 0)) | ((fm < 0) & (fm0 >= 0))).sum())
    line = {"metric": "stream_bind_latency_ms", "config": {"workload": f"c5: Poisson {args.rate:.0f} pods/s for {args.seconds}s vs {args.nodes} nodes",
            "world": world, "rank": rank, "policy": args.policy},
            "p50_ms": float(np.percentile(lat_ms, 50)), "p99_ms": float(np.percentile(lat_ms, 99)), "max_ms": float(lat_ms.max()),
            "pods": int(len(ids)), "bound": bound, "binds_per_s": bound / total, "mean_batch": float(np.mean(batches)),
            "max_batch": int(np.max(batches)), "mean_rounds": float(np.mean(rounds_l)),
            "nodes_oversubscribed_by_stream": newly_over, "min_free_cpu_after": int(fc.min()), "min_free_cpu_before": int(fc0.min()), "replica_checksum": int((fc.sum() * 31 + fm.sum()) % (1 << 61)),
            "gpu_launches": ks.launch_count()}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
