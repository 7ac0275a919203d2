#!/usr/bin/env python
"""bench_stream.py — BASELINE.json config C5: streaming reconcile.  Poisson arrivals (default 10k pods/s for 10 s,
SURVEY.md §8d) of unschedulable pods against a resident 50k-node snapshot; bind latency = time the binding is
available to the caller - arrival time.  Prints one JSON line with p50/p99/max latency.

  1 GPU      the asynchronous surface (ks_stream_submit / ks_stream_poll, include/ksched.h): the producer submits each
             pod at its arrival time, the library's dispatcher drains what has arrived into one micro-batch and runs the
             device-side round loop (k_stream_batch: one cooperative launch per batch, capacity committed on the device).
             `--sync` uses the blocking ks_stream_bind from the producer thread instead (round 1's loop).
  N GPUs     `--mode replicas` (default): capacity commit is inherently sequential and at this rate one GPU is idle
             >90 % of the time, so rank 0 schedules every arrival exactly as in the 1-GPU run and the other ranks are
             hot replicas: every few ms rank 0 broadcasts the claims accepted since the last sync and each replica
             commits them (ks_snapshot_commit_claims); all replicas must end with rank 0's free[] (checksum printed).
             `--mode lockstep`: round 1's protocol for comparison - arrivals round-robin over the ranks, one all-gather
             of the claims per round, every rank commits the union in arrival order.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def percentiles(lat_ms):
    return {"p50_ms": float(np.percentile(lat_ms, 50)), "p99_ms": float(np.percentile(lat_ms, 99)),
            "p999_ms": float(np.percentile(lat_ms, 99.9)), "max_ms": float(lat_ms.max()), "mean_ms": float(lat_ms.mean())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rate", type=float, default=10000.0)
    ap.add_argument("--seconds", type=float, default=10.0)  # >= 10 s of arrivals
    ap.add_argument("--nodes", type=int, default=50000)
    ap.add_argument("--policy", default="leftover")
    ap.add_argument("--sync", action="store_true", help="1 GPU: blocking ks_stream_bind instead of submit/poll")
    ap.add_argument("--mode", default="replicas", choices=["replicas", "lockstep"])
    ap.add_argument("--sync-ms", type=float, default=5.0, help="replicas: period of the claim broadcast")
    args = ap.parse_args()
    import torch
    import ksched_pkg
    ks = ksched_pkg.load()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
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
    snap.select(rc[:64], rm[:64], sel[:64], flags=1)  # warm-up (allocations, module load); no commit
    fc0, fm0 = cl.free()  # the synthetic cluster starts with a few oversubscribed nodes by construction
    if world > 1:
        dist.barrier()
    lockstep = world > 1 and args.mode == "lockstep"
    lat = np.zeros(n)
    node = np.full(n, -2, np.int32)
    extra = {}

    if lockstep:  # ---- round 1's distributed protocol (all ranks schedule; one all-gather of claims per round) ----
        ids = np.nonzero(np.arange(n) % world == rank)[0]
        batches, rounds_l, nxt = [], [], 0
        t0 = time.perf_counter()
        while True:
            now = time.perf_counter() - t0
            hi = int(np.searchsorted(t_arr[ids], now, side="right"))
            b = ids[nxt:hi]
            idx, rounds, all_done = ks.multigpu.stream_bind_distributed(snap, rc[b], rm[b], sel[b], b, policy=policy,
                                                                        done=hi >= len(ids))
            done = time.perf_counter() - t0
            if len(b):
                lat[b] = done - t_arr[b]
                node[b] = idx
                batches.append(len(b))
                rounds_l.append(rounds)
            nxt = hi
            if all_done:
                break
        total = time.perf_counter() - t0
        mine = ids
        extra = {"mean_batch": float(np.mean(batches)), "max_batch": int(np.max(batches)), "mean_rounds": float(np.mean(rounds_l))}
    elif rank == 0:  # ---- the scheduler: 1 GPU, or rank 0 of the replica set ----
        log_n, log_c, log_m = [], [], []  # accepted claims not yet sent to the replicas
        t_sync = 0.0
        n_sync = 0

        def replicate(final=False):
            nonlocal n_sync
            k = len(log_n)
            head = torch.tensor([k, 1 if final else 0], dtype=torch.int64, device=dev)
            dist.broadcast(head, src=0)
            if k:
                payload = torch.from_numpy(np.stack([np.asarray(log_n, np.int64), np.asarray(log_c, np.int64),
                                                     np.asarray(log_m, np.int64)])).to(dev)
                dist.broadcast(payload, src=0)
                log_n.clear(), log_c.clear(), log_m.clear()
            n_sync += 1

        def record(t, i, now):
            if len(t) == 0:
                return
            t = t.astype(np.int64)
            lat[t] = now - t_arr[t]
            node[t] = i
            if world > 1:
                ok = i >= 0
                log_n.extend(i[ok].tolist()), log_c.extend(rc[t[ok]].tolist()), log_m.extend(rm[t[ok]].tolist())

        t0 = time.perf_counter()
        nxt = 0
        if args.sync:
            batches, rounds_l = [], []
            while nxt < n:
                now = time.perf_counter() - t0
                hi = int(np.searchsorted(t_arr, now, side="right"))
                if hi == nxt:
                    continue
                b = np.arange(nxt, hi)
                idx, _, rounds = snap.stream_bind(rc[b], rm[b], sel[b], policy=policy)
                record(b.astype(np.uint64), idx, time.perf_counter() - t0)
                batches.append(len(b))
                rounds_l.append(rounds)
                nxt = hi
                if world > 1 and time.perf_counter() - t0 - t_sync > args.sync_ms * 1e-3:
                    replicate()
                    t_sync = time.perf_counter() - t0
            extra = {"mean_batch": float(np.mean(batches)), "max_batch": int(np.max(batches)), "mean_rounds": float(np.mean(rounds_l)),
                     "surface": "ks_stream_bind (blocking)"}
        else:
            tickets = np.arange(n, dtype=np.uint64)
            got = 0
            with ks.Stream(snap, policy=policy) as q:
                while got < n:
                    now = time.perf_counter() - t0
                    hi = int(np.searchsorted(t_arr, now, side="right")) if nxt < n else n
                    if hi > nxt:
                        q.submit(rc[nxt:hi], rm[nxt:hi], sel[nxt:hi], tickets[nxt:hi])
                        nxt = hi
                    t, i, _ = q.poll()
                    if len(t):
                        record(t, i, time.perf_counter() - t0)
                        got += len(t)
                    if world > 1 and time.perf_counter() - t0 - t_sync > args.sync_ms * 1e-3:
                        replicate()
                        t_sync = time.perf_counter() - t0
                nb, nr, mx = q.stats()
            extra = {"mean_batch": n / max(nb, 1), "max_batch": mx, "mean_rounds": nr / max(nb, 1), "batches": nb,
                     "surface": "ks_stream_submit / ks_stream_poll (dispatcher thread + k_stream_batch)"}
        total = time.perf_counter() - t0
        if world > 1:
            replicate(final=True)
            extra["replica_syncs"] = n_sync
        mine = np.arange(n)
    else:  # ---- a hot replica: apply what rank 0 accepted ----
        t0 = time.perf_counter()
        applied = 0
        while True:
            head = torch.zeros(2, dtype=torch.int64, device=dev)
            dist.broadcast(head, src=0)
            k, final = int(head[0].item()), int(head[1].item())
            if k:
                payload = torch.empty((3, k), dtype=torch.int64, device=dev)
                dist.broadcast(payload, src=0)
                p = payload.cpu().numpy()
                acc = snap.commit_claims(p[0].astype(np.int32), p[1], p[2])
                assert acc.all(), "a claim accepted by the scheduler rank did not fit on the replica"
                applied += k
            if final:
                break
        total = time.perf_counter() - t0
        mine = np.arange(0)
        extra = {"claims_applied": applied}

    fc, fm = snap.free()
    checksum = int((int(fc.sum()) * 31 + int(fm.sum())) % (1 << 61))
    if world > 1:
        cs = torch.tensor([checksum], dtype=torch.int64, device=dev)
        allcs = [torch.zeros_like(cs) for _ in range(world)]
        dist.all_gather(allcs, cs)
        replicas_equal = len({int(c.item()) for c in allcs}) == 1
    else:
        replicas_equal = True
    if lockstep and world > 1:  # latency over all pods: gather the per-rank vectors on rank 0
        lat_t = torch.from_numpy(lat).to(dev)
        dist.all_reduce(lat_t)
        lat = lat_t.cpu().numpy()
        node_t = torch.from_numpy(np.where(node == -2, 0, node).astype(np.int64)).to(dev)
        dist.all_reduce(node_t)
        node = node_t.cpu().numpy().astype(np.int32)
        mine = np.arange(n)
    if rank == 0:
        lat_ms = lat[mine] * 1e3
        bound = int((node[mine] >= 0).sum())
        newly_over = int((((fc < 0) & (fc0 >= 0)) | ((fm < 0) & (fm0 >= 0))).sum())
        line = {"metric": "stream_bind_latency_ms",
                "config": {"workload": f"c5: Poisson {args.rate:.0f} pods/s for {args.seconds}s vs {args.nodes} nodes", "world": world,
                           "mode": ("lockstep" if lockstep else "replicas") if world > 1 else "single", "policy": args.policy},
                **percentiles(lat_ms), "pods": int(len(mine)), "bound": bound, "binds_per_s": bound / total,
                "nodes_oversubscribed_by_stream": newly_over, "replicas_equal": replicas_equal, "replica_checksum": checksum,
                "gpu_launches": ks.launch_count(), **extra}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
