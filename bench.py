#!/usr/bin/env python
"""bench.py — (pod,node) predicate-cells/sec of the fused scheduling pass on N B200s (BASELINE.json metric).

One "step" = one pass of the hot path over one batch of synthetic pods: feasible mask (resource_fits +
nodeSelector) + feasible count + argmax-score binding for every pod of the batch against the resident node
snapshot.  Workload at N=1 = BASELINE.json configs[1]: 100k pods x 10k nodes (1e9 cells), SURVEY.md §8d
generator, seed 0xB2000002.  With N>1 every rank takes its own 100k-pod shard of an N*100k-pod batch against
the replicated node table (pods-dimension sharding, weak scaling) and the step ends with ONE NCCL all-gather
of the packed per-pod bindings.

  value     : cells/s with the pod batch already resident in HBM (device-space ks_select), CUDA events,
              max over ranks; L2 is flushed between timed iterations.
  e2e       : the same pass through the C ABI with HOST buffers (pinned): H2D of the pod batch and D2H of the
              bindings inside the timed region (the 125 MB feasible mask is produced but stays in HBM).
  roofline  : dominant kernel's algorithmic bytes / its CUDA-event duration vs MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline / --impl reference : the CPU restatement of the reference's per-cell path (oracle/, string
              parsing + per-cell re-summation of bound pods) on the host cores.  The Rust reference itself
              cannot be built in this image (no rustc/cargo), so kind = "port".
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "pod_node_predicate_cells_per_sec"
UNIT = "cells/s"
FALLBACK_HBM_GBS = 6650.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3"])
    ap.add_argument("--path", default="auto", choices=["auto", "direct", "bitpar"])
    ap.add_argument("--policy", default="leftover", choices=["leftover", "least_allocated"])
    ap.add_argument("--no-mask", action="store_true", help="do not emit the feasible mask (bindings only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work budget of the baseline sample")
    return ap.parse_args()


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(prefix="ks_clocks_", suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            out["reasons"] = ["nvidia-smi unavailable"]
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    smax.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(smax), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def algorithmic_bytes(P, N, W, B, mask):
    """every input read once, every output written once."""
    pods = P * (16 + 8 * W)
    nodes = N * (32 + 8 * W)
    bound = B * 20
    outs = P * 16
    m = P * ((N + 7) // 8) if mask else 0
    return {"pods": pods, "nodes": nodes, "bound": bound, "bindings": outs, "mask": m,
            "step_total": pods + nodes + bound + outs + m,
            "dominant_kernel": pods + nodes + m + P * 4}  # mask kernel: pods + node index in, mask + counts out


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_arm(cl, ks, seconds, policy, threads=0):
    """Time the oracle's faithful per-cell path (string parse + bound re-sum per cell, like
    src/predicates.rs:20-77) on a bounded pod sample of the SAME workload.  Returns (cells/s, cores, text)."""
    from oracle import orc
    cores = orc.lib.orc_online_cores() if threads <= 0 else threads
    nodes_s, bound_s, _ = ks.objects.cluster_specs(cl, pod_slice=slice(0, 0))
    arena = ks.objects.ObjectArena()
    nodes, bound = arena.nodes(nodes_s), arena.pods(bound_s)
    oc = orc.Cluster(nodes, cl.N, bound, cl.B)
    probe_n = max(4 * cores, 16)

    def run(first, count):
        _, _, pods_s = ks.objects.cluster_specs(cl, pod_slice=slice(first, first + count))
        pods = arena.pods(pods_s)
        t0 = time.perf_counter()
        oc.run(pods, count, policy=policy, want_mask=True, nthreads=cores)
        return time.perf_counter() - t0

    t_probe = run(0, probe_n)
    per_pod = t_probe / probe_n
    n = int(max(probe_n, min(cl.P - probe_n, seconds / max(per_pod, 1e-9))))
    t = run(probe_n, n)
    cells = n * cl.N
    return cells / t, cores, n, t


def reference_main(args):
    """--impl reference: the CPU restatement of the reference's own per-cell path on the host cores.
    Under torchrun only rank 0 works."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import ksched_pkg
    ks = ksched_pkg.load()  # host-side object rendering only; no GPU entry point is called on this arm
    from oracle import orc
    cl = ks.synth.config(args.workload)
    policy = 0 if args.policy == "leftover" else 1
    cores = orc.lib.orc_online_cores()
    nodes_s, bound_s, _ = ks.objects.cluster_specs(cl, pod_slice=slice(0, 0))
    arena = ks.objects.ObjectArena()
    nodes, bound = arena.nodes(nodes_s), arena.pods(bound_s)
    oc = orc.Cluster(nodes, cl.N, bound, cl.B)
    # size one step to ~4 s of wall time with all cores
    probe = max(cores, 4)
    _, _, ps = ks.objects.cluster_specs(cl, pod_slice=slice(0, probe))
    pp = arena.pods(ps)
    t0 = time.perf_counter()
    oc.run(pp, probe, policy=policy, nthreads=cores)
    per_pod = (time.perf_counter() - t0) / probe
    budget = min(4.0, 150.0 / max(1, args.steps + args.warmup))
    n = int(max(probe, min(2000, budget / max(per_pod, 1e-9))))
    times = []
    for it in range(args.warmup + args.steps):
        first = (probe + it * n) % max(1, cl.P - n)
        _, _, ps = ks.objects.cluster_specs(cl, pod_slice=slice(first, first + n))
        pods = arena.pods(ps)
        t0 = time.perf_counter()
        oc.run(pods, n, policy=policy, want_mask=True, nthreads=cores)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = n * cl.N / (ms * 1e-3)
    sample = f"{n} pods x {cl.N} nodes per step ({n * cl.N} cells) of workload {args.workload}, all {cores} host threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {cl.P} pods x {cl.N} nodes, resource_fits + nodeSelector, policy {args.policy}",
                   "note": "CPU restatement (oracle/) of the reference per-cell path; the Rust reference cannot be "
                           "built here (no rustc/cargo); each step is a bounded pod sample"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    args = parse_args()
    if args.impl == "reference":
        return reference_main(args)

    import torch
    import torch.distributed as dist
    import ksched_pkg
    ks = ksched_pkg.load()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or ks.device_count() == 0:
        print(json.dumps({"error": "no CUDA device: bench.py has no CPU fallback for the product arm"}), flush=True)
        return 2
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    policy = ks.KS_SCORE_LEFTOVER if args.policy == "leftover" else ks.KS_SCORE_LEAST_ALLOCATED
    flags = {"auto": ks.KS_SELECT_AUTO, "direct": ks.KS_SELECT_FORCE_DIRECT, "bitpar": ks.KS_SELECT_FORCE_BITPAR}[args.path]
    emit_mask = not args.no_mask

    # ---- workload: node snapshot replicated, pods sharded (rank r takes shard r of world*P pods) ----
    P_shard, N = ks.synth.SHAPES[args.workload]
    seed = ks.synth.SEEDS[args.workload]
    cl_all = ks.synth.make(P_shard * world, N, seed)
    cl = cl_all.take_pods(rank * P_shard, P_shard)
    ac, am, lab, bn, bc, bm, rc, rm, sel = cl.packed()
    W = cl.label_words
    snap = ks.Snapshot(local)
    snap.set_nodes(ac, am, lab)
    snap.set_bound(bn, bc, bm)
    P = cl.P

    stream = torch.cuda.Stream()
    row = ks.mask_row_bytes(N)
    d_rc = torch.from_numpy(rc).to(dev)
    d_rm = torch.from_numpy(rm).to(dev)
    d_sel = torch.from_numpy(sel.view(np.int64)).to(dev)
    # bindings of one shard packed in one buffer: [score i64 | node_idx i32 | feasible_cnt u32] = 16 B per pod
    d_bind = torch.empty(P * 16, dtype=torch.uint8, device=dev)
    p_score = d_bind.data_ptr()
    p_idx = p_score + 8 * P
    p_cnt = p_idx + 4 * P
    d_mask = torch.empty((P, row), dtype=torch.uint8, device=dev) if emit_mask else None
    # the all-gather moves the bindings proper (score i64 | node_idx i32 = the first 12 B per pod of the shard buffer);
    # feasible counts and the mask stay sharded
    d_all = torch.empty(world * P * 12, dtype=torch.uint8, device=dev) if world > 1 else None
    side = torch.cuda.Stream() if world > 1 else None
    ev_ready = torch.cuda.Event() if world > 1 else None
    ev_gathered = torch.cuda.Event() if world > 1 else None
    if world > 1:
        ev_ready.record(stream)  # materialise the underlying cudaEvent handles
        ev_gathered.record(stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    h_rc = torch.from_numpy(rc).pin_memory()
    h_rm = torch.from_numpy(rm).pin_memory()
    h_sel = torch.from_numpy(sel.view(np.int64)).pin_memory()
    h_bind = torch.empty(P * 16, dtype=torch.uint8).pin_memory()

    def step_resident(timing):
        snap.select_raw(P, d_rc, d_rm, d_sel, ks.KS_MEM_DEVICE, p_idx, p_score, p_cnt, ks.KS_MEM_DEVICE,
                        mask=d_mask, mask_row_bytes=row if emit_mask else 0, mask_space=ks.KS_MEM_DEVICE,
                        policy=policy, flags=flags | (ks.KS_SELECT_TIMING if timing else 0), stream=stream.cuda_stream,
                        ready_event=ev_ready.cuda_event if world > 1 else None)
        if world > 1:
            # the ONE collective of the step runs on a side stream as soon as the library signals that node_idx and
            # score are final, i.e. under the mask kernel; the step ends when both the select and the gather are done
            side.wait_event(ev_ready)
            with torch.cuda.stream(side):
                ks.multigpu.all_gather_bindings(d_bind[:12 * P], d_all)
                ev_gathered.record(side)
            stream.wait_event(ev_gathered)

    def step_e2e():
        hp = h_bind.data_ptr()
        snap.select_raw(P, h_rc, h_rm, h_sel, ks.KS_MEM_HOST, hp + 8 * P, hp, hp + 12 * P, ks.KS_MEM_HOST,
                        mask=d_mask, mask_row_bytes=row if emit_mask else 0, mask_space=ks.KS_MEM_DEVICE,
                        policy=policy, flags=flags, stream=stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # clocks are sampled over warm-up + both timed loops (the timed regions alone are a few ms)
    sampler = ClockSampler(local) if rank == 0 else None
    # ---- warm-up ----
    for _ in range(max(args.warmup, 3)):
        step_resident(False)
        step_e2e()
    # The timed region lasts only a few ms, shorter than nvidia-smi's sampling period, so the same step is kept
    # running (untimed) for ~0.4 s right before it: the clock / throttle samples are taken under exactly this load.
    # (the iteration count is fixed by rank 0 and broadcast: every rank must issue the same number of collectives)
    t_soak = time.perf_counter()
    for _ in range(5):
        step_resident(False)
    stream.synchronize()
    n_soak = torch.tensor([max(5, min(20000, int(0.4 / max((time.perf_counter() - t_soak) / 5, 1e-6))))],
                          dtype=torch.int64, device=dev)
    if world > 1:
        dist.broadcast(n_soak, src=0)
    for _ in range(int(n_soak.item())):
        step_resident(False)
    stream.synchronize()
    barrier()

    # ---- timed: K resident steps (CUDA events on the launching stream, L2 flushed between iterations) ----
    launches0 = ks.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kern_ms, scan_ms, call_ms = [], [], []
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        with torch.cuda.stream(stream):
            flush.fill_(k & 0xFF)
            ev[k][0].record(stream)
        step_resident(False)  # all-device call: the library replays its cached CUDA graph
        with torch.cuda.stream(stream):
            ev[k][1].record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = ks.launch_count() - launches0
    # same K steps again with per-kernel CUDA events inside the library (dominant-kernel duration for the roofline;
    # in this mode the library runs the argmax scan after the mask kernel instead of beside it, so the event pair
    # times k_mask_bitpar alone - the measured HBM peak it is compared with is also a kernel timed alone)
    for k in range(args.steps):
        with torch.cuda.stream(stream):
            flush.fill_(k & 0xFF)
        step_resident(True)
        stream.synchronize()
        t = snap.last_timings()
        kern_ms.append(t[0])
        scan_ms.append(t[1])
        call_ms.append(t[2])
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(total_ms.item()) / args.steps
    cells_per_step = world * P * N
    value = cells_per_step / (ms_per_step * 1e-3)

    # ---- timed: K end-to-end steps (host buffers, copies inside; wall clock around the blocking call) ----
    barrier()
    e2e_t = []
    for k in range(args.steps):
        with torch.cuda.stream(stream):
            flush.fill_(k & 0xFF)
        stream.synchronize()
        t0 = time.perf_counter()
        step_e2e()
        e2e_t.append(time.perf_counter() - t0)
    barrier()
    e2e_ms = torch.tensor([1e3 * sum(e2e_t)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = cells_per_step / (float(e2e_ms.item()) / args.steps * 1e-3)
    clocks = sampler.stop() if sampler else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # sanity: the resident and e2e passes produced the same bindings (cheap guard against a skipped pass)
    torch.cuda.synchronize()
    assert torch.equal(d_bind.cpu(), h_bind), "resident and e2e bindings differ"
    if world > 1:  # and the all-gather delivered every shard: rank 0's own slice matches, every pod was decided
        g = d_all.cpu().numpy().reshape(world, 12 * P)
        mine = d_bind.cpu().numpy()
        assert np.array_equal(g[0], mine[:12 * P]), "all-gather did not deliver rank 0's own bindings"
        mi, _, mc = ks.multigpu.unpack_bindings(mine, P, 1)
        assert np.array_equal(mi < 0, mc == 0)
        for r in range(world):  # every shard arrived: node indices are in range, scores of unbound pods are zero
            gi = g[r, 8 * P:12 * P].view(np.int32)
            gs = g[r, :8 * P].view(np.int64)
            assert gi.min() >= -1 and gi.max() < N and (gs[gi < 0] == 0).all()

    ab = algorithmic_bytes(P, N, W, cl.B, emit_mask)
    peak, peak_src = hbm_peak()
    k_ms = sum(kern_ms) / len(kern_ms)
    achieved = ab["dominant_kernel"] / (k_ms * 1e-3) / 1e9
    traffic = None  # dram__bytes_read+write of the dominant kernel from the committed ncu --set full capture
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(f"{args.workload}_{snap.last_path()}")
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": snap.last_path(), "kernel_ms": k_ms, "rest_of_step_ms": sum(scan_ms) / len(scan_ms),
                "algorithmic_bytes": ab["dominant_kernel"], "peak_source": peak_src}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64", "data": "synthetic",
        "config": {
            "workload": f"{args.workload}: {P} pods/GPU x {N} nodes ({P * N:.3g} cells/GPU), resource_fits + nodeSelector "
                        f"+ argmax score ({args.policy}), mask {'emitted' if emit_mask else 'not emitted'}",
            "label_words": W, "bound_pods": cl.B, "seed": hex(seed), "path": snap.last_path(),
            "parallelism": f"pods sharded x{world}, node table replicated" + (", 1 NCCL all-gather of bindings/step (side stream, under the mask kernel)" if world > 1 else ""),
            "l2": "256 MiB flush write between timed iterations", "wall_s_timed_region": t_wall,
            "clocks_window": "0.4 s untimed soak of the same step + both timed loops (timed region alone is a few ms)",
            "call_ms_inside_library": sum(call_ms) / len(call_ms),
        },
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": P * (16 + 8 * W), "d2h_bytes_per_step": P * 16,
                "note": "host pinned pods in, bindings out; feasible mask written to HBM, not copied to host"},
        "gpu_launches": int(launches),
        "roofline": roofline,
    }
    if world == 1:
        # for information: the reference's own policy (<=5 seeded draws per pod, src/main.rs:49-71) on the same batch,
        # host buffers in and out; it evaluates <=5 cells per pod, so it is quoted in pods/s, not cells/s
        snap.select_sampling(rc[:1024], rm[:1024], sel[:1024], seed=seed)
        t0 = time.perf_counter()
        s_idx, s_used, _, _ = snap.select_sampling(rc, rm, sel, seed=seed)
        t_s = time.perf_counter() - t0
        a_idx = h_bind[8 * P:12 * P].numpy().view(np.int32)
        line["reference_policy"] = {"pods_per_s": P / t_s, "cells_evaluated": int(s_used.sum()),
                                    "bound_frac": float((s_idx >= 0).mean()), "argmax_bound_frac": float((a_idx >= 0).mean()),
                                    "note": "ks_select_sampling, ATTEMPTS=5, seeded; not part of value/e2e"}
    if world == 1 and not args.no_cpu_baseline:
        # second CPU figure (BASELINE.md §3 "CPU-packed"): the same SoA int64 + bitmask algorithm the GPU runs, C,
        # all host threads, on a pod sample — what a well-written CPU scheduler core could do with packed inputs
        from oracle import orc
        n_pk = min(P, 20000)
        fc_h, fm_h = cl.free()
        t0 = time.perf_counter()
        orc.run_packed(fc_h, fm_h, ac, am, lab, rc[:n_pk], rm[:n_pk], sel[:n_pk], policy=policy, want_mask=True, nthreads=0)
        t_pk = time.perf_counter() - t0
        cps, cores, n, t = cpu_reference_arm(cl, ks, args.cpu_seconds, policy)
        line["cpu_baseline"] = {"value": cps, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"first {n} pods x {N} nodes ({n * N} cells, {t:.1f} s) of the same workload, "
                                          f"faithful per-cell path (quantity parse + bound-pod re-sum per cell)",
                                "packed_soa": {"value": n_pk * N / t_pk, "unit": UNIT, "cores": cores,
                                               "sample": f"first {n_pk} pods x {N} nodes, same SoA/bitmask algorithm in C "
                                                         f"(oracle packed flavour), mask + argmax, {t_pk:.2f} s"}}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
