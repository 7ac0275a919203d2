"""N>1 host-side logic on CPU: world_size-2 gloo ranks shard the pods, evaluate their shard (the CPU oracle
stands in for the GPU kernel here — this test is about sharding + the single all-gather, not about kernels),
gather the packed bindings and must reproduce the unsharded answer on every rank."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, P, N, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import ksched_pkg
    ks = ksched_pkg.load()
    from oracle import orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cl = ks.synth.make(P, N, seed=4242)
    ac, am, lab, bn, bc, bm, rc, rm, sel = cl.packed()
    fc, fm = orc.free_reduce(ac, am, bn, bc, bm)
    lo, hi = ks.multigpu.shard_bounds(P, world, rank)
    idx, score, cnt, _, _ = orc.run_packed(fc, fm, ac, am, lab, rc[lo:hi], rm[lo:hi], sel[lo:hi], want_mask=False,
                                           nthreads=1)
    cap = ks.multigpu.shard_capacity(P, world)
    local = torch.from_numpy(ks.multigpu.pack_bindings(cap, idx, score, cnt))
    gathered = ks.multigpu.all_gather_bindings(local)
    gi, gs, gc = ks.multigpu.unpack_bindings(gathered.numpy(), P, world)
    fi, fs, fcn, _, _ = orc.run_packed(fc, fm, ac, am, lab, rc, rm, sel, want_mask=False, nthreads=1)
    ok = np.array_equal(gi, fi) and np.array_equal(gs, fs) and np.array_equal(gc, fcn)
    q.put((rank, bool(ok), int(hi - lo)))
    dist.destroy_process_group()


@pytest.mark.parametrize("P", [1001, 64])
def test_two_rank_gloo_shard_and_gather(P):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + P % 7
    procs = [ctx.Process(target=_worker, args=(r, 2, port, P, 300, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res)
    assert sum(r[2] for r in res) == P


def test_shard_bounds_cover_exactly(ks):
    for n in (0, 1, 7, 100, 1_000_003):
        for world in (1, 2, 3, 8):
            spans = [ks.multigpu.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
            assert max(h - l for l, h in spans) <= ks.multigpu.shard_capacity(n, world)
