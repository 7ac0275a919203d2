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


class _OracleSnap:
    """Stand-in for ks.Snapshot in the distributed streaming protocol test (CPU, no kernels)."""

    def __init__(self, orc, cl):
        self.orc = orc
        ac, am, lab, bn, bc, bm, *_ = cl.packed()
        self.ac, self.am, self.lab = ac, am, lab
        self.fc, self.fm = orc.free_reduce(ac, am, bn, bc, bm)

    def select(self, rc, rm, sel, policy=0, flags=0):
        idx, score, cnt, _, _ = self.orc.run_packed(self.fc, self.fm, self.ac, self.am, self.lab, rc, rm, sel,
                                                     policy=policy, want_mask=False, nthreads=1)

        class R:
            node_idx = idx
        return R

    def commit_claims(self, node, cpu, mem):
        return self.orc.commit_claims(self.fc, self.fm, node, cpu, mem)


def _stream_worker(rank, world, port, P, N, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import ksched_pkg
    ks = ksched_pkg.load()
    from oracle import orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cl = ks.synth.make(P, N, seed=777, bound_per_node=3)
    _, _, _, _, _, _, rc, rm, sel = cl.packed()
    snap = _OracleSnap(orc, cl)
    mine = np.arange(P)[np.arange(P) % world == rank]
    half = len(mine) // 2
    out = np.full(P, -1, np.int32)
    # two lockstep micro-batches; rank 1 has nothing in the second one
    b1 = mine[:half] if rank == 0 else mine
    b2 = mine[half:] if rank == 0 else mine[:0]
    i1, _, d1 = ks.multigpu.stream_bind_distributed(snap, rc[b1], rm[b1], sel[b1], b1, done=(rank != 0))
    out[b1] = i1
    i2, _, d2 = ks.multigpu.stream_bind_distributed(snap, rc[b2], rm[b2], sel[b2], b2, done=True)
    out[b2] = i2
    q.put((rank, out, snap.fc.copy(), snap.fm.copy(), bool(d1), bool(d2), b1.tolist(), b2.tolist()))
    dist.destroy_process_group()


def test_two_rank_streaming_protocol_matches_single_stream(ks, orc):
    """Union of the ranks' claims committed in global arrival order == one process streaming the merged batches."""
    import torch.multiprocessing as mp
    P, N = 400, 60
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_stream_worker, args=(r, 2, port, P, N, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
    # replicas identical
    assert np.array_equal(res[0][2], res[1][2]) and np.array_equal(res[0][3], res[1][3])
    assert res[0][4] is False and res[0][5] is True  # rank 0 still had arrivals after the first batch
    # single-process reference: batch 1 = union of first batches in arrival order, then batch 2
    cl = ks.synth.make(P, N, seed=777, bound_per_node=3)
    ac, am, lab, bn, bc, bm, rc, rm, sel = cl.packed()
    fc, fm = orc.free_reduce(ac, am, bn, bc, bm)
    exp = np.full(P, -1, np.int32)
    for batch in (sorted(res[0][6] + res[1][6]), sorted(res[0][7] + res[1][7])):
        b = np.asarray(batch, np.int64)
        if len(b):
            exp[b], _, _ = orc.stream_bind_packed(fc, fm, ac, am, lab, rc[b], rm[b], sel[b])
    got = np.where(res[0][1] >= 0, res[0][1], res[1][1])
    assert np.array_equal(got, exp)
    assert np.array_equal(res[0][2], fc) and np.array_equal(res[0][3], fm)
