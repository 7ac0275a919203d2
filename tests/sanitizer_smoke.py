"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): both kernel paths, streaming,
host layer, on shapes that finish quickly under the tool.  Usage on a GPU box:
    compute-sanitizer --tool memcheck python tests/sanitizer_smoke.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ksched_pkg  # noqa: E402

ks = ksched_pkg.load()
from oracle import orc  # noqa: E402

for (P, N, keys) in [(300, 700, 8), (70000, 300, 8), (129, 2049, 32)]:
    cl = ks.synth.make(P, N, seed=99 + P, n_keys=keys, bound_per_node=3)
    ac, am, lab, bn, bc, bm, rc, rm, sel = cl.packed()
    with ks.Snapshot(0) as snap:
        snap.set_nodes(ac, am, lab)
        snap.set_bound(bn, bc, bm)
        fc, fm = snap.free()
        o = orc.run_packed(fc, fm, ac, am, lab, rc, rm, sel)
        for flag in (ks.KS_SELECT_FORCE_DIRECT, ks.KS_SELECT_FORCE_BITPAR):
            r = snap.select(rc, rm, sel, flags=flag, want_mask=True)
            assert np.array_equal(r.node_idx, o[0]) and np.array_equal(r.mask, o[3]) and np.array_equal(r.feasible_cnt, o[2])
        for flag in (ks.KS_SELECT_FORCE_DIRECT, ks.KS_SELECT_FORCE_BITPAR):  # non-separable score on both paths
            r = snap.select(rc[:50], rm[:50], sel[:50], policy=ks.KS_SCORE_LEAST_ALLOCATED, flags=flag)
        s_idx, s_used, s_dn, s_dc = snap.select_sampling(rc, rm, sel, seed=7)  # the reference's own policy (K1s)
        assert ((s_idx >= 0) | (s_used == 5)).all()
        idx, score, rounds = snap.stream_bind(rc[:200], rm[:200], sel[:200])  # device-side loop (k_stream_batch)
        assert rounds >= 1 and snap.last_path() == "stream_batch"
        with ks.Stream(snap) as q:  # async surface
            k0 = min(200, P - 60)
            q.submit(rc[k0:k0 + 60], rm[k0:k0 + 60], sel[k0:k0 + 60], np.arange(60, dtype=np.uint64))
            q.flush()
            assert len(q.poll()[0]) == 60
# object level: pack -> upload -> micro-batch loop -> commits
cl = ks.synth.make(300, 40, seed=5, bound_per_node=2)
nodes_s, bound_s, pods_s = ks.objects.cluster_specs(cl)
arena = ks.objects.ObjectArena()
with ks.host.Context(0) as ctx:
    ctx.set_nodes(arena.nodes(nodes_s), cl.N)
    ctx.set_cluster_pods(arena.pods(bound_s), cl.B)
    status, node, bodies, rounds = ctx.reconcile_batch(arena.pods(pods_s), cl.P)
    assert rounds >= 1 and (node >= 0).any()
print("sanitizer smoke ok")
