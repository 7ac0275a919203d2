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
        r = snap.select(rc[:50], rm[:50], sel[:50], policy=ks.KS_SCORE_LEAST_ALLOCATED)
        idx, score, rounds = snap.stream_bind(rc[:200], rm[:200], sel[:200])
        assert rounds >= 1
print("sanitizer smoke ok")
