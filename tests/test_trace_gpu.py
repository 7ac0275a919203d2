"""ks_last_trace: the step timeline stamped by the kernels themselves (KS_TRACE=1), include/ksched.h."""
import os

import numpy as np
import pytest


@pytest.mark.gpu
def test_step_trace_is_ordered_and_results_unchanged(ks, orc):
    cl = ks.synth.make(6000, 5000, seed=77, bound_per_node=2)
    ac, am, lab, bn, bc, bm, rc, rm, sel = cl.packed()
    with ks.Snapshot(0) as plain:
        plain.set_nodes(ac, am, lab)
        plain.set_bound(bn, bc, bm)
        ref = plain.select(rc, rm, sel, flags=ks.KS_SELECT_FORCE_BITPAR, want_mask=True)
        with pytest.raises(ks.KsError):
            plain.last_trace()  # the process was not started with KS_TRACE=1 and this snapshot is already prepared
    os.environ["KS_TRACE"] = "1"  # read when a snapshot prepares its first bit-parallel select
    try:
        with ks.Snapshot(0) as snap:
            snap.set_nodes(ac, am, lab)
            snap.set_bound(bn, bc, bm)
            for policy in (ks.KS_SCORE_LEFTOVER, ks.KS_SCORE_LEAST_ALLOCATED):
                res = snap.select(rc, rm, sel, policy=policy, flags=ks.KS_SELECT_FORCE_BITPAR, want_mask=True)
                tr = snap.last_trace()
                assert res.path == "bitpar"
                assert np.array_equal(res.mask, ref.mask) and np.array_equal(res.feasible_cnt, ref.feasible_cnt)
                if policy == ks.KS_SCORE_LEFTOVER:
                    assert np.array_equal(res.node_idx, ref.node_idx) and np.array_equal(res.score, ref.score)
                for k in ("pod_ranks", "argmax1", "mask"):
                    a, b = tr[k]
                    assert 0 <= a <= b < 1e6, (k, tr)  # microseconds from the first stamp; a step takes well under a second
                assert tr["pod_ranks"][0] == 0.0
                assert tr["mask"][0] >= tr["pod_ranks"][1] - 1.0 and tr["argmax1"][0] >= tr["pod_ranks"][1] - 1.0
                assert tr["mask"][0] <= tr["mask_first_cta_end"] <= tr["mask"][1]
                if policy == ks.KS_SCORE_LEFTOVER:
                    assert tr["argmax2"][0] >= tr["argmax1"][1] - 1.0  # the tail kernel follows the head kernel
                else:
                    assert "argmax2" not in tr
    finally:
        del os.environ["KS_TRACE"]
