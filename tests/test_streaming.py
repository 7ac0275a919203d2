"""Config C5 (streaming reconcile): K3 claim resolution + capacity commit vs the oracle restatement."""
import numpy as np
import pytest


def test_oracle_commit_never_overcommits(orc):
    fc = np.array([1000, 500], np.int64)
    fm = np.array([100, 100], np.int64)
    acc = orc.commit_claims(fc, fm, [0, 0, 1, 0, -1, 1], [600, 600, 500, 400, 1, 1], [10, 10, 100, 10, 1, 1])
    assert list(acc) == [1, 0, 1, 1, 0, 0]
    assert list(fc) == [0, 0] and list(fm) == [80, 0]


@pytest.mark.gpu
@pytest.mark.parametrize("P,N,seed", [(50, 40, 1), (600, 300, 2), (5000, 64, 3), (4500, 3000, 4)])
def test_commit_claims_matches_oracle(ks, orc, P, N, seed):
    cl = ks.synth.make(P, N, seed=seed, bound_per_node=2)
    ac, am, lab, bn, bc, bm, rc, rm, sel = cl.packed()
    rng = np.random.default_rng(seed)
    claims = rng.integers(-1, N, size=P).astype(np.int32)  # many pods pile onto few nodes
    with ks.Snapshot(0) as snap:
        snap.set_nodes(ac, am, lab)
        snap.set_bound(bn, bc, bm)
        fc, fm = snap.free()
        acc = snap.commit_claims(claims, rc, rm)
        oacc = orc.commit_claims(fc, fm, claims, rc, rm)
        gfc, gfm = snap.free()
    assert np.array_equal(acc, oacc)
    assert np.array_equal(gfc, fc) and np.array_equal(gfm, fm)
    assert acc.sum() > 0 and (acc == 0).sum() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("policy", [0, 1])
@pytest.mark.parametrize("P,N,seed,keys", [(40, 12, 5, 8), (800, 200, 6, 8), (3000, 1500, 7, 8), (1024, 9, 11, 8), (1, 20000, 12, 8),
                                           (700, 2600, 13, 32)])
def test_stream_bind_matches_oracle_and_never_overcommits(ks, orc, P, N, seed, keys, policy):
    """Batches of <= 1024 pods run the device-side loop (k_stream_batch: one cooperative launch for all rounds), larger
    ones the host-driven loop (per-cell select + K3 per round); both must equal the oracle's round protocol."""
    cl = ks.synth.make(P, N, seed=seed, bound_per_node=3, n_keys=keys)
    ac, am, lab, bn, bc, bm, rc, rm, sel = cl.packed()
    with ks.Snapshot(0) as snap:
        snap.set_nodes(ac, am, lab)
        snap.set_bound(bn, bc, bm)
        fc0, fm0 = snap.free()
        idx, score, rounds = snap.stream_bind(rc, rm, sel, policy=policy)
        fc1, fm1 = snap.free()
        # a second micro-batch sees the committed capacity
        idx2, _, _ = snap.stream_bind(rc[:100], rm[:100], sel[:100], policy=policy)
        assert snap.last_path() == "stream_batch"
        fc2, fm2 = snap.free()
    ofc, ofm = fc0.copy(), fm0.copy()
    oidx, oscore, orounds = orc.stream_bind_packed(ofc, ofm, ac, am, lab, rc, rm, sel, policy=policy)
    assert np.array_equal(idx, oidx) and np.array_equal(score, oscore) and rounds == orounds
    assert np.array_equal(fc1, ofc) and np.array_equal(fm1, ofm)
    oidx2, _, _ = orc.stream_bind_packed(ofc, ofm, ac, am, lab, rc[:100], rm[:100], sel[:100], policy=policy)
    assert np.array_equal(idx2, oidx2) and np.array_equal(fc2, ofc) and np.array_equal(fm2, ofm)
    # capacity never goes negative through this path (nodes that started >= 0 stay >= 0)
    ok = fc0 >= 0
    assert (fc1[ok] >= 0).all() and (fm1[fm0 >= 0] >= 0).all()
    # conservation: what left free[] is exactly what the bound pods requested
    bound = idx >= 0
    assert fc0.sum() - fc1.sum() == rc[bound].sum() and fm0.sum() - fm1.sum() == rm[bound].sum()
    assert rounds >= 1 and bound.sum() > 0


@pytest.mark.gpu
def test_async_stream_submit_poll(ks, orc):
    """ks_stream (the reference's Controller queue, src/main.rs:73,141-148): tickets come back exactly once; with a
    flush after every submit the batch boundaries are the submit calls, so the result equals the oracle's round
    protocol batch by batch; free-running submits never oversubscribe and conserve capacity."""
    cl = ks.synth.make(2400, 500, seed=91, bound_per_node=3)
    ac, am, lab, bn, bc, bm, rc, rm, sel = cl.packed()
    with ks.Snapshot(0) as snap:
        snap.set_nodes(ac, am, lab)
        snap.set_bound(bn, bc, bm)
        fc0, fm0 = snap.free()
        ofc, ofm = fc0.copy(), fm0.copy()
        got = np.full(cl.P, -2, np.int32)
        with ks.Stream(snap) as q:
            lo = 0
            for size in (1, 7, 300, 900, 2, 190):  # deterministic boundaries: flush after each submit
                hi = lo + size
                q.submit(rc[lo:hi], rm[lo:hi], sel[lo:hi], np.arange(lo, hi, dtype=np.uint64))
                q.flush()
                t, i, s = q.poll()
                assert sorted(t.tolist()) == list(range(lo, hi))
                oidx, oscore, _ = orc.stream_bind_packed(ofc, ofm, ac, am, lab, rc[lo:hi], rm[lo:hi], sel[lo:hi])
                order = np.argsort(t)
                assert np.array_equal(i[order], oidx) and np.array_equal(s[order], oscore)
                got[lo:hi] = oidx
                lo = hi
            # free-running: many small submits, poll while the dispatcher works
            seen = []
            for a in range(lo, cl.P, 37):
                b = min(cl.P, a + 37)
                q.submit(rc[a:b], rm[a:b], sel[a:b], np.arange(a, b, dtype=np.uint64))
                t, i, _ = q.poll()
                seen += list(zip(t.tolist(), i.tolist()))
            q.flush()
            t, i, _ = q.poll()
            seen += list(zip(t.tolist(), i.tolist()))
            assert sorted(x[0] for x in seen) == list(range(lo, cl.P))
            for tk, nd in seen:
                got[tk] = nd
            batches, rounds, max_seen = q.stats()
            assert batches >= 7 and max_seen <= 1024
        fc1, fm1 = snap.free()
    bound = got >= 0
    assert (got >= -1).all() and bound.sum() > 0
    assert fc0.sum() - fc1.sum() == rc[bound].sum() and fm0.sum() - fm1.sum() == rm[bound].sum()
    assert (fc1[fc0 >= 0] >= 0).all() and (fm1[fm0 >= 0] >= 0).all()
    used_c = np.zeros(cl.N, np.int64)
    np.add.at(used_c, got[bound], rc[bound])
    assert np.array_equal(fc0 - used_c, fc1)
