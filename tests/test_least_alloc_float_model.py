"""CPU model of k_least_alloc's single-precision scoring (csrc/ks_bitpar.cu: k_build_eval + the per-tile loop), in numpy
float32 = the same IEEE operations the kernel issues.  It pins the two claims the kernel's shortcut rests on, against the
oracle's integer arithmetic (test infrastructure, like the oracle itself; no GPU):

  1. for every feasible cell the interval [lo, hi] the float path derives contains the exact LeastAllocated score, and
     lo == hi (score "known" without the 64-bit row) only where it EQUALS the exact score;
  2. the windowed, bound-ordered scan that scores only candidates (hi >= max(lo) of the tile and >= the best so far; known
     scores as they are, unknown ones by the exact arithmetic) returns the oracle's binding - node and score - for every pod.

The GPU tests (test_gpu_parity.py: random, adversarial, near-ties, round fills) check the kernel itself; this file checks
the reasoning, at sizes the model finishes in seconds."""
import numpy as np
import pytest

f32 = np.float32
INF = f32(np.inf)


def _hint(fc, fm, ac, am):
    """k_build_eval: (x, y, z, w) = (free_cpu, alloc_cpu, free_mem*100/alloc_mem, 100/alloc_mem) as floats."""
    x = np.where(np.abs(fc) < (1 << 24), fc.astype(np.float64), np.nan).astype(f32)
    y = np.where(ac <= 0, np.inf, np.where(ac < (1 << 24), ac.astype(np.float64), np.nan)).astype(f32)
    inv_am = np.where(am > 0, 1.0 / np.where(am > 0, am, 1).astype(np.float64), 0.0)
    z = np.where(am > 0, fm.astype(np.float64) * 100.0 * inv_am, 0.0).astype(f32)
    w = (100.0 * inv_am).astype(f32)
    return x, y, z, w


def _intervals(h, idx, rc, rm):
    """The per-slot float arithmetic of the kernel's tile loop: (lo, hi, ok)."""
    x, y, z, w = (a[idx] for a in h)
    rcf, rmf = f32(rc), f32(rm)
    rc_exact = abs(int(rc)) < (1 << 24)
    with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
        t = (x - rcf) * f32(100.0)
        pc = np.floor((t / y).astype(f32))
        tm = rmf * w
        qm = z - tm
        em = f32(2e-6) * (np.abs(z) + np.abs(tm)) + f32(1e-3)
        zero = w == 0
        pm_lo = np.where(zero, f32(0), np.floor(qm - em))
        pm_hi = np.where(zero, f32(0), np.floor(qm + em))
        ok = rc_exact & (t >= 0) & (t < f32(16777216.0)) & (pc < f32(4e6)) & ((np.abs(qm) + em) < f32(4e6))
        lo = np.floor((pc + pm_lo) * f32(0.5))
        hi = np.where(ok, np.floor((pc + pm_hi) * f32(0.5)), INF)
    return np.where(ok, lo, -INF), hi, ok


def _tdiv(a, b):
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b > 0) else -q


def _exact(fc, fm, ac, am, rc, rm):
    """The oracle's expression (oracle.c, KS_SCORE_LEAST_ALLOCATED): truncating int64 divisions, in Python integers."""
    pc = _tdiv((fc - rc) * 100, ac) if ac > 0 else 0
    pm = _tdiv((fm - rm) * 100, am) if am > 0 else 0
    return _tdiv(pc + pm, 2)


def _scan(ac, am, lab, fc, fm, rc, rm, sel):
    """The kernel's control flow: nodes in descending bound order, windows of 8 then 32 tiles of 256, early exit."""
    N = len(ac)
    bound = np.array([_exact(int(a), int(b), int(c), int(d), 0, 0) for a, b, c, d in zip(fc, fm, ac, am)], np.int64)
    order = np.lexsort((np.arange(N), -bound))
    top = bound[order]
    nt = (N + 255) // 256
    amax_c, amax_m = int(ac.max()), int(am.max())
    h = _hint(fc, fm, ac, am)
    out_i, out_s, n_cand, n_exact = [], [], 0, 0
    for p in range(len(rc)):
        r_c, r_m = int(rc[p]), int(rm[p])
        feas = (fc >= r_c) & (fm >= r_m) & np.all((sel[p][None, :] & ~lab) == 0, axis=1)
        bounded = r_c >= 0 and r_m >= 0
        D = (((r_c * 100) // amax_c if amax_c > 0 else 0) + ((r_m * 100) // amax_m if amax_m > 0 else 0)) // 2 if bounded else 0
        best, bidx, k0, win, done = None, -1, 0, 8, False
        while not done and k0 < nt:
            if k0 > 0 and bounded and bidx >= 0 and best > top[k0 * 256] - D:
                break
            for k in range(k0, min(k0 + win, nt)):
                idx = order[k * 256:(k + 1) * 256]
                idx = idx[feas[idx]]
                if len(idx) == 0:
                    continue
                if bounded and bidx >= 0 and best > top[k * 256] - D:
                    done = True
                    break
                lo, hi, ok = _intervals(h, idx, r_c, r_m)
                known = ok & (lo == hi)
                cand = hi >= (lo.max() if ok.any() else -INF)
                if bidx >= 0:
                    cand &= ~(hi < f32(best))
                n_cand += int(cand.sum())
                n_exact += int((cand & ~known).sum())
                for c in np.nonzero(cand)[0]:
                    n = int(idx[c])
                    sc = int(hi[c]) if known[c] else _exact(int(fc[n]), int(fm[n]), int(ac[n]), int(am[n]), r_c, r_m)
                    if bidx < 0 or sc > best or (sc == best and n < bidx):
                        best, bidx = sc, n
            k0 += win
            win = 32
        out_i.append(bidx)
        out_s.append(best if bidx >= 0 else 0)
    return np.array(out_i, np.int32), np.array(out_s, np.int64), n_cand, n_exact


def _families(rng, N, P):
    """(name, fc, fm, ac, am, rc, rm): realistic, round, limit and degenerate value ranges."""
    ac = rng.choice(np.array([4000, 8000, 16000, 32000, 64000, 96000], np.int64), N)
    am = rng.choice(np.array([16, 32, 64, 128, 256, 384], np.int64) << 30, N)
    yield ("realistic", ac * rng.integers(0, 100, N) // 100, (am * rng.random(N)).astype(np.int64), ac, am,
           rng.choice(np.array([0, 50, 100, 250, 500, 1000, 2000, 4000], np.int64), P), rng.integers(0, 1 << 33, P))
    yield ("round fills", ac // 4 * rng.integers(0, 5, N), am // 4 * rng.integers(0, 5, N), ac, am,
           rng.choice(np.array([0, 250, 1000], np.int64), P), rng.choice(np.array([0, 1 << 30, 1 << 31], np.int64), P))
    yield ("api limits", rng.integers(-(1 << 36), 1 << 36, N), rng.integers(-(1 << 55), 1 << 55, N), rng.integers(1, 1 << 36, N),
           rng.integers(1, 1 << 55, N), rng.integers(-(1 << 36), 1 << 36, P), rng.integers(-(1 << 55), 1 << 55, P))
    yield ("tiny allocatable", rng.integers(0, 1 << 25, N), rng.integers(0, 1 << 30, N), rng.integers(1, 4, N), rng.integers(1, 4, N),
           rng.integers(0, 1 << 24, P), rng.integers(0, 1 << 28, P))
    yield ("zero and negative allocatable", rng.integers(-100, 100, N), rng.integers(-100, 1 << 30, N), rng.integers(-5, 5, N),
           rng.integers(-5, 1 << 30, N), rng.integers(-100, 100, P), rng.integers(-100, 1 << 29, P))
    yield ("24-bit edge", rng.integers(0, 1 << 24, N), rng.integers(0, 1 << 40, N), rng.integers(1, 1 << 24, N),
           rng.integers(1, 1 << 40, N), rng.integers(0, 1 << 18, P), rng.integers(0, 1 << 38, P))
    yield ("quotient near 2^24", 167772 - rng.integers(0, 3, N), np.full(N, 1 << 29), rng.integers(1, 200, N), np.full(N, 1 << 30),
           rng.integers(0, 5, P), np.zeros(P, np.int64))


def test_float_intervals_contain_the_exact_score():
    rng = np.random.default_rng(3)
    N, P = 1500, 40
    for name, fc, fm, ac, am, rc, rm in _families(rng, N, P):
        fc, fm, ac, am = (np.asarray(a, np.int64) for a in (fc, fm, ac, am))
        h = _hint(fc, fm, ac, am)
        known_total = 0
        for p in range(P):
            feas = np.nonzero((fc >= rc[p]) & (fm >= rm[p]))[0]
            if len(feas) == 0:
                continue
            lo, hi, ok = _intervals(h, feas, int(rc[p]), int(rm[p]))
            ex = np.array([float(_exact(int(fc[n]), int(fm[n]), int(ac[n]), int(am[n]), int(rc[p]), int(rm[p]))) for n in feas])
            assert (lo.astype(np.float64) <= ex).all() and (ex <= hi.astype(np.float64)).all(), name
            known = ok & (lo == hi)
            assert (hi[known].astype(np.float64) == ex[known]).all(), name
            known_total += int(known.sum())
        if name == "realistic":
            assert known_total > 0  # the shortcut is actually taken there


@pytest.mark.parametrize("case", ["synthetic", "near ties", "adversarial", "round fills"])
def test_scan_model_equals_the_oracle(orc, case):
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("synth_standalone", os.path.join(root, "kube-scheduler-rs-reference_b200", "synth.py"))
    synth = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(synth)
    rng = np.random.default_rng(17)
    if case == "synthetic":
        ac, am, lab, bn, bc, bm, rc, rm, sel = synth.make(150, 6000, 5, bound_per_node=4).packed()
    elif case == "near ties":
        N, P = 4000, 150
        ac = rng.choice(np.array([64000, 64000, 64001, 63999, 32000], np.int64), N)
        am = rng.choice(np.array([1 << 38, (1 << 38) + 4096, 1 << 37], np.int64), N)
        lab = np.zeros((N, 1), np.uint64)
        lab[:, 0] = rng.choice(np.array([1, 3, 7], np.uint64), N)
        bn = np.repeat(np.arange(N, dtype=np.int32), 2)
        bc = (ac[bn] // 4 + rng.integers(0, 400, 2 * N)).astype(np.int64)
        bm = (am[bn] // 4 + rng.integers(0, 1 << 30, 2 * N)).astype(np.int64)
        rc = rng.integers(-50, 16000, P).astype(np.int64)
        rm = rng.integers(0, 1 << 35, P).astype(np.int64)
        sel = np.zeros((P, 1), np.uint64)
        sel[:, 0] = rng.choice(np.array([0, 1, 2, 4, 8], np.uint64), P)
    elif case == "adversarial":
        N, P, lim_c, lim_m = 2500, 150, 1 << 36, 1 << 55
        ac = rng.choice(np.array([-5, 0, 1, 250, 1000, 1000, 4000, 64000, lim_c // 2], np.int64), N)
        am = rng.choice(np.array([-1, 0, 1, 1 << 20, 1 << 30, 1 << 30, 1 << 34, lim_m // 2], np.int64), N)
        lab = np.zeros((N, 1), np.uint64)
        bn = rng.integers(0, N, 3 * N).astype(np.int32)
        bc = rng.choice(np.array([0, 0, 100, 1000, -100], np.int64), 3 * N)
        bm = rng.choice(np.array([0, 1, 1 << 20, 1 << 28, -1], np.int64), 3 * N)
        rc = rng.choice(np.array([-1000, 0, 0, 1, 250, 1000, 1001, 4000, 63999, lim_c], np.int64), P)
        rm = rng.choice(np.array([-1, 0, 1, (1 << 20) - 1, 1 << 20, 1 << 30, (1 << 30) + 1, lim_m], np.int64), P)
        sel = np.zeros((P, 1), np.uint64)
    else:
        N, P = 3000, 150
        ac = rng.choice(np.array([4000, 8000, 16000], np.int64), N)
        am = rng.choice(np.array([16, 32, 64], np.int64) << 30, N)
        lab = np.zeros((N, 1), np.uint64)
        bn = np.arange(N, dtype=np.int32)
        bc = (ac // 4 * rng.integers(0, 4, N)).astype(np.int64)
        bm = (am // 4 * rng.integers(0, 4, N)).astype(np.int64)
        rc = rng.choice(np.array([0, 250, 1000], np.int64), P)
        rm = rng.choice(np.array([0, 1 << 30, 1 << 32], np.int64), P)
        sel = np.zeros((P, 1), np.uint64)
    fc, fm = orc.free_reduce(ac, am, bn, bc, bm)
    o = orc.run_packed(fc, fm, ac, am, lab, rc, rm, sel, policy=1, want_mask=False)
    idx, score, n_cand, n_exact = _scan(ac, am, lab, fc, fm, rc, rm, sel)
    assert np.array_equal(score, o[1]), case
    assert np.array_equal(idx, o[0]), case
    if case == "synthetic":
        assert n_exact * 20 < n_cand  # almost every candidate's score is known from the float path alone
