"""CPU model of the bit-parallel path's exactness argument (csrc/ks_bitpar.cu): `request <= free`
(/root/reference/src/predicates.rs:42) depends only on ORDER, so a pod's request can be replaced by a rank threshold and a
tile's feasibility by one row of a per-tile prefix table.  numpy restatement of the data structures the kernels build -
global sort positions, thresholds by lower bound, tile-local ranks, prefix-table rows, label-pair columns - checked
against the per-cell predicate on tie-heavy, negative and sentinel-padded inputs.  Test infrastructure, no GPU: the GPU
tests check the kernels, this file checks the reasoning they implement."""
import numpy as np
import pytest

TILE = 256
I64_MIN = np.iinfo(np.int64).min


def _build(free):
    """Per resource: global sort position of every node (ties by node index), the sorted values, and per tile the
    nodes' tile-local ranks."""
    n = len(free)
    order = np.lexsort((np.arange(n), free))
    gpos = np.empty(n, np.int64)
    gpos[order] = np.arange(n)
    nt = n // TILE
    local_rank = np.empty(n, np.int64)  # rank of the node among its tile's nodes by gpos
    tile_sorted = np.empty((nt, TILE), np.int64)  # gpos of the tile's nodes, ascending (what the rank tables search)
    for t in range(nt):
        g = gpos[t * TILE:(t + 1) * TILE]
        o = np.argsort(g)
        lr = np.empty(TILE, np.int64)
        lr[o] = np.arange(TILE)
        local_rank[t * TILE:(t + 1) * TILE] = lr
        tile_sorted[t] = g[o]
    return free[order], gpos, local_rank, tile_sorted


def _mask_via_ranks(free_c, free_m, lab, rc, rm, sel):
    sc, gc, lrc, tsc = _build(free_c)
    sm, gm, lrm, tsm = _build(free_m)
    n = len(free_c)
    nt = n // TILE
    W = lab.shape[1]
    out = np.zeros((len(rc), n), bool)
    for p in range(len(rc)):
        g_c = np.searchsorted(sc, rc[p], side="left")  # number of nodes with free < request
        g_m = np.searchsorted(sm, rm[p], side="left")
        row = np.zeros(n, bool)
        for t in range(nt):
            r_c = np.searchsorted(tsc[t], g_c, side="left")  # tile-local rank: tile nodes with gpos < g
            r_m = np.searchsorted(tsm[t], g_m, side="left")
            s = slice(t * TILE, (t + 1) * TILE)
            row[s] = (lrc[s] >= r_c) & (lrm[s] >= r_m)  # row r of the tile's prefix table, both resources
        for w in range(W):  # AND of the node columns of the required (key,value) pairs
            bits = int(sel[p, w])
            b = 0
            while bits:
                if bits & 1:
                    row &= ((lab[:, w] >> np.uint64(b)) & np.uint64(1)).astype(bool)
                bits >>= 1
                b += 1
        out[p] = row
    return out


def _cells(free_c, free_m, lab, rc, rm, sel):
    fit = (rc[:, None] <= free_c[None, :]) & (rm[:, None] <= free_m[None, :])
    match = np.all((sel[:, None, :] & ~lab[None, :, :]) == 0, axis=2)
    return fit & match


@pytest.mark.parametrize("seed,kind", [(1, "ties"), (2, "random"), (3, "negative"), (4, "extremes")])
def test_rank_thresholds_reproduce_the_cell_predicate(seed, kind):
    rng = np.random.default_rng(seed)
    n_real, P, W = 1900, 120, 2
    n = (n_real + TILE - 1) // TILE * TILE
    if kind == "ties":
        pool_c, pool_m = np.array([0, 250, 1000, 1000, 4000]), np.array([0, 1 << 20, 1 << 30, 1 << 30])
    elif kind == "random":
        pool_c, pool_m = rng.integers(0, 96000, 500), rng.integers(0, 1 << 38, 500)
    elif kind == "negative":
        pool_c, pool_m = rng.integers(-5000, 5000, 50), rng.integers(-(1 << 30), 1 << 30, 50)
    else:
        pool_c, pool_m = np.array([-(1 << 36), -1, 0, 1, 1 << 36]), np.array([-(1 << 55), -1, 0, 1, 1 << 55])
    free_c = np.full(n, I64_MIN, np.int64)  # the padding of the node table: never-feasible sentinels
    free_m = np.full(n, I64_MIN, np.int64)
    free_c[:n_real] = rng.choice(pool_c, n_real)
    free_m[:n_real] = rng.choice(pool_m, n_real)
    lab = np.zeros((n, W), np.uint64)
    lab[:n_real] = rng.integers(0, 1 << 63, size=(n_real, W), dtype=np.uint64) | rng.integers(0, 1 << 63, size=(n_real, W), dtype=np.uint64)
    rc = np.concatenate([rng.choice(pool_c, P - 4), [pool_c.min() - 1, pool_c.max() + 1, pool_c.min(), pool_c.max()]]).astype(np.int64)
    rm = np.concatenate([rng.choice(pool_m, P - 4), [pool_m.max(), pool_m.min(), pool_m.min() - 1, pool_m.max() + 1]]).astype(np.int64)
    sel = np.zeros((P, W), np.uint64)
    for p in range(P):
        for _ in range(rng.integers(0, 4)):
            sel[p, rng.integers(0, W)] |= np.uint64(1) << np.uint64(rng.integers(0, 63))
    got = _mask_via_ranks(free_c, free_m, lab, rc, rm, sel)
    want = _cells(free_c, free_m, lab, rc, rm, sel)
    assert np.array_equal(got, want)
    assert not got[:, n_real:].any()  # sentinels stay infeasible, also for the smallest requests
