"""CPU tests: pin the oracle against every golden vector the reference supplies (the three nodeSelector tests,
src/predicates/test.rs:42-58) and against the hand-derived vectors of SURVEY.md §8c; check the two oracle
flavours (faithful object model vs packed SoA) against each other on random clusters."""
import ctypes as C

import numpy as np
import pytest

from helpers import load_golden, mask_bits, rows_to_str


def _addr(arr, i=0):
    return C.addressof(arr) + i * C.sizeof(arr._type_)


def test_reference_selector_tests_and_derived_kats(ks, orc):
    kats = load_golden("selector_kats.json")["cases"]
    assert sum(1 for c in kats if c["id"].startswith("S")) == 3  # the reference's own three tests
    for case in kats:
        arena = ks.objects.ObjectArena()
        pods = arena.pods([case["pod"]])
        nodes = arena.nodes([case["node"]])
        got = orc.lib.orc_does_node_selector_match(_addr(pods), _addr(nodes))
        assert bool(got) == case["expect"], case["id"]


def test_quantity_grammar(orc):
    kats = load_golden("quantity_kats.json")
    names = {"INEXACT": orc.ERR_INEXACT, "RANGE": orc.ERR_RANGE}
    for s, want in kats["ok"]:
        rc, v = orc.parse_quantity(s)
        if isinstance(want, str):
            assert rc == names[want], (s, rc)
        else:
            assert rc == 0 and v == want, (s, rc, v)
    for s in kats["bad"]:
        rc, _ = orc.parse_quantity(s)
        assert rc == orc.ERR_PARSE, (s, rc)


def _gv1(ks, orc):
    g = load_golden("gv1.json")
    arena = ks.objects.ObjectArena()
    nodes = arena.nodes(g["nodes"])
    allp = arena.pods(g["bound_pods"])
    pods = arena.pods(g["pods"])
    cl = orc.Cluster(nodes, len(g["nodes"]), allp, len(g["bound_pods"]))
    return g, arena, nodes, pods, cl


def test_gv1_free_and_requests(ks, orc):
    g, arena, nodes, pods, cl = _gv1(ks, orc)
    for n in range(5):
        rc, (c, m) = cl.available(n)
        assert rc == 0
        assert c == g["expected_free_cpu_milli"][n]
        assert m == g["expected_free_mem_bytes"][n] * 1000  # oracle memory unit is milli-bytes
    for p in range(10):
        out = (C.c_int64 * 2)()
        assert orc.lib.orc_total_pod_resources(_addr(pods, p), out) == 0
        assert out[0] == g["expected_req_cpu_milli"][p]
        assert out[1] == g["expected_req_mem_bytes"][p] * 1000


def test_gv1_feasible_rows_and_argmax(ks, orc):
    g, arena, nodes, pods, cl = _gv1(ks, orc)
    idx, score, cnt, mask, codes = cl.run(pods, 10, policy=orc.ORC_SCORE_LEFTOVER, want_codes=True, nthreads=2)
    assert rows_to_str(mask_bits(mask, 5)) == g["expected_feasible_rows"]
    assert list(idx) == g["expected_node_idx_leftover"]
    assert list(cnt) == [r.count("1") for r in g["expected_feasible_rows"]]
    # resource_fits only: every infeasible cell is NotEnoughResources
    assert set(np.unique(codes)) <= {0, 1}
    fc, fm = g["expected_free_cpu_milli"], g["expected_free_mem_bytes"]
    for p in range(10):
        if idx[p] >= 0:
            n = idx[p]
            want = (fc[n] - g["expected_req_cpu_milli"][p]) * (1 << 22) + (fm[n] - g["expected_req_mem_bytes"][p])
            assert score[p] == want


def test_reason_precedence_fit_before_selector(ks, orc):
    """A cell failing both predicates reports NotEnoughResources (src/predicates.rs:68-74)."""
    arena = ks.objects.ObjectArena()
    nodes = arena.nodes([{"name": "n", "labels": {"a": "1"}, "allocatable": {"cpu": "1", "memory": "100"}}])
    pods = arena.pods([
        {"name": "both", "containers": [{"cpu": "2", "memory": "1"}], "selector": {"a": "2"}},
        {"name": "sel", "containers": [{"cpu": "1", "memory": "100"}], "selector": {"a": "2"}},
        {"name": "ok", "containers": [{"cpu": "1", "memory": "100"}], "selector": {"a": "1"}},
        {"name": "mem", "containers": [{"cpu": "1", "memory": "101"}], "selector": {"a": "1"}},
    ])
    none = arena.pods([])
    cl = orc.Cluster(nodes, 1, none, 0)
    assert [cl.check(pods, i, 0) for i in range(4)] == [1, 2, 0, 1]


def test_reference_panics_become_errors(ks, orc):
    arena = ks.objects.ObjectArena()
    nodes = arena.nodes([{"name": "n", "allocatable": {"cpu": "1"}},           # memory key missing (:30)
                         {"name": "m", "allocatable": {"cpu": "x1", "memory": "1"}}])  # unparsable (:29)
    pods = arena.pods([{"name": "p", "containers": [{"cpu": "1"}]},
                       {"name": "q", "containers": [{"cpu": "1 core"}]}])
    cl = orc.Cluster(nodes, 2, arena.pods([]), 0)
    assert cl.check(pods, 0, 0) == orc.ERR_MISSING
    assert cl.check(pods, 0, 1) == orc.ERR_PARSE
    out = (C.c_int64 * 2)()
    assert orc.lib.orc_total_pod_resources(_addr(pods, 1), out) == orc.ERR_PARSE


def test_is_pod_bound(ks, orc):
    arena = ks.objects.ObjectArena()
    pods = arena.pods([{"name": "a"}, {"name": "b", "node_name": "n1"}, {"name": "c", "spec": False}])
    assert [orc.lib.orc_is_pod_bound(_addr(pods, i)) for i in range(3)] == [0, 1, 0]


@pytest.mark.parametrize("seed,P,N,keys", [(1, 60, 37, 8), (2, 33, 300, 8), (3, 20, 64, 32)])
@pytest.mark.parametrize("policy", [0, 1])
def test_faithful_equals_packed(ks, orc, seed, P, N, keys, policy):
    """Oracle (a) object model with strings == oracle (b) packed SoA (first parity gate)."""
    cl = ks.synth.make(P, N, seed, n_keys=keys, bound_per_node=3)
    nodes_s, bound_s, pods_s = ks.objects.cluster_specs(cl)
    arena = ks.objects.ObjectArena()
    nodes, bound, pods = arena.nodes(nodes_s), arena.pods(bound_s), arena.pods(pods_s)
    oc = orc.Cluster(nodes, cl.N, bound, cl.B)
    f = oc.run(pods, P, policy=policy, want_codes=True, nthreads=2)
    ac, am, lab, bn, bc, bm, rc, rm, sel = cl.packed()
    fc, fm = orc.free_reduce(ac, am, bn, bc, bm)
    cfc, cfm = cl.free()
    assert np.array_equal(fc, cfc) and np.array_equal(fm, cfm)
    pk = orc.run_packed(fc, fm, ac, am, lab, rc, rm, sel, policy=policy, want_codes=True, nthreads=2)
    for a, b, what in zip(f, pk, ("node_idx", "score", "cnt", "mask", "codes")):
        assert np.array_equal(a, b), what
    assert (f[2] > 0).any() and (f[2] == 0).any()  # the generator exercises both outcomes


def test_sampling_policy_picks_from_feasible_set(ks, orc):
    """A6: the reference's <=5 random draws can only return a node of the feasible set; None is allowed even
    when the set is non-empty (src/main.rs:49-71)."""
    cl = ks.synth.make(40, 50, 7, bound_per_node=2)
    nodes_s, bound_s, pods_s = ks.objects.cluster_specs(cl)
    arena = ks.objects.ObjectArena()
    nodes, bound, pods = arena.nodes(nodes_s), arena.pods(bound_s), arena.pods(pods_s)
    oc = orc.Cluster(nodes, cl.N, bound, cl.B)
    _, _, cnt, mask, _ = oc.run(pods, 40, nthreads=2)
    bits = mask_bits(mask, cl.N)
    hits = 0
    for p in range(40):
        n, cells = oc.sampling(pods, p, 5, seed=1234 + p)
        assert 0 <= cells <= 5
        if n >= 0:
            assert bits[p, n]
            hits += 1
        if cnt[p] == 0:
            assert n == -1
    assert hits > 0


def test_sampling_policy_packed_equals_faithful(ks, orc):
    """The packed restatement of the reference policy draws, checks and stops exactly like the object-model one."""
    cl = ks.synth.make(80, 60, 11, bound_per_node=4)
    nodes_s, bound_s, pods_s = ks.objects.cluster_specs(cl)
    arena = ks.objects.ObjectArena()
    nodes, bound, pods = arena.nodes(nodes_s), arena.pods(bound_s), arena.pods(pods_s)
    oc = orc.Cluster(nodes, cl.N, bound, cl.B)
    ac, am, lab, bn, bc, bm, rc, rm, sel = cl.packed()
    fc, fm = orc.free_reduce(ac, am, bn, bc, bm)
    states = np.array([ks.capi.sampling_stream(5, p) for p in range(cl.P)], np.uint64)
    idx, cells, dn, dc = orc.sampling_packed(fc, fm, lab, rc, rm, sel, 5, states)
    for p in range(cl.P):
        n, c = oc.sampling(pods, p, 5, seed=int(states[p]))
        assert (idx[p], cells[p]) == (n, c)
        assert (dn[p, c:] == -1).all() and (dc[p, c:] == 0xff).all()
        if n >= 0:
            assert dn[p, c - 1] == n and dc[p, c - 1] == 0
    assert (idx >= 0).any() and (idx < 0).any()
