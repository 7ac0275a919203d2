"""Property tests (hypothesis, CPU): the product's quantity parser (C++, libksched) and the oracle's (C) are two
independent restatements of the Kubernetes quantity grammar; they must agree on every string, and on the exact
domain both must equal plain integer arithmetic.  Also: the faithful and packed oracle flavours agree on small
random clusters built directly from hypothesis data (not from the synth generator)."""
import numpy as np
from hypothesis import given, settings, strategies as st

SUFFIXES = ["", "m", "k", "M", "G", "T", "Ki", "Mi", "Gi", "Ti", "u", "n", "e3", "E2", "e-3", "P", "Pi"]


@st.composite
def quantity(draw):
    sign = draw(st.sampled_from(["", "", "", "+", "-"]))
    ip = draw(st.integers(0, 10**7))
    frac = draw(st.one_of(st.just(None), st.integers(0, 999)))
    s = f"{sign}{ip}" + ("" if frac is None else f".{frac:03d}".rstrip("0") if frac else ".")
    return s + draw(st.sampled_from(SUFFIXES))


def _exact(q):
    """Third, independent restatement (Python rationals) of the quantity value; None if malformed."""
    import re
    from fractions import Fraction
    m = re.fullmatch(r"([+-]?)(\d*)(?:\.(\d*))?(Ki|Mi|Gi|Ti|Pi|Ei|[numkMGTPE]|[eE][+-]?\d+)?", q)
    if not m or not (m.group(2) or m.group(3)):
        return None
    v = Fraction(int(m.group(2) or "0")) + (Fraction(int(m.group(3)), 10 ** len(m.group(3))) if m.group(3) else 0)
    suf = m.group(4) or ""
    binary = {"Ki": 10, "Mi": 20, "Gi": 30, "Ti": 40, "Pi": 50, "Ei": 60}
    dec = {"n": -9, "u": -6, "m": -3, "k": 3, "M": 6, "G": 9, "T": 12, "P": 15, "E": 18}
    if suf in binary:
        v *= 2 ** binary[suf]
    elif suf in dec:
        v *= Fraction(10) ** dec[suf]
    elif suf:
        v *= Fraction(10) ** int(suf[1:])
    return -v if m.group(1) == "-" else v


def _expect(v, scale):
    """(rc, value) a parser with `scale` sub-units per unit must return."""
    x = v * scale
    if x.denominator != 1:
        return -6, None   # inexact
    if abs(x.numerator) > (1 << 63) - 1:
        return -5, None   # range
    return 0, x.numerator


@settings(max_examples=400, deadline=None)
@given(quantity())
def test_parsers_agree_on_random_quantities(ks, orc, q):
    v = _exact(q)
    assert v is not None
    for (rc, got), scale in ((orc.parse_quantity(q), 1000), (ks.host.parse_cpu_millicores(q), 1000),
                             (ks.host.parse_memory_bytes(q), 1)):
        erc, ev = _expect(v, scale)
        assert rc == erc, (q, scale, rc, erc)
        if erc == 0:
            assert got == ev, (q, scale)


@settings(max_examples=300, deadline=None)
@given(st.integers(0, 10**9), st.booleans())
def test_exact_domain_is_integer_arithmetic(ks, orc, v, as_milli):
    """SURVEY §8c exact domain: integer cores, integer millicores, integer bytes."""
    if as_milli:
        assert ks.host.parse_cpu_millicores(f"{v}m") == (0, v)
        assert orc.parse_quantity(f"{v}m") == (0, v)
    else:
        assert ks.host.parse_cpu_millicores(str(v)) == (0, v * 1000)
        assert ks.host.parse_memory_bytes(str(v)) == (0, v)
        assert orc.parse_quantity(str(v)) == (0, v * 1000)


@st.composite
def small_cluster(draw):
    n = draw(st.integers(1, 6))
    p = draw(st.integers(1, 6))
    keys = ["a", "b", "c"]
    vals = ["x", "y"]
    nodes = []
    for i in range(n):
        labels = draw(st.one_of(st.none(), st.dictionaries(st.sampled_from(keys), st.sampled_from(vals), max_size=3)))
        alloc = draw(st.one_of(st.none(), st.tuples(st.integers(0, 8), st.integers(0, 1 << 20))))
        nodes.append({"name": f"n{i}", "labels": labels,
                      "allocatable": None if alloc is None else {"cpu": str(alloc[0]), "memory": str(alloc[1])}})
    def pod(i, bound):
        conts = draw(st.lists(st.one_of(st.none(), st.tuples(st.integers(0, 3000), st.integers(0, 1 << 19))), max_size=3))
        sel = draw(st.one_of(st.none(), st.dictionaries(st.sampled_from(keys), st.sampled_from(vals), max_size=2)))
        return {"name": f"p{i}", "node_name": f"n{draw(st.integers(0, n))}" if bound else None,
                "containers": [None if c is None else {"cpu": f"{c[0]}m", "memory": str(c[1])} for c in conts],
                "selector": sel}
    bound = [pod(100 + i, True) for i in range(draw(st.integers(0, 5)))]
    pods = [pod(i, False) for i in range(p)]
    return nodes, bound, pods


@settings(max_examples=60, deadline=None)
@given(small_cluster(), st.sampled_from([0, 1]))
def test_faithful_and_packed_oracles_agree_on_hypothesis_clusters(ks, orc, cluster, policy):
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from helpers import pack_by_dictionary
    nodes_s, bound_s, pods_s = cluster
    arena = ks.objects.ObjectArena()
    nodes, bound, pods = arena.nodes(nodes_s), arena.pods(bound_s), arena.pods(pods_s)
    oc = orc.Cluster(nodes, len(nodes_s), bound, len(bound_s))
    f = oc.run(pods, len(pods_s), policy=policy, want_codes=True, nthreads=1)
    # packed form built by the TEST (independent of the C++ packer)
    ac = np.array([0 if n["allocatable"] is None else int(n["allocatable"]["cpu"]) * 1000 for n in nodes_s], np.int64)
    am = np.array([0 if n["allocatable"] is None else int(n["allocatable"]["memory"]) for n in nodes_s], np.int64)
    names = {n["name"]: i for i, n in enumerate(nodes_s)}
    def tot(p):
        c = sum(int(x["cpu"][:-1]) for x in p["containers"] if x)
        m = sum(int(x["memory"]) for x in p["containers"] if x)
        return c, m
    b = [(names[p["node_name"]],) + tot(p) for p in bound_s if p["node_name"] in names]
    fc, fm = orc.free_reduce(ac, am, np.array([x[0] for x in b], np.int32), np.array([x[1] for x in b], np.int64),
                             np.array([x[2] for x in b], np.int64)) if b else (ac.copy(), am.copy())
    labels, sel = pack_by_dictionary([n["labels"] for n in nodes_s], [p["selector"] for p in pods_s])
    rc = np.array([tot(p)[0] for p in pods_s], np.int64)
    rm = np.array([tot(p)[1] for p in pods_s], np.int64)
    pk = orc.run_packed(fc, fm, ac, am, labels, rc, rm, sel, policy=policy, want_codes=True, nthreads=1)
    for a, b_, what in zip(f, pk, ("node_idx", "score", "cnt", "mask", "codes")):
        assert np.array_equal(a, b_), what


@st.composite
def tricky_cluster(draw):
    """Like small_cluster, with label keys/values that stress the string handling of the packer: empty strings, shared
    prefixes, separators, non-ASCII, long values."""
    nodes_s, bound_s, pods_s = draw(small_cluster())
    words = ["", "a", "ab", "a/b", "a=b", "zone", "Zone", "zöne", "x" * 70, "a\tb", "k:v", " "]
    pick = st.sampled_from(words)
    for n in nodes_s:
        if draw(st.booleans()):
            n["labels"] = draw(st.dictionaries(pick, pick, max_size=4))
    for p in pods_s:
        if draw(st.booleans()):
            p["selector"] = draw(st.dictionaries(pick, pick, max_size=3))
    if draw(st.booleans()) and len(nodes_s) > 1:
        nodes_s[-1]["name"] = nodes_s[0]["name"]      # outside the reference's domain (its store is keyed by name):
        #                                                  oracle and product both charge the first node of the name
    return nodes_s, bound_s, pods_s


@settings(max_examples=120, deadline=None)
@given(tricky_cluster(), st.sampled_from([0, 1]))
def test_product_packer_equals_faithful_oracle_on_hypothesis_clusters(ks, orc, cluster, policy):
    """The C++ packer (packing-only context, no device) turns hypothesis-generated objects into arrays whose packed-oracle
    answer equals the object-model oracle's answer on the same objects."""
    nodes_s, bound_s, pods_s = cluster
    arena = ks.objects.ObjectArena()
    nodes, bound, pods = arena.nodes(nodes_s), arena.pods(bound_s), arena.pods(pods_s)
    oc = orc.Cluster(nodes, len(nodes_s), bound, len(bound_s))
    want = oc.run(pods, len(pods_s), policy=policy, want_codes=True, nthreads=1)
    with ks.host.Context(ks.host.KSH_DEVICE_NONE) as ctx:
        ctx.set_nodes(nodes, len(nodes_s))
        ctx.set_cluster_pods(bound, len(bound_s))
        rc, rm, sel = ctx.pack_pods(pods, len(pods_s))
        ac, am, lab, bn, bc, bm = ctx.export_packed()
    fc, fm = orc.free_reduce(ac, am, bn, bc, bm)
    got = orc.run_packed(fc, fm, ac, am, lab, rc, rm, sel, policy=policy, want_codes=True, nthreads=1)
    for g, w, what in zip(got, want, ("node_idx", "score", "cnt", "mask", "codes")):
        assert np.array_equal(g, w), what
