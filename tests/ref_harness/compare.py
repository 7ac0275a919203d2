#!/usr/bin/env python
"""Compare the output of tests/ref_harness/golden_dump.rs (run inside the REAL reference crate, see its header) with
this repository's hand-derived golden vectors.  Prints PINNED when the crate agrees on GV-1 (free values, pod totals,
feasible rows) and on every exact-domain quantity; the entries outside the exact domain (Ki/Mi/Gi, fractions, k) are
reported for information - SURVEY.md §8c expects the crate's f32 format conversion to disagree on some of them."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
EXACT_DOMAIN = {"0", "1", "4", "250m", "1500m", "0m", "100", "1073741824"}


def main(path):
    docs = [json.loads(line) for line in open(path) if line.strip()]
    by_kind = {d["kind"]: d for d in docs}
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "gv1.json")))
    ok = True
    gv = by_kind.get("gv1")
    if gv is None:
        print("missing gv1 record")
        return 1
    if not all(gv["free_equals_expected"]):
        print("free values differ:", gv["free_display"], "expected", g["expected_free_cpu_milli"], g["expected_free_mem_bytes"])
        ok = False
    if not all(gv["req_equals_expected"]):
        print("pod totals differ:", gv["req_display"])
        ok = False
    if gv["feasible_rows"] != g["expected_feasible_rows"]:
        print("feasible rows differ:", gv["feasible_rows"], "expected", g["expected_feasible_rows"])
        ok = False
    qs = by_kind.get("quantities", {"rows": []})
    for s, shown, direct, accumulated, le in qs["rows"]:
        inside = s in EXACT_DOMAIN
        good = direct and accumulated and le
        print(f"  quantity {s!r:14} -> {shown!s:22} {'exact domain' if inside else 'outside     '} {'agrees' if good else 'DIFFERS'}")
        if inside and not good:
            ok = False
    print("PINNED: the crate reproduces GV-1 and the exact-domain quantities" if ok else "NOT PINNED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main(sys.argv[1]))
