#!/bin/bash
# A/B of the mask-kernel variants (KS_BP_VARIANT, csrc/ks_bitpar.cu) on ONE GPU box: parity first, then alternating bench runs.
#   gpurun --timeout 2400 -- 'bash scripts/ab_variants.sh "1 2 3 4 5"'
# Results: gpurun_out/ab_*.json (bench lines), gpurun_out/ab_parity_v*.log, summary on stdout.
VARIANTS=${1:-"1 2 3 4 5"}
mkdir -p gpurun_out
OK=""
for v in $VARIANTS; do
    KS_BP_VARIANT=$v timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/ab_parity_v$v.log 2>&1
    rc=$?
    echo "variant $v parity rc=$rc: $(tail -1 gpurun_out/ab_parity_v$v.log)"
    [ $rc -eq 0 ] && OK="$OK $v"
done
echo "benchmarking variants:$OK"
for rep in 1 2; do
    for w in c2 c3; do
        for v in $OK; do
            KS_BP_VARIANT=$v timeout 600 python bench.py --workload $w --no-cpu-baseline > gpurun_out/ab_${w}_v${v}_r$rep.json 2> gpurun_out/ab_${w}_v${v}_r$rep.err
            python - "$w" "$v" "$rep" <<'PY'
import json, sys
w, v, rep = sys.argv[1:4]
try:
    d = json.load(open(f"gpurun_out/ab_{w}_v{v}_r{rep}.json"))
    r = d["roofline"]
    print(f"{w} variant {v} run {rep}: K2 {1e3 * r['kernel_ms']:.2f} us  frac {r['frac']:.4f}  step {1e3 * d['ms_per_step']:.1f} us  e2e {d['e2e']['value']:.3g}")
except Exception as e:
    print(f"{w} variant {v} run {rep}: FAILED ({e})")
PY
        done
    done
done
