#!/bin/bash
# Round-2 GPU session D (1 GPU): mask-kernel variants (block size, pipeline depth, argmax beside the mask kernel), parity under each.
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-secondary --no-objects"
for v in "1024 1" "1024 2" "768 1" "768 2"; do
  set -- $v
  export KS_ROWS_THREADS=$1 KS_ROWS_PIPE=$2
  timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=3 -k "device_buffers or random_clusters_leftover or adversarial or c2_full or bindings_ready or graph_replay" > gpurun_out/d_pytest_$1_$2.log 2>&1
  echo "pytest threads $1 pipe $2: rc=$? $(tail -1 gpurun_out/d_pytest_$1_$2.log)"
done
for rep in 1 2; do
  for w in c3 c2; do
    for v in "1024 1" "1024 2" "768 1" "768 2"; do
      set -- $v
      export KS_ROWS_THREADS=$1 KS_ROWS_PIPE=$2
      timeout 240 $B --workload $w > gpurun_out/d_${w}_$1_$2_r$rep.json 2> gpurun_out/d_${w}_$1_$2_r$rep.err
      python - "$w" "$1" "$2" "$rep" <<'PY'
import json, sys
w, th, pp, rep = sys.argv[1:5]
try:
    d = json.load(open(f"gpurun_out/d_{w}_{th}_{pp}_r{rep}.json"))
    r = d["roofline"]
    print(f"{w} threads {th} pipe {pp} run {rep}: K2 {1e3 * r['kernel_ms']:.2f} us  frac {r['frac']:.4f}  step {1e3 * d['ms_per_step']:.1f} us  e2e {d['e2e']['value']:.3g}")
except Exception as e:
    print(f"{w} threads {th} pipe {pp} run {rep}: FAILED ({e})")
PY
    done
  done
done
unset KS_ROWS_THREADS KS_ROWS_PIPE
timeout 300 python -m pytest tests -m gpu -q --maxfail=5 > gpurun_out/d_pytest_all.log 2>&1
echo "pytest all rc=$? $(tail -1 gpurun_out/d_pytest_all.log)"
