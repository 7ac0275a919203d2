#!/bin/bash
# Round-2 GPU session B (1 GPU): write-only bandwidth ceiling, full GPU test suite, sort-mode A/B, the default bench line.
#   gpurun --timeout 1200 -- 'bash scripts/gpu_round2_b.sh'
mkdir -p gpurun_out
timeout 120 scripts/build/write_bw > gpurun_out/b_write_bw.json 2> gpurun_out/b_write_bw.err
echo "write_bw: $(cat gpurun_out/b_write_bw.json)"
timeout 600 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/b_pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/b_pytest.log)"
B="python bench.py --no-cpu-baseline --no-secondary --no-objects"
for rep in 1 2; do
  for w in c3 c2; do
    for k in grid exact; do
      export KS_ROWS_SORT=0
      [ $k = exact ] && export KS_ROWS_SORT=1
      timeout 240 $B --workload $w > gpurun_out/b_${w}_${k}_r$rep.json 2> gpurun_out/b_${w}_${k}_r$rep.err
      python - "$w" "$k" "$rep" <<'PY'
import json, sys
w, k, rep = sys.argv[1:4]
try:
    d = json.load(open(f"gpurun_out/b_{w}_{k}_r{rep}.json"))
    r = d["roofline"]
    print(f"{w} {k} run {rep}: K2 {1e3 * r['kernel_ms']:.2f} us  frac {r['frac']:.4f}  step {1e3 * d['ms_per_step']:.1f} us  e2e {d['e2e']['value']:.3g}")
except Exception as e:
    print(f"{w} {k} run {rep}: FAILED ({e})")
PY
    done
  done
done
unset KS_ROWS_SORT
timeout 400 python bench.py > gpurun_out/b_bench_default.json 2> gpurun_out/b_bench_default.err
echo "default bench: $(cut -c1-500 gpurun_out/b_bench_default.json)"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/b_bench_reference.json 2> gpurun_out/b_bench_reference.err
echo "reference arm: $(cut -c1-300 gpurun_out/b_bench_reference.json)"
timeout 200 $B --policy least_allocated > gpurun_out/b_c3_least.json 2> gpurun_out/b_c3_least.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/b_c3_least.json")); print("least_allocated c3: step", d["ms_per_step"], "ms; rest", d["roofline"]["rest_of_step_ms"])
except Exception as e:
    print("least FAILED", e)
PY
timeout 120 python bench_stream.py --seconds 10 > gpurun_out/b_stream_async.json 2> gpurun_out/b_stream_async.err
echo "stream async: $(cut -c1-330 gpurun_out/b_stream_async.json)"
timeout 120 python bench_stream.py --seconds 10 --sync > gpurun_out/b_stream_sync.json 2> gpurun_out/b_stream_sync.err
echo "stream sync: $(cut -c1-330 gpurun_out/b_stream_sync.json)"
timeout 120 python bench_stream.py --seconds 5 --rate 100000 > gpurun_out/b_stream_100k.json 2> gpurun_out/b_stream_100k.err
echo "stream 100k/s: $(cut -c1-330 gpurun_out/b_stream_100k.json)"
ls gpurun_out | wc -l
