#!/bin/bash
# Round-2 GPU session F (1 GPU): the full GPU suite, the default bench line, step timelines and the ncu evidence of the
# shipped build (dynamic work distribution, 960-thread mask CTAs, SM sharing policy of select_w).
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --maxfail=5 > gpurun_out/f_pytest_all.log 2>&1
echo "pytest all rc=$? $(tail -1 gpurun_out/f_pytest_all.log)"
timeout 400 python bench.py > gpurun_out/f_bench_default.json 2> gpurun_out/f_bench_default.err
echo "default bench: $(cut -c1-300 gpurun_out/f_bench_default.json)"
B="python bench.py --no-cpu-baseline --no-secondary --no-objects"
for wl in c3 c2; do
  KS_TRACE=1 timeout 240 $B --workload $wl > gpurun_out/f_${wl}_trace.json 2> gpurun_out/f_${wl}_trace.err
  python - gpurun_out/f_${wl}_trace.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d['roofline']
print(sys.argv[1], 'K2', round(1e3*r['kernel_ms'],1), 'us frac', round(r['frac'],3), 'step', round(1e3*d['ms_per_step'],1), 'us', d['config'].get('trace_us_rank0_last_timed_step'))
PY
done
timeout 200 $B --workload c2 > gpurun_out/f_c2.json 2> gpurun_out/f_c2.err
echo "c2: $(cut -c1-200 gpurun_out/f_c2.json)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_mask_rows -s 3 -c 1 -o gpurun_out/r02_ncu_rows_c3_shipped \
    $B --workload c3 --steps 1 --warmup 1 > gpurun_out/f_ncu_c3.log 2>&1
echo "ncu c3 rc=$?"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_mask_rows -s 3 -c 1 -o gpurun_out/r02_ncu_rows_c2_shipped \
    $B --workload c2 --steps 1 --warmup 1 > gpurun_out/f_ncu_c2.log 2>&1
echo "ncu c2 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_c3_shipped.csv \
    $B --workload c3 --steps 2 --warmup 1 > gpurun_out/f_launches_c3.log 2>&1
echo "launch list rc=$?"
timeout 200 $B --policy least_allocated > gpurun_out/f_c3_least.json 2> gpurun_out/f_c3_least.err
python -c "
import json; d=json.loads(open('gpurun_out/f_c3_least.json').read().strip().splitlines()[-1]); print('least_allocated c3: step', d['ms_per_step'], 'ms')"
timeout 120 python bench_stream.py --seconds 10 > gpurun_out/f_stream_async.json 2> gpurun_out/f_stream_async.err
echo "stream async: $(cut -c1-330 gpurun_out/f_stream_async.json)"
timeout 150 compute-sanitizer --tool memcheck python tests/sanitizer_smoke.py > gpurun_out/f_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$? $(tail -2 gpurun_out/f_sanitizer_memcheck.log | tr '\n' ' ')"
timeout 200 compute-sanitizer --tool racecheck python tests/sanitizer_smoke.py > gpurun_out/f_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$? $(tail -2 gpurun_out/f_sanitizer_racecheck.log | tr '\n' ' ')"
