#!/bin/bash
# Round-2 GPU session K (1 GPU, the round's last GPU minutes): validation of the build at HEAD (k_least_alloc single-precision
# pre-filter, k_pod_ranks with 8192 splitters in dynamic shared memory) - full GPU suite first, then the default bench line,
# the LeastAllocated line, timelines, the launch list and memcheck, most important first (the call may be cut short).
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/k_pytest_all.log 2>&1
echo "pytest all rc=$? $(tail -1 gpurun_out/k_pytest_all.log)"
grep -E "FAILED|ERROR" gpurun_out/k_pytest_all.log | head -10
timeout 300 python bench.py > gpurun_out/k_bench_default.json 2> gpurun_out/k_bench_default.err
echo "default bench: $(cut -c1-260 gpurun_out/k_bench_default.json)"
B="python bench.py --no-cpu-baseline --no-secondary --no-objects"
show() { python - "$1" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d['roofline']
print(sys.argv[1], 'K2', round(1e3*r['kernel_ms'],1), 'us frac', round(r['frac'],3), 'step', round(1e3*d['ms_per_step'],1), 'us', d['config'].get('trace_us_rank0_last_timed_step'))
PY
}
timeout 150 $B --policy least_allocated > gpurun_out/k_c3_least.json 2> gpurun_out/k_c3_least.err; show gpurun_out/k_c3_least.json
KS_TRACE=1 timeout 150 $B --workload c3 > gpurun_out/k_c3_trace.json 2> gpurun_out/k_c3_trace.err; show gpurun_out/k_c3_trace.json
KS_TRACE=1 timeout 150 $B --workload c2 > gpurun_out/k_c2_trace.json 2> gpurun_out/k_c2_trace.err; show gpurun_out/k_c2_trace.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_c3_head.csv \
    $B --workload c3 --steps 2 --warmup 1 > gpurun_out/k_launches_c3.log 2>&1
echo "launch list rc=$?"
KS_TRACE=1 timeout 150 $B --policy least_allocated > gpurun_out/k_c3_least_trace.json 2> gpurun_out/k_c3_least_trace.err; show gpurun_out/k_c3_least_trace.json
timeout 150 compute-sanitizer --tool memcheck python tests/sanitizer_smoke.py > gpurun_out/k_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$? $(tail -2 gpurun_out/k_sanitizer_memcheck.log | tr '\n' ' ')"
