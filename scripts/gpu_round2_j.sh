#!/bin/bash
# Round-2 GPU session J (1 GPU): validation of the last two changes (k_least_alloc single-precision pre-filter, k_pod_ranks
# with 8192 splitters in dynamic shared memory) - full GPU suite, default bench line, timelines.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/j_pytest_all.log 2>&1
echo "pytest all rc=$? $(tail -1 gpurun_out/j_pytest_all.log)"
grep -E "FAILED|ERROR" gpurun_out/j_pytest_all.log | head -10
timeout 400 python bench.py > gpurun_out/j_bench_default.json 2> gpurun_out/j_bench_default.err
echo "default bench: $(cut -c1-260 gpurun_out/j_bench_default.json)"
B="python bench.py --no-cpu-baseline --no-secondary --no-objects"
show() { python - "$1" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d['roofline']
print(sys.argv[1], 'K2', round(1e3*r['kernel_ms'],1), 'us frac', round(r['frac'],3), 'step', round(1e3*d['ms_per_step'],1), 'us', d['config'].get('trace_us_rank0_last_timed_step'))
PY
}
KS_TRACE=1 timeout 200 $B --workload c3 > gpurun_out/j_c3_trace.json 2> gpurun_out/j_c3_trace.err; show gpurun_out/j_c3_trace.json
KS_TRACE=1 timeout 200 $B --workload c2 > gpurun_out/j_c2_trace.json 2> gpurun_out/j_c2_trace.err; show gpurun_out/j_c2_trace.json
KS_TRACE=1 timeout 200 $B --policy least_allocated > gpurun_out/j_c3_least_trace.json 2> gpurun_out/j_c3_least_trace.err; show gpurun_out/j_c3_least_trace.json
timeout 200 $B --policy least_allocated > gpurun_out/j_c3_least.json 2> gpurun_out/j_c3_least.err; show gpurun_out/j_c3_least.json
