#!/bin/bash
# Round-2 GPU session L (1 GPU): k_least_alloc with the windowed scan (phase A: 8 / 32 tile masks per warp in parallel,
# phase B: only the tiles with a feasible slot) - LeastAllocated parity first, its bench line with the timeline, then the
# whole GPU suite of the same build.
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q -k "least or adversarial or near_ties" > gpurun_out/l_pytest_least.log 2>&1
echo "pytest least rc=$? $(tail -1 gpurun_out/l_pytest_least.log)"
grep -E "FAILED|ERROR" gpurun_out/l_pytest_least.log | head -10
B="python bench.py --no-cpu-baseline --no-secondary --no-objects"
show() { python - "$1" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d['roofline']
print(sys.argv[1], 'K2', round(1e3*r['kernel_ms'],1), 'us frac', round(r['frac'],3), 'step', round(1e3*d['ms_per_step'],1), 'us', d['config'].get('trace_us_rank0_last_timed_step'))
PY
}
KS_TRACE=1 timeout 150 $B --policy least_allocated > gpurun_out/l_c3_least_trace.json 2> gpurun_out/l_c3_least_trace.err; show gpurun_out/l_c3_least_trace.json
timeout 420 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/l_pytest_all.log 2>&1
echo "pytest all rc=$? $(tail -1 gpurun_out/l_pytest_all.log)"
grep -E "FAILED|ERROR" gpurun_out/l_pytest_all.log | head -10
timeout 150 $B --policy least_allocated > gpurun_out/l_c3_least.json 2> gpurun_out/l_c3_least.err; show gpurun_out/l_c3_least.json
