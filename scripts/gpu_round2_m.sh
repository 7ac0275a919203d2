#!/bin/bash
# Round-2 GPU session M (1 GPU, the round's last GPU minutes): k_least_alloc with exact single-precision scores (cpu quotient by
# IEEE division of < 2^24 integers, memory quotient by an estimate whose floor is unambiguous away from integers; only ties load
# the node index, only unknown scores the 64-byte row) - LeastAllocated parity, its bench line with the timeline, the whole suite.
mkdir -p gpurun_out
timeout 120 python -m pytest tests -m gpu -q -k "least or adversarial or near_ties" > gpurun_out/m_pytest_least.log 2>&1
echo "pytest least rc=$? $(tail -1 gpurun_out/m_pytest_least.log)"
grep -E "FAILED|ERROR" gpurun_out/m_pytest_least.log | head -10
B="python bench.py --no-cpu-baseline --no-secondary --no-objects"
KS_TRACE=1 timeout 100 $B --policy least_allocated > gpurun_out/m_c3_least_trace.json 2> gpurun_out/m_c3_least_trace.err
python - gpurun_out/m_c3_least_trace.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d['roofline']
print(sys.argv[1], 'K2', round(1e3*r['kernel_ms'],1), 'us frac', round(r['frac'],3), 'step', round(1e3*d['ms_per_step'],1), 'us', d['config'].get('trace_us_rank0_last_timed_step'))
PY
timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/m_pytest_all.log 2>&1
echo "pytest all rc=$? $(tail -1 gpurun_out/m_pytest_all.log)"
grep -E "FAILED|ERROR" gpurun_out/m_pytest_all.log | head -10
timeout 60 python __graft_entry__.py smoke 2>&1 | tail -2
