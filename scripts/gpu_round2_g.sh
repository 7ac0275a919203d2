#!/bin/bash
# Round-2 GPU session G (1 GPU): step timelines from the kernels' own %globaltimer stamps (KS_TRACE=1) and the
# 128-thread argmax CTAs that fit beside an 896-thread mask CTA.
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-secondary --no-objects"
show() { python - "$1" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d['roofline']
print(sys.argv[1], 'K2', round(1e3*r['kernel_ms'],1), 'us step', round(1e3*d['ms_per_step'],1), 'us', d['config'].get('trace_us_rank0_last_timed_step'))
PY
}
for wl in c2 c3; do
  for at in 256 128; do
    KS_TRACE=1 KS_ARGMAX_THREADS=$at timeout 240 $B --workload $wl > gpurun_out/g_${wl}_trace_a$at.json 2> gpurun_out/g_${wl}_trace_a$at.err
    show gpurun_out/g_${wl}_trace_a$at.json
  done
done
for rep in 1 2; do
  for wl in c2 c3; do
    for at in 256 128; do
      KS_ARGMAX_THREADS=$at timeout 240 $B --workload $wl > gpurun_out/g_${wl}_a${at}_$rep.json 2> gpurun_out/g_${wl}_a${at}_$rep.err
      show gpurun_out/g_${wl}_a${at}_$rep.json
    done
  done
done
KS_TRACE=1 timeout 240 $B --workload c3 --policy least_allocated > gpurun_out/g_c3_least_trace.json 2> gpurun_out/g_c3_least_trace.err
show gpurun_out/g_c3_least_trace.json
