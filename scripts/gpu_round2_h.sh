#!/bin/bash
# Round-2 GPU session H (1 GPU): chunked work stealing in the mask kernel - parity first, then timelines and A/B of the
# argmax block size (128-thread CTAs run beside the mask CTAs, 256-thread ones after them).
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_trace_gpu.py -m gpu -q -x \
    -k "dynamic_work or c3_full or c2_full or step_trace or device_buffers or random_clusters_leftover" > gpurun_out/h_pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/h_pytest.log)"
B="python bench.py --no-cpu-baseline --no-secondary --no-objects"
show() { python - "$1" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d['roofline']
print(sys.argv[1], 'K2', round(1e3*r['kernel_ms'],1), 'us frac', round(r['frac'],3), 'step', round(1e3*d['ms_per_step'],1), 'us', d['config'].get('trace_us_rank0_last_timed_step'))
PY
}
for wl in c3 c2; do
  for at in 256 128; do
    KS_TRACE=1 KS_ARGMAX_THREADS=$at timeout 240 $B --workload $wl > gpurun_out/h_${wl}_trace_a$at.json 2> gpurun_out/h_${wl}_trace_a$at.err
    show gpurun_out/h_${wl}_trace_a$at.json
  done
done
for rep in 1 2; do
  for wl in c3 c2; do
    for at in 256 128; do
      KS_ARGMAX_THREADS=$at timeout 240 $B --workload $wl > gpurun_out/h_${wl}_a${at}_$rep.json 2> gpurun_out/h_${wl}_a${at}_$rep.err
      show gpurun_out/h_${wl}_a${at}_$rep.json
    done
  done
done
for t in 768 832 960; do
  KS_ARGMAX_THREADS=128 KS_ROWS_THREADS=$t timeout 240 $B --workload c3 > gpurun_out/h_c3_a128_t$t.json 2> gpurun_out/h_c3_a128_t$t.err
  show gpurun_out/h_c3_a128_t$t.json
done
