#!/bin/bash
# Round-2 GPU session I (2 GPUs): where the ~23 us of the fused exchange go - kernel stamps (KS_TRACE) of both regimes,
# the same step without any exchange, and the NCCL variant for comparison.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
port=29700
run() { # name, env..., then bench args after --
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  port=$((port + 1))
  env "${envs[@]}" timeout 240 $TR --master-port $port bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-secondary --no-objects "$@" \
      > gpurun_out/i_$name.json 2> gpurun_out/i_$name.err
  python - gpurun_out/i_$name.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d['roofline']
    print(sys.argv[1], 'K2', round(1e3*r['kernel_ms'],1), 'us step', round(1e3*d['ms_per_step'],1), 'us', d['config'].get('trace_us_rank0_last_timed_step'))
except Exception as e:
    print(sys.argv[1], 'unreadable:', e)
PY
}
run c2w_p2p_trace_a256 KS_TRACE=1 -- --workload c2
run c4_p2p_trace_a256 KS_TRACE=1 --
run c2w_p2p_trace_a128 KS_TRACE=1 KS_ARGMAX_THREADS=128 -- --workload c2
run c4_p2p_trace_a128 KS_TRACE=1 KS_ARGMAX_THREADS=128 --
run c2w_none KS_X=0 -- --workload c2 --exchange none
run c4_none KS_X=0 -- --exchange none
run c4_p2p KS_X=0 --
run c4_p2p_a128 KS_ARGMAX_THREADS=128 --
# block-size check of the mask kernel after the dynamic work distribution (1 GPU)
B="python bench.py --no-cpu-baseline --no-secondary --no-objects"
for t in 896 960 1024; do
  for wl in c3 c2; do
    KS_ROWS_THREADS=$t timeout 240 $B --workload $wl > gpurun_out/i_${wl}_t$t.json 2> gpurun_out/i_${wl}_t$t.err
    python - gpurun_out/i_${wl}_t$t.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d['roofline']
print(sys.argv[1], 'K2', round(1e3*r['kernel_ms'],1), 'us frac', round(r['frac'],3), 'step', round(1e3*d['ms_per_step'],1), 'us')
PY
  done
done
