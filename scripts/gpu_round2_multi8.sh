#!/bin/bash
# Round-2 8-GPU session (gpurun --gpus 8; charged 8x, so only what needs 8 GPUs and tight timeouts): BASELINE C4 with the
# fused exchange, the C2 weak-scaling run, C5 streaming as 8 replicas and in lockstep, C4 with NCCL if time is left.
N=${1:-8}
mkdir -p gpurun_out
T0=$SECONDS
nvidia-smi topo -m > gpurun_out/m${N}_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29800
run() { # name, timeout, args...
  name=$1; to=$2; shift 2
  port=$((port + 1))
  timeout $to $TR --nproc-per-node $N --master-port $port "$@" > gpurun_out/m${N}_$name.json 2> gpurun_out/m${N}_$name.err
  echo "$name rc=$? t=$((SECONDS - T0))s $(tail -1 gpurun_out/m${N}_$name.json | cut -c1-200)"
}
KS_TRACE=1 run c4_p2p_trace 110 bench.py --gpus $N --steps 20 --warmup 5
run c2w_p2p 80 bench.py --gpus $N --steps 20 --warmup 5 --workload c2
run stream_replicas 50 bench_stream.py --seconds 6
run stream_lockstep 50 bench_stream.py --seconds 3 --mode lockstep
[ $((SECONDS - T0)) -lt 170 ] && run c4_p2p 100 bench.py --gpus $N --steps 20 --warmup 5
[ $((SECONDS - T0)) -lt 230 ] && run c4_nccl 100 bench.py --gpus $N --steps 20 --warmup 5 --exchange nccl
echo "total $((SECONDS - T0))s"
