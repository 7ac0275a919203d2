#!/bin/bash
# Round-2 8-GPU session (gpurun --gpus 8; charged 8x, so only what needs 8 GPUs): BASELINE C4 with the fused exchange and
# with NCCL, the C2 weak-scaling run, C5 streaming as 8 replicas and in lockstep.
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/m${N}_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29800
run() { # name, timeout, args...
  name=$1; to=$2; shift 2
  port=$((port + 1))
  timeout $to $TR --nproc-per-node $N --master-port $port "$@" > gpurun_out/m${N}_$name.json 2> gpurun_out/m${N}_$name.err
  echo "$name rc=$? $(tail -1 gpurun_out/m${N}_$name.json | cut -c1-200)"
}
run c4_p2p 200 bench.py --gpus $N --steps 20 --warmup 5
KS_TRACE=1 run c4_p2p_trace 200 bench.py --gpus $N --steps 20 --warmup 5
run c2w_p2p 150 bench.py --gpus $N --steps 20 --warmup 5 --workload c2
run c4_nccl 200 bench.py --gpus $N --steps 20 --warmup 5 --exchange nccl
run stream_replicas 100 bench_stream.py --seconds 8
run stream_lockstep 100 bench_stream.py --seconds 4 --mode lockstep
