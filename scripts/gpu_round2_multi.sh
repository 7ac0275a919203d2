#!/bin/bash
# Round-2 multi-GPU session (one gpurun call with --gpus N): fused NVLink exchange vs NCCL on BASELINE C4, C5 on N GPUs.
#   gpurun --gpus 8 --timeout 900 -- 'bash scripts/gpu_round2_multi.sh 8'
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/m${N}_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29600
run() { # name, nproc, args...
  name=$1; np=$2; shift 2
  port=$((port + 1))
  timeout 300 $TR --nproc-per-node $np --master-port $port "$@" > gpurun_out/m${N}_$name.json 2> gpurun_out/m${N}_$name.err
  echo "$name rc=$? $(tail -1 gpurun_out/m${N}_$name.json | cut -c1-260)"
}
# the two-rank exchange test uses two different GPUs here (stores cross NVLink)
timeout 300 python -m pytest tests/test_exchange_gpu.py -m gpu -q > gpurun_out/m${N}_pytest_exchange.log 2>&1
echo "pytest exchange rc=$? $(tail -1 gpurun_out/m${N}_pytest_exchange.log)"
for np in 2 4 8; do
  [ $np -le $N ] || continue
  run c4_p2p_$np $np bench.py --gpus $np --steps 20 --warmup 5
  run c4_nccl_$np $np bench.py --gpus $np --steps 20 --warmup 5 --exchange nccl
  run c2w_p2p_$np $np bench.py --gpus $np --steps 20 --warmup 5 --workload c2
done
timeout 200 python bench.py --no-cpu-baseline --no-objects > gpurun_out/m${N}_c3_1gpu.json 2> gpurun_out/m${N}_c3_1gpu.err
echo "1 gpu: $(cut -c1-260 gpurun_out/m${N}_c3_1gpu.json)"
run stream_replicas $N bench_stream.py --seconds 10
run stream_lockstep $N bench_stream.py --seconds 5 --mode lockstep
timeout 120 python bench_stream.py --seconds 10 > gpurun_out/m${N}_stream_1gpu.json 2> gpurun_out/m${N}_stream_1gpu.err
echo "stream 1 gpu: $(cut -c1-300 gpurun_out/m${N}_stream_1gpu.json)"
ls -la gpurun_out | tail -30
