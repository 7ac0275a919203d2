#!/bin/bash
# Round-2 GPU session A (one gpurun call, 1 GPU): parity of the rows kernel, A/B against the round-1 kernel, ncu capture.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round2_a.sh'
mkdir -p gpurun_out
nproc > gpurun_out/a_nproc.txt
B="python bench.py --no-cpu-baseline --no-secondary --no-objects"
# 1. parity first (the rows kernel is the default); stop at the first failure but keep going with the measurements
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_streaming.py tests/test_abi.py tests/test_host_layer.py -m gpu -q --maxfail=8 \
    --deselect tests/test_gpu_parity.py::test_config_c3_full_size_bit_exact > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/a_pytest.log)"
# 2. A/B as run in session A (the quads kernel and the carry-save count were deleted afterwards; kept for the record)
for rep in 1 2; do
  for w in c3 c2; do
    for k in quads rows rows_csa rows_hint; do
      export KS_MASK_KERNEL=rows KS_ROWS_COUNT=0 KS_ROWS_HINT=0
      [ $k = quads ] && export KS_MASK_KERNEL=quads
      [ $k = rows_csa ] && export KS_ROWS_COUNT=1
      [ $k = rows_hint ] && export KS_ROWS_HINT=1
      timeout 240 $B --workload $w > gpurun_out/a_${w}_${k}_r$rep.json 2> gpurun_out/a_${w}_${k}_r$rep.err
      python - "$w" "$k" "$rep" <<'PY'
import json, sys
w, k, rep = sys.argv[1:4]
try:
    d = json.load(open(f"gpurun_out/a_{w}_{k}_r{rep}.json"))
    r = d["roofline"]
    print(f"{w} {k} run {rep}: K2 {1e3 * r['kernel_ms']:.2f} us  frac {r['frac']:.4f}  step {1e3 * d['ms_per_step']:.1f} us  e2e {d['e2e']['value']:.3g}")
except Exception as e:
    print(f"{w} {k} run {rep}: FAILED ({e})")
PY
    done
  done
done
export KS_MASK_KERNEL=rows KS_ROWS_COUNT=0 KS_ROWS_HINT=0
# 3. ncu: full capture of the rows kernel on C3 and C2, then the launch list of one C3 bench command
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_mask_rows -s 3 -c 1 -o gpurun_out/r02_ncu_rows_c3 \
    $B --workload c3 --steps 1 --warmup 1 > gpurun_out/a_ncu_c3.log 2>&1
echo "ncu c3 rc=$?"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_mask_rows -s 3 -c 1 -o gpurun_out/r02_ncu_rows_c2 \
    $B --workload c2 --steps 1 --warmup 1 > gpurun_out/a_ncu_c2.log 2>&1
echo "ncu c2 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_c3.csv \
    $B --workload c3 --steps 2 --warmup 1 > gpurun_out/a_launches_c3.log 2>&1
echo "launch list rc=$?"
# 4. the remaining GPU tests (two-rank exchange on one GPU, the full C3 mask) and the other benches
timeout 420 python -m pytest tests/test_exchange_gpu.py tests/test_gpu_parity.py::test_config_c3_full_size_bit_exact -m gpu -q > gpurun_out/a_pytest2.log 2>&1
echo "pytest2 rc=$? $(tail -1 gpurun_out/a_pytest2.log)"
timeout 120 python bench_stream.py --seconds 5 > gpurun_out/a_stream_async.json 2> gpurun_out/a_stream_async.err
echo "stream async: $(cut -c1-400 gpurun_out/a_stream_async.json)"
timeout 120 python bench_stream.py --seconds 5 --sync > gpurun_out/a_stream_sync.json 2> gpurun_out/a_stream_sync.err
echo "stream sync: $(cut -c1-400 gpurun_out/a_stream_sync.json)"
KS_STREAM_HOST_LOOP=1 timeout 120 python bench_stream.py --seconds 5 --sync > gpurun_out/a_stream_hostloop.json 2> gpurun_out/a_stream_hostloop.err
echo "stream host loop: $(cut -c1-300 gpurun_out/a_stream_hostloop.json)"
timeout 200 python bench.py --policy least_allocated --no-cpu-baseline --no-secondary --no-objects > gpurun_out/a_c3_least.json 2> gpurun_out/a_c3_least.err
echo "least allocated c3: $(cut -c1-300 gpurun_out/a_c3_least.json)"
ls -la gpurun_out | tail -40
