#!/bin/bash
# Round-2 GPU session C (1 GPU): aligned mask pitch A/B, parity of the new tests, ncu capture + launch list of the shipped kernel.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=5 -k "device_buffers or c3_full or least_allocated or random_clusters_leftover or adversarial" > gpurun_out/c_pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/c_pytest.log)"
B="python bench.py --no-cpu-baseline --no-secondary --no-objects"
for rep in 1 2; do
  for k in minimal aligned aligned_exact aligned_unsorted; do
    export KS_ROWS_SORT=0
    pitch=aligned; [ $k = minimal ] && pitch=minimal
    [ $k = aligned_exact ] && export KS_ROWS_SORT=1
    [ $k = aligned_unsorted ] && export KS_ROWS_SORT=2
    timeout 240 $B --workload c3 --mask-pitch $pitch > gpurun_out/c_c3_${k}_r$rep.json 2> gpurun_out/c_c3_${k}_r$rep.err
    python - "$k" "$rep" <<'PY'
import json, sys
k, rep = sys.argv[1:3]
try:
    d = json.load(open(f"gpurun_out/c_c3_{k}_r{rep}.json"))
    r = d["roofline"]
    print(f"c3 {k} run {rep}: K2 {1e3 * r['kernel_ms']:.2f} us  frac {r['frac']:.4f}  write_peak {r.get('write_peak')}  step {1e3 * d['ms_per_step']:.1f} us  e2e {d['e2e']['value']:.3g}")
except Exception as e:
    print(f"c3 {k} run {rep}: FAILED ({e})")
PY
  done
done
for srt in 0 2; do
  export KS_ROWS_SORT=$srt
  timeout 200 $B --workload c2 > gpurun_out/c_c2_sort$srt.json 2> gpurun_out/c_c2_sort$srt.err
  python -c "
import json; d=json.load(open('gpurun_out/c_c2_sort$srt.json')); r=d['roofline']; print('c2 sort $srt: K2', 1e3*r['kernel_ms'], 'us frac', r['frac'], 'step', 1e3*d['ms_per_step'], 'us')"
done
KS_ROWS_SORT=2 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=5 -k "device_buffers or random_clusters_leftover or adversarial or c2_full" > gpurun_out/c_pytest_unsorted.log 2>&1
echo "pytest unsorted rc=$? $(tail -1 gpurun_out/c_pytest_unsorted.log)"
unset KS_ROWS_SORT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_mask_rows -s 3 -c 1 -o gpurun_out/r02_ncu_rows_c3_final \
    $B --workload c3 --steps 1 --warmup 1 > gpurun_out/c_ncu_c3.log 2>&1
echo "ncu c3 rc=$?"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_mask_rows -s 3 -c 1 -o gpurun_out/r02_ncu_rows_c2_final \
    $B --workload c2 --steps 1 --warmup 1 > gpurun_out/c_ncu_c2.log 2>&1
echo "ncu c2 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_c3_final.csv \
    $B --workload c3 --steps 2 --warmup 1 > gpurun_out/c_launches_c3.log 2>&1
echo "launch list rc=$?"
timeout 200 $B --policy least_allocated > gpurun_out/c_c3_least.json 2> gpurun_out/c_c3_least.err
python -c "
import json; d=json.load(open('gpurun_out/c_c3_least.json')); print('least_allocated c3: step', d['ms_per_step'], 'ms; rest', d['roofline']['rest_of_step_ms'])"
