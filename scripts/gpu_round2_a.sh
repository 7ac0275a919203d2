#!/bin/bash
# Round-2 GPU session A (one gpurun call): parity of the rows kernel, A/B against the round-1 kernel, ncu capture.
#   gpurun --timeout 2400 -- 'bash scripts/gpu_round2_a.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
nproc > gpurun_out/a_nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest_rows.log 2>&1
echo "pytest rows rc=$? $(tail -1 gpurun_out/a_pytest_rows.log)"
for rep in 1 2; do
  for w in c3 c2; do
    for k in quads rows rows1; do
      if [ $k = rows1 ]; then export KS_MASK_KERNEL=rows KS_ROWS_COUNT=1; else export KS_MASK_KERNEL=$k KS_ROWS_COUNT=0; fi
      timeout 600 python bench.py --workload $w --no-cpu-baseline > gpurun_out/a_${w}_${k}_r$rep.json 2> gpurun_out/a_${w}_${k}_r$rep.err
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
export KS_MASK_KERNEL=rows KS_ROWS_COUNT=0
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mask_rows -s 3 -c 1 -o gpurun_out/r02_ncu_rows_c3 \
    python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/a_ncu_c3.log 2>&1
echo "ncu c3 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mask_rows -s 3 -c 1 -o gpurun_out/r02_ncu_rows_c2 \
    python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/a_ncu_c2.log 2>&1
echo "ncu c2 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_c3.csv \
    python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/a_launches_c3.log 2>&1
echo "launch list rc=$?"
ls -la gpurun_out | tail -30
