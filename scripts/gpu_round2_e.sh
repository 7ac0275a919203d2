#!/bin/bash
# Round-2 GPU session E (1 GPU): block size of the mask kernel, launch order of mask / argmax kernels, strata.
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-secondary --no-objects"
run() { # tag workload
  timeout 240 $B --workload $2 > gpurun_out/e_$2_$1.json 2> gpurun_out/e_$2_$1.err
  python - "$1" "$2" <<'PY'
import json, sys
tag, w = sys.argv[1:3]
try:
    d = json.load(open(f"gpurun_out/e_{w}_{tag}.json"))
    r = d["roofline"]
    print(f"{w} {tag}: K2 {1e3 * r['kernel_ms']:.2f} us  frac {r['frac']:.4f}  step {1e3 * d['ms_per_step']:.1f} us  e2e {d['e2e']['value']:.3g}")
except Exception as e:
    print(f"{w} {tag}: FAILED ({e})")
PY
}
for t in 512 640 768 896; do KS_ROWS_THREADS=$t run t$t c3; done
KS_ROWS_MASK_FIRST=0 run t768_argmaxfirst c3
KS_ROWS_MASK_FIRST=1 run t768_maskfirst c3
KS_ROWS_STRATA=0 run t768_nostrata c3
KS_ROWS_THREADS=512 KS_ROWS_MASK_FIRST=0 run t512_argmaxfirst c3
for t in 512 640 768; do KS_ROWS_THREADS=$t run t$t c2; done
KS_ROWS_STRATA=0 run t768_nostrata c2
KS_ROWS_THREADS=512 timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=3 -k "device_buffers or random_clusters_leftover or adversarial or c2_full" > gpurun_out/e_pytest_512.log 2>&1
echo "pytest 512: rc=$? $(tail -1 gpurun_out/e_pytest_512.log)"
KS_ROWS_STRATA=0 timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=3 -k "device_buffers or random_clusters_leftover or adversarial or c2_full" > gpurun_out/e_pytest_nostrata.log 2>&1
echo "pytest nostrata: rc=$? $(tail -1 gpurun_out/e_pytest_nostrata.log)"
