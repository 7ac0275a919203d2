#!/bin/bash
mkdir -p gpurun_out
timeout 100 scripts/build/write_bw > gpurun_out/w_write_bw2.json 2> gpurun_out/w_write_bw2.err
cat gpurun_out/w_write_bw2.err
