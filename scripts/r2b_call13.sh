#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
MAML_B200_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:export_kernel -c 2 -o $O/prof_r2b_export -f python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > $O/ncu_r2b_export.log 2>&1
ls -la $O/prof_r2b_export.ncu-rep; tail -3 $O/ncu_r2b_export.log | cut -c1-300
