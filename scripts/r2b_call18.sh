#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tiny_pp or trace or bern" > $O/r2b_t11.log 2>&1
timeout 300 python scripts/ab_inproc.py --steps 20 --rounds 2 "" > $O/ab18_headline.txt 2>&1
timeout 120 python scripts/trace_timeline.py > $O/trace_r2e_summary.txt 2>&1
tail -3 $O/r2b_t11.log; tail -1 $O/ab18_headline.txt; grep -E "entries" $O/trace_r2e_summary.txt
