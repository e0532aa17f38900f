#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tiny" > $O/r2b_t9.log 2>&1
MAML_B200_PDL=0 MAML_B200_ONE_STREAM=1 timeout 120 python scripts/trace_timeline.py --full > $O/trace_serial_r2d.txt 2>/dev/null
timeout 600 python scripts/ab_inproc.py --steps 20 --rounds 3 --out $O/ab14_headline.json "" > $O/ab14_headline.txt 2>&1
timeout 300 python scripts/ab_inproc.py --config mini_imagenet_mamlpp_5w1s --steps 6 --warmup 3 --rounds 1 "" > $O/ab14_cfg3.txt 2>&1
timeout 300 python scripts/ab_inproc.py --config omniglot_mamlpp_20w5s --batch-size 8 --steps 6 --warmup 3 --rounds 1 "" > $O/ab14_cfg5.txt 2>&1
tail -3 $O/r2b_t9.log; grep -E " export | adam |entries" $O/trace_serial_r2d.txt | tail -4; tail -2 $O/ab14_headline.txt; tail -1 $O/ab14_cfg3.txt; tail -1 $O/ab14_cfg5.txt
