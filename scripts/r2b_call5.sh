#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tiny" > $O/r2b_t2.log 2>&1
rm -f $O/tc_timeline_r2c.txt
for b in 1 2 3; do timeout 100 python scripts/tc_timeline.py omniglot_mamlpp_5w1s eval $b >> $O/tc_timeline_r2c.txt 2>/dev/null; done
timeout 600 python scripts/ab_inproc.py --steps 20 --rounds 2 --out $O/ab5_headline.json "" "MAML_B200_TC_PUSH=0" > $O/ab5_headline.txt 2>&1
timeout 300 python scripts/ab_inproc.py --config mini_imagenet_mamlpp_5w1s --steps 6 --warmup 3 --rounds 1 "" "MAML_B200_TC_PUSH=0" > $O/ab5_cfg3.txt 2>&1
tail -3 $O/r2b_t2.log; cat $O/tc_timeline_r2c.txt; tail -4 $O/ab5_headline.txt; tail -3 $O/ab5_cfg3.txt
