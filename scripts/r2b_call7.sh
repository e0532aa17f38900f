#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tiny" > $O/r2b_t3.log 2>&1
rm -f $O/tc_timeline_r2e.txt
for m in eval train; do for b in 1 3; do timeout 100 python scripts/tc_timeline.py omniglot_mamlpp_5w1s $m $b >> $O/tc_timeline_r2e.txt 2>/dev/null; done; done
timeout 600 python scripts/ab_inproc.py --steps 20 --rounds 2 --out $O/ab7_headline.json "" "MAML_B200_TC_PUSH=0" "MAML_B200_TC_SPLIT=4" > $O/ab7_headline.txt 2>&1
timeout 300 python scripts/ab_inproc.py --config mini_imagenet_mamlpp_5w1s --steps 6 --warmup 3 --rounds 1 "" > $O/ab7_cfg3.txt 2>&1
timeout 300 python scripts/ab_inproc.py --config omniglot_mamlpp_20w5s --batch-size 8 --steps 6 --warmup 3 --rounds 1 "" > $O/ab7_cfg5.txt 2>&1
tail -3 $O/r2b_t3.log; cat $O/tc_timeline_r2e.txt; tail -5 $O/ab7_headline.txt; tail -2 $O/ab7_cfg3.txt; tail -2 $O/ab7_cfg5.txt
