#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/tc_timeline_r2d.txt
for p in 1 0; do for m in eval train; do for b in 1 3; do MAML_B200_TC_PUSH=$p timeout 100 python scripts/tc_timeline.py omniglot_mamlpp_5w1s $m $b >> $O/tc_timeline_r2d.txt 2>/dev/null; done; done; done
timeout 600 python scripts/ab_inproc.py --steps 20 --rounds 2 --out $O/ab6_headline.json "" "MAML_B200_TC_PUSH=0" > $O/ab6_headline.txt 2>&1
cat $O/tc_timeline_r2d.txt; tail -4 $O/ab6_headline.txt
