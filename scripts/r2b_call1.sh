#!/bin/bash
# GPU call: realised critical chain of the current build + in-process A/B of the contention switches
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
MAML_B200_GRAPH_DOT=$O/graph_r2b.dot timeout 120 python scripts/trace_timeline.py > /dev/null 2> $O/trace_dot.err
timeout 120 python scripts/trace_timeline.py --full > $O/trace_streams_r2b.txt 2>> $O/trace_dot.err
MAML_B200_ONE_STREAM=1 timeout 120 python scripts/trace_timeline.py --full > $O/trace_serial_r2b.txt 2>> $O/trace_dot.err
timeout 60 python scripts/real_critical_path.py $O/graph_r2b.dot $O/trace_streams_r2b.txt $O/trace_serial_r2b.txt > $O/realised_chain_r2b.txt 2>&1
timeout 600 python scripts/ab_inproc.py --steps 20 --rounds 2 --out $O/ab_contention_headline.json \
  "" "MAML_B200_TC_SPLIT_SIDE=1" "MAML_B200_TC_SPLIT_SIDE=2" "MAML_B200_TC_SPLIT_L1=1" \
  "MAML_B200_TC_SPLIT_SIDE=1 MAML_B200_TC_SPLIT_L1=1" "MAML_B200_PDL=2" "MAML_B200_PDL=1" \
  "MAML_B200_BN_SIDE_CAP=148" "MAML_B200_BN_SIDE_CAP=296" "MAML_B200_TGT_SLOTS=1" "MAML_B200_NO_PRIO=1" "MAML_B200_PRE_ON_WG=1" \
  > $O/ab_contention_headline.txt 2>&1
timeout 300 python scripts/ab_inproc.py --config omniglot_mamlpp_20w5s --batch-size 8 --steps 6 --warmup 3 --rounds 1 --out $O/ab_contention_cfg5.json \
  "" "MAML_B200_TC_SPLIT_SIDE=1" "MAML_B200_PDL=2" "MAML_B200_BN_SIDE_CAP=296" > $O/ab_contention_cfg5.txt 2>&1
tail -16 $O/ab_contention_headline.txt; tail -6 $O/ab_contention_cfg5.txt; head -40 $O/realised_chain_r2b.txt
