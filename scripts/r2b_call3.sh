#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python scripts/ab_inproc.py --steps 20 --rounds 2 --out $O/ab3_headline.json \
  "" "MAML_B200_PDL=0" "MAML_B200_PDL=0 MAML_B200_TC_NB_FIT=0" "MAML_B200_TC_NB_FIT=0" \
  "MAML_B200_TC_NB_SIDE=2" "MAML_B200_TC_NB_SIDE=3" "MAML_B200_TC_NB_SIDE=4" "MAML_B200_WG_NSTAGE=2" \
  "MAML_B200_WG_NSTAGE=2 MAML_B200_TC_NB_SIDE=3" "MAML_B200_WG_NSTAGE=2 MAML_B200_TC_NB_SIDE=3 MAML_B200_BN_SIDE_CAP=148" \
  "MAML_B200_WG_NSTAGE=2 MAML_B200_TC_NB_SIDE=3 MAML_B200_TC_SPLIT_SIDE=2" \
  > $O/ab3_headline.txt 2>&1
timeout 300 python scripts/ab_inproc.py --config omniglot_mamlpp_20w5s --batch-size 8 --steps 6 --warmup 3 --rounds 1 --out $O/ab3_cfg5.json \
  "" "MAML_B200_TC_NB=4" "MAML_B200_TC_NB=4 MAML_B200_TC_NB_SIDE=3" "MAML_B200_TC_NB=3" "MAML_B200_TC_NB=4 MAML_B200_WG_NSTAGE=2" "MAML_B200_TC_NB=5" > $O/ab3_cfg5.txt 2>&1
timeout 300 python scripts/ab_inproc.py --config mini_imagenet_mamlpp_5w1s --steps 6 --warmup 3 --rounds 1 --out $O/ab3_cfg3.json \
  "" "MAML_B200_TC_NB=4" "MAML_B200_TC_NB=4 MAML_B200_TC_NB_SIDE=3" "MAML_B200_TC_NB=3" "MAML_B200_TC_NB=4 MAML_B200_WG_NSTAGE=2" "MAML_B200_TC_NB=5" > $O/ab3_cfg3.txt 2>&1
tail -14 $O/ab3_headline.txt; tail -8 $O/ab3_cfg5.txt; tail -8 $O/ab3_cfg3.txt
