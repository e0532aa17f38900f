#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python scripts/ab_inproc.py --steps 20 --rounds 2 --out $O/ab2_headline.json \
  "" "MAML_B200_PDL=2" "MAML_B200_PDL=2 MAML_B200_PDL_CLUSTER=0" "MAML_B200_PDL=3" \
  "MAML_B200_TC_NB=2" "MAML_B200_TC_NB=3" "MAML_B200_TC_NB=4" "MAML_B200_WG_NSTAGE=2" "MAML_B200_WG_NSTAGE=3" \
  "MAML_B200_TC_NB=3 MAML_B200_WG_NSTAGE=2" "MAML_B200_PDL=2 MAML_B200_TC_NB=3 MAML_B200_WG_NSTAGE=2" \
  "MAML_B200_PDL=2 MAML_B200_TC_NB=3 MAML_B200_WG_NSTAGE=2 MAML_B200_BN_SIDE_CAP=148 MAML_B200_TC_SPLIT_SIDE=2" \
  > $O/ab2_headline.txt 2>&1
timeout 300 python scripts/ab_inproc.py --config omniglot_mamlpp_20w5s --batch-size 8 --steps 6 --warmup 3 --rounds 1 --out $O/ab2_cfg5.json \
  "" "MAML_B200_TC_NB=3" "MAML_B200_TC_NB=4" "MAML_B200_WG_NSTAGE=2" "MAML_B200_WG_NSTAGE=3" > $O/ab2_cfg5.txt 2>&1
timeout 300 python scripts/ab_inproc.py --config mini_imagenet_mamlpp_5w1s --steps 6 --warmup 3 --rounds 1 --out $O/ab2_cfg3.json \
  "" "MAML_B200_TC_NB=3" "MAML_B200_WG_NSTAGE=2" "MAML_B200_PDL=2" > $O/ab2_cfg3.txt 2>&1
MAML_B200_PDL=2 MAML_B200_TC_NB=3 MAML_B200_WG_NSTAGE=2 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tiny" > $O/ab2_tests.txt 2>&1
tail -16 $O/ab2_headline.txt; tail -7 $O/ab2_cfg5.txt; tail -6 $O/ab2_cfg3.txt; tail -3 $O/ab2_tests.txt
