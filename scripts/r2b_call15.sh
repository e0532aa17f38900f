#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tiny_pp or trace" > $O/r2b_t10.log 2>&1
timeout 600 python scripts/ab_inproc.py --steps 20 --rounds 3 --out $O/ab15_headline.json "" > $O/ab15_headline.txt 2>&1
timeout 300 python scripts/ab_inproc.py --config mini_imagenet_mamlpp_5w1s --steps 6 --warmup 3 --rounds 1 "" > $O/ab15_cfg3.txt 2>&1
timeout 300 python scripts/ab_inproc.py --config omniglot_mamlpp_20w5s --batch-size 8 --steps 6 --warmup 3 --rounds 1 "" > $O/ab15_cfg5.txt 2>&1
tail -3 $O/r2b_t10.log; tail -2 $O/ab15_headline.txt; tail -1 $O/ab15_cfg3.txt; tail -1 $O/ab15_cfg5.txt
