#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tiny" > $O/r2b_t7.log 2>&1
timeout 600 python scripts/ab_inproc.py --steps 20 --rounds 3 --out $O/ab11_headline.json "" "MAML_B200_TAIL_ONCHIP=0" "MAML_B200_TAIL_FUSE=0" > $O/ab11_headline.txt 2>&1
tail -3 $O/r2b_t7.log; tail -5 $O/ab11_headline.txt
