#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tiny or plain" > $O/r2b_t8.log 2>&1
timeout 600 python scripts/ab_inproc.py --steps 20 --rounds 3 --out $O/ab12_headline.json "" "MAML_B200_TAIL_ONCHIP=1" > $O/ab12_headline.txt 2>&1
tail -3 $O/r2b_t8.log; grep "^round" $O/ab12_headline.txt; tail -3 $O/ab12_headline.txt
