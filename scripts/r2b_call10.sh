#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tiny or builder or checkpoint" > $O/r2b_t6.log 2>&1
timeout 600 python scripts/ab_inproc.py --steps 20 --rounds 2 --out $O/ab10_headline.json "" "MAML_B200_PDL_CLUSTER=1" "MAML_B200_PDL_CLUSTER=2" > $O/ab10_headline.txt 2>&1
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > $O/bench_r2c_quick.json 2> $O/bench_r2c_quick.err
tail -3 $O/r2b_t6.log; tail -5 $O/ab10_headline.txt; python -c "
import json; d=json.load(open('gpurun_out/bench_r2c_quick.json')); print(d['value'], d['ms_per_step'], d['e2e'])"
