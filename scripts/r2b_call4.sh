#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > $O/r2b_t1.log 2>&1
for b in 1 2 3; do timeout 100 python scripts/tc_timeline.py omniglot_mamlpp_5w1s eval $b >> $O/tc_timeline_r2b.txt 2>/dev/null; done
for b in 1 2 3; do timeout 100 python scripts/tc_timeline.py omniglot_mamlpp_5w1s train $b >> $O/tc_timeline_r2b.txt 2>/dev/null; done
timeout 600 python bench.py --no-cpu-baseline > $O/bench_r2b_1gpu.json 2> $O/bench_r2b_1gpu.err
tail -3 $O/r2b_t1.log; cat $O/tc_timeline_r2b.txt; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2b_1gpu.json"))
print(d["value"], d["ms_per_step"], d["e2e"], d["roofline"]["frac"], d["roofline"]["by_class_ms_per_step"])
for o in d["other_configs"]: print(o["config"], o.get("value"), o.get("ms_per_step"))
PY
