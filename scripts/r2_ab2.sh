#!/bin/bash
# A/B on the two FLOP-heavy workloads only: usage scripts/r2_ab2.sh TAG ENVVAR "v0 v1"
TAG=$1; VAR=$2; VALS=$3
for v in $VALS; do
  for cfg in "mini_imagenet_mamlpp_5w1s 6 0" "omniglot_mamlpp_20w5s 6 8" "mini_imagenet_mamlpp_5w5s 6 2"; do
    set -- $cfg
    extra=""; [ "$3" != "0" ] && extra="--batch-size $3"
    env $VAR=$v timeout 300 python bench.py --config $1 --steps $2 --warmup 3 --no-cpu-baseline --no-extras $extra > gpurun_out/ab_${TAG}_${v}_$1.json 2> gpurun_out/ab_${TAG}_${v}_$1.err
  done
done
python - <<PY
import json, glob, os
for f in sorted(glob.glob("gpurun_out/ab_${TAG}_*.json")):
    try:
        d = json.load(open(f))
        print(os.path.basename(f), "value %.1f ms %.3f" % (d["value"], d["ms_per_step"]), {k: round(v, 3) for k, v in d["roofline"]["by_class_ms_per_step"].items()})
    except Exception as e:
        print(os.path.basename(f), "ERR", e)
PY
