# usage: bash scripts/ab_bench.sh "ENV1=.. ENV2=.." "ENV=.." ...   (one bench.py run per argument; prints ms/step, tasks/s)
run() { echo "== $1"; env $1 timeout 200 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"; }
for a in "$@"; do run "$a"; done
