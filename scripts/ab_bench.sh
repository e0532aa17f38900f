run() { echo "== $1"; env $1 timeout 200 python bench.py 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"; }
run "X=1"
run "MAML_B200_LAUNCH_PRIO=1"
run "MAML_B200_WG_ROWS=256"
run "MAML_B200_WG_ROWS=512"
run "MAML_B200_WG_ROWS=512 MAML_B200_LAUNCH_PRIO=1"
run "MAML_B200_WG_ROWS=1280 MAML_B200_LAUNCH_PRIO=1"
