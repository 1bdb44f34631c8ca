(python -m pytest tests -m gpu -x -q 2>&1 | tail -3) 
run() { echo "== $*"; env "$@" python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), {k:round(v['ms'],3) for k,v in d['roofline']['all_kernels'].items()})"; }
run X=1
PGEOF_KNN_STATS=1 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu 2>&1 | grep "pgeof knn tile" | head -1; CONFIGS=3 PGEOF_KNN_STATS=1 python tools/configs_probe.py 2>&1 | awk "!seen[\$0]++" | tail -3
