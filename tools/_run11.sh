(python -m pytest tests -m gpu -x -q 2>&1 | tail -3) 
run() { echo "== $*"; env "$@" python bench.py --steps 12 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), d['step_ms_min_median_max'], {k:round(v['ms'],3) for k,v in d['roofline']['all_kernels'].items()})"; }
run X=1
PGEOF_KNN_STATS=1 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu 2>&1 | grep "pgeof knn tile" | head -1
run PGEOF_FEATURES_HINT=3
run PGEOF_FEATURES_HINT=4
run PGEOF_FEATURES_HINT=4 PGEOF_FEATURES_UNPERMUTE=1
