(python -m pytest tests -m gpu -x -q 2>&1 | tail -3) 
run() { echo "== $*"; env "$@" python bench.py --steps 12 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), d['step_ms_min_median_max'], {k:round(v['ms'],3) for k,v in d['roofline']['all_kernels'].items()})"; }
run X=1
run PGEOF_KNN_WARPS=4
echo "== 1M points"; python bench.py --points 1000000 --steps 12 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), d['step_ms_min_median_max'], {k:round(v['ms'],3) for k,v in d['roofline']['all_kernels'].items()})"
