run() { echo "== $*"; env "$@" python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), {k:round(v['ms'],3) for k,v in d['roofline']['all_kernels'].items()})"; }
run X=1
run PGEOF_KNN_FLAGS=1
run PGEOF_KNN_FLAGS=2
run PGEOF_KNN_FLAGS=3
run PGEOF_KNN_WARPS=2
