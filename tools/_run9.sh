run() { echo "== $*"; env "$@" python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), d['step_ms_min_median_max'])"; }
run X=1
run PGEOF_BENCH_SYNC=1
run PGEOF_BENCH_NO_CLOCKS=1 PGEOF_BENCH_SYNC=1
run PYTORCH_CUDA_ALLOC_CONF=backend:native
run PGEOF_BENCH_NO_CLOCKS=1
