(python -m pytest tests -m gpu -x -q 2>&1 | tail -3) 
run() { echo "== $*"; env "$@" python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), {k:round(v['ms'],3) for k,v in d['roofline']['all_kernels'].items()})"; }
run PGEOF_FEATURES_WAVES=1
run PGEOF_FEATURES_WAVES=4
run PGEOF_FEATURES_WAVES=16
run PGEOF_FEATURES_WAVES=1000
run PGEOF_FEATURES_CTA=128 PGEOF_FEATURES_WAVES=1000
run PGEOF_FEATURES_CTA=256 PGEOF_FEATURES_WAVES=8
run PGEOF_FEATURES_CTA=1024 PGEOF_FEATURES_WAVES=8
