(python -m pytest tests -m gpu -x -q 2>&1 | tail -3) 
run() { echo "== $*"; env "$@" python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), d['step_ms_min_median_max'], {k:round(v['ms'],3) for k,v in d['roofline']['all_kernels'].items()})"; }
run X=1
run PGEOF_FEATURES_CTA=128
run PGEOF_FEATURES_CTA=256
run PGEOF_FEATURES_CTA=1024
run PGEOF_FEATURES_WAVES=4
run PGEOF_FEATURES_HINT=4
ncu --set full --clock-control none --import-source on -k regex:"features_direct" -s 1 -c 1 -o gpurun_out/r1h_prof python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/t13_ncu.log 2>&1; tail -2 gpurun_out/t13_ncu.log
