(python -m pytest tests -m gpu -x -q 2>&1 | tail -3) 
run() { echo "== $*"; env "$@" python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), d['step_ms_min_median_max'], {k:round(v['ms'],3) for k,v in d['roofline']['all_kernels'].items()})"; }
run X=1
run PGEOF_FEATURES_HINT=1
ncu --set full --clock-control none --import-source on -k regex:"features_direct|unpermute|row_count|row_scatter|pad_xyz" -s 5 -c 5 -o gpurun_out/r1g_prof python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/t10_ncu.log 2>&1; tail -2 gpurun_out/t10_ncu.log
