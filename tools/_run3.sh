mkdir -p gpurun_out
(python -m pytest tests -m gpu -x -q 2>&1 | tail -3) 
run() { echo "== $*"; env "$@" PGEOF_KNN_STATS=1 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu 2>&1 | grep "pgeof knn tile" | head -1; env "$@" python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), {k:round(v['ms'],3) for k,v in d['roofline']['all_kernels'].items()})"; }
run X=1
run PGEOF_KNN_Z=2.9
run PGEOF_KNN_Z=2.3
ncu --set full --clock-control none --import-source on -k regex:"knn_tile_kernel|knn_slow" -s 4 -c 2 -o gpurun_out/r1e_prof python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/t3_ncu.log 2>&1; tail -2 gpurun_out/t3_ncu.log
