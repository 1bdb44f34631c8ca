(python -m pytest tests -m gpu -x -q 2>&1 | tail -2) 
run() { echo "== $*"; env "$@" python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), {k:round(v['ms'],3) for k,v in d['roofline']['all_kernels'].items()})"; }
run X=1
run PGEOF_KNN_FLAGS=2
ncu --set full --clock-control none --import-source on -k regex:"knn_tile_kernel" -s 3 -c 1 -o gpurun_out/r1i_prof python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/t18_ncu.log 2>&1; tail -1 gpurun_out/t18_ncu.log
