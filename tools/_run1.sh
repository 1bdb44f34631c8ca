set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
(time python -m pytest tests -m gpu -x -q) > gpurun_out/t1_pytest.log 2>&1; tail -5 gpurun_out/t1_pytest.log
PGEOF_KNN_STATS=1 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/t1_stats.log 2>&1; grep "pgeof knn tile" gpurun_out/t1_stats.log | head -3
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/t1_bench.json 2> gpurun_out/t1_bench.err; cat gpurun_out/t1_bench.json
for z in 2.5 3.3; do PGEOF_KNN_Z=$z PGEOF_KNN_STATS=1 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu 2>&1 | grep "pgeof knn tile" | head -1; PGEOF_KNN_Z=$z python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('z=$z', d['ms_per_step'], d['roofline']['all_kernels'])"; done
ncu --set full --clock-control none --import-source on -k regex:"knn_tile_kernel|features_kernel" -s 6 -c 2 -o gpurun_out/r1c_prof python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/t1_ncu.log 2>&1; tail -3 gpurun_out/t1_ncu.log
