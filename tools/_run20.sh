set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py > gpurun_out/r1c_bench_n1.json 2> gpurun_out/r1c_bench_n1.err; cat gpurun_out/r1c_bench_n1.json | head -c 400; echo
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1c_bench_ref.json 2>/dev/null; cat gpurun_out/r1c_bench_ref.json | head -c 300; echo
ncu --set full --clock-control none --import-source on -k regex:"knn_tile_kernel|features_direct" -s 2 -c 2 -o gpurun_out/r1c_prof python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/t20_ncu.log 2>&1; tail -1 gpurun_out/t20_ncu.log
