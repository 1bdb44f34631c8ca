python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py > gpurun_out/r1d_bench_n1.json 2> gpurun_out/r1d_bench_n1.err; head -c 300 gpurun_out/r1d_bench_n1.json; echo
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1d_bench_ref.json 2>/dev/null; head -c 200 gpurun_out/r1d_bench_ref.json; echo
