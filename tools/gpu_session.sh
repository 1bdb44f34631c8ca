#!/bin/bash
# One gpurun call: tests, bench lines, ncu launch list + full capture.  Usage: tools/gpu_session.sh TAG [steps...]
# (steps: tests bench configs ab ncu; default all)
TAG=${1:-r2a}; shift
STEPS=${@:-tests bench configs ab ncu}
OUT=gpurun_out
mkdir -p $OUT
python build_native.py > $OUT/${TAG}_build.log 2>&1
for s in $STEPS; do case $s in
tests)
  timeout 900 python -m pytest tests -m gpu -q -s --maxfail=40 -rf --timeout=600 > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -5 $OUT/${TAG}_pytest.log;;
bench)
  timeout 600 python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; cat $OUT/${TAG}_bench_n1.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('M', d['ms_per_step'], d['value'], d['roofline']['all_kernels'], d['e2e']['ms_per_step'] if d['e2e'] else None, d['cpu_baseline'])"
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err;;
configs)
  for c in C2 C3 C4; do timeout 900 python bench.py --config $c --steps 5 --e2e-steps 2 > $OUT/${TAG}_bench_$c.json 2> $OUT/${TAG}_bench_$c.err; python -c "import sys,json; d=json.loads(open('$OUT/${TAG}_bench_$c.json').read()); print('$c', d['ms_per_step'], d['value'], d['roofline']['all_kernels'], d['e2e']['ms_per_step'] if d['e2e'] else None)"; done
  timeout 1200 python bench.py --config C5 --steps 3 --e2e-steps 1 > $OUT/${TAG}_bench_C5.json 2> $OUT/${TAG}_bench_C5.err; tail -c 600 $OUT/${TAG}_bench_C5.json; tail -3 $OUT/${TAG}_bench_C5.err;;
ab)
  for v in "PGEOF_FEATURES_CTA=256" "PGEOF_FEATURES_CTA=128"; do env $v timeout 300 python bench.py --steps 10 --no-e2e --no-cpu > $OUT/${TAG}_ab.json 2>/dev/null; python -c "import sys,json; d=json.loads(open('$OUT/${TAG}_ab.json').read()); print('$v', d['ms_per_step'], d['roofline']['all_kernels'])"; done;;
ncu)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"knn_tile_kernel|features_direct" -s 6 -c 2 -o $OUT/${TAG}_prof -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/${TAG}_ncu.log 2>&1; tail -2 $OUT/${TAG}_ncu.log
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"optimal_scan" -s 1 -c 1 -o $OUT/${TAG}_opt -f python bench.py --config C5 --points 2000000 --steps 1 --warmup 1 --no-e2e --no-cpu > $OUT/${TAG}_ncu_opt.log 2>&1; tail -1 $OUT/${TAG}_ncu_opt.log
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"multiscale_moments|multiscale_eigen" -s 2 -c 2 -o $OUT/${TAG}_ms -f python bench.py --config C4 --points 2000000 --steps 1 --warmup 1 --no-e2e --no-cpu > $OUT/${TAG}_ncu_ms.log 2>&1; tail -1 $OUT/${TAG}_ncu_ms.log;;
esac; done
