python -m pytest tests -m gpu -x -q 2>&1 | tail -2
ITS=12 python tools/e2e_probe.py 2>&1 | tail -10
python bench.py --no-cpu 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), d['value'], d['step_ms_min_median_max'], {k:round(v['ms'],3) for k,v in d['roofline']['all_kernels'].items()}, d['e2e']['value'], d['e2e']['step_ms_min_median_max'])"
PGEOF_BENCH_DEBUG=1 python bench.py --no-cpu --no-e2e --steps 30 2>&1 >/dev/null | awk '{ if ($4+0 > 12.6) print }' | head
