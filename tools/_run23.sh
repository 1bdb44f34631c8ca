run() { echo "== $*"; env "$@" python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), {k:round(v['ms'],3) for k,v in d['roofline']['all_kernels'].items()})"; }
python -c "import torch; p=torch.cuda.get_device_properties(0); print(p.L2_cache_size)"
run X=1
run PGEOF_L2_PERSIST_MB=0
run PGEOF_L2_PERSIST_MB=32
run PGEOF_L2_PERSIST_MB=64
