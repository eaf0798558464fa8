# Dev script: kernel-2 start stagger sweep (bench.py, resident batch)
for s in 0 8 16; do echo "stagger=$s"; MVG_K2_STAGGER=$s python bench.py --no-cpu-baseline --e2e-frames 16 --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], {k:round(v['ms_per_launch'],3) for k,v in d['kernels'].items()})"; done
