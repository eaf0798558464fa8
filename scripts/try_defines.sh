#!/bin/bash
# Development aid: rebuild libmvgpu.so with extra -D defines and print the kernel times of a short bench.
#   scripts/try_defines.sh "MVG_LT_STRIDE=48" "K2_POLL_NS=200" ...   (one build + bench per argument; "" = defaults)
for defs in "$@"; do
    MVG_EXTRA_DEFINES="$defs" python -c "from minivideo_b200 import build; build.build_gpu(True)" > /dev/null 2>&1 || { echo "build failed: $defs"; continue; }
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-frames 8 --stream-frames 0 2>/dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.readline())
k = d['kernels']
print('%-40s k1 %.3f k2 %.3f k3 %.3f ms  value %.0f' % ('$defs', k['k1']['ms_per_launch'], k['k2']['ms_per_launch'], k['k3']['ms_per_launch'], d['value']))"
done
python -c "from minivideo_b200 import build; build.build_gpu(True)" > /dev/null 2>&1
