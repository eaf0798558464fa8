#!/bin/bash
# Development aid: rebuild libmvgpu.so with extra -D defines and time the fused kernel (tests/tools/kf_time.py).
#   scripts/try_kf.sh "" "KF_POLL_NS=200" "KF_WARPS=32 KF_GROUP=2" ...   (one build + run per argument; "" = defaults)
# KF_FRAMES (1000) pictures, KF_LINES (2) output lines per variant, KF_FULL=1 also times the split pipeline.
for defs in "$@"; do
    MVG_EXTRA_DEFINES="$defs" python -c "from minivideo_b200 import build; build.build_gpu(True)" > /dev/null 2>&1 || { echo "build failed: $defs"; continue; }
    echo "== [$defs] $(grep -A2 'kf_reconILi1' minivideo_b200/libmvgpu.build.log | grep -o 'Used [0-9]* registers' | head -1) $(grep -A1 'kf_reconILi1' minivideo_b200/libmvgpu.build.log | grep -o '[0-9]* bytes spill stores' | head -1)"
    python tests/tools/kf_time.py ${KF_FRAMES:-1000} 5 $([ -z "$KF_FULL" ] && echo quick) 2>&1 | head -${KF_LINES:-3}
done
python -c "from minivideo_b200 import build; build.build_gpu(True)" > /dev/null 2>&1
