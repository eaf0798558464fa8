#!/bin/bash
# Development aid: build with the given defines, then one `ncu --set full` capture of a warm kf_recon launch
# (RGB mode = the 4th launch of tests/tools/kf_time.py) and one of a warm tiles launch, at the benchmark size.  Output: gpurun_out/<tag>.ncu-rep
#   scripts/ncu_kf.sh <tag> "<defines>"
tag=$1; defs=$2
MVG_EXTRA_DEFINES="$defs" python -c "from minivideo_b200 import build; build.build_gpu(True)" > /dev/null 2>&1 || { echo "build failed"; exit 1; }
cp minivideo_b200/libmvgpu.so gpurun_out/$tag.libmvgpu.so
mkdir -p gpurun_out/$tag.src && cp minivideo_b200/csrc/mvg_kernels.cuh minivideo_b200/csrc/mvg_fused.cuh gpurun_out/$tag.src/
python tests/tools/kf_time.py ${KF_FRAMES:-1000} 2 quick > gpurun_out/$tag.time.txt 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:kf_recon --launch-skip 3 -c 1 -f -o gpurun_out/${tag}_rgb python tests/tools/kf_time.py ${KF_FRAMES:-1000} 1 quick > gpurun_out/$tag.ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kf_recon --launch-skip 7 -c 1 -f -o gpurun_out/${tag}_tiles python tests/tools/kf_time.py ${KF_FRAMES:-1000} 1 quick >> gpurun_out/$tag.ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kf_recon --launch-skip 11 -c 1 -f -o gpurun_out/${tag}_thumbs python tests/tools/kf_time.py ${KF_FRAMES:-1000} 1 quick >> gpurun_out/$tag.ncu.log 2>&1
cat gpurun_out/$tag.time.txt
