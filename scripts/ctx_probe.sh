#!/bin/bash
# how long does the first CUDA call take with all GPUs visible vs one? (CLI start-up)
for vis in "" "0"; do
  if [ -n "$vis" ]; then export CUDA_VISIBLE_DEVICES=$vis; else unset CUDA_VISIBLE_DEVICES; fi
  python - <<'PY'
import time, ctypes, os
t=time.time()
lib=ctypes.CDLL("minivideo_b200/libmvgpu.so")
n=lib.mvg_device_count()
t1=time.time()
h=ctypes.c_void_p()
lib.mvg_create(ctypes.byref(h),0,240,135,24)
t2=time.time()
print("visible=%r devices=%d device_count %.3f s, mvg_create(2160p x24) %.3f s"%(os.environ.get("CUDA_VISIBLE_DEVICES"),n,t1-t,t2-t1),flush=True)
PY
done
