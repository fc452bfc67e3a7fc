#!/bin/bash
# Lane occupancy / timeline of the bounce and primary stages, from a tools build of the library (never the product):
#   tools/walk_profile.sh TAG [config=c2|c3|c4|c1] [spp=1] [mode=TIMELINE|PROFILE]
#     TIMELINE  -DRT_WALK_TIMELINE: %globaltimer stamps per warp / per packet, printed by rt_destroy (undistorted timing)
#     PROFILE   -DRT_WALK_PROFILE:  lanes per phase of the lock-step walk (atomics per iteration: counts only, not times)
# Run on the GPU box (gpurun); the result goes to gpurun_out/TAG_walk_profile.log.  The product library is put back.
set -u
tag=${1:-wp}; cfg=${2:-c2}; spp=${3:-1}; mode=${4:-TIMELINE}
mkdir -p gpurun_out tools/ab
so=tools/ab/librt_b200_${mode}.so
if [ ! -f $so ]; then
  (cd raytracer.js_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared \
     -Xcompiler -ffp-contract=off -cudart static -DRT_WALK_${mode} -o ../../$so rt_b200.cu) || exit 1
fi
cp raytracer.js_b200/librt_b200.so /tmp/librt_b200.keep
cp $so raytracer.js_b200/librt_b200.so
timeout 600 python tools/walk_profile.py $cfg $spp > gpurun_out/${tag}_walk_profile.log 2>&1
cp /tmp/librt_b200.keep raytracer.js_b200/librt_b200.so
tail -5 gpurun_out/${tag}_walk_profile.log
