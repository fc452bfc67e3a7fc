#!/bin/bash
# usage (on the GPU box): tools/walk_profile.sh <tag>; needs tools/ab/librt_b200_prof.so (nvcc ... -DRT_WALK_PROFILE)
set -u
tag=${1:-wp}
mkdir -p gpurun_out
cp raytracer.js_b200/librt_b200.so /tmp/librt_b200.keep
cp tools/ab/librt_b200_prof.so raytracer.js_b200/librt_b200.so
timeout 600 python tools/walk_profile.py ${2:-c2} ${3:-1} > gpurun_out/${tag}_walk_profile.log 2>&1
cp /tmp/librt_b200.keep raytracer.js_b200/librt_b200.so
tail -5 gpurun_out/${tag}_walk_profile.log
