#!/bin/bash
# usage (GPU box): tools/ab_run.sh <variant.so> <command...>: runs the command with tools/ab/<variant.so> in place of the library
set -u
so=$1; shift
cp raytracer.js_b200/librt_b200.so /tmp/librt_b200.keep
cp tools/ab/$so raytracer.js_b200/librt_b200.so
"$@"
rc=$?
cp /tmp/librt_b200.keep raytracer.js_b200/librt_b200.so
exit $rc
