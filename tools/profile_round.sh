#!/bin/bash
# Round profiling pass on one B200 (run under gpurun): launch lists and one full capture of the dominant
# kernels: headline config (primary stage), config 2 at 1 spp (bounce stage) and at 16 spp (resample stage).
# usage: tools/profile_round.sh TAG
TAG=${1:-rX}
O=gpurun_out
set -x
timeout 200 python bench.py --steps 20 --warmup 3 > $O/${TAG}_bench.log 2>$O/${TAG}_bench.err || exit 1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-group > $O/${TAG}_ncu_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rt_primary_kernel -s 4 -c 1 -f -o $O/prof_${TAG}_primary \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-group > $O/${TAG}_ncu_primary.log 2>&1
timeout 100 python tools/config_bench.py c2 --spp 1 > $O/${TAG}_c2_1spp.log 2>&1 || exit 1
timeout 100 python tools/config_bench.py c2 > $O/${TAG}_c2_16spp.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/${TAG}_c2_launches.csv \
    python tools/config_bench.py c2 --steps 1 > $O/${TAG}_c2_ncu_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rt_bounce_kernel -s 1 -c 1 -f -o $O/prof_${TAG}_c2_bounce \
    python tools/config_bench.py c2 --spp 1 --steps 1 > $O/${TAG}_c2_ncu_bounce.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rt_resample_kernel -s 0 -c 1 -f -o $O/prof_${TAG}_c2_resample \
    python tools/config_bench.py c2 --steps 1 > $O/${TAG}_c2_ncu_resample.log 2>&1
