#!/bin/bash
# Round profiling pass on one B200 (run under gpurun): launch lists and one full capture of the dominant
# kernel of the headline config (primary stage) and of config 2 (bounce stage).  usage: tools/profile_round.sh TAG
TAG=${1:-rX}
O=gpurun_out
set -x
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rt_primary_kernel -s 4 -c 1 -f -o $O/prof_${TAG}_primary \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_primary.log 2>&1
python tools/config_bench.py c2 --spp 1 > $O/${TAG}_c2_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/${TAG}_c2_launches.csv \
    python tools/config_bench.py c2 --spp 1 --steps 1 > $O/${TAG}_c2_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rt_bounce_kernel -s 1 -c 1 -f -o $O/prof_${TAG}_c2_bounce \
    python tools/config_bench.py c2 --spp 1 --steps 1 > $O/${TAG}_c2_ncu_bounce.log 2>&1
