#!/bin/bash
# A/B of the bounce stage's resident-CTAs knob (and with it the registers per thread and the walk-stack capacity in
# shared memory) on configs[2] at 1 and 16 spp.  One line each into gpurun_out/TAG_bounce_sweep.jsonl.
TAG=${1:-rX}
O=gpurun_out
: > $O/${TAG}_bounce_sweep.jsonl
for minb in 8 6 5 4; do
  for spp in 1 16; do
    RT_B200_BOUNCE_MINB=$minb timeout 200 python tools/config_bench.py c2 --spp $spp 2>>$O/${TAG}_bounce_sweep.err \
      | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'minb': $minb, 'spp': $spp, 'frame_ms': d['frame_ms'], 'Mrays_per_s': d['Mrays_per_s']}))" >> $O/${TAG}_bounce_sweep.jsonl
  done
done
cat $O/${TAG}_bounce_sweep.jsonl
