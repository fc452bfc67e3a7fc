#!/bin/bash
# A/B of the primary stage's tuning knobs on one B200 (run under gpurun): rays per lane (packet size) x resident
# CTAs per SM the kernel is compiled for.  One bench line each into gpurun_out/TAG_sweep.jsonl.
TAG=${1:-rX}
O=gpurun_out
: > $O/${TAG}_sweep.jsonl
for ppl in 4 8; do
  for minb in 4 5 6; do
    RT_B200_PPL=$ppl RT_B200_PRIMARY_MINB=$minb timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-group 2>>$O/${TAG}_sweep.err \
      | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'ppl': $ppl, 'minb': $minb, 'ms_per_step': d['ms_per_step'], 'value': d['value'], 'stage_ms': d['roofline']['stage_ms_live'], 'replay_ms': d['config']['replay_ms_per_step_camera_standing_still'], 'e2e_ms': d['e2e']['frame_ms']}))" >> $O/${TAG}_sweep.jsonl
  done
done
cat $O/${TAG}_sweep.jsonl
