#!/bin/bash
# config 3 with different distance-window sets (XP_WINDOWS = inner boundaries)
for w in "2,4,8,16,32,64,256" "2,4,8,16,32,64,128,256,512" "3,6,12,24,48,96,384" "2,4,6,8,12,16,24,32,64,256" "4,8,16,32,64,256" "2,4,8,12,16,24,32,48,64,128,512" "3,6,9,12,18,24,36,48,96,384"; do
  XP_WINDOWS=$w timeout 300 python bench.py --config 3 --no-cpu 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())['detail']
print('$w', 'frame %.2f' % d['ms_per_frame'], {k: round(v,2) for k,v in d['sweep_stage_ms'].items()}, 'cand %.1f' % d['candidates_per_sphere'], 'hit', round(d['hit_fraction'],6))"
done
