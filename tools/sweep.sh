#!/bin/bash
# usage: tools_sweep.sh VAR v1 v2 ... : runs a short bench per value of the env var and prints stage times
VAR=$1; shift
for v in "$@"; do
  env $VAR=$v python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.readline())
print('$VAR=$v', 'pairs/s %.0f' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'e2e %.0f' % d['e2e']['value'], json.dumps(d['stage_ms_per_step']), 'roofline %.3f' % d['roofline']['frac'])
"
done
