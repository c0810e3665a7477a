#!/bin/bash
# throughput sweep: windows per launch x CTAs per window group (measurement aid)
for b in 16 32 64; do
  for g in 0 8 16 24 37; do
    MSS_GROUP_CTAS=$g MSS_WATCHDOG_MS=5000 timeout 300 python bench.py --batch $b --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > /tmp/b.json 2>/tmp/b.err || { echo "batch $b group $g FAILED"; tail -3 /tmp/b.err; continue; }
    python -c "import json; d=json.load(open('/tmp/b.json')); print('batch', $b, 'group_ctas', $g, 'windows/s %.0f' % d['value'], 'ms/step %.3f' % d['ms_per_step'], 'grid', d['config']['grid_ctas'])"
  done
done
