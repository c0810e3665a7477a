#!/bin/bash
# throughput vs windows per launch (measurement aid)
for b in 32 64 96 128; do
  MSS_WATCHDOG_MS=5000 timeout 400 python bench.py --batch $b --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > /tmp/b.json 2>/tmp/b.err || { echo "batch $b FAILED"; tail -3 /tmp/b.err; continue; }
  python -c "import json; d=json.load(open('/tmp/b.json')); print('batch', $b, 'windows/s %.0f' % d['value'], 'ms/step %.3f' % d['ms_per_step'], 'grid', d['config']['grid_ctas'])"
done
