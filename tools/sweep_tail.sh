#!/bin/bash
# sweep of the shared-memory tail thresholds (measurement aid)
for tv in 0 64 256 1024; do
  echo "== tail_vars $tv"
  MSS_TAIL_VARS=$tv MSS_TAIL_ENTS=$((tv*4)) MSS_WATCHDOG_MS=3000 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > /tmp/b.json 2>&1; python -c "import json; d=json.load(open('/tmp/b.json')); print('windows/s', d['value'], 'ms/step', d['ms_per_step'])"
  MSS_TAIL_VARS=$tv MSS_TAIL_ENTS=$((tv*4)) timeout 100 python tools/trace_window.py c2 1 > /tmp/t.txt 2>&1; sed -n 2p /tmp/t.txt
done
