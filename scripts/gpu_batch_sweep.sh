#!/bin/bash
# usage: gpu_batch_sweep.sh <batch_reads> ...   -- bench throughput and the solo k_extend pass for several batch sizes
for b in "$@"; do
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --batch-reads $b 2>/dev/null > /tmp/sweep.json
  [ -s /tmp/sweep.json ] && python -c "import json; d=json.load(open('/tmp/sweep.json')); r=d['roofline']; print('batch', $b, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],1), 'k_extend ms', round(r['ms_per_launch'],1), 'gcups', round(r['gcups']), 'frac', round(r['integer']['frac'],3))"
done
