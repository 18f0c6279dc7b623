#!/bin/bash
# usage: gpu_host_threads.sh <threads> ...   -- bench sensitivity to the host post-processing threads per context
for t in "$@"; do
  MAB_HOST_THREADS=$t python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null > /tmp/sweep.json
  [ -s /tmp/sweep.json ] && python -c "import json; d=json.load(open('/tmp/sweep.json')); print('host_threads', $t, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],1))"
done
