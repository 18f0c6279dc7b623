#!/bin/bash
# usage: gpu_sweep.sh "<contexts> <ext_ctas>" ...   -- bench throughput for pipeline depth / resident k_extend CTAs per SM
for cfg in "$@"; do
  set -- $cfg
  MAB_EXT_CTAS=$2 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --contexts $1 2>/dev/null > /tmp/sweep.json
  [ -s /tmp/sweep.json ] && python -c "import json; d=json.load(open('/tmp/sweep.json')); print('ctx', d['config']['contexts_per_gpu'], 'ext_ctas', '$2', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],1))"
done
