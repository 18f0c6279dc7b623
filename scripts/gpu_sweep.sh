#!/bin/bash
# contexts / resident extend CTAs / chunk size sweep of bench.py (one line per run: value, e2e, ms per step)
for cfg in "3 5 16384" "4 5 16384" "5 5 16384" "4 4 16384" "3 6 16384" "4 6 16384" "3 5 24576" "4 5 24576"; do
  set -- $cfg
  MAB_EXT_CTAS=$2 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --contexts $1 --batch-reads $3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ctx',d['config']['contexts_per_gpu'],'ctas',d['config']['extend_ctas_per_sm'],'reads',d['config']['batch_reads'],'value %.0f e2e %.0f ms/step %.1f' % (d['value'],d['e2e']['value'],d['ms_per_step']))"
done
