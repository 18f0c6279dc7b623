#!/bin/bash
# bench each occupancy variant of the extend kernel (MAB_EXT_CTAS_PER_SM = 4, 5, 6)
for LB in 4 5 6; do
  cp minialign_b200/libminialign_b200_lb$LB.so minialign_b200/libminialign_b200.so
  MAB_BENCH_VERBOSE=1 timeout 400 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_lb$LB.json 2> gpurun_out/bench_lb$LB.err
  echo "LB=$LB"; grep "step 1 device=True" gpurun_out/bench_lb$LB.err | cut -c1-330; python -c "import json;d=json.load(open('gpurun_out/bench_lb$LB.json'));print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],1))"
done
cp minialign_b200/libminialign_b200_lb4.so minialign_b200/libminialign_b200.so
timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
