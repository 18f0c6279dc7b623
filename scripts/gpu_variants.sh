#!/bin/bash
# bench each occupancy variant of the extend kernel (MAB_EXT_CTAS_PER_SM = resident CTAs per SM); usage: gpu_variants.sh 5 6 7 8
cp minialign_b200/libminialign_b200.so /tmp/lib_keep.so
for LB in "$@"; do
  cp minialign_b200/libminialign_b200_lb$LB.so minialign_b200/libminialign_b200.so
  MAB_BENCH_VERBOSE=1 timeout 400 python bench.py --steps 2 --warmup 2 --contexts 1 --no-cpu-baseline > gpurun_out/bench_lb$LB.json 2> gpurun_out/bench_lb$LB.err
  echo "LB=$LB"; grep "device=True" gpurun_out/bench_lb$LB.err | tail -1 | cut -c1-200; python -c "import json;d=json.load(open('gpurun_out/bench_lb$LB.json'));print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],1))"
done
cp /tmp/lib_keep.so minialign_b200/libminialign_b200.so
