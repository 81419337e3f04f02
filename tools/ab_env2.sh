#!/bin/bash
# GAN-step bench for several values of one environment knob: tools/ab_env2.sh VAR v1 v2 ...
VAR=$1; shift
mkdir -p gpurun_out
B="--steps 20 --warmup 5 --no-cpu-baseline --no-extras"
for rep in 1 2; do
for v in "$@"; do
  env $VAR=$v timeout 600 python bench.py --workload gan $B > gpurun_out/ab2_${v}_${rep}.json 2> gpurun_out/ab2_${v}_${rep}.err
  python -c "
import json,sys
l=[x for x in open('gpurun_out/ab2_${v}_${rep}.json') if x.startswith('{')]
d=json.loads(l[-1]); r=d['roofline']
print('$VAR=$v', round(d['ms_per_step'],3), round(d['value'],1), 'igemm', round(r['achieved']), round(r['avg_launch_us'],1))"
done
done
