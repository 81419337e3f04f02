#!/bin/bash
# Same-box A/B of an environment knob: tools/ab_env.sh PESR_NO_FUSED_BIAS  -> bench with VAR=1 ("off") and unset ("on").
VAR=$1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/ab_tests.log
B="--steps 20 --warmup 5 --no-cpu-baseline --no-extras"
for w in gan pretrain; do
  env $VAR=1 timeout 600 python bench.py --workload $w $B > gpurun_out/ab_${w}_off.json 2> gpurun_out/ab_${w}_off.err
  timeout 600 python bench.py --workload $w $B > gpurun_out/ab_${w}_on.json 2> gpurun_out/ab_${w}_on.err
done
cat gpurun_out/ab_tests.log
for f in gpurun_out/ab_*_off.json gpurun_out/ab_*_on.json; do python -c "
import json,sys
l=[x for x in open('$f') if x.startswith('{')]
d=json.loads(l[-1]); r=d['roofline']
print('$f', round(d['ms_per_step'],3), round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'host', round(d.get('host_enqueue_ms_per_step',0),1), 'igemm', round(r['achieved']), round(r['avg_launch_us'],1), 'wgrad', round(r['wgrad']['achieved']), d['clocks']['reasons'])"; done
