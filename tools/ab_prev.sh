#!/bin/bash
# Same-box A/B of the in-tree library against pesr_b200/libpesr_b200_prev.so (built from an earlier commit):
# full GPU test suite on the in-tree build, then the GAN and pretrain step on both, alternating.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/ab_tests.log
B="--steps 20 --warmup 5 --no-cpu-baseline --no-extras"
for w in gan pretrain; do
  PESR_B200_LIB=$PWD/pesr_b200/libpesr_b200_prev.so timeout 600 python bench.py --workload $w $B > gpurun_out/ab_${w}_prev.json 2> gpurun_out/ab_${w}_prev.err
  timeout 600 python bench.py --workload $w $B > gpurun_out/ab_${w}_new.json 2> gpurun_out/ab_${w}_new.err
done
cat gpurun_out/ab_tests.log
for f in gpurun_out/ab_*_prev.json gpurun_out/ab_*_new.json; do python -c "
import json,sys
l=[x for x in open('$f') if x.startswith('{')]
d=json.loads(l[-1]); r=d['roofline']
print('$f', round(d['ms_per_step'],3), round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'host', round(d.get('host_enqueue_ms_per_step',0),1), 'igemm', round(r['achieved']), round(r['avg_launch_us'],1), 'wgrad', round(r['wgrad']['achieved']), d['clocks']['reasons'])"; done
