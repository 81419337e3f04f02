#!/bin/bash
# Same-box A/B on the GAN / pretrain step (run on the GPU box through gpurun):
#   prev  = library built from the previous commit (pesr_b200/libpesr_b200_prev.so, if present)
#   nopdl = current library with programmatic dependent launch disabled, pdl = current default
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/ab_tests.log
B="--steps 20 --warmup 5 --no-cpu-baseline --no-extras"
for w in gan pretrain; do
  if [ -f pesr_b200/libpesr_b200_prev.so ]; then
    PESR_B200_LIB=$PWD/pesr_b200/libpesr_b200_prev.so timeout 600 python bench.py --workload $w $B > gpurun_out/ab_${w}_prev.json 2> gpurun_out/ab_${w}_prev.err
  fi
  PESR_NO_PDL=1 timeout 600 python bench.py --workload $w $B > gpurun_out/ab_${w}_nopdl.json 2> gpurun_out/ab_${w}_nopdl.err
  timeout 600 python bench.py --workload $w $B > gpurun_out/ab_${w}_pdl.json 2> gpurun_out/ab_${w}_pdl.err
done
cat gpurun_out/ab_tests.log
for f in gpurun_out/ab_*_*.json; do python -c "
import json,sys
l=[x for x in open('$f') if x.startswith('{')]
d=json.loads(l[-1]); print('$f', round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1), d['clocks']['reasons'])"; done
