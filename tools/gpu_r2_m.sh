#!/bin/bash
# single GPU: what do reserved SMs cost / gain on their own?  (graph replay, 50 steps)
mkdir -p gpurun_out
{
for r in 0 2 4; do
  echo "== GAN reserve $r"
  PESR_RESERVE_SMS=$r timeout 300 python bench.py --steps 50 --warmup 3 --no-extras --no-cpu-baseline 2>> gpurun_out/r2m.err | tee gpurun_out/r2m_gan_r$r.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['sustained']['ms_per_step'] if 'ms_per_step' in d.get('sustained',{}) else d.get('sustained'), d['clocks'])"
done
for r in 0 4; do
  echo "== pretrain reserve $r"
  PESR_RESERVE_SMS=$r timeout 300 python bench.py --workload pretrain --steps 50 --warmup 3 --no-extras --no-cpu-baseline 2>> gpurun_out/r2m.err | tee gpurun_out/r2m_pre_r$r.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['clocks'])"
done
} > gpurun_out/r2m.log 2>&1
cat gpurun_out/r2m.log | cut -c1-300
