#!/bin/bash
# 8 GPUs: weak-scaling bench with the new data-parallel defaults vs the round-2-start behaviour
mkdir -p gpurun_out
run8() {  # name, env...
  name=$1; shift
  env "$@" PESR_DDP_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 8 --steps 30 --warmup 3 > gpurun_out/r2s_$name.json 2> gpurun_out/r2s_$name.err; echo "$name exit=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2s_$name.json') if l.startswith('{')][-1])
    print('$name: N=8 ms/step', round(d['ms_per_step'],3), 'value', round(d['value'],1), 'sustained', round(d['sustained']['ms_per_step'],3), 'host', round(d['host_enqueue_ms_per_step'],2), {k: round(v,3) for k,v in (d['ddp']['exposed_wait_ms_per_step'] or {}).items()}, d['clocks'])
except Exception as e: print('$name failed', e)
PY
  grep -v "OMP_NUM\|^\*\*\*" gpurun_out/r2s_$name.err | tail -3
}
{
nvidia-smi -L | wc -l
run8 new
run8 old PESR_NO_FC1_GATHER=1 PESR_RESERVE_SMS=0 NCCL_MAX_CTAS=32 PESR_DDP_BUCKET_MB=32 PESR_DDP_LOW_PRIORITY=1
} > gpurun_out/r2s.log 2>&1
cat gpurun_out/r2s.log | cut -c1-500
