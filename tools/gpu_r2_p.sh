#!/bin/bash
# 2 GPUs: where does the N=2 step wait?  PESR_DDP_TRACE=1 measures the compute stream's waits for NCCL on the device
mkdir -p gpurun_out
run2() {  # name, env...
  name=$1; shift
  env "$@" PESR_DDP_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 --steps 40 --warmup 3 > gpurun_out/r2p_$name.json 2> gpurun_out/r2p_$name.err; echo "$name exit=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2p_$name.json') if l.startswith('{')][-1])
    print('$name: N=2 ms/step', round(d['ms_per_step'],3), 'value', round(d['value'],1), 'sustained', round(d['sustained']['ms_per_step'],3), 'host', round(d['host_enqueue_ms_per_step'],2), d['ddp'])
except Exception as e: print('$name failed', e)
PY
  grep -v "OMP_NUM\|^\*\*\*" gpurun_out/r2p_$name.err | tail -3
}
{
run2 new
run2 ctas8 PESR_RESERVE_SMS=4 NCCL_MAX_CTAS=8
run2 ctas2 PESR_RESERVE_SMS=2 NCCL_MAX_CTAS=2
run2 old PESR_NO_FC1_GATHER=1 PESR_RESERVE_SMS=0 NCCL_MAX_CTAS=32 PESR_DDP_BUCKET_MB=32
timeout 300 python tools/phase_times.py gan 20
} > gpurun_out/r2p.log 2>&1
tail -60 gpurun_out/r2p.log | cut -c1-500
