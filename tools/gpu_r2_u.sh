#!/bin/bash
mkdir -p gpurun_out
run2() {  # name, env...
  name=$1; shift
  env "$@" PESR_DDP_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 --steps 40 --warmup 3 > gpurun_out/r2u_$name.json 2> gpurun_out/r2u_$name.err; echo "$name exit=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2u_$name.json') if l.startswith('{')][-1])
    print('$name: N=2 ms/step', round(d['ms_per_step'],3), 'sustained', round(d['sustained']['ms_per_step'],3), 'host', round(d['host_enqueue_ms_per_step'],2), {k: round(v,3) for k,v in d['ddp']['exposed_wait_ms_per_step'].items()})
except Exception as e: print('$name failed', e)
PY
  grep -v "OMP_NUM\|^\*\*\*" gpurun_out/r2u_$name.err | tail -5
}
{
run2 p2p
run2 p2p_nomm PESR_DDP_NO_MULTIMEM=1
run2 symm_nccl PESR_DDP_SYMM_NCCL=1
run2 nccl PESR_DDP_NCCL_ONLY=1
} > gpurun_out/r2u.log 2>&1
cat gpurun_out/r2u.log | cut -c1-400
