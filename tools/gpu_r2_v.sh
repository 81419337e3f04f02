#!/bin/bash
# 8 GPUs: which gradient-reduction path?  (a) peer-memory multimem kernel, (b) NCCL 4 CTAs / 16 MB buckets,
# (c) NCCL 8 CTAs over 4 reserved SMs, (d) NCCL NVLS with 4 CTAs
mkdir -p gpurun_out
run8() {  # name, env...
  name=$1; shift
  env "$@" PESR_DDP_TRACE=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 8 --steps 30 --warmup 3 > gpurun_out/r2v_$name.json 2> gpurun_out/r2v_$name.err; echo "$name exit=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2v_$name.json') if l.startswith('{')][-1])
    print('$name: N=8 ms/step', round(d['ms_per_step'],3), 'value', round(d['value'],1), 'sustained', round(d['sustained']['ms_per_step'],3), 'host', round(d['host_enqueue_ms_per_step'],2), {k: round(v,3) for k,v in (d['ddp']['exposed_wait_ms_per_step'] or {}).items()})
except Exception as e: print('$name failed', e)
PY
  grep -v "OMP_NUM\|^\*\*\*" gpurun_out/r2v_$name.err | tail -4 | cut -c1-300
}
{
run8 p2p
run8 nccl_b16 PESR_DDP_NCCL_ONLY=1 PESR_DDP_BUCKET_MB=16
run8 nccl_c8 PESR_DDP_NCCL_ONLY=1 NCCL_MAX_CTAS=8 PESR_RESERVE_SMS=4
run8 nccl_nvls PESR_DDP_NCCL_ONLY=1 NCCL_ALGO=allreduce:NVLS
} > gpurun_out/r2v.log 2>&1
cat gpurun_out/r2v.log | cut -c1-500
