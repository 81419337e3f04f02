#!/bin/bash
# 2 GPUs: NCCL parity of the data-parallel step, then the weak-scaling bench with (a) the new defaults (FC1 factors
# all-gathered, 4 SMs left to NCCL, NCCL_MAX_CTAS=4), (b) the round-2-start behaviour, (c) 1 GPU on the same box
mkdir -p gpurun_out
run2() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 --steps 40 --warmup 3 > gpurun_out/r2n_$name.json 2> gpurun_out/r2n_$name.err; echo "$name exit=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2n_$name.json'))
    print('$name: N=2 ms/step', round(d['ms_per_step'],3), 'value', round(d['value'],1), 'sustained', d['sustained']['ms_per_step'], d['clocks'])
except Exception as e: print('$name failed', e)
PY
  tail -3 gpurun_out/r2n_$name.err
}
{
nvidia-smi -L
timeout 900 python -m pytest tests/test_parallel_gpu.py -q -m gpu -s -p no:cacheprovider 2>&1 | tail -8
run2 new
run2 old PESR_NO_FC1_GATHER=1 PESR_RESERVE_SMS=0 NCCL_MAX_CTAS=32
run2 gather_only PESR_RESERVE_SMS=0 NCCL_MAX_CTAS=32
run2 reserve8 PESR_RESERVE_SMS=8 NCCL_MAX_CTAS=8
timeout 300 python bench.py --steps 40 --warmup 3 --no-extras --no-cpu-baseline 2> gpurun_out/r2n_1gpu.err | tee gpurun_out/r2n_1gpu.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('1 GPU', d['ms_per_step'], d['sustained']['ms_per_step'])"
} > gpurun_out/r2n.log 2>&1
tail -60 gpurun_out/r2n.log | cut -c1-400
