#!/bin/bash
mkdir -p gpurun_out
{
timeout 300 python -m pytest tests/test_parallel_gpu.py -q -m gpu -s -p no:cacheprovider -k "peer-memory-all-reduce" 2>&1 | grep -v "^$" | tail -4
PESR_DDP_TRACE=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 30 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('N=2', d['ms_per_step'], d['value'], d['ddp'])"
} > gpurun_out/r2ddpfin.log 2>&1
cat gpurun_out/r2ddpfin.log | cut -c1-400
