#!/bin/bash
mkdir -p gpurun_out
{
nvidia-smi -L
timeout 900 python -m pytest tests/test_parallel_gpu.py -q -m gpu -s -p no:cacheprovider 2>&1 | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/r2_2gpu_bench.json 2> gpurun_out/r2_2gpu_bench.err; echo "bench2 exit=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r2_2gpu_bench.json'))
print('N=2 ms/step', round(d['ms_per_step'],3), 'value', round(d['value'],1), d['config']['launch'], 'sustained', d['sustained']['ms_per_step'])
PY
tail -5 gpurun_out/r2_2gpu_bench.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 3 --graph-multi > gpurun_out/r2_2gpu_bench_graph.json 2> gpurun_out/r2_2gpu_bench_graph.err; echo "bench2 graph exit=$?"
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_2gpu_bench_graph.json'))
    print('N=2 graph ms/step', round(d['ms_per_step'],3), 'value', round(d['value'],1), d['config']['launch'], d.get('graph_note'))
except Exception as e: print('graph multi failed', e)
PY
tail -5 gpurun_out/r2_2gpu_bench_graph.err
} > gpurun_out/r2_2gpu.log 2>&1
tail -45 gpurun_out/r2_2gpu.log
