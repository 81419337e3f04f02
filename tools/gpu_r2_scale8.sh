#!/bin/bash
# the driver's 8-GPU launch, verbatim
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 30 --warmup 3 > gpurun_out/r02b_bench_gan_8gpu.json 2> gpurun_out/r02b_bench_gan_8gpu.err
echo "exit=$? stdout lines: $(wc -l < gpurun_out/r02b_bench_gan_8gpu.json)"
head -c 700 gpurun_out/r02b_bench_gan_8gpu.json; echo
grep -v "OMP_NUM\|^\*\*\*" gpurun_out/r02b_bench_gan_8gpu.err | tail -5
