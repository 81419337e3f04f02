#!/bin/bash
mkdir -p gpurun_out
python tools/phase_times.py gan 20 > gpurun_out/r2k_phase_gan.txt 2>&1
python tools/phase_times.py gan 10 detail > gpurun_out/r2k_phase_gan_detail.txt 2>&1
python tools/phase_times.py pretrain 20 > gpurun_out/r2k_phase_pretrain.txt 2>&1
cat gpurun_out/r2k_phase_gan.txt gpurun_out/r2k_phase_pretrain.txt
tail -5 gpurun_out/r2k_phase_gan_detail.txt
