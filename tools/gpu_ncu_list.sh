#!/bin/bash
# launch list of one GAN step: gpurun_out/$1_launches_gan.{csv,txt}
tag=${1:-cur}
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_gan.csv python tools/profile_step.py gan > gpurun_out/${tag}_prof.log 2>&1; echo ncu=$?
python tools/summarize_launches.py gpurun_out/${tag}_launches_gan.csv > gpurun_out/${tag}_launches_gan.txt; head -45 gpurun_out/${tag}_launches_gan.txt
