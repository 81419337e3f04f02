#!/bin/bash
mkdir -p gpurun_out
{
timeout 3000 python -m pytest tests -q -m gpu -s -p no:cacheprovider > gpurun_out/r2c_pytest.log 2>&1; echo "pytest exit=$?"
grep -E "passed|failed" gpurun_out/r2c_pytest.log | tail -3
grep -E "^FAILED|^ERROR" gpurun_out/r2c_pytest.log | head -20
timeout 1200 python bench.py --steps 50 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench exit=$?"
cut -c1-1500 gpurun_out/r2c_bench.json; tail -5 gpurun_out/r2c_bench.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_launches_gan.csv python tools/profile_step.py gan > gpurun_out/r2c_prof.log 2>&1; echo ncu=$?
python tools/summarize_launches.py gpurun_out/r2c_launches_gan.csv > gpurun_out/r2c_launches_gan.txt; head -30 gpurun_out/r2c_launches_gan.txt
} > gpurun_out/r2c.log 2>&1
tail -70 gpurun_out/r2c.log
