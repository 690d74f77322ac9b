#!/bin/bash
# round 2, first GPU call: full GPU suite, the new bench line (2-D + 3-D + big arms + CPU baseline), launch list
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02a_pytest.log; tail -5 gpurun_out/r02a_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; tail -c 600 gpurun_out/r02a_bench.err; head -c 3000 gpurun_out/r02a_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02a_launches_bench2d.csv python bench.py --steps 2 --warmup 1 --skip-3d --skip-big --no-cpu-baseline > gpurun_out/r02a_ncu_bench.log 2>&1; tail -2 gpurun_out/r02a_ncu_bench.log
