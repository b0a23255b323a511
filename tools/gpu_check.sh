#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list + one full capture of the gather kernel.
# usage (from the repo root, under gpurun):  bash tools/gpu_check.sh [quick]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/nproc.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -4 gpurun_out/bench.log
if [ "$1" != "quick" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 1 --pool 2 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:rgcn_gather_d200 -s 4 -c 4 -f -o gpurun_out/prof_gather \
      python bench.py --steps 2 --warmup 1 --pool 2 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
  ls -la gpurun_out
fi
