#!/bin/bash
# Round-2 evidence run (one GPU): launch list of the default bench + one full capture per hot kernel.
# Outputs under gpurun_out/; tools/ncu_summary.py turns the reports into profiles/r02_*.txt.
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 2 --warmup 1 --pool 2 --no-cpu-baseline --no-train > gpurun_out/r02_launch_bench.log 2>&1
cap() {  # name regex skip count mode-args...
  name=$1; regex=$2; skip=$3; cnt=$4; shift 4
  ncu --set full --import-source on --clock-control none -k "regex:$regex" -s $skip -c $cnt -o gpurun_out/r02_$name \
    python bench.py "$@" > gpurun_out/r02_$name.log 2>&1
}
cap gather 'rgcn_gather_d200' 8 2 --steps 2 --warmup 1 --pool 2 --no-cpu-baseline --no-e2e --no-train
cap gemm 'umma_gemm_packed' 8 2 --steps 2 --warmup 1 --pool 2 --no-cpu-baseline --no-e2e --no-train
cap gru 'gru_recur' 1 1 --steps 2 --warmup 1 --pool 2 --no-cpu-baseline --no-e2e --no-train
cap induce 'induce_' 10 5 --steps 4 --warmup 1 --pool 2 --no-cpu-baseline --no-train
cap bwd 'rgcn_dh_tile|rgcn_dw_d200|gru_gate_bwd|sgemm_tn_splitk|adam_step' 20 8 --mode train --steps 1 --warmup 3 --pool 2
compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_rgcn.py -q -x -k "golden or zero_edge" > gpurun_out/r02_racecheck.log 2>&1
compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_rgcn.py tests/test_gpu_gru_renet.py -q -x -k "golden" > gpurun_out/r02_memcheck.log 2>&1
ls -la gpurun_out | tail -20
