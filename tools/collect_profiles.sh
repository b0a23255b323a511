#!/bin/bash
# Round-2 evidence run (one GPU): launch list of the default bench + one full capture per hot kernel.
# Outputs under gpurun_out/; tools/ncu_summary.py turns the reports into profiles/r02_*.txt.   usage: collect_profiles.sh [what...]
set -x
WHAT=${@:-launches gather gemm gru induce bwd sanitize}
cap() {  # name regex skip count mode-args...
  name=$1; regex=$2; skip=$3; cnt=$4; shift 4
  ncu --set full --import-source on --clock-control none -k "regex:$regex" -s $skip -c $cnt -f -o gpurun_out/r02_$name \
    python bench.py "$@" > gpurun_out/r02_$name.log 2>&1
}
for w in $WHAT; do case $w in
launches) ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 2 --warmup 1 --pool 2 --no-cpu-baseline --no-train > gpurun_out/r02_launch_bench.log 2>&1 ;;
gather) cap gather 'rgcn_gather_stream|rgcn_gather_d200' 8 4 --steps 2 --warmup 1 --pool 2 --no-cpu-baseline --no-e2e --no-train ;;
gemm) cap gemm 'umma_gemm_packed' 8 2 --steps 2 --warmup 1 --pool 2 --no-cpu-baseline --no-e2e --no-train ;;
gru) cap gru 'gru_recur' 1 1 --steps 2 --warmup 1 --pool 2 --no-cpu-baseline --no-e2e --no-train ;;
induce) cap induce 'induce_' 10 5 --steps 4 --warmup 1 --pool 2 --no-cpu-baseline --no-train ;;
bwd) cap bwd 'rgcn_gather_stream_kernel<.*1>|rgcn_dh_tile|rgcn_dw_d200|gru_gate_bwd|sgemm_tn_splitk|adam_step' 20 8 --mode train --steps 1 --warmup 3 --pool 2 ;;
sanitize)
  compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_rgcn.py -q -x -k "golden or zero_edge or edge_cases" > gpurun_out/r02_racecheck.log 2>&1
  compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_rgcn.py tests/test_gpu_gru_renet.py -q -x -k "golden or edge_cases" > gpurun_out/r02_memcheck.log 2>&1 ;;
esac; done
ls -la gpurun_out | tail -20
