#!/bin/bash
# A/B of the gather kernels on one GPU box: parity tests, then the micro-benchmark per kernel / stream configuration.
# usage: gather_ab.sh [preset] [stream cfgs...]   (cfg "Nh" = configuration N with the dataset's hot-relation list passed)
P=${1:-icews18}; shift
timeout 900 python -m pytest tests/test_gpu_rgcn.py -x -q > gpurun_out/ab_rgcn.log 2>&1; tail -4 gpurun_out/ab_rgcn.log
RENET_GATHER_KERNEL=tile timeout 300 python tools/bench_gather.py $P > gpurun_out/ab_bg_tile.log 2>&1; grep layer gpurun_out/ab_bg_tile.log
for c in ${@:-0}; do
  H=0; [[ $c == *h ]] && H=1
  RENET_HOT_LIST=$H RENET_STREAM_CFG=${c%h} RENET_GATHER_KERNEL=stream timeout 300 python tools/bench_gather.py $P > gpurun_out/ab_bg_stream$c.log 2>&1; echo cfg $c; grep "layer\|repro" gpurun_out/ab_bg_stream$c.log
done
