#!/bin/bash
# ncu launch list of the training step (one GPU): per-kernel time shares of fwd+bwd+optimiser.  usage: train_launches.sh [tag] [dropout]
TAG=${1:-r2_train}; DROP=${2:-0.0}
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --mode train --steps 2 --warmup 3 --pool 2 --timestamps 60 --dropout $DROP > gpurun_out/${TAG}_ncu.log 2>&1
python tools/launch_shares.py gpurun_out/${TAG}_launches.csv "bench.py --mode train --steps 2 --warmup 3 --dropout $DROP (5 training steps)" > gpurun_out/${TAG}_shares.txt 2>&1 || true
