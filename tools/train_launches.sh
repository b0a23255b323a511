#!/bin/bash
# ncu launch list of the training step (one GPU): per-kernel time shares of fwd+bwd+optimiser
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_train_launches.csv \
  python bench.py --mode train --steps 2 --warmup 3 --pool 2 --timestamps 60 > gpurun_out/r2_train_ncu.log 2>&1
python tools/launch_shares.py gpurun_out/r2_train_launches.csv "bench.py --mode train --steps 2 --warmup 3 (5 training steps)" > gpurun_out/r2_train_shares.txt 2>&1 || true
