#!/bin/bash
# usage: run_xp.sh <label> [ENV=VAL ...]
label=$1; shift
port=$((29600 + RANDOM % 300))
for r in 0 1 2 3; do
  env "$@" RANK=$r WORLD_SIZE=4 MASTER_ADDR=127.0.0.1 MASTER_PORT=$port OMP_NUM_THREADS=1 timeout 150 python tests/_xp_worker.py > gpurun_out/xp_${label}_$r.log 2>&1 &
done
t0=$(date +%s.%N); wait; t1=$(date +%s.%N)
echo "$label: $(python -c "print(round($t1 - $t0, 1))") s; $(tail -1 gpurun_out/xp_${label}_0.log | cut -c1-120)"
