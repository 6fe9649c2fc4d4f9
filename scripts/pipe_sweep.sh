#!/bin/bash
# sweep of the frame-pipeline tuning switches (graph mode): us/frame of the c2 stream
for cfg in "444 32 1" "296 32 1" "222 32 1" "148 32 1" "296 16 1" "296 48 1" "444 32 0" "296 32 0" "148 32 0"; do
  set -- $cfg
  echo -n "fold_blocks=$1 long_blocks=$2 exclusive=$3 : "
  GEM_B200_FOLD_BLOCKS=$1 GEM_B200_LONG_BLOCKS=$2 GEM_B200_EXCLUSIVE=$3 timeout 100 python scripts/pipe_ab.py graph 2>&1 | tail -1 | cut -c1-75
done
