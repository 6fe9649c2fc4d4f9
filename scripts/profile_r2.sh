#!/bin/bash
# Round-2 ncu evidence, one GPU.  Everything lands in gpurun_out/; scripts/ncu_rep_summary.py / ncu_summarize.py /
# ncu_lines.py turn it into the text files under profiles/.  (ncu flushes caches between replays: durations are cold-cache.)
set -x
export ADD_LOOP_PROFILE=0
O=gpurun_out
# 1. launch list of the c2 add loop (serial schedule: plain gem_add_points)
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file $O/r2_launches.csv python scripts/add_loop.py 60 plain > /dev/null 2>&1
# 2. the three add kernels, full set with source
for k in k_bin k_fold_long k_fold; do   # regex anchored at both ends: "k_fold" must not match k_fold_long
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^$k\$" -s 20 -c 2 -f -o $O/r2_$k python scripts/add_loop.py 30 plain > /dev/null 2>&1
done
# 3. the 1 M-point batch (gem_add_points_multi into 8192^2): DRAM bytes per launch
timeout 400 ncu --set full --clock-control none -k regex:"k_bin|k_fold" -s 36 -c 6 -f -o $O/r2_batch python scripts/batch_bench.py 20 > /dev/null 2>&1
# 4. the frame's other kernels
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_features|k_ray_trace|k_ray_collect|k_lowest_bitmap|k_export_colmajor" -s 30 -c 10 -f -o $O/r2_frame python scripts/frame_breakdown.py > /dev/null 2>&1
# 5. the tiled step's kernels on a one-rank "world" (the world = 1 baseline of the tiled path)
timeout 400 ncu --target-processes all --set full --clock-control none -k regex:"k_route_peer|k_bin_peer" -s 40 -c 4 -f -o $O/r2_tiled \
  python -m torch.distributed.run --nnodes=1 --nproc-per-node=1 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --steps 40 --warmup 5 > $O/r2_tiled_w1_under_ncu.log 2>&1
# 6. the tiled path at world = 1, not under ncu: the baseline the scaling curve should be read against
GEM_B200_BENCH_PARITY=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=1 --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus 2 --steps 500 --warmup 10 2>/dev/null | tail -1 > $O/r2_bench_tiled_world1.json
ls -la $O/*.ncu-rep
