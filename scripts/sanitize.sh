#!/bin/bash
# compute-sanitizer over a selection of the GPU suite that launches every kernel of the library (one GPU, ~1 min per tool).
# Logs -> gpurun_out/r2_{memcheck,racecheck,synccheck}.log; the summary lines are kept in profiles/r2_sanitizers.txt.
SEL="test_fused_add_c1_all_layers or test_order_dependence_dense_collisions or test_very_long_cell_lists or test_stream_mode_overlapping or test_multi_frame_stream_with_scroll or test_loop_closure or test_multi_segment or test_two_tiles_on_one_gpu or test_pcl_record"
for tool in ${@:-memcheck racecheck synccheck}; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 99 python -m pytest tests -m gpu -q -k "$SEL" > gpurun_out/r2_$tool.log 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/r2_$tool.log
done
