#!/bin/bash
# One `ncu --set full` capture of the hot kernels of a depth-1 training step (same per-layer shapes as the bench
# config), exported as CSV text on the box: the .ncu-rep itself is too large to travel back through gpurun_out/.
set -u
OUT=gpurun_out
REP=/tmp/prof_hot
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off \
  -k regex:'gemm_ffn_up_kernel|ffn_mid_bwd_walk_kernel|gemm_bf16_kernel|attn_fwd_tc_kernel|attn_bwd_tc_kernel|layernorm_bwd_kernel|ffn_norm_fwd_kernel|adamw' \
  -f -o $REP python tools/profile_step.py --ncu --depth 1 > $OUT/ncu_hot.log 2>&1
tail -2 $OUT/ncu_hot.log
ncu -i $REP.ncu-rep --page raw --csv > $OUT/ncu_hot_raw.csv 2>/dev/null
ncu -i $REP.ncu-rep --page source --csv --kernel-name regex:ffn_mid_bwd_walk_kernel > $OUT/ncu_src_walk.csv 2>/dev/null
ncu -i $REP.ncu-rep --page source --csv --kernel-name regex:gemm_ffn_up_kernel > $OUT/ncu_src_ffnup.csv 2>/dev/null
ls -la $OUT/ncu_hot_raw.csv $OUT/ncu_src_walk.csv $OUT/ncu_src_ffnup.csv
# launch list of the full-depth step (one pass per kernel)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $OUT/launches_step.csv python tools/profile_step.py --ncu > /dev/null 2>&1
wc -l $OUT/launches_step.csv
du -sh $OUT
