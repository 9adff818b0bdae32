#!/bin/bash
# Round evidence on one B200: (1) launch list of a full cfg2 training step (one pass per kernel), (2) ncu --set full of
# EVERY GEMM-family launch of that step (DRAM traffic, tensor-pipe activity), (3) ncu --set full of the other hot
# kernels of a depth-1 step (same per-layer shapes).  CSV text only travels back (the .ncu-rep files are too large).
set -u
OUT=gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $OUT/launches_step.csv python tools/profile_step.py --ncu > /dev/null 2>&1
wc -l $OUT/launches_step.csv
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:'gemm_bf16_kernel|gemm_ffn_up_kernel' \
  -f -o /tmp/prof_gemm python tools/profile_step.py --ncu > $OUT/ncu_gemm.log 2>&1
tail -1 $OUT/ncu_gemm.log
ncu -i /tmp/prof_gemm.ncu-rep --page raw --csv > $OUT/ncu_gemm_raw.csv 2>/dev/null
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off \
  -k regex:'attn_fwd_tc_kernel|attn_bwd_tc_kernel|attn_bwd_tc_dsum|ffn_mid_bwd_walk_kernel|layernorm_bwd_kernel|layernorm_fwd_kernel|ffn_norm_fwd_kernel|qk_l2norm|adamw|pack_multi|sumsq' \
  -f -o /tmp/prof_hot python tools/profile_step.py --ncu --depth 1 > $OUT/ncu_hot.log 2>&1
tail -1 $OUT/ncu_hot.log
ncu -i /tmp/prof_hot.ncu-rep --page raw --csv > $OUT/ncu_hot_raw.csv 2>/dev/null
ls -la $OUT/ncu_gemm_raw.csv $OUT/ncu_hot_raw.csv
du -sh $OUT
