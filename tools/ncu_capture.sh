#!/bin/bash
# Round-2 ncu evidence (run under gpurun on ONE B200): the launch list of a short bench run, one `--set full` capture of every
# kernel of the cfg-2 training step and of the eval top-k path.  Outputs in gpurun_out/ (raw CSV pages; tools/ncu_traffic.py
# turns them into profiles/r2_ncu_*.{json,md}).  Numbers printed by runs under ncu are never bench values.
set -x
OUT=gpurun_out
TRAIN="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras"
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file $OUT/r2_launches.csv $TRAIN > $OUT/r2_launches.log 2>&1
K='regex:sasrec_fwd_fused_kernel|sasrec_bwd_ffn_fused_kernel|attn_bwd_tc2_kernel|wgrad_tc_kernel|gemm_tc_kernel|adam_kernel|table_grad_kernel|score_bce_kernel|prep_scan_kernel|pos_grad_kernel|colsum_kernel|reduce_segments_kernel|fused_tiles_kernel|weight_image_kernel|neg_sample_kernel|sum_kernel|scale_grads_kernel'
ncu --set full --clock-control none --import-source off -k "$K" -s 300 -c 36 -o $OUT/r2_prof_train $TRAIN > $OUT/r2_prof_train.log 2>&1
ncu -i $OUT/r2_prof_train.ncu-rep --page raw --csv > $OUT/r2_prof_train_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source off -k 'regex:logits_topk_kernel|select_kernel|table_image_kernel' -s 12 -c 5 -o $OUT/r2_prof_eval python bench.py --mode eval > $OUT/r2_prof_eval.log 2>&1
ncu -i $OUT/r2_prof_eval.ncu-rep --page raw --csv > $OUT/r2_prof_eval_raw.csv 2>/dev/null
ls -la $OUT/r2_prof_* $OUT/r2_launches.csv
