#!/bin/bash
# ncu --set full of the cfg-3 kernels (GRU4Rec recurrence, FMLP filter / LayerNorm), one B200; raw CSV pages -> gpurun_out/
set -x
OUT=gpurun_out
ncu --set full --clock-control none --import-source off -k 'regex:gru_fwd_tc_kernel|gru_bwd_kernel|gru_order_kernel' -s 12 -c 5 -o $OUT/r2_prof_gru \
    python bench.py --model GRU4Rec --steps 2 --warmup 3 --no-extras > $OUT/r2_prof_gru.log 2>&1
ncu -i $OUT/r2_prof_gru.ncu-rep --page raw --csv > $OUT/r2_prof_gru_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source off -k 'regex:filter_|ln_fwd_kernel|ln_bwd_kernel|embed_dense_kernel' -s 40 -c 14 -o $OUT/r2_prof_fmlp \
    python bench.py --model FMLP --steps 2 --warmup 3 --no-extras > $OUT/r2_prof_fmlp.log 2>&1
ncu -i $OUT/r2_prof_fmlp.ncu-rep --page raw --csv > $OUT/r2_prof_fmlp_raw.csv 2>/dev/null
ls -la $OUT/r2_prof_gru* $OUT/r2_prof_fmlp*
