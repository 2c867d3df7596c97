#!/bin/bash
O=gpurun_out/r2l
mkdir -p $O
DZ_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tc2 -s 4 -c 1 -o $O/attn python scripts/seg_one.py 96 > $O/ncu_attn.log 2>&1; echo "ncu attn rc=$?" >> $O/summary.txt
DZ_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:layernorm_rows_fast -s 10 -c 2 -o $O/ln python scripts/seg_one.py 96 > $O/ncu_ln.log 2>&1; echo "ncu ln rc=$?" >> $O/summary.txt
cat $O/summary.txt
