#!/bin/bash
O=gpurun_out/r2u
mkdir -p $O
DZ_REPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:glu_dwconv -c 1 -o $O/dwconv python scripts/seg_one.py 96 > $O/ncu_dw.log 2>&1; echo "ncu dwconv rc=$?" >> $O/summary.txt
DZ_REPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:relpos_gate -s 5 -c 1 -o $O/gate python scripts/seg_one.py 96 > $O/ncu_gate.log 2>&1; echo "ncu gate rc=$?" >> $O/summary.txt
DZ_REPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'fbank_kernel|emb_conv1_kernel|stats_pool' -c 3 -o $O/embfront DZ_PROFILE=0 python scripts/emb_one.py > $O/ncu_emb.log 2>&1; echo "ncu emb rc=$?" >> $O/summary.txt
cat $O/summary.txt; ls -la $O
