#!/bin/bash
# round-2 second GPU job: full GPU suite, the new bench line, ncu captures of the top kernels
O=gpurun_out/r2b
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > $O/t_all.log 2>&1; echo "all gpu rc=$?" >> $O/summary.txt
tail -5 $O/t_all.log
timeout 900 python bench.py --steps 3 --warmup 3 --profile-out $O/prof.json > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
tail -c 600 $O/bench.json
# launch list of one short recording (shares) + full captures of the kernels that matter
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches_seg.csv python scripts/seg_one.py 96 > $O/ncu_launch.log 2>&1; echo "ncu launches rc=$?" >> $O/summary.txt
DZ_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv0_tc -c 1 -o $O/conv0tc python scripts/seg_one.py 96 > $O/ncu_conv0.log 2>&1; echo "ncu conv0 rc=$?" >> $O/summary.txt
DZ_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_tma -s 60 -c 6 -o $O/gemm python scripts/seg_one.py 96 > $O/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?" >> $O/summary.txt
DZ_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:glu_dwconv -c 1 -o $O/dwconv python scripts/seg_one.py 96 > $O/ncu_dw.log 2>&1; echo "ncu dwconv rc=$?" >> $O/summary.txt
cat $O/summary.txt
