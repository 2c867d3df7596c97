#!/bin/bash
# end-of-round evidence: smoke, ncu launch list of a 3-minute recording, ncu --set full of the kernels rewritten late in the round
O=gpurun_out/r2x
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/summary.txt; tail -2 $O/smoke.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches_pipeline.csv python bench.py --minutes 3 --steps 1 --warmup 3 --no-cpu-baseline --no-sub-records > $O/ncu_bench.log 2>&1; echo "ncu launches rc=$?" >> $O/summary.txt
DZ_PROFILE=0 timeout 200 ncu --set full --clock-control none --import-source on -k regex:'fbank2_kernel|emb_conv1_rows_kernel|stats_pool' -c 3 -o $O/embfront python scripts/emb_one.py > $O/ncu_emb.log 2>&1; echo "ncu emb rc=$?" >> $O/summary.txt
LINK_DATA=hard timeout 200 ncu --set full --clock-control none --import-source on -k regex:'linkage_centroid_lazy|linkage_nn_init' -c 2 -o $O/linkage python scripts/linkage_time.py 8964 > $O/ncu_link.log 2>&1; echo "ncu linkage rc=$?" >> $O/summary.txt
cat $O/summary.txt; ls -la $O | head -20
