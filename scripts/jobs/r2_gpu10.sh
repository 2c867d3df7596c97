#!/bin/bash
O=gpurun_out/r2i
mkdir -p $O
timeout 900 python -m pytest tests/test_seg_mc_gpu.py -q --timeout 600 -x > $O/t_mc.log 2>&1; echo "mc tests rc=$?" >> $O/summary.txt
tail -30 $O/t_mc.log
cat $O/summary.txt
