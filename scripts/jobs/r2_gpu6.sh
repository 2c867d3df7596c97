#!/bin/bash
O=gpurun_out/r2f
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -s > $O/t_all.log 2>&1; echo "all gpu rc=$?" >> $O/summary.txt
grep -E "passed|failed|FAILED|\[large_s80|\[tiny" $O/t_all.log | tail -30
cat $O/summary.txt
