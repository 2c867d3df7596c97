#!/bin/bash
O=gpurun_out/r2r
mkdir -p $O
{
timeout 300 python -m pytest tests/test_post_gpu.py -q --timeout 280 2>&1 | tail -3
echo "== lazy 1024 x 8, hard"; timeout 120 python scripts/linkage_time.py 8964 check
echo "== lazy 512 x 16, hard"; DZ_LINKAGE_NT=512 timeout 120 python scripts/linkage_time.py 8964
for d in easy uniform; do
echo "== lazy 1024 x 8, $d"; LINK_DATA=$d timeout 120 python scripts/linkage_time.py 8964 check
echo "== lazy 512 x 16, $d"; LINK_DATA=$d DZ_LINKAGE_NT=512 timeout 120 python scripts/linkage_time.py 8964
done
echo "== N=2556"; timeout 120 python scripts/linkage_time.py 2556 check
echo "== N=15000 (global state)"; timeout 200 python scripts/linkage_time.py 15000 check
} > $O/linkage2.log 2>&1
cat $O/linkage2.log
