#!/bin/bash
# round-2 first GPU job: parity at the benchmarked sizes, A/B of the opt-in kernels, current profile
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt 2>&1
timeout 900 python -m pytest tests/test_bench_config_gpu.py -x -q > $O/t_benchcfg.log 2>&1; echo "benchcfg rc=$?" >> $O/summary.txt
DZ_CONV0_TC=1 timeout 600 python -m pytest tests/test_seg_gpu.py "tests/test_bench_config_gpu.py::test_large_s80_16s" -q > $O/t_conv0tc.log 2>&1; echo "conv0tc rc=$?" >> $O/summary.txt
DZ_LINKAGE_V2=1 timeout 600 python -m pytest tests/test_post_gpu.py tests/test_bench_config_gpu.py -k linkage -q > $O/t_linkv2.log 2>&1; echo "linkv2 rc=$?" >> $O/summary.txt
DZ_DWCONV_V2=1 timeout 600 python -m pytest tests/test_seg_gpu.py -q > $O/t_dwv2.log 2>&1; echo "dwv2 rc=$?" >> $O/summary.txt
timeout 600 python bench.py --minutes 20 --steps 2 --warmup 3 --no-cpu-baseline --profile-out $O/prof_default.json > $O/b_default.json 2> $O/b_default.err; echo "bench default rc=$?" >> $O/summary.txt
DZ_CONV0_TC=1 timeout 600 python bench.py --minutes 20 --steps 2 --warmup 3 --no-cpu-baseline --profile-out $O/prof_conv0tc.json > $O/b_conv0tc.json 2> $O/b_conv0tc.err; echo "bench conv0tc rc=$?" >> $O/summary.txt
DZ_LINKAGE_V2=1 DZ_DWCONV_V2=1 DZ_TIMING=1 timeout 600 python bench.py --minutes 60 --steps 2 --warmup 3 --no-cpu-baseline --profile-out $O/prof_v2.json > $O/b_v2.json 2> $O/b_v2.err; echo "bench v2 rc=$?" >> $O/summary.txt
DZ_TIMING=1 timeout 600 python bench.py --minutes 60 --steps 2 --warmup 3 --no-cpu-baseline > $O/b_60.json 2> $O/b_60.err; echo "bench 60 rc=$?" >> $O/summary.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/t_all.log 2>&1; echo "all gpu rc=$?" >> $O/summary.txt
cat $O/summary.txt
