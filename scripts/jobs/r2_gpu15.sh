#!/bin/bash
O=gpurun_out/r2m
mkdir -p $O
timeout 600 python -m pytest tests/test_attention_gpu.py -q --timeout 300 > $O/t_attn.log 2>&1; echo "attention tests rc=$?" >> $O/summary.txt
grep -E "passed|failed|FAILED|Error|error" $O/t_attn.log | tail -8
timeout 300 python scripts/attn_bench2.py 2>&1 | tee $O/attn_split.log
DZ_ATTN_ONE_THREAD_PER_ROW=1 timeout 300 python scripts/attn_bench2.py 2>&1 | tee $O/attn_one.log
timeout 900 python -m pytest tests/test_seg_gpu.py tests/test_bench_config_gpu.py tests/test_gemm_gpu.py -q --timeout 600 > $O/t_seg.log 2>&1; echo "seg tests rc=$?" >> $O/summary.txt
grep -E "passed|failed|FAILED" $O/t_seg.log | tail -8
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sub-records > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2m/bench.json"))
print(round(d["value"],1), round(d["ms_per_step"],1), d["breakdown"].get("stages_ms"), d["breakdown"]["seg:attention"], d["roofline"]["groups"].get("cnn_conv1-6"))
PY
cat $O/summary.txt
