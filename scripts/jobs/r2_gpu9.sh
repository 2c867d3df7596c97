#!/bin/bash
O=gpurun_out/r2h
mkdir -p $O
timeout 900 python -m pytest tests/test_attention_gpu.py -q --timeout 600 > $O/t_attn.log 2>&1; echo "attention tests rc=$?" >> $O/summary.txt
grep -E "passed|failed|FAILED|Error" $O/t_attn.log | tail -12
timeout 1500 python -m pytest tests/test_seg_gpu.py tests/test_bench_config_gpu.py tests/test_glue_golden_gpu.py tests/test_pipeline_gpu.py tests/test_loader_gpu.py -q --timeout 900 -s > $O/t_rest.log 2>&1; echo "other tests rc=$?" >> $O/summary.txt
grep -E "passed|failed|FAILED|\[large_s80|\[tiny" $O/t_rest.log | tail -20
timeout 600 python bench.py --steps 2 --warmup 3 --precision bf16x3 --no-cpu-baseline --no-sub-records > $O/bench_bf16x3.json 2> $O/bench_bf16x3.err; echo "bench bf16x3 rc=$?" >> $O/summary.txt
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2h/bench_bf16x3.json"))
print("bf16x3", round(d["value"],1), round(d["ms_per_step"],1), {k:(v["ms_per_recording"] if isinstance(v,dict) and "ms_per_recording" in v else v) for k,v in d["breakdown"].items()})
PY
cat $O/summary.txt
