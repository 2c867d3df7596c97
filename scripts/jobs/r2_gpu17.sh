#!/bin/bash
O=gpurun_out/r2o
mkdir -p $O
timeout 1200 python -m pytest tests/test_pipeline_gpu.py tests/test_glue_golden_gpu.py tests/test_loader_gpu.py tests/test_seg_mc_gpu.py -q --timeout 900 > $O/t_pipe.log 2>&1; echo "pipeline tests rc=$?" >> $O/summary.txt
grep -E "passed|failed|FAILED" $O/t_pipe.log | tail -8
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sub-records > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2o/bench.json"))
print(round(d["value"],1), round(d["ms_per_step"],1), d["breakdown"].get("stages_ms"), d["config"]["engine_windows_per_call"])
PY
cat $O/summary.txt
