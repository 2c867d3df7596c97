#!/bin/bash
O=gpurun_out/r2s
mkdir -p $O
timeout 900 python -m pytest tests/test_post_gpu.py tests/test_glue_golden_gpu.py tests/test_pipeline_gpu.py -q --timeout 600 > $O/t.log 2>&1; echo "tests rc=$?" >> $O/summary.txt
grep -E "passed|failed|FAILED" $O/t.log | tail -8
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sub-records > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2s/bench.json"))
print(round(d["value"],1), round(d["ms_per_step"],1), d["breakdown"].get("stages_ms"))
PY
cat $O/summary.txt
