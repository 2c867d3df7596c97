#!/bin/bash
O=gpurun_out/r2c
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -s > $O/t_all.log 2>&1; echo "all gpu rc=$?" >> $O/summary.txt
grep -E "passed|failed|FAILED|\[large_s80|\[tiny" $O/t_all.log | tail -30
timeout 900 python bench.py --steps 3 --warmup 3 --profile-out $O/prof.json > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2c/bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["clusters_found"], d["breakdown"].get("stages_ms"))
PY
DZ_CONV_LN_UNFUSED=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sub-records > $O/bench_unfused.json 2> $O/bench_unfused.err; echo "bench unfused rc=$?" >> $O/summary.txt
cat $O/summary.txt
