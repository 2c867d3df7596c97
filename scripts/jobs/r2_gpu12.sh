#!/bin/bash
O=gpurun_out/r2j
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -s > $O/t_all.log 2>&1; echo "all gpu rc=$?" >> $O/summary.txt
grep -E "passed|failed|FAILED|\[large_s80|\[tiny" $O/t_all.log | tail -20
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/summary.txt; tail -2 $O/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 --profile-out $O/prof.json > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2j/bench.json"))
print(round(d["value"],1), round(d["ms_per_step"],1), round(d["e2e"]["value"],1), d["config"]["clusters_found"], d["breakdown"].get("stages_ms"), d["sub_records"], d["cpu_baseline"]["value"])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "bench reference rc=$?" >> $O/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_pipeline.csv python bench.py --minutes 3 --steps 1 --warmup 3 --no-cpu-baseline --no-sub-records > $O/ncu_bench.log 2>&1; echo "ncu launches rc=$?" >> $O/summary.txt
cat $O/summary.txt
