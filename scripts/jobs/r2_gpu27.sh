#!/bin/bash
O=gpurun_out/r2w
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x > $O/t_all.log 2>&1; echo "all gpu rc=$?" >> $O/summary.txt
grep -E "passed|failed|FAILED|Error" $O/t_all.log | tail -8
timeout 900 python bench.py --profile-out $O/prof.json > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2w/bench.json"))
print(round(d["value"],1), round(d["ms_per_step"],1), d["e2e"], d["breakdown"].get("stages_ms"), d["roofline"]["frac"], d["cpu_baseline"])
p=json.load(open("gpurun_out/r2w/prof.json"))
print({x["name"]: round(x["ms"],3) for x in p["emb"] if x["name"] in ("fbank","fbank_mean","conv1","stats_pool","seg_1")}, "emb total", round(sum(x["ms"] for x in p["emb"]),2), "seg total", round(sum(x["ms"] for x in p["seg"]),2))
PY
cat $O/summary.txt
