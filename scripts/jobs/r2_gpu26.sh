#!/bin/bash
O=gpurun_out/r2v
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x > $O/t_all.log 2>&1; echo "all gpu rc=$?" >> $O/summary.txt
grep -E "passed|failed|FAILED|Error" $O/t_all.log | tail -8
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sub-records --profile-out $O/prof.json > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
python - <<'PY'
import json, re, collections
d=json.load(open("gpurun_out/r2v/bench.json"))
print(round(d["value"],1), round(d["ms_per_step"],1), d["breakdown"].get("stages_ms"))
p=json.load(open("gpurun_out/r2v/prof.json"))
agg=collections.OrderedDict()
for x in p['seg']:
    k=re.sub(r'^(L|C)\d+_','\\1*_',x['name'])
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=x['ms']
for k in ("L*_ln1","L*_ln2","L*_gate","C*_dwconv","pc_stage","mix_bf","wave_stats","C*_ln","C*_mha_ln","conv4_ln","conv5_ln"):
    print(k, agg.get(k))
print("seg total", round(sum(v[1] for v in agg.values()),2))
PY
cat $O/summary.txt
python - <<'PY'
import json
p=json.load(open("gpurun_out/r2v/prof.json"))
print({x["name"]: round(x["ms"],3) for x in p["emb"] if x["name"] in ("fbank","fbank_mean","conv1","stats_pool","seg_1")}, "emb total", round(sum(x["ms"] for x in p["emb"]),2))
PY
