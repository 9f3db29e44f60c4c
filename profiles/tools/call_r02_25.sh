#!/bin/bash
# round 2, GPU call 25: bench.py wall clock and the Env_2 entries timed as two env ranges
cd /root/repo; mkdir -p gpurun_out
T0=$(date +%s); timeout 900 python bench.py > gpurun_out/r02_25_bench.json 2> gpurun_out/r02_25_bench.err; echo "bench rc=$?"
echo "bench wall clock: $(( $(date +%s) - T0 )) s"; tail -2 gpurun_out/r02_25_bench.err
python - <<'PY'
import json
d = json.loads(open('/root/repo/gpurun_out/r02_25_bench.json').read().strip().splitlines()[-1])
print("value", round(d["value"]/1e9, 2), "frac", round(d["roofline"]["frac"], 3), "traffic", d["roofline"]["traffic"], d["roofline"]["traffic_source"])
for k, v in d["configs"].items():
    print(k, round(v["value"]/1e9, 2), "G", round(v["us_per_step"], 1), "us", round(v["frac"], 3), v["variant"], v.get("one_launch"))
print("rollout", round(d["rollout"]["value"]/1e9, 2), "e2e", round(d["e2e"]["value"]/1e9, 3))
PY
