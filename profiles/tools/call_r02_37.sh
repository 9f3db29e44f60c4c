#!/bin/bash
# round 2, GPU call 37: whole GPU suite + smoke + bench on the final library (tensor policy default = fused kernel)
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_37_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r02_37_gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_37_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_37_smoke.log
timeout 900 python bench.py > gpurun_out/r02_37_bench.json 2> gpurun_out/r02_37_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('/root/repo/gpurun_out/r02_37_bench.json').read().strip().splitlines()[-1])
print("value", round(d["value"]/1e9, 2), "frac", round(d["roofline"]["frac"], 3), "e2e", round(d["e2e"]["value"]/1e9, 3), "rollout", round(d["rollout"]["value"]/1e9, 2))
for k, v in d["configs"].items(): print(k, round(v["value"]/1e9, 2), "G", round(v["us_per_step"], 1), "us", round(v["frac"], 3), v["variant"])
PY
