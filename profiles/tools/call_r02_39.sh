#!/bin/bash
# round 2, GPU call 39: the driver's short form of the bench (--steps 20 --warmup 3), both arms
cd /root/repo; mkdir -p gpurun_out
T0=$(date +%s); timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r02_39_bench20.json 2> gpurun_out/r02_39_bench20.err; echo "bench rc=$? in $(( $(date +%s) - T0 )) s"
T0=$(date +%s); timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/r02_39_ref20.json 2> gpurun_out/r02_39_ref20.err; echo "ref rc=$? in $(( $(date +%s) - T0 )) s"
python - <<'PY'
import json
d = json.loads(open('/root/repo/gpurun_out/r02_39_bench20.json').read().strip().splitlines()[-1])
r = json.loads(open('/root/repo/gpurun_out/r02_39_ref20.json').read().strip().splitlines()[-1])
print("value", round(d["value"]/1e9, 2), "G  steps", d["steps"], " frac", round(d["roofline"]["frac"], 3), " e2e", round(d["e2e"]["value"]/1e9, 3), "G  launches", d["gpu_launches"])
print("reference arm:", r["impl"], round(r["value"]), r["unit"], r["cpu_baseline"]["kind"], r["cpu_baseline"]["cores"], "cores; e2e ratio", round(d["e2e"]["value"] / r["value"]))
PY
