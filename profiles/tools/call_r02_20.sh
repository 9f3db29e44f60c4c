#!/bin/bash
# round 2, multi-GPU call: bench.py on N GPUs (weak scaling, default run) + the PCIe table with all ranks copying at once
# usage: call_r02_20.sh N
N=${1:-8}
cd /root/repo; mkdir -p gpurun_out
P=29517
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N \
  > gpurun_out/bench_r02_${N}gpu.json 2> gpurun_out/bench_r02_${N}gpu.err; echo "bench $N rc=$?"; tail -2 gpurun_out/bench_r02_${N}gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+1)) profiles/tools/bw_pcie.py \
  > gpurun_out/pcie_r02_${N}gpu.txt 2>&1; echo "pcie rc=$?"; tail -12 gpurun_out/pcie_r02_${N}gpu.txt
python - <<PY
import json
d = json.loads(open('/root/repo/gpurun_out/bench_r02_${N}gpu.json').read().strip().splitlines()[-1])
print("N", d["n_gpus"], "value", round(d["value"]/1e9, 2), "G", round(d["ms_per_step"]*1e3, 2), "us  frac", round(d["roofline"]["frac"], 3),
      " e2e", round(d["e2e"]["value"]/1e9, 3), "G  rollout", round(d["rollout"]["value"]/1e9, 2), "G")
PY
