#!/bin/bash
# round 2, GPU call 4: 2-round-trip tensor-core policy (6 vs 5 CTAs/SM), full GPU test suite, new bench.py (1 GPU)
cd /root/repo; mkdir -p gpurun_out
V=marl-sortingenv_b200/csrc/variants; L=marl-sortingenv_b200/csrc/libmsort.so
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_4_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r02_4_gpu_tests.log
{
for lib in $L $V/libmsort_tc5.so; do
  N=1048576 TENSOR=1 timeout 200 python profiles/tools/time_variant.py $lib press 2>&1 | tail -2
  N=262144 TENSOR=1 timeout 200 python profiles/tools/time_variant.py $lib press 2>&1 | tail -1
done
for c in 4 5; do N=1048576 TENSOR=1 CTAS=$c timeout 200 python profiles/tools/time_variant.py $L press 2>&1 | tail -1; done
timeout 120 python profiles/tools/time_tc_logits.py $L 2>&1 | tail -1
} | tee gpurun_out/r02_4_press.txt
timeout 900 python bench.py --steps 400 --warmup 5 > gpurun_out/r02_4_bench.json 2> gpurun_out/r02_4_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_4_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_4_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'])
print('e2e', d['e2e']); print('rollout', d.get('rollout',{}).get('value')); print('shard', d['shard_invariance']); print('numa', d['numa'])
for k,v in (d.get('configs') or {}).items(): print(k, {a:v[a] for a in ('value','us_per_step','frac','variant')})
print('cpu', d.get('cpu_baseline'), d.get('cpu_reference'))
PY
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_4_bench_ref.json 2> gpurun_out/r02_4_bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/r02_4_bench_ref.json
