#!/bin/bash
# round 2, GPU call 50: flakiness check — the whole GPU suite three times on one box, plus smoke
cd /root/repo; mkdir -p gpurun_out
for k in 1 2 3; do timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r02_50_run$k.log 2>&1; echo "run $k rc=$? $(tail -1 gpurun_out/r02_50_run$k.log)"; done
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_50_smoke.log 2>&1; echo "smoke rc=$?"
