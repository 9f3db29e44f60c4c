#!/bin/bash
# round 2, GPU call 5: PCIe table (1 GPU) and the chunk sweep of the host-buffer step
cd /root/repo; mkdir -p gpurun_out
timeout 300 python profiles/tools/bw_pcie.py 2>&1 | tee gpurun_out/r02_5_pcie_1gpu.txt
timeout 600 python profiles/tools/time_e2e.py mono 2>&1 | tee gpurun_out/r02_5_e2e_sweep.txt
nvidia-smi topo -m 2>&1 | head -20 | tee gpurun_out/r02_5_topo.txt
lscpu | head -25 | tee -a gpurun_out/r02_5_topo.txt
