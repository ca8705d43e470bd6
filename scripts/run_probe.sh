mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python scripts/gpu_probe.py 2>&1 | tee gpurun_out/probe.log
for kc in 13 7 4 2; do QMCB_KC=$kc timeout 200 python scripts/gpu_probe.py one 2>&1 | grep "W=" ; done | tee -a gpurun_out/probe.log
for nt in 160 256; do QMCB_NT=$nt timeout 200 python scripts/gpu_probe.py one 2>&1 | grep "W=" ; done | tee -a gpurun_out/probe.log
