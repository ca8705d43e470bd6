mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/gpu_probe.py one 2>&1 | tee gpurun_out/probe.log
timeout 300 python scripts/gpu_probe.py 2>&1 | grep "W=" | tee -a gpurun_out/probe.log
