mkdir -p gpurun_out
TAG=${1:-x}
timeout 300 python scripts/gpu_probe.py one > gpurun_out/probe_plain_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dmc_step -s 20 -c 1 \
    -f -o gpurun_out/prof_probe_$TAG python scripts/gpu_probe.py one > gpurun_out/ncu_probe_$TAG.log 2>&1
tail -3 gpurun_out/probe_plain_$TAG.log; tail -2 gpurun_out/ncu_probe_$TAG.log
