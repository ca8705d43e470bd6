# ncu --set full capture of the batched VMC block kernel (BASELINE configs[1]:
# N=50, 1e5 chains, M=50).  Usage: gpurun -- 'bash scripts/run_ncu_vmc.sh TAG'
mkdir -p gpurun_out
TAG=${1:-x}
CMD="python scripts/bench_configs.py c2s"
timeout 300 $CMD > gpurun_out/c2_plain_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"vmc_block" -s 1 -c 1 \
    -f -o gpurun_out/prof_vmc_$TAG $CMD > gpurun_out/ncu_vmc_$TAG.log 2>&1
cut -c1-300 gpurun_out/c2_plain_$TAG.log; tail -2 gpurun_out/ncu_vmc_$TAG.log
