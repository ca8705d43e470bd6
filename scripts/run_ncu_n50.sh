# ncu --set full capture of the step kernel at N=50 (large population).
mkdir -p gpurun_out
TAG=${1:-x}
CMD="python scripts/bench_configs.py c3big"
timeout 300 $CMD > gpurun_out/c3big_plain_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"dmc_step" -s 140 -c 1 \
    -f -o gpurun_out/prof_n50_$TAG $CMD > gpurun_out/ncu_n50_$TAG.log 2>&1
cut -c1-300 gpurun_out/c3big_plain_$TAG.log; tail -2 gpurun_out/ncu_n50_$TAG.log
