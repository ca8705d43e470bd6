# ncu launch list + --set full capture of the estimator kernels on the C5
# workload (N=200, S(k) M=400 + density B=6400).
# Usage: gpurun -- 'bash scripts/run_ncu_est.sh TAG'
mkdir -p gpurun_out
TAG=${1:-x}
CMD="python scripts/bench_configs.py c5e"
timeout 300 $CMD > gpurun_out/c5e_plain_$TAG.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_c5e_$TAG.csv $CMD > gpurun_out/ncu_c5e_$TAG.log 2>&1
timeout 300 $CMD > /dev/null 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"ssf_eval|density_hist|colsum" -s 12 -c 6 \
    -f -o gpurun_out/prof_est_$TAG $CMD > gpurun_out/ncu_est_$TAG.log 2>&1
cut -c1-300 gpurun_out/c5e_plain_$TAG.log; tail -2 gpurun_out/ncu_est_$TAG.log
