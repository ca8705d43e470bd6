# ncu launch list + one full capture (with source) of the step kernel.
# Usage: gpurun -- 'bash scripts/run_ncu_step.sh TAG [extra bench args]'
TAG=${1:-r02}; shift
OUT=gpurun_out; mkdir -p $OUT
SHORT="python bench.py --steps 1 --warmup 1 --nts 8 --no-cpu $@"
timeout 300 $SHORT > $OUT/short_plain_$TAG.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"dmc_|branch_" -c 300 --csv \
    --log-file $OUT/launches_$TAG.csv $SHORT > $OUT/ncu_launch_$TAG.log 2>&1
timeout 300 $SHORT > $OUT/short_plain2_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dmc_step -s 9 -c 1 \
    --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum \
    -f -o $OUT/prof_step_$TAG $SHORT > $OUT/ncu_full_$TAG.log 2>&1
tail -3 $OUT/ncu_full_$TAG.log
ls -la $OUT | tail -5
