# ncu launch lists + one full capture of the dominant kernel for every bench
# config.  Usage: gpurun -- 'bash scripts/run_ncu_configs.sh TAG'
TAG=${1:-r02}
OUT=gpurun_out; mkdir -p $OUT
MET=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum
run() {  # name, kernel regex for the full capture, skip count, bench args...
  local name=$1 kre=$2 skip=$3; shift 3
  local SHORT="python bench.py --config $name --steps 1 --warmup 1 --no-cpu $@"
  timeout 300 $SHORT > $OUT/short_${name}_$TAG.log 2>&1 || { echo "$name plain run failed"; return; }
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"dmc_|branch_|ssf_|density_|colsum_|rows_|vmc_|multi_" -c 400 --csv \
      --log-file $OUT/launches_${name}_$TAG.csv $SHORT > $OUT/ncu_launch_${name}_$TAG.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kre -s $skip -c 1 --metrics $MET \
      -f -o $OUT/prof_${name}_$TAG $SHORT > $OUT/ncu_full_${name}_$TAG.log 2>&1
  tail -2 $OUT/ncu_full_${name}_$TAG.log
}
run c4 dmc_step 9 --nts 8
run c3_dmc50 dmc_step 20 --nts 16
run c5_est dmc_step 5 --nts 4
run c2_vmc vmc_block 2 --nts 32
# the S(k) kernel of c5 as well
SHORT="python bench.py --config c5_est --steps 1 --warmup 1 --no-cpu --nts 4"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ssf_eval -s 5 -c 1 --metrics $MET \
    -f -o $OUT/prof_c5_ssf_$TAG $SHORT > $OUT/ncu_full_c5_ssf_$TAG.log 2>&1
ls -la $OUT | grep $TAG | tail -20
