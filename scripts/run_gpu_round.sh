# GPU-box session: parity tests, smoke, bench, ncu launch list + full capture
# of the step kernel.  Usage: gpurun -- 'bash scripts/run_gpu_round.sh TAG'
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $OUT/pytest_gpu_$TAG.log
tail -3 $OUT/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke_$TAG.log 2>&1; tail -2 $OUT/smoke_$TAG.log
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; tail -c 3000 $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
SHORT="python bench.py --steps 1 --warmup 1 --nts 8 --no-cpu"
timeout 300 $SHORT > $OUT/short_plain_$TAG.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"dmc_|branch_" -c 300 --csv \
    --log-file $OUT/launches_$TAG.csv $SHORT > $OUT/ncu_launch_$TAG.log 2>&1
timeout 300 $SHORT > $OUT/short_plain2_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dmc_step -s 9 -c 2 \
    -f -o $OUT/prof_step_$TAG $SHORT > $OUT/ncu_full_$TAG.log 2>&1
tail -3 $OUT/ncu_full_$TAG.log
ls -la $OUT
