# GPU-box session: parity tests, smoke, the four bench lines, the reference arm.
# Usage: gpurun -- 'bash scripts/run_gpu_round.sh TAG'
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $OUT/pytest_gpu_$TAG.log
tail -3 $OUT/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke_$TAG.log 2>&1; tail -1 $OUT/smoke_$TAG.log
timeout 900 python bench.py > $OUT/bench_${TAG}_c4.json 2> $OUT/bench_${TAG}_c4.err
for c in c3_dmc50 c5_est c2_vmc; do
  timeout 900 python bench.py --config $c > $OUT/bench_${TAG}_$c.json 2> $OUT/bench_${TAG}_$c.err
done
timeout 900 python bench.py --impl reference --steps 8 --warmup 3 > $OUT/bench_${TAG}_reference.json 2> $OUT/bench_${TAG}_reference.err
python - <<PY
import json
for c in ('c4', 'c3_dmc50', 'c5_est', 'c2_vmc', 'reference'):
    try:
        d = json.load(open('$OUT/bench_${TAG}_%s.json' % c))
        cb = d.get('cpu_baseline') or {}
        print(c, '%.4g' % d['value'], d['unit'], 'ms/step %.2f' % d['ms_per_step'],
              'e2e %.4g' % d['e2e']['value'], 'frac', d.get('roofline', {}).get('frac'),
              'cpu', cb.get('kind'), cb.get('value'), cb.get('cores'))
    except Exception as exc:
        print(c, 'FAILED', exc)
PY
