mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in "$@"; do
  for il in 1 0; do
  echo "== $v interleave=$il"; QMCB_INTERLEAVE=$il QMCB_LIB=$PWD/phd_qmclib_b200/variant_$v.so timeout 200 python scripts/gpu_probe.py one 2>&1 | grep -E "N=100|Error|error"
  done
done | tee gpurun_out/variants.log
