# A/B runs of tuning variants built by scripts/build_variants.py.
# Usage: gpurun -- 'bash scripts/run_variants.sh base v1 v2 ...'
mkdir -p gpurun_out
for v in "$@"; do
  echo "== $v"
  if [ "$v" = base ]; then LIBV=$PWD/phd_qmclib_b200/libqmcb200.so; else LIBV=$PWD/phd_qmclib_b200/variant_$v.so; fi
  QMCB_LIB=$LIBV timeout 300 python scripts/gpu_probe.py one 2>&1 | grep -E "shuffled|Error|error"
done | tee gpurun_out/variants.log
