mkdir -p gpurun_out
for v in "$@"; do
  echo "== $v"; QMCB_LIB=$PWD/phd_qmclib_b200/variant_$v.so timeout 200 python scripts/gpu_probe.py one 2>&1 | grep -E "N=100|Error|error"
done | tee gpurun_out/variants.log
QMCB_KC=1 QMCB_LIB=$PWD/phd_qmclib_b200/variant_qatom.so timeout 200 python scripts/gpu_probe.py one 2>&1 | grep -E "shuffled=False"
