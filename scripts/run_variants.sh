mkdir -p gpurun_out
v=lean4
for odd in 0 1; do for kc in 1 2 3 4 7; do
  echo "== $v odd=$odd kc=$kc"; QMCB_ODD_ROWS=$odd QMCB_KC=$kc QMCB_LIB=$PWD/phd_qmclib_b200/variant_$v.so timeout 200 python scripts/gpu_probe.py one 2>&1 | grep -E "shuffled=False|Error|error"
done; done | tee gpurun_out/variants.log
