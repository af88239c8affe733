#!/bin/bash
# Build experiment variants of the library (role ablations, wait-hint, timeline trace) next to the product build.
cd "$(dirname "$0")/.."
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared"
for v in "noload:-DXTB_EXP_NO_LOAD" "nomma:-DXTB_EXP_NO_MMA" "noepi:-DXTB_EXP_NO_EPI" "nohint:-DXTB_BP_WAIT_HINT=0u" "trace:-DXTB_BP_TRACE"; do
  name=${v%%:*}; flag=${v#*:}
  nvcc $F $flag -o xingtian_b200/lib/libxtb200_$name.so xingtian_b200/csrc/xtb_engine.cu 2>&1 | grep -i "error" &
done
wait
ls -la xingtian_b200/lib/
