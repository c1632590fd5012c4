#!/bin/bash
# A/B of variant builds (tools/variants.sh) on the local passes: tools/gpu_kvariants.sh "<variant names>" "<lengths>" [modes]
out=gpurun_out/kvar.log; mkdir -p gpurun_out; : > $out
for v in ${1:-default}; do
  lib=offt_b200/lib/variants/lib_$v.so; [ $v = default ] && lib=offt_b200/lib/libofft_b200.so
  for n in ${2:-512 1024}; do
    echo "== $v $n" >> $out
    OFFTB_LIB=$PWD/$lib timeout 100 python tools/kbench.py $n 64 --modes ${3:-z,y,x,xs} --clogs=-1 2>&1 | grep -E "fft|copy_" >> $out
  done
done
cat $out | cut -c1-120
