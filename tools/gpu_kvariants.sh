#!/bin/bash
# strided passes with the experiment builds of tools/variants.sh (L2 prefetch-size hint on the ring fills, runs of adjacent tiles per CTA)
out=gpurun_out/kvar.log; mkdir -p gpurun_out; : > $out
for v in default l2_128 l2_256 q1 q1_l2_128 q1_l2_256 q2_l2_256; do
  lib=offt_b200/lib/variants/lib_$v.so; [ $v = default ] && lib=offt_b200/lib/libofft_b200.so
  for n in 512 1024; do
    echo "== $v $n" >> $out
    OFFTB_LIB=$PWD/$lib timeout 100 python tools/kbench.py $n 64 --modes z,y,x,xs --clogs=-1 2>&1 | grep -E "fft|copy_" >> $out
  done
done
cat $out | cut -c1-120
