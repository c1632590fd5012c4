#!/bin/bash
out=gpurun_out; mkdir -p $out
V=offt_b200/lib/variants
{
for d in 0 1; do echo "# OFFTB_DEPTH=$d (0 = auto)"
OFFTB_DEPTH=$d OFFTB_LIB=$V/lib_lb1024_512.so python tools/kbench.py 512 64 --modes y,x,xt --clogs 3,4 | grep -v "^P1\|^M1\|torch"
done
} > $out/exp14.log 2>&1
cat $out/exp14.log
