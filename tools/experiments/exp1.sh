#!/bin/bash
out=gpurun_out; mkdir -p $out
V=offt_b200/lib/variants
{
python __graft_entry__.py smoke 2>&1 | tail -1
python tools/kbench.py 512 64 --copy --modes z,y,x --clogs 0,1,2,3
OFFTB_LIB=$V/lib_lb256x4.so python tools/kbench.py 512 64 --modes z,y,x,xt --clogs 0,1,2
OFFTB_LIB=$V/lib_lb512x2.so python tools/kbench.py 512 64 --modes y,x,xt --clogs 2,3
OFFTB_LIB=$V/lib_lb128x8.so python tools/kbench.py 512 64 --modes z,y,x --clogs 0,1
} > $out/exp1.log 2>&1
cat $out/exp1.log
