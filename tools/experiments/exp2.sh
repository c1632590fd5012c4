#!/bin/bash
out=gpurun_out; mkdir -p $out
V=offt_b200/lib/variants
{
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for d in 1 2 3; do
echo "### default lib depth=$d"
OFFTB_DEPTH=$d python tools/kbench.py 512 64 --modes z --clogs 0,1,2 | grep -v "^P1\|^M1\|torch"
OFFTB_DEPTH=$d python tools/kbench.py 512 64 --modes y,x,xt --clogs 2,3 | grep -v "^P1\|^M1\|torch"
done
for d in 1 2; do
echo "### lb512x2 depth=$d"
OFFTB_DEPTH=$d OFFTB_LIB=$V/lib_lb512x2.so python tools/kbench.py 512 64 --modes z --clogs 0,1,2 | grep -v "^P1\|^M1\|torch"
OFFTB_DEPTH=$d OFFTB_LIB=$V/lib_lb512x2.so python tools/kbench.py 512 64 --modes y,x,xt --clogs 2,3 | grep -v "^P1\|^M1\|torch"
done
} > $out/exp2.log 2>&1
cat $out/exp2.log
