#!/bin/bash
out=gpurun_out; mkdir -p $out
V=offt_b200/lib/variants
{
echo "### default"; python tools/kbench.py 1024 64 --modes z,y,x,xt --clogs -1 | grep -v "^P1\|^M1\|torch"
for v in r32_1024 lb3_1024; do
echo "### $v"
OFFTB_LIB=$V/lib_$v.so python tools/kbench.py 1024 64 --modes z --clogs 0,1 | grep -v "^P1\|^M1\|torch"
for d in 1 2; do echo "# depth $d"; OFFTB_DEPTH=$d OFFTB_LIB=$V/lib_$v.so python tools/kbench.py 1024 64 --modes y,x,xt --clogs 2,3 | grep -v "^P1\|^M1\|torch\|^lib"; done
done
OFFTB_LIB=$V/lib_r32_1024.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rows_contiguous and 1024" 2>&1 | tail -2
} > $out/exp12.log 2>&1
cat $out/exp12.log
