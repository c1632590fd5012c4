#!/bin/bash
out=gpurun_out; mkdir -p $out
{
python tools/kbench.py 512 64 --modes x --clogs 2,3 --xpad 8,64,520,4104 | grep -v "^P1\|^M1\|torch"
echo "### 1024"
python tools/kbench.py 1024 64 --modes z,y,x,xt --clogs 0,1,2,3 | grep -v "^P1\|^M1"
OFFTB_DEPTH=2 python tools/kbench.py 1024 64 --modes z,y,x,xt --clogs 0,1,2,3 | grep -v "^P1\|^M1"
OFFTB_DEPTH=1 python tools/kbench.py 1024 64 --modes z,y,x,xt --clogs 0,1,2,3 | grep -v "^P1\|^M1"
} > $out/exp3.log 2>&1
cat $out/exp3.log
