#!/bin/bash
out=gpurun_out; mkdir -p $out
{
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_dropin.py -m gpu -x -q 2>&1 | tail -4
for tma in 1 0; do echo "### OFFTB_TMA=$tma"
OFFTB_TMA=$tma python tools/kbench.py 512 64 --clogs -1 | grep -v "^P1\|^M1\|torch"
OFFTB_TMA=$tma python tools/kbench.py 1024 64 --clogs -1 | grep -v "^P1\|^M1\|torch"
done
echo "### TMA depth sweep 512"
for d in 1 2 3; do OFFTB_DEPTH=$d python tools/kbench.py 512 64 --modes z,y,x --clogs 0,1,2,3 | grep -v "^P1\|^M1\|torch\|^lib" | sed "s/^/d$d /"; done
} > $out/exp15.log 2>&1
cat $out/exp15.log
